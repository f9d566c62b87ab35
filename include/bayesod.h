/*
 * bayesod.h — C ABI of the B200-native BayesOD post-head path.
 *
 * The reference (asharakeh/bayes-od-rc) has no FFI of its own: the boundary it
 * offers is the Python pair
 *     inference_utils.bayes_od_inference   (src/retina_net/experiments/inference_utils.py:13-217)
 *     inference_utils.bayes_od_clustering  (src/retina_net/experiments/inference_utils.py:285-364)
 * piped into each other by run_inference.py:137-149.  This header declares the
 * entry points a ctypes binding for that pair needs; every
 * function notes which reference lines it stands in for.
 *
 * Conventions
 *   - plain C, no C++/torch types; device pointers are CUDA global-memory
 *     addresses (fp32, row-major, contiguous), owned by the caller, read-only.
 *   - every function returns 0 (BOD_OK) or a negative bod_status; the message of
 *     the last failure on a context is available through bod_last_error().
 *   - a bod_ctx is bound to one device and owns its workspace; contexts are
 *     independent (one per GPU / per pipeline slot), a single context is not
 *     thread-safe.
 *   - there is NO CPU fallback: without a CUDA device bod_create fails.
 */
#ifndef BAYESOD_H
#define BAYESOD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOD_ABI_VERSION 8

typedef enum bod_status {
    BOD_OK            = 0,
    BOD_ERR_INVALID   = -1, /* bad argument / unsupported configuration          */
    BOD_ERR_CUDA      = -2, /* a CUDA runtime call failed (see bod_last_error)   */
    BOD_ERR_NOMEM     = -3, /* workspace allocation failed                       */
    BOD_ERR_STATE     = -4, /* call order violated (fetch before run, ...)       */
    BOD_ERR_OVERFLOW  = -5  /* an image produced more survivors than max_survivors */
} bod_status;

/* cov_layout — how `anchors_box_covar_predictions` is handed over
 * (retinanet_model.py:103-112) */
#define BOD_COV_NONE    0 /* key absent: aleatoric term = 0 (inference_utils.py:83-84) */
#define BOD_COV_FULL16  1 /* [B,N,A,4,4] after tfp.math.fill_triangular              */
#define BOD_COV_PACKED10 2 /* [B,N,A,10] raw head output, fill_triangular order      */

#define BOD_PRIOR_NONE 0
#define BOD_DIRICHLET_NON_INFORMATIVE 1 /* retinanet_bdd.yaml:135 */
#define BOD_GAUSSIAN_ISOTROPIC 1        /* retinanet_bdd.yaml:138 */

#define BOD_RANK_SCORE 0          /* inference_utils.py:202 */
#define BOD_RANK_JOINT_ENTROPY 1  /* inference_utils.py:169-200 */

#define BOD_ANCHORS_TENSOR 0     /* anchors passed as [A,4] (v,u,h,w) */
#define BOD_ANCHORS_GENERATE 1   /* FPN anchors regenerated in-kernel from (im_h, im_w) */

/* Everything run_inference.py:25-29 reads from testing_config, as a plain
 * struct, plus the shapes and the extension knobs of SURVEY.md §8(d). */
typedef struct bod_config {
    int32_t  B;                  /* images per call (the reference is B=1: run_inference.py:68) */
    int32_t  N;                  /* mc_dropout_samples (retinanet_bdd.yaml:68); >= 2 for
                                    bod_run (sample covariance), ignored by bod_validate_run */
    int32_t  A;                  /* anchors per image                                      */
    int32_t  K;                  /* logits per anchor = classes + background (last)        */
    int32_t  cov_layout;         /* BOD_COV_*                                              */
    int32_t  use_full_covar;     /* testing_config.use_full_covar                          */
    int32_t  dirichlet_prior;    /* bayes_od_config.dirichlet_prior.type                   */
    int32_t  gaussian_prior;     /* bayes_od_config.gaussian_prior.type                    */
    float    isotropic_variance; /* bayes_od_config.gaussian_prior.isotropic_variance      */
    int32_t  ranking_method;     /* bayes_od_config.ranking_method                         */
    int32_t  max_output_size;    /* nms_config.max_output_size (<= 255)                    */
    float    iou_threshold;      /* nms_config.iou_threshold; also the affinity threshold
                                    of bayes_od_clustering (run_inference.py:149)          */
    float    soft_nms_sigma;     /* nms_config.soft_nms_sigma                              */
    float    scale_v, scale_u;   /* KITTI rescale orig/resized (inference_utils.py:147-167);
                                    1,1 elsewhere                                          */
    float    cov_calibration;    /* the literal 70 of inference_utils.py:361               */
    int32_t  num_draws;          /* the literal 30 of inference_utils.py:42                */
    uint64_t seed;               /* Philox key; used only when counts == NULL              */
    uint32_t image_id_base;      /* global id of image 0 of this context's shard: RNG
                                    counters are keyed by global image id, so results do
                                    not depend on how a batch is sharded over GPUs        */
    float    score_threshold;    /* EXTENSION (default -INFINITY = off)                    */
    int32_t  pre_nms_top_k;      /* EXTENSION (default 0 = off)                            */
    int32_t  anchor_mode;        /* BOD_ANCHORS_*                                          */
    int32_t  im_h, im_w;         /* image shape for BOD_ANCHORS_GENERATE                   */
    int32_t  max_survivors;      /* per-image survivor capacity; 0 => A                    */
    int32_t  emit_probs;         /* keep the [B,A,K] mean class probabilities (parity)     */
    int32_t  pipeline_depth;     /* 0/1: every bod_run is issued whole, in order, on the
                                    caller's stream.  L = 2..16: L sets of buffers (lanes);
                                    run i+1 streams its logits (K1, scan) on the context's own
                                    stream while runs i, i-1, .. are still computing posteriors,
                                    selecting centres and fusing (K2, soft-NMS, K4), each on its
                                    lane's own stream.  3-4 lanes are enough once a batch fills
                                    the GPU (soft-NMS holds one SM per image); small batches want
                                    more.  The caller's stream then only waits until the logits
                                    have been consumed: `box`, `cov` and `anchors` of run i must
                                    stay untouched until run i's results are complete (bod_fetch
                                    or bod_wait_results), i.e. a streaming producer keeps L sets
                                    of input buffers in flight                                    */
    int32_t  n_levels;           /* 0/1: the head outputs come as one tensor per kind.  2..8:
                                    bod_run_levels may hand them over per FPN level, i.e. BEFORE
                                    the tf.concat(axis=1) of retinanet_model.py:89-112; level l
                                    holds level_anchors[l] anchors, sum = A, order P3 -> P7       */
    int32_t  level_anchors[8];
} bod_config;

typedef struct bod_ctx bod_ctx;

/* Results of bayes_od_clustering for the whole batch, padded to
 * Dmax = max_output_size rows per image (rows >= num_dets[b] are zero).
 * Row d of image b corresponds to nms_indices[b][d] (inference_utils.py:312).
 * Host pointers, caller-allocated (pinned memory makes the copies async);
 * NULL members are skipped. */
typedef struct bod_host_results {
    int32_t* num_dets;          /* [B]          D                                         */
    int32_t* num_survivors;     /* [B]          S                                         */
    float*   means;             /* [B,Dmax,4]   final_box_means (v,u,h,w)                 */
    float*   covs;              /* [B,Dmax,16]  final_box_covs, already x cov_calibration */
    float*   cat_param;         /* [B,Dmax,K]   final_box_class_scores                    */
    float*   cat_count;         /* [B,Dmax,K]   final_box_class_counts                    */
    int32_t* nms_indices;       /* [B,Dmax]     survivor rank of each centre              */
    int32_t* centre_anchor_idx; /* [B,Dmax]     original anchor index of each centre      */
    float*   centre_scores;     /* [B,Dmax]     soft-NMS score at selection time          */
} bod_host_results;

/* The same blocks as device pointers (valid until the next bod_run on ctx; with
 * pipeline_depth = L for the next L-1 runs),
 * for consumers that stay on the GPU (e.g. an NCCL all-gather of detections). */
typedef struct bod_device_results {
    const int32_t* num_dets;
    const int32_t* num_survivors;
    const float*   means;
    const float*   covs;
    const float*   cat_param;
    const float*   cat_count;
    const int32_t* nms_indices;
    const int32_t* centre_anchor_idx;
    const float*   centre_scores;
} bod_device_results;

/* Per-image intermediates of bayes_od_inference, exposed for parity tests and
 * for the drop-in's five return values (inference_utils.py:217).  Host
 * pointers, each sized for `capacity` survivors; NULL members are skipped. */
typedef struct bod_host_survivors {
    int32_t  capacity;       /* in:  rows available in the arrays below                  */
    int32_t  count;          /* out: S                                                   */
    int32_t* anchor_idx;     /* [S]     kept anchor indices, ascending (boolean_mask)    */
    float*   counts;         /* [S,K]   dirichlit_posterior_count  (:89-94)              */
    float*   means;          /* [S,4]   gaussian_posterior_means   (:141-145,160)        */
    float*   covs;           /* [S,16]  gaussian_posterior_covs    (:129,162)            */
    float*   scores;         /* [S]     ranking_scores             (:169-202)            */
    float*   corners;        /* [S,4]   predicted_boxes_corners    (:204)                */
} bod_host_survivors;

int         bod_abi_version(void);
const char* bod_status_string(int status);

/* Allocate a context + workspace on `device` for the given shapes/config. */
int  bod_create(bod_ctx** out, int device, const bod_config* cfg);
void bod_destroy(bod_ctx* ctx);
const char* bod_last_error(const bod_ctx* ctx);
/* Bytes of device workspace owned by the context. */
int64_t bod_workspace_bytes(const bod_ctx* ctx);

/*
 * The hot path: inference_utils.py:25-217 (minus the model call) followed by
 * bayes_od_clustering (:285-364) for B images, entirely on the GPU,
 * asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream).
 *   cls     [B,N,A,K]            anchors_class_predictions (logits)
 *   box     [B,N,A,4]            anchors_box_predictions (deltas)
 *   cov     [B,N,A,16|10] / NULL anchors_box_covar_predictions (see cov_layout)
 *   anchors [A,4] (v,u,h,w), shared by the batch / NULL when anchor_mode=GENERATE
 *   counts  [B,A,K] categorical sample counts to inject (parity mode: replaces the
 *           unseeded tfp Categorical.sample(30) of :37-46) / NULL = in-kernel
 *           Philox4x32-10 sampler keyed by (seed, global image id, anchor)
 */
int bod_run(bod_ctx* ctx, const float* cls, const float* box, const float* cov,
            const float* anchors, const float* counts, void* cuda_stream);

/* Per-run parameters of an existing context (they do not change the workspace, so a caller that streams
 * single images does not need one context per image): the sampler stream of the next runs -- image b of a
 * run draws from Philox counters keyed by (seed, image_id_base + b, anchor) -- and the KITTI rescale factors
 * (inference_utils.py:147-167: original size / network input size, which follow the image). */
int bod_set_sampler_stream(bod_ctx* ctx, uint64_t seed, uint32_t image_id_base);
int bod_set_image_scale(bod_ctx* ctx, float scale_v, float scale_u);
/* Pipelined contexts.  By default bod_run makes the caller's stream wait until the run's logits have been consumed,
 * so the caller may overwrite `cls` right behind the call (box / cov / anchors have to stay untouched until the results
 * are complete in any case).  A streaming producer that keeps ALL inputs of a run untouched until its results are
 * complete (bod_ticket_wait / bod_fetch / bod_wait_results) says so here: the caller's stream is then not touched, runs
 * issued from one stream no longer depend on each other through it, and the moments kernels of consecutive short runs
 * (< 0.8 GB of logits) alternate between two internal streams, so that the next one's CTAs move in while the previous
 * one drains (4 images per run: +10 % images/s; 8: +4 %).  No effect on serial contexts. */
int bod_set_input_hold(bod_ctx* ctx, int enabled);

/* Make `cuda_stream` wait (on the device, without blocking the host) until the
 * results of every bod_run issued so far (and the copies bod_fetch_async enqueued
 * behind them) are complete.  Only needed with pipeline_depth >= 2
 * by consumers that read bod_device_results_of on their own stream; a no-op
 * otherwise (the run is already ordered on the caller's stream). */
int bod_wait_results(bod_ctx* ctx, void* cuda_stream);

/*
 * bod_run with the head outputs still split per FPN level: what the
 * reference's headers return for each pyramid layer before retinanet_model.py:89 / :99 / :109
 * concatenates them along the anchor axis.  cls[l] [B,N,A_l,K], box[l] [B,N,A_l,4],
 * cov[l] [B,N,A_l,16|10] (cov may be NULL when cov_layout = NONE), l < cfg.n_levels, device
 * memory, 16-byte aligned.  The concat copies (the whole [N,A,K+4+16] volume, once more through
 * HBM) are never made; everything else is bod_run.  The context's tile grid follows the level
 * structure, and bod_run (one tensor per kind) keeps working on the same context.
 */
int bod_run_levels(bod_ctx* ctx, const float* const* cls, const float* const* box, const float* const* cov,
                   const float* anchors, const float* counts, void* cuda_stream);

/*
 * The validation post-process: validation_utils.post_process_predictions
 * (src/retina_net/experiments/validation_utils.py:10-77, called at
 * run_validation.py:143-144) for B images, one deterministic sample each, no
 * MC-dropout moments, no fusion: softmax -> arg-max filter -> decode -> the same
 * soft-NMS (ctx's max_output_size / iou_threshold / soft_nms_sigma;
 * validation_utils.py:47-52 hard-codes 100 / 0.5 / 0.5) -> gather.
 *   cls [B,A,K] logits, box [B,A,4] deltas, anchors [A,4] (v,u,h,w), device memory
 *   scaling: NULL (bdd) or how :54-66 rescales the selected corners
 * Results come back through bod_fetch / bod_device_results_of:
 *   cat_param [B,Dmax,K] = predicted_boxes_classes_out   (softmax rows of the selected boxes)
 *   means     [B,Dmax,4] = predicted_boxes_corners_out   (v_min,u_min,v_max,u_max, rescaled)
 *   num_dets, nms_indices, centre_anchor_idx, centre_scores, num_survivors as for bod_run;
 *   covs and cat_count are zero.  bod_fetch_survivors gives the per-survivor softmax rows
 *   (counts), unscaled corners and scores.  Uses the context's single-run buffers
 *   (not pipelined); N of the context is ignored.
 */
#define BOD_VAL_SCALE_NONE  0 /* bdd: corners as decoded (validation_utils.py:65-66)              */
#define BOD_VAL_SCALE_KITTI 1 /* (corners / [h,w,h,w]) * [H0,W0,H0,W0]            (:54-59)        */
#define BOD_VAL_SCALE_COCO  2 /* ((corners - padding) / [h',w',h',w']) * [H0,W0,H0,W0]  (:60-64)  */
typedef struct bod_val_scaling {
    int32_t mode;             /* BOD_VAL_SCALE_*                                                   */
    float   shift[4];         /* coco: sample_dict[IMAGE_PADDING_KEY][0]                           */
    float   norm_h, norm_w;   /* normalize_2d_bounding_boxes divisor (box_utils.py:195-205)        */
    float   scale_h, scale_w; /* expand_2d_bounding_boxes factor     (box_utils.py:208-220)        */
} bod_val_scaling;
int bod_validate_run(bod_ctx* ctx, const float* cls, const float* box, const float* anchors,
                     const bod_val_scaling* scaling, void* cuda_stream);

/* Same call with HOST buffers (what a caller holding numpy arrays makes):
 * stages `cls` host->device in image chunks overlapped with compute, runs the
 * path and copies the padded results back into `out`.  `box` and `cov` are only
 * needed for the survivors: if they live in pinned (device-mapped) host memory
 * their rows are gathered in place over PCIe, otherwise they are copied too.
 * Synchronous. */
int bod_run_host(bod_ctx* ctx, const float* cls, const float* box, const float* cov,
                 const float* anchors, const float* counts, bod_host_results* out);

/* Bytes the last bod_run_host moved: copied host->device, read in place from
 * pinned host memory by the survivor gather, copied device->host. */
int bod_last_host_traffic(const bod_ctx* ctx, int64_t* h2d_copied, int64_t* h2d_gathered, int64_t* d2h);

/* Second half of the drop-in on its own: bayes_od_clustering(...) for ONE image
 * from host arrays (inference_utils.py:285-364).  `affinity` is the [S,S]
 * matrix of the reference signature (only the D centre columns are read, :316). */
int bod_cluster_host(bod_ctx* ctx, int32_t S, const float* counts /*[S,K]*/,
                     const float* means /*[S,4]*/, const float* covs /*[S,16]*/,
                     int32_t D, const int32_t* centres /*[D]*/,
                     const float* affinity /*[S,S]*/, float affinity_threshold,
                     bod_host_results* out /* B=1 blocks */);

/* Wait for the last bod_run and copy the padded result blocks to the host. */
int bod_fetch(bod_ctx* ctx, bod_host_results* out);

/*
 * Streaming retrieval (every run's results, not only the last one's).  A pipelined
 * context keeps pipeline_depth runs in flight, each on its own lane of buffers; the
 * reference hands every image's result on as soon as it exists
 * (run_inference.py:141-161), so a streaming caller needs the same without
 * draining the pipeline:
 *   t = bod_last_ticket(ctx)           ticket of the run just issued (1, 2, 3, ...)
 *   bod_fetch_async(ctx, t, &blocks)   enqueue the device->host copies of run t's padded
 *                                      result blocks behind that run, on its lane's stream
 *                                      (`blocks` should be pinned memory: bod_host_alloc);
 *                                      nothing blocks, later runs keep streaming
 *   bod_ticket_wait(ctx, t)            block the host until run t (and its copies) are complete;
 *                                      returns BOD_ERR_OVERFLOW like bod_fetch
 * A run's device results stay valid until pipeline_depth - 1 further runs have been
 * issued (then its lane is reused): bod_fetch_async / bod_device_results_at fail with
 * BOD_ERR_STATE for tickets older than that.  Copies enqueued by bod_fetch_async are
 * ordered before the lane's next run, so they never see a later run's data.
 */
/* One-copy variant: a lane's result arrays are ONE contiguous device block; bod_result_block_layout gives the
 * byte offsets of num_dets, num_survivors, means, covs, cat_param, cat_count, nms_indices, centre_anchor_idx,
 * centre_scores and the status word inside it (each 16-byte aligned) and its size; bod_fetch_block_async copies
 * the whole block of run `ticket` into `host_block` (pinned, block-sized) with a single cudaMemcpyAsync. */
int bod_result_block_layout(const bod_ctx* ctx, int64_t offsets[10], int64_t* bytes);
int bod_fetch_block_async(bod_ctx* ctx, int64_t ticket, void* host_block);
int64_t bod_last_ticket(const bod_ctx* ctx);
int bod_fetch_async(bod_ctx* ctx, int64_t ticket, bod_host_results* out);
int bod_ticket_wait(bod_ctx* ctx, int64_t ticket);
int bod_device_results_at(bod_ctx* ctx, int64_t ticket, bod_device_results* out);
/* Page-locked host memory for result blocks (cudaMallocHost / cudaFreeHost). */
void* bod_host_alloc(size_t bytes);
void bod_host_free(void* p);
/* Device pointers of the padded result blocks (no sync, no copy). */
int bod_device_results_of(bod_ctx* ctx, bod_device_results* out);

/* Parity / drop-in intermediates of image `b` of the last run (sync + D2H). */
int bod_fetch_survivors(bod_ctx* ctx, int32_t b, bod_host_survivors* out);
/* Cluster membership bitmask of image b: row d (centre d in NMS order) holds
 * ceil(S/32) words, bit s set <=> bbox_iou_vuvu(survivor s, centre d) > thr
 * (inference_utils.py:316).  `words_per_row` is the row pitch of `mask`. */
int bod_fetch_members(bod_ctx* ctx, int32_t b, uint32_t* mask, int32_t words_per_row);
/* [A,K] mean class probabilities of image b (needs cfg.emit_probs). */
int bod_fetch_probs(bod_ctx* ctx, int32_t b, float* probs);
/* [A,K] sampled categorical counts of image b as the Philox sampler drew them
 * (needs cfg.emit_probs and counts == NULL in the last run). */
int bod_fetch_sampled_counts(bod_ctx* ctx, int32_t b, float* counts);

/* cudaDeviceSynchronize() on the context's device.  For producers that do not
 * expose their stream (TensorFlow eager tensors exported through DLPack): call
 * it between the producer and bod_run. */
int bod_synchronize(bod_ctx* ctx);

/* Device-time of the last run per stage in milliseconds (events recorded on the
 * run's stream): [0]=moments/filter, [1]=scan, [2]=posterior, [3]=soft-NMS,
 * [4]=fusion, [5]=total.  Syncs on the run.  Pipelined contexts (pipeline_depth
 * > 1) time the moments kernel only, with the kernel's own launch clock (see
 * bod_moments_clock_accum): [1..5] read as zero there (the stages of
 * consecutive runs overlap, and event records cost a one-image run a visible
 * share of its host time). */
int bod_last_stage_ms(bod_ctx* ctx, float ms[6]);
/* Stage events are recorded by default (6 cudaEventRecord per run); switch them
 * off for latency-critical callers. */
int bod_set_stage_timing(bod_ctx* ctx, int enabled);
/* Sum of the per-stage device times (same layout as bod_last_stage_ms) over the
 * runs issued since the previous call (at most the last 128), and how many runs
 * that was.  Syncs on the last run. */
int bod_stage_ms_accum(bod_ctx* ctx, float sum_ms[6], int32_t* runs);
/* Sum of the device-clock durations (%globaltimer: start of the first CTA to the
 * end of the last one, written by the kernel itself) of the moments-kernel
 * launches since the previous call -- at most the last 64 per lane -- and how
 * many launches that was.  This is what stage [0] of a pipelined context is
 * measured with (timing events around a kernel cost ~10 us between two short
 * launches); serial contexts have it beside their CUDA events.  Syncs. */
int bod_moments_clock_accum(bod_ctx* ctx, double* sum_ms, int32_t* runs);
/* Number of kernels the last bod_run launched. */
int bod_last_launch_count(const bod_ctx* ctx);

/* Batched result writers (host code): the np.save x 4 that ends every iteration of
 * run_inference.py's loop (:241-244), for the B images of a fetched result block, on
 * `nthreads` host threads.  Image b writes <dir>/<sample_ids[b]>.npy in each of the four
 * directories: means [D,4], covs [D,4,4], cat_param [D,K], cat_count [D,K] (float32,
 * byte-identical to numpy.save); an image without detections writes the empty
 * (0,4,1) array four times, as run_inference.py:153-161 does. */
int bod_write_results_npy(const bod_host_results* res, int32_t B, int32_t Dmax, int32_t K,
                          const char* mean_dir, const char* cov_dir, const char* cat_param_dir,
                          const char* cat_count_dir, const char* const* sample_ids, int32_t nthreads);
/* One float32 array as a .npy file (numpy format 1.0, C order). */
int bod_write_npy(const char* path, const float* data, const int32_t* shape, int32_t ndim);

/* BDD / COCO / Pascal predictions.json (host code).  Replaces, for the images of
 * fetched result blocks, validation_utils.py:183-213 (predictions_to_bdd_format: one
 * entry per detection whose first-maximum class is < n_categories; "bbox" =
 * [u_min, v_min, u_max, v_max] of box_utils.vuhw_to_vuvu_np in float32; "score" = that
 * class probability), run_inference.py:206-212 (final_results_list.extend per image)
 * and :258-260 (json.dump(..., indent=4, separators=(',', ': '))).  The file is
 * byte-identical to the reference's: floats are written as Python's float.__repr__
 * of the binary64 value, strings as json's ensure_ascii encoder writes them.
 * open -> append once per result block, in dataset order -> close (which terminates
 * the list; a writer that saw no detection writes "[]").  `res->means` and
 * `res->cat_param` are read ([B,Dmax,4], [B,Dmax,K]); pass the class block
 * map_dataset_classes produced when training and test data sets differ. */
typedef struct bod_json_writer bod_json_writer;
int bod_bdd_json_open(bod_json_writer** out, const char* path, const char* const* categories, int32_t n_categories);
int bod_bdd_json_append(bod_json_writer* w, const bod_host_results* res, int32_t B, int32_t Dmax, int32_t K,
                        const char* const* sample_ids);
int bod_bdd_json_close(bod_json_writer* w);

/* KITTI label files (host code): <dir>/<sample_ids[b]>.txt for every image of a result
 * block, byte-identical to run_inference.py:176-201 writing the rows of
 * validation_utils.py:216-272 (predictions_to_kitti_format) with
 * np.savetxt(..., newline='\r\n', fmt='%s'): "Car" / "Pedestrian" rows for detections
 * whose first-maximum class is 0 / 1, "-1 -1 -10 u_min v_min u_max v_max -10 x7 score",
 * numbers as str(numpy.float32); an image without such a detection gets an empty file. */
int bod_write_results_kitti_txt(const bod_host_results* res, int32_t B, int32_t Dmax, int32_t K, const char* dir,
                                const char* const* sample_ids, int32_t nthreads);
/* The two number formats the text writers use, exposed for tests: style 0 = Python
 * float.__repr__ (json spellings for non-finite values), style 1 = str(numpy.float32(value)).
 * Returns the length written (NUL-terminated) or BOD_ERR_INVALID if `cap` is too small. */
int bod_format_float(double value, int32_t style, char* out, int32_t cap);

/* ---------------------------------------------------------------------------------------
 * PDQ spatial quality: the Gaussian-corner heat maps of
 * offline_eval/pdq_data_holders.py:92-247 (PBoxDetInst.calc_heatmap, find_roi,
 * gen_single_heatmap) and the foreground / background loss sums of offline_eval/pdq.py:199-230,
 * for box-shaped ground truth as bdd/compute_pdq.py:93-124 and kitti/compute_pdq.py build it.
 * A context is bound to one image size; calls are synchronous (results are on the host when
 * they return) and one at a time per context.  All pointers are HOST pointers unless stated.
 *   boxes [D,4] int32   detection corners [x1, y1, x2, y2] (compute_pdq.py:118-120)
 *   covs  [D,2,2,2] f64 the two corner covariances [[var_x, c], [c, var_y]] (top-left, bottom-right; :121)
 * Errors: BOD_ERR_INVALID also where the reference itself raises (a corner mean outside its
 * own 5-sigma window clipped to the image) or a variance is not positive; bod_pdq_last_error
 * names the detection. */
typedef struct bod_pdq_ctx bod_pdq_ctx;
int  bod_pdq_create(bod_pdq_ctx** out, int device, int32_t im_h, int32_t im_w);
void bod_pdq_destroy(bod_pdq_ctx* ctx);
const char* bod_pdq_last_error(const bod_pdq_ctx* ctx);   /* ctx == NULL: why the last bod_pdq_create failed */
/* Dense heat maps, out [D, im_h, im_w] float32 = calc_heatmap of every detection; `out` is a
 * device pointer when out_on_device != 0. */
int bod_pdq_heatmaps(bod_pdq_ctx* ctx, int32_t D, const int32_t* boxes, const double* covs, float* out, int32_t out_on_device);
/* Loss sums for a batch of images without materialising any map.  Image b owns detections
 * [det_offsets[b], det_offsets[b+1]) and ground-truth boxes [gt_offsets[b], gt_offsets[b+1])
 * (gt_boxes [G,4] int32 [x1, y1, x2, y2]: foreground = rows [y1,y2) x columns [x1,x2) as
 * compute_pdq.py:108-110 fills the mask, background = everything outside the inclusive box,
 * pdq.py:162-165).  Outputs: fg_loss / bg_loss = the [G_b, D_b] matrices of _calc_fg_loss /
 * _calc_bg_loss, row-major, concatenated in image order; bg_total [D] = the background term
 * over the whole image (pdq.py:423-424, false-positive spatial quality).  binary64 sums of
 * binary32 terms. */
int bod_pdq_losses(bod_pdq_ctx* ctx, int32_t n_images, const int32_t* det_offsets, const int32_t* boxes, const double* covs,
                   const int32_t* gt_offsets, const int32_t* gt_boxes, double* fg_loss, double* bg_loss, double* bg_total);
/* Device time of the last call in ms: [0] ROI kernel, [1] ROI read-back + table layout + CDF
 * tables, [2] loss sums or dense maps; floats in the CDF tables; kernels launched. */
int bod_pdq_last_ms(const bod_pdq_ctx* ctx, float ms[3], int64_t* table_floats, int64_t* launches);
/* The device's bivariate normal CDF P(X <= h, Y <= k; r) on n points (test hook). */
int bod_pdq_bvn_cdf(bod_pdq_ctx* ctx, int32_t n, const double* h, const double* k, const double* r, double* out);

/*
 * Uncertainty scoring of fused detections (second half): what
 * src/retina_net/offline_eval/{bdd,kitti}/compute_uncertainty_error.py:91-132 computes in Python
 * loops over every detection of the validation set.  Host arrays in and out.
 *
 * bod_entropies: evaluation_utils_2d.py:280-285 compute_gaussian_entropy_np of n covariances
 *   covs [n,4,4] f32 -> gaussian_out [n] f64 (determinant, np.round(., 5) + 1e-12 and log in binary32 as numpy
 *   does on binary32 arrays, then the binary64 constant is added), and
 *   :288-290 compute_categorical_entropy_np of n parameter vectors cat_params [n,K] f32 ->
 *   categorical_out [n] f32.  Either input may be NULL.
 * bod_mu_error: evaluation_utils_2d.py:129-212 compute_mu_error for the predictions of ONE category:
 *   pred_boxes [n,4] f64 [x1,y1,x2,y2], pred_image [n] (image ids 0 .. n_images-1), the ranking
 *   `order` [n] = stable arg-sort of the entropy scores (ascending; rank -> prediction), by_image /
 *   img_off = the ranks grouped by image (img_off [n_images+1]; inside an image ascending),
 *   ground truth as CSR rows per image (gt_off [n_images+1], gt_boxes [G,4] f64), IoU thresholds.
 *   Outputs: min over the whole [n, n_thr] uncertainty-error matrix, its first flat arg-min (the
 *   reference indexes its score list with it, :211-213), and optionally (TP, FP) totals per threshold.
 */
int bod_entropies(int device, int32_t n, int32_t K, const float* covs, const float* cat_params,
                  double* gaussian_out, float* categorical_out);
int bod_mu_error(int device, int32_t n, const double* pred_boxes, const int32_t* pred_image, const int32_t* order,
                 int32_t n_images, const int32_t* img_off, const int32_t* by_image,
                 const int32_t* gt_off, const double* gt_boxes, int32_t n_thr, const double* thresholds,
                 double* min_u_error, int64_t* argmin_flat, double* totals);
const char* bod_uncertainty_last_error(void);

/* FPN anchors exactly as fpn_anchor_generator.py:21-59 produces them for levels
 * 3..7, 3 aspect ratios x 3 scales, concatenated P3->P7
 * (bdd_dataset_handler.py:161-186).  Writes [A,4] to device memory `anchors`
 * and returns A (or a negative status).  Pass anchors=NULL to query A. */
int bod_generate_anchors(int32_t im_h, int32_t im_w, float* anchors_dev, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* BAYESOD_H */
