// bod_kernels.h — argument blocks and launchers of the stage kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

namespace bod {

// ---- launch-time attribute caches: one-image runs issue a kernel every few microseconds, where a
// cudaFuncSetAttribute / occupancy query per launch is a visible share of the host time ----
// cudaFuncAttributeMaxDynamicSharedMemorySize of `fn` on the current device is at least `bytes` afterwards
inline cudaError_t ensure_dyn_smem(const void* fn, size_t bytes) {
    struct Entry { const void* fn; int dev; size_t bytes; };
    static std::mutex mu;
    static std::vector<Entry> seen;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    for (Entry& e : seen)
        if (e.fn == fn && e.dev == dev) {
            if (e.bytes >= bytes) return cudaSuccess;
            cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (r == cudaSuccess) e.bytes = bytes;
            return r;
        }
    cudaError_t r = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (r == cudaSuccess) seen.push_back(Entry{fn, dev, bytes});
    return r;
}
// SMs of the current device
inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < 64) cached[dev] = sms;
    return sms;
}

// Head outputs either as one tensor per kind ([B,N,A,*], n = 1) or as one tensor per FPN level
// ([B,N,A_l,*], concatenated P3 -> P7 along the anchor axis by the reference, retinanet_model.py:82-112):
// level l covers the global anchors [first_anchor[l], first_anchor[l+1]) and the tiles
// [first_tile[l], first_tile[l+1]) of the per-image tile grid (tiles never straddle levels).
constexpr int kMaxLevels = 8;
constexpr int kClkSlots = 64;            // launches a K1 launch clock remembers
struct LevelTable {
    int n;
    int first_anchor[kMaxLevels + 1];   // entries >= n: INT_MAX (a level is found by counting entries <= the index)
    int first_tile[kMaxLevels + 1];
    int count[kMaxLevels];              // anchors of level l
    // where level l lives: tensors of `rows[l]` anchors per sample, the level's first anchor at row `row0[l]`
    // (separate tensors: rows = A_l, row0 = 0; one concatenated tensor viewed per level: rows = A, row0 = first_anchor)
    int rows[kMaxLevels], row0[kMaxLevels];
    const float* cls[kMaxLevels];    // [B,N,rows,K]
    const float* box[kMaxLevels];    // [B,N,rows,4]
    const float* cov[kMaxLevels];    // [B,N,rows,16|10] or nullptr
};

// ---- K1: moments + counts + filter + per-tile compaction ------------------
struct K1Args {
    LevelTable lv;           // cls per level
    const float* counts_in;  // [B,A,K] or nullptr (Philox)
    float* probs_out;        // [B,A,K] or nullptr
    float* sampled_out;      // [B,A,K] or nullptr (Philox counts, parity)
    int32_t* slot_anchor;    // [B,tiles*TILE]    survivors of tile t at [t*TILE, t*TILE+count)
    float* slot_counts;      // [B,tiles*TILE,K]
    int32_t* tile_count;     // [B,tiles]
    int B, N, A, K, tiles;   // tiles per image; the slot lists have tiles*TILE rows per image
    int num_draws;
    uint64_t seed;
    uint32_t image_id_base;
    int debug;               // BOD_DIAGNOSTICS builds only: 1 = skip sampler/compaction, 2 = also skip the softmax
    int ratio_slot;          // filled by launch_k1: slot of the binomial ratio table for num_draws (-1: divide)
    int leave_room;          // pipelined context: size the ring so that other stages' CTAs fit beside K1's
    uint32_t* ticket;        // dynamic tile scheduler: global ticket counter (never reset) ...
    uint32_t ticket_base;    // ... and its value when this launch starts
    unsigned long long* tl;     // BOD_DIAGNOSTICS builds only: the kernel's three timeline stamps (bod_common.cuh), or nullptr
    // Launch clock (pipelined contexts; nullptr: none): clk[0] = launches so far, clk[2 + 2*(i % 64)] / clk[3 + 2*(i % 64)] =
    // %globaltimer at the start of CTA 0 / at the end of the CTA that ended last, of launch i.  A launch's duration costs
    // the host nothing this way; two timing events around the kernel add ~10 us between two short launches (and ~25 us
    // as event nodes of a graph).
    unsigned long long* clk;
};
// tickets a launch of launch_k1 consumes (tiles + one failing fetch per CTA); 0 for the non-pipelined fallback
uint32_t k1_tickets_per_launch(const K1Args& a);
cudaError_t launch_k1(const K1Args& a, cudaStream_t st);
bool k1_supports(int K);
// CUDA graphs: the kernel launch_k1 launches for these arguments; re-pointing a captured launch at other input tensors
const void* k1_kernel_func(const K1Args& a);
cudaError_t k1_graph_update(cudaGraphExec_t exec, cudaGraphNode_t node, const K1Args& a);

// ---- K1b: per-image exclusive scan of tile counts (+ optional pre-NMS top-k) --
struct ScanArgs {
    const int32_t* tile_count;  // [B,tiles]
    int32_t* tile_off;          // [B,tiles+1]
    int32_t* num_survivors;     // [B]
    int32_t* status;            // [1] sticky error flags
    int B, tiles, capacity;
    unsigned long long* tl;     // BOD_DIAGNOSTICS builds only: the kernel's three timeline stamps (bod_common.cuh), or nullptr
};
cudaError_t launch_scan(const ScanArgs& a, cudaStream_t st, bool beside_k1 = false);

// ---- K1c: pre-NMS filter on the slot lists (extension knobs score_threshold / pre_nms_top_k) --
struct PrefilterArgs {
    const int32_t* slot_anchor; // [B,A]   K1's slot lists ...
    const float* slot_counts;   // [B,A,K]
    const int32_t* tile_count;  // [B,tiles]
    const int32_t* tile_off;    // [B,tiles+1] exclusive scan of tile_count
    int32_t* out_anchor;        // [B,A]   ... and the filtered ones (same layout)
    float* out_counts;          // [B,A,K]
    int32_t* out_tile_count;    // [B,tiles]
    unsigned long long* key;    // [B,A] scratch
    unsigned long long* thr_key;  // [B] scratch
    int B, A, K, tiles, slot_stride;   // slot_stride = tiles*TILE rows per image in the slot lists / key
    int dirichlet;              // counts + 1/K before normalising (non_informative prior)
    float score_threshold;      // keep iff score > threshold (-inf: all)
    int top_k;                  // 0 = off
};
cudaError_t launch_prefilter(const PrefilterArgs& a, cudaStream_t st);

// ---- K2: per-survivor posterior -------------------------------------------
struct K2Args {
    LevelTable lv;              // box / cov per level (cov entries nullptr when cov_layout = none)
    const float* anchors;       // [A,4] or nullptr (generate)
    const int32_t* slot_anchor; // [B,tiles*TILE]
    const float* slot_counts;   // [B,tiles*TILE,K]
    const int32_t* tile_off;    // [B,tiles+1]
    const int32_t* num_survivors;  // [B]
    // outputs, [B,cap,*]
    int32_t* surv_anchor;
    float* cnt_post;   // [B,cap,K]
    float* mu_post;    // [B,cap,4]
    float* sig_post;   // [B,cap,16]
    float* score;      // [B,cap]
    float4* corners;   // [B,cap]
    float* info;       // [B,cap,2] (gaussian, categorical) information gains (joint_entropy ranking)
    int B, N, A, K, tiles, capacity;
    int cov_layout, use_full_covar, dirichlet_prior, gaussian_prior, ranking_method;
    float isotropic_variance, scale_v, scale_u;
    int anchor_mode, im_h, im_w;
    unsigned long long* tl;     // BOD_DIAGNOSTICS builds only: the kernel's three timeline stamps (bod_common.cuh), or nullptr
};
cudaError_t launch_k2(const K2Args& a, cudaStream_t st);
const void* k2_kernel_func(const K2Args& a);
cudaError_t k2_graph_update(cudaGraphExec_t exec, cudaGraphNode_t node, const K2Args& a);
// joint_entropy ranking: normalise the information gains over each image's survivors
cudaError_t launch_rank_normalise(const K2Args& a, cudaStream_t st);

// ---- K3: soft-NMS centre selection ------------------------------------------
struct K3Args {
    const float4* corners;        // [B,cap]
    const float* score;           // [B,cap]
    const int32_t* num_survivors; // [B]
    const int32_t* surv_anchor;   // [B,cap]
    // scratch per candidate, [B,cap]
    float* stale;     // score as of the candidate's last queue update
    float* cur;       // up-to-date score
    int32_t* begin;   // suppress_begin_index (generic kernel) / pending-entry counts (round kernel, state in global memory)
    uint32_t* pend;   // [B,cap,kPendStride] pending-selection bitmask (generic kernel)
    float* pw;        // [B,pw_rows,pstride] spill rows of the pending soft-NMS weights (round kernel)
    int pw_rows;      // rows per image in pw: min(capacity, 65535)
    int max_rows;     // survivors the round kernel takes (<= pw_rows); more: the literal kernel
    int pstride;
    // outputs
    int32_t* nms_idx;           // [B,Dmax]
    float* nms_score;           // [B,Dmax]
    int32_t* centre_anchor;     // [B,Dmax]
    int32_t* num_dets;          // [B]
    int B, capacity, Dmax;
    float iou_threshold, soft_nms_sigma;
    int threads;                // CTA size: 256, 512 (default) or 1024
    int force_big;              // tests: per-candidate state in global memory whatever the survivor count
    int psm_max;                // cap on the pending weights kept in shared memory per candidate (-1: none; tests)
    int seg_cap;                // pairs per warp list segment (-1: the kernel's own; tests shrink it, >= 32, to reach the piecewise path)
    long long* dbg;             // BOD_DIAGNOSTICS builds only: [B][32 warps][12] phase cycle counters, or nullptr
    unsigned long long* tl;     // BOD_DIAGNOSTICS builds only: the kernel's three timeline stamps (bod_common.cuh), or nullptr
};
cudaError_t launch_k3(const K3Args& a, cudaStream_t st);

// ---- K4: cluster membership + per-cluster Bayesian fusion -------------------
struct K4Args {
    const float* cnt_post;   // [B,cap,K]
    const float* mu_post;    // [B,cap,4]
    const float* sig_post;   // [B,cap,16]
    const int32_t* num_survivors;
    const int32_t* nms_idx;  // [B,Dmax]
    const int32_t* num_dets; // [B]
    const float4* corners;   // [B,cap] survivor corners: membership rows are computed here (and written to `member`);
                             // nullptr: `member` is an input (bod_cluster_host: the caller's affinity matrix)
    uint32_t* member;        // [B,Dmax,words] bit s of row d <=> bbox_iou_vuvu(s, centre d) > iou_threshold
    float* out_means;        // [B,Dmax,4]
    float* out_covs;         // [B,Dmax,16]
    float* out_param;        // [B,Dmax,K]
    float* out_count;        // [B,Dmax,K]
    int B, K, capacity, Dmax, words;
    float calibration, iou_threshold;
    const int32_t* status_in;   // the context's status word ...
    int32_t* status_out;        // ... copied into the lane's result block by the kernel's first thread (or nullptr)
    unsigned long long* tl;     // BOD_DIAGNOSTICS builds only: the kernel's three timeline stamps (bod_common.cuh), or nullptr
};
cudaError_t launch_k4(const K4Args& a, cudaStream_t st);

// ---- validation post-process (validation_utils.py:10-77): V1 filter, V2 survivors, V3 gather ----
struct ValArgs {
    const float* cls;        // [B,A,K] logits (one sample)
    const float* box;        // [B,A,4] deltas
    const float* anchors;    // [A,4]
    int32_t* slot_anchor; float* slot_counts; int32_t* tile_count;      // V1 out (slot_counts holds the probabilities); tiles*TILE rows per image
    const int32_t* tile_off; const int32_t* num_survivors;              // scan out
    int32_t* surv_anchor; float* cnt_post; float* mu_post; float* score; float4* corners;   // V2 out
    const int32_t* nms_idx; const int32_t* num_dets;                    // K3 out
    float* out_means; float* out_covs; float* out_param; float* out_count;                  // V3 out
    int B, A, K, tiles, capacity, Dmax;
    int scale_mode;          // 0 none, 1 kitti, 2 coco (validation_utils.py:54-66)
    float shift[4]; float norm_h, norm_w, scale_h, scale_w;
};
cudaError_t launch_val_filter(const ValArgs& a, cudaStream_t st);
cudaError_t launch_val_survivors(const ValArgs& a, cudaStream_t st);
cudaError_t launch_val_gather(const ValArgs& a, cudaStream_t st);

// ---- anchors ----------------------------------------------------------------
cudaError_t launch_generate_anchors(int im_h, int im_w, float* anchors, cudaStream_t st);
int count_anchors(int im_h, int im_w);

}  // namespace bod
