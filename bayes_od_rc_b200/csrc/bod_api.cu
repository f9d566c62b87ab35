// bod_api.cu — the C ABI of include/bayesod.h: context, workspace, launch
// sequence of the stage kernels, result transfer.  No CPU fallback anywhere:
// without a CUDA device bod_create fails with BOD_ERR_CUDA.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/bayesod.h"
#include "bod_common.cuh"
#include "bod_kernels.h"

using namespace bod;

// Everything one run writes.  A context owns one lane (cfg.pipeline_depth <= 1) or two: with two, run i+1
// may stream its logits (K1, scan, K2 on the head stream) while run i is still selecting centres and fusing
// clusters (K3, membership, K4 on the tail stream), each on its own set of buffers.
struct Lane {
    // K1 outputs
    int32_t* slot_anchor = nullptr; float* slot_counts = nullptr; int32_t* tile_count = nullptr;
    int32_t* tile_off = nullptr; int32_t* num_survivors = nullptr;
    // K2 outputs
    int32_t* surv_anchor = nullptr; float* cnt_post = nullptr; float* mu_post = nullptr; float* sig_post = nullptr;
    float* score = nullptr; float4* corners = nullptr; float* info = nullptr;
    // K3 scratch + outputs
    float* stale = nullptr; float* cur = nullptr; int32_t* begin = nullptr; uint32_t* pend = nullptr;
    float* pw = nullptr;
    int32_t* nms_idx = nullptr; float* nms_score = nullptr; int32_t* centre_anchor = nullptr; int32_t* num_dets = nullptr;
    uint32_t* member = nullptr;
    // K4 outputs
    float* out_means = nullptr; float* out_covs = nullptr; float* out_param = nullptr; float* out_count = nullptr;
    cudaEvent_t head_done = nullptr, tail_done = nullptr;
    cudaStream_t tail_stream = nullptr;   // the lane's own tail stream: tails of different lanes run concurrently
    uint32_t* tile_ticket = nullptr;      // the lane's own ticket counter of the moments kernel's tile scheduler
    bool tail_pending = false;        // a tail has been issued on this lane (tail_done is meaningful)
    // CUDA-graph replay of a whole run on this lane (pipelined contexts; see issue_run)
    cudaGraph_t graph[2] = {nullptr, nullptr}; cudaGraphExec_t gexec[2] = {nullptr, nullptr};   // head, tail
    cudaGraphNode_t gk1 = nullptr, gk2 = nullptr;     // the moments / posterior kernel nodes (their inputs change per run)
    K1Args ga1{}; K2Args ga2{};                       // what those nodes currently point at
    int uses = 0;                                     // runs issued on this lane (the first one goes through the streams: lazy one-time setup)
    long long ticket = 0;                             // the run whose results the lane holds (0: none yet)
    cudaEvent_t fetch_done = nullptr;                 // after the copies of the lane's last bod_fetch_async
    bool fetch_pending = false;
    int32_t* h_status = nullptr;                      // pinned: the context's status word as of that fetch
    unsigned char* block = nullptr;                   // the lane's result arrays are one contiguous block (bod_result_block_layout)
    int32_t* block_status = nullptr;                  // its last word: the context's status as of the end of the lane's last tail
    const int32_t* h_block_status = nullptr;          // where that word lands on the host (last bod_fetch_block_async)
    unsigned long long* clk = nullptr; // launch clock of the lane's moments kernels (K1Args::clk): [2 + 2 * kClkSlots]
    unsigned long long clk_read = 0;   // launches already folded into a report
    unsigned long long* tl = nullptr; // BOD_TIMELINE (diagnostic builds only): [5 kernels][4] timeline stamps of the lane's last run
};

// what run_range does differently while a lane's graphs are being captured: only one half of the run is
// issued (1: the head -- ticket reset, moments kernel, scan; 2: the tail -- posterior, soft-NMS, fusion)
struct GraphHooks {
    int phase;
    K1Args* k1_out; K2Args* k2_out;   // the arguments the captured moments / posterior kernels were given
};

struct bod_ctx {
    bod_config cfg;
    int device = 0;
    int tiles = 0, capacity = 0, words = 0, Dmax = 0;
    LevelTable levels{};              // level structure (anchors, tiles); the pointers are filled per run
    char err[512] = {0};
    // one slab of device memory, carved up below
    unsigned char* slab = nullptr;
    size_t slab_bytes = 0;
    static constexpr int kMaxLanes = 16;
    Lane lane[kMaxLanes];
    int nlanes = 1, cur = 0;          // cur: lane of the last issued run
    int32_t* status = nullptr;
    // pre-NMS filter (only with score_threshold / pre_nms_top_k): key scratch [B,A], threshold keys [B], and the
    // filtered slot lists K2 reads instead of K1's (the head stream serialises their use across lanes)
    unsigned long long* pf_key = nullptr; unsigned long long* pf_thr = nullptr;
    int32_t* pf_anchor = nullptr; float* pf_counts = nullptr; int32_t* pf_tile_count = nullptr;
    bool prefilter = false;
    float* probs = nullptr; float* sampled = nullptr;
    int pstride = 0, pw_rows = 0, k3_rows = 0, k3_threads = 512, k3_force_big = 0;
    // device staging of host inputs (bod_run_host), allocated on first use
    float* in_cls = nullptr; float* in_box = nullptr; float* in_cov = nullptr; float* in_anchors = nullptr; float* in_counts = nullptr;
    cudaStream_t own_stream = nullptr, copy_stream = nullptr;
    cudaStream_t head_a = nullptr, head_b = nullptr;   // the two head streams of short runs with held inputs (bod_set_input_hold)
    int head_flip = 0;
    bool hold_inputs = false;
    cudaStream_t last_stream = nullptr;
    cudaEvent_t ev_in = nullptr;
    // stage-timing events: a ring of the last kEvRing runs, 7 events each
    // (0 start, 1 K1, 2 scan, 3 K2, 4 soft-NMS + membership, 5 K4, 6 start of the tail)
    static constexpr int kEvRing = 128;
    static constexpr int kEvPerRun = 7;
    cudaEvent_t evring[kEvRing][kEvPerRun] = {{nullptr}};
    cudaEvent_t* ev = nullptr;            // event set of the run being issued
    long long runs_recorded = 0, runs_reported = 0;
    cudaEvent_t ev_copy[4] = {nullptr};
    bool ran = false, used_sampler = false, timing = true, last_timed = false;
    int launches = 0;
    int64_t h2d_copied = 0, h2d_mapped_rows = 0, d2h_copied = 0;   // traffic of the last bod_run_host
    bool host_copy_all = false;       // BOD_HOST_COPY_ALL: never read box/cov in place from pinned host memory
#ifdef BOD_DIAGNOSTICS
    int k1_debug = 0;                 // BOD_K1_DEBUG (diagnostic builds only)
    int skip_mask = 0;                // BOD_DEBUG_SKIP (diagnostic builds only, results invalid): 1 = no K2, 2 = no soft-NMS, 4 = no K4
    long long* k3_dbg = nullptr;      // BOD_K3_DEBUG (diagnostic builds only): [B][32][12] cycle counters
#endif
    bool k2_on_tail = true;           // pipelined contexts: K2 rides with the tail (see run_range); BOD_K2_TAIL=0 keeps it on the head
    bool use_graphs = true;           // pipelined contexts replay each lane's run as a CUDA graph (BOD_GRAPHS=0: stream launches)
    bool force_graphs = false;        // BOD_GRAPHS=2: also for short runs
    long long next_ticket = 0;        // tickets of issued runs: 1, 2, 3, ...
    int64_t block_off[10] = {0};      // result block: offsets of num_dets, num_survivors, means, covs, cat_param, cat_count,
    int64_t block_bytes = 0;          //   nms_indices, centre_anchor_idx, centre_scores, status; total size
    int k3_psm_max = -1, k3_seg_cap = -1;   // BOD_K3_PSM_MAX / BOD_K3_SEGCAP (tests: reach the spill rows / the piecewise pass B on small inputs)
};

static int fail(bod_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap; va_start(ap, fmt);
        vsnprintf(c->err, sizeof c->err, fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CU(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, BOD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" int bod_abi_version(void) { return BOD_ABI_VERSION; }

extern "C" const char* bod_status_string(int s) {
    switch (s) {
        case BOD_OK: return "ok";
        case BOD_ERR_INVALID: return "invalid argument or unsupported configuration";
        case BOD_ERR_CUDA: return "CUDA error";
        case BOD_ERR_NOMEM: return "out of device memory";
        case BOD_ERR_STATE: return "call order violated";
        case BOD_ERR_OVERFLOW: return "more survivors than max_survivors";
        default: return "unknown status";
    }
}

// message of the last failed bod_create on this thread (there is no context to ask)
static thread_local char create_err[512] = {0};
extern "C" const char* bod_last_error(const bod_ctx* ctx) { return ctx ? ctx->err : create_err; }
extern "C" int64_t bod_workspace_bytes(const bod_ctx* ctx) { return ctx ? (int64_t)ctx->slab_bytes : 0; }
extern "C" int bod_last_launch_count(const bod_ctx* ctx) { return ctx ? ctx->launches : 0; }

static int cov_width(int layout) { return layout == BOD_COV_FULL16 ? 16 : (layout == BOD_COV_PACKED10 ? 10 : 0); }

extern "C" int bod_create(bod_ctx** out, int device, const bod_config* cfg) {
    if (!out || !cfg) return BOD_ERR_INVALID;
    *out = nullptr;
    bod_ctx* c = new (std::nothrow) bod_ctx();
    if (!c) return BOD_ERR_NOMEM;
    c->cfg = *cfg;
    c->device = device;
    auto bad = [&](const char* msg) { snprintf(create_err, sizeof create_err, "%s", msg); delete c; return BOD_ERR_INVALID; };
    if (cfg->B < 1 || cfg->B > 128) return bad("B must be in [1,128]");
    if (cfg->N < 1) return bad("N (mc_dropout_samples) must be >= 1");
    if (cfg->A < 1) return bad("A must be positive");
    if (!k1_supports(cfg->K))
        return bad("unsupported K (classes + background): the kernels are instantiated for K = 2..13, 16, 21 and 32 -- BDD (11), KITTI (4 / 8), "
                   "Pascal VOC (21); COCO's 81 columns are not (one thread holds an anchor's K probabilities in registers)");
    if (cfg->max_output_size < 1 || cfg->max_output_size > 255) return bad("max_output_size must be in [1,255]");
    if (cfg->cov_layout < 0 || cfg->cov_layout > 2) return bad("bad cov_layout");
    if (!(cfg->iou_threshold >= 0.0f)) return bad("iou_threshold must be >= 0");
    if (!(cfg->soft_nms_sigma >= 0.0f)) return bad("soft_nms_sigma must be >= 0");
    if (cfg->num_draws < 1 || cfg->num_draws > 4096) return bad("num_draws must be in [1,4096]");
    if (cfg->pre_nms_top_k < 0) return bad("pre_nms_top_k must be >= 0");
    if (cfg->n_levels < 0 || cfg->n_levels > kMaxLevels) return bad("n_levels must be in [0,8]");
    if (cfg->n_levels > 1) {
        long long sum = 0;
        for (int l = 0; l < cfg->n_levels; ++l) {
            if (cfg->level_anchors[l] < 1) return bad("level_anchors must be positive");
            sum += cfg->level_anchors[l];
        }
        if (sum != cfg->A) return bad("level_anchors do not sum to A");
    }
    if (cfg->anchor_mode == BOD_ANCHORS_GENERATE && count_anchors(cfg->im_h, cfg->im_w) != cfg->A)
        return bad("anchor_mode=GENERATE: A does not match the FPN anchor count of (im_h, im_w)");

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { snprintf(create_err, sizeof create_err, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e)); delete c; return BOD_ERR_CUDA; }

    const int B = cfg->B, A = cfg->A, K = cfg->K;
    {
        LevelTable& lv = c->levels;
        lv.n = cfg->n_levels < 1 ? 1 : cfg->n_levels;
        int an = 0, tl = 0;
        for (int l = 0; l < lv.n; ++l) {
            const int A_l = (cfg->n_levels < 1 || cfg->n_levels == 1) ? A : cfg->level_anchors[l];
            lv.first_anchor[l] = an; lv.first_tile[l] = tl; lv.count[l] = A_l;
            an += A_l; tl += (A_l + kTileAnchors - 1) / kTileAnchors;
        }
        lv.first_anchor[lv.n] = an; lv.first_tile[lv.n] = tl;
        // the kernels find a tile's / an anchor's level by counting first_* entries <= it over entries 1..7:
        // entry n (the total) and everything past it must never count
        const int total_anchors = an, total_tiles = tl;
        for (int l = lv.n; l <= kMaxLevels; ++l) { lv.first_anchor[l] = 0x7fffffff; lv.first_tile[l] = 0x7fffffff; }
        c->tiles = total_tiles; (void)total_anchors;
    }
    c->capacity = (cfg->max_survivors > 0 && cfg->max_survivors < A) ? cfg->max_survivors : A;
    c->capacity = (c->capacity + 31) & ~31;
    c->words = c->capacity / 32;
    c->Dmax = cfg->max_output_size;
    const size_t cap = (size_t)c->capacity, D = (size_t)c->Dmax;

    // carve the slab
    c->nlanes = cfg->pipeline_depth < 1 ? 1 : (cfg->pipeline_depth > bod_ctx::kMaxLanes ? bod_ctx::kMaxLanes : cfg->pipeline_depth);
    c->pstride = (c->Dmax + 3) & ~3;
    c->pw_rows = c->capacity < 65535 ? c->capacity : 65535;
    c->k3_rows = c->pw_rows;
    if (const char* md = getenv("BOD_K3_MODE")) {          // tests: force the soft-NMS variants on small inputs
        if (!strcmp(md, "big")) c->k3_force_big = 1;                        // per-candidate state in global memory
        else if (!strcmp(md, "generic")) c->k3_rows = 0;                    // the literal round-per-selection kernel
    }
    if (const char* t = getenv("BOD_K3_THREADS")) { const int v = atoi(t); if (v == 256 || v == 512 || v == 1024) c->k3_threads = v; }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    struct Piece { void** p; size_t o; };
    std::vector<Piece> pieces;
#define TAKE(ptr, bytes) pieces.push_back(Piece{reinterpret_cast<void**>(&ptr), take(bytes)})
    TAKE(c->status, 256);
    for (int l = 0; l < bod_ctx::kMaxLanes; ++l) TAKE(c->lane[l].tile_ticket, 256);
    c->prefilter = cfg->pre_nms_top_k > 0 || cfg->score_threshold > -INFINITY;
    if (c->prefilter) {
        const size_t slots = (size_t)c->tiles * kTileAnchors;
        TAKE(c->pf_key, B * slots * 8); TAKE(c->pf_thr, (size_t)B * 8);
        TAKE(c->pf_anchor, B * slots * 4); TAKE(c->pf_counts, B * slots * K * 4);
        TAKE(c->pf_tile_count, (size_t)B * c->tiles * 4);
    }
    if (cfg->emit_probs) { TAKE(c->probs, (size_t)B * A * K * 4); TAKE(c->sampled, (size_t)B * A * K * 4); }
    {
        // every lane's result arrays in one block, so that a run's results leave the device in ONE copy
        const size_t sz[10] = {(size_t)B * 4, (size_t)B * 4, (size_t)B * D * 16, (size_t)B * D * 64, (size_t)B * D * K * 4,
                               (size_t)B * D * K * 4, (size_t)B * D * 4, (size_t)B * D * 4, (size_t)B * D * 4, 16};
        int64_t o = 0;
        for (int i = 0; i < 10; ++i) { c->block_off[i] = o; o += (int64_t)((sz[i] + 15) & ~(size_t)15); }
        c->block_bytes = o;
    }
    for (int l = 0; l < c->nlanes; ++l) {
        Lane& L = c->lane[l];
        TAKE(L.slot_anchor, (size_t)B * c->tiles * kTileAnchors * 4);
        TAKE(L.slot_counts, (size_t)B * c->tiles * kTileAnchors * K * 4);
        TAKE(L.tile_count, (size_t)B * c->tiles * 4);
        TAKE(L.tile_off, (size_t)B * (c->tiles + 1) * 4);
        TAKE(L.surv_anchor, B * cap * 4);
        TAKE(L.cnt_post, B * cap * K * 4);
        TAKE(L.mu_post, B * cap * 16);
        TAKE(L.sig_post, B * cap * 64);
        TAKE(L.score, B * cap * 4);
        TAKE(L.corners, B * cap * 16);
        TAKE(L.info, B * cap * 8);
        TAKE(L.stale, B * cap * 4);
        TAKE(L.cur, B * cap * 4);
        TAKE(L.begin, B * cap * 4);
        TAKE(L.pend, B * cap * kPendStride * 4);
        TAKE(L.pw, (size_t)B * c->pw_rows * c->pstride * 4);
        TAKE(L.block, (size_t)c->block_bytes);
        TAKE(L.member, B * D * c->words * 4);
    }
#undef TAKE
    c->slab_bytes = off;
    e = cudaMalloc(&c->slab, off);
    if (e != cudaSuccess) {
        snprintf(create_err, sizeof create_err, "cudaMalloc(%zu): %s", off, cudaGetErrorString(e));
        cudaGetLastError(); delete c; return BOD_ERR_NOMEM;
    }
    for (auto& p : pieces) *p.p = c->slab + p.o;
    for (int l = 0; l < c->nlanes; ++l) {
        Lane& L = c->lane[l];
        unsigned char* b = L.block;
        L.num_dets = reinterpret_cast<int32_t*>(b + c->block_off[0]); L.num_survivors = reinterpret_cast<int32_t*>(b + c->block_off[1]);
        L.out_means = reinterpret_cast<float*>(b + c->block_off[2]); L.out_covs = reinterpret_cast<float*>(b + c->block_off[3]);
        L.out_param = reinterpret_cast<float*>(b + c->block_off[4]); L.out_count = reinterpret_cast<float*>(b + c->block_off[5]);
        L.nms_idx = reinterpret_cast<int32_t*>(b + c->block_off[6]); L.centre_anchor = reinterpret_cast<int32_t*>(b + c->block_off[7]);
        L.nms_score = reinterpret_cast<float*>(b + c->block_off[8]); L.block_status = reinterpret_cast<int32_t*>(b + c->block_off[9]);
    }
    cudaMemset(c->slab, 0, off);
    cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (c->nlanes > 1) {
        // two more head streams, of different priority, for short runs with held inputs (bod_set_input_hold): the next
        // run's moments kernel is queued while the current one still runs, and its CTAs move in as the current one's
        // retire; the priorities break the tie when two grids become eligible at the same moment
        cudaStreamCreateWithPriority(&c->head_a, cudaStreamNonBlocking, -1);
        cudaStreamCreateWithPriority(&c->head_b, cudaStreamNonBlocking, 0);
    }
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    {
        int lo = 0, hi = 0;                                  // the tail is latency-bound and short: let its CTAs go first
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (int l = 0; l < c->nlanes; ++l) cudaStreamCreateWithPriority(&c->lane[l].tail_stream, cudaStreamNonBlocking, hi);
    }
    cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming);
    for (int l = 0; l < c->nlanes; ++l) {
        cudaEventCreateWithFlags(&c->lane[l].head_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->lane[l].tail_done, cudaEventDisableTiming);
    }
    for (auto& set : c->evring) for (auto& ev : set) cudaEventCreate(&ev);
    for (auto& ev : c->ev_copy) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (int l = 0; l < c->nlanes; ++l) {
        cudaMalloc(&c->lane[l].clk, (2 + 2 * kClkSlots) * sizeof(unsigned long long));
        cudaMemset(c->lane[l].clk, 0, (2 + 2 * kClkSlots) * sizeof(unsigned long long));
    }
    c->timing = getenv("BOD_NO_STAGE_EVENTS") == nullptr;
#ifdef BOD_DIAGNOSTICS
    if (const char* d = getenv("BOD_K1_DEBUG")) c->k1_debug = atoi(d);
    if (const char* d = getenv("BOD_DEBUG_SKIP")) c->skip_mask = atoi(d);
    if (getenv("BOD_TIMELINE"))
        for (int l = 0; l < c->nlanes; ++l) { cudaMalloc(&c->lane[l].tl, 20 * sizeof(unsigned long long)); cudaMemset(c->lane[l].tl, 0, 20 * sizeof(unsigned long long)); }
    if (getenv("BOD_K3_DEBUG")) { cudaMalloc(&c->k3_dbg, ((size_t)B * 384 + 8) * sizeof(long long)); cudaMemset(c->k3_dbg, 0, ((size_t)B * 384 + 8) * sizeof(long long)); }
#endif
    c->host_copy_all = getenv("BOD_HOST_COPY_ALL") != nullptr;
    if (const char* d = getenv("BOD_K2_TAIL")) c->k2_on_tail = atoi(d) != 0;
    if (const char* d = getenv("BOD_GRAPHS")) { c->use_graphs = atoi(d) != 0; c->force_graphs = atoi(d) >= 2; }
    for (int l = 0; l < c->nlanes; ++l) {
        cudaEventCreateWithFlags(&c->lane[l].fetch_done, cudaEventDisableTiming);
        cudaMallocHost(&c->lane[l].h_status, sizeof(int32_t));
        if (c->lane[l].h_status) *c->lane[l].h_status = 0;
    }
    if (const char* d = getenv("BOD_K3_PSM_MAX")) c->k3_psm_max = atoi(d);
    if (const char* d = getenv("BOD_K3_SEGCAP")) c->k3_seg_cap = atoi(d);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { snprintf(create_err, sizeof create_err, "init: %s", cudaGetErrorString(e)); bod_destroy(c); return BOD_ERR_CUDA; }
    *out = c;
    return BOD_OK;
}

extern "C" void bod_destroy(bod_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->slab) cudaFree(c->slab);
    for (float* p : {c->in_cls, c->in_box, c->in_cov, c->in_anchors, c->in_counts}) if (p) cudaFree(p);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->head_a) cudaStreamDestroy(c->head_a);
    if (c->head_b) cudaStreamDestroy(c->head_b);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto& L : c->lane) if (L.tail_stream) cudaStreamDestroy(L.tail_stream);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    for (auto& L : c->lane) {
        if (L.head_done) cudaEventDestroy(L.head_done);
        if (L.tail_done) cudaEventDestroy(L.tail_done);
        if (L.clk) cudaFree(L.clk);
        if (L.fetch_done) cudaEventDestroy(L.fetch_done);
        if (L.h_status) cudaFreeHost(L.h_status);
        for (int i = 0; i < 2; ++i) {
            if (L.gexec[i]) cudaGraphExecDestroy(L.gexec[i]);
            if (L.graph[i]) cudaGraphDestroy(L.graph[i]);
        }
    }
    for (auto& set : c->evring) for (auto& ev : set) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_copy) if (ev) cudaEventDestroy(ev);
    delete c;
}

// Launch the stage kernels for images [b0, b0+nb) of the context's batch on lane L.  The head (K1, scan,
// K2) goes to stream `hs`, the tail (soft-NMS, membership, K4) to `ts`; hs == ts runs them back to back.
// one tensor per kind (already offset to the first image of the range), seen through the context's level structure
static LevelTable levels_concat(const bod_ctx* c, const float* cls, const float* box, const float* cov) {
    LevelTable lv = c->levels;
    for (int l = 0; l < lv.n; ++l) {
        lv.rows[l] = c->cfg.A; lv.row0[l] = lv.first_anchor[l];
        lv.cls[l] = cls; lv.box[l] = box; lv.cov[l] = cov;
    }
    return lv;
}
// one tensor per kind and FPN level
static LevelTable levels_split(const bod_ctx* c, const float* const* cls, const float* const* box, const float* const* cov) {
    LevelTable lv = c->levels;
    for (int l = 0; l < lv.n; ++l) {
        lv.rows[l] = lv.count[l]; lv.row0[l] = 0;
        lv.cls[l] = cls[l]; lv.box[l] = box[l]; lv.cov[l] = cov ? cov[l] : nullptr;
    }
    return lv;
}

static int run_range(bod_ctx* c, Lane& L, int b0, int nb, const LevelTable& lv,
                     const float* anchors, const float* counts, cudaStream_t hs, cudaStream_t ts, int record,
                     const GraphHooks* gh = nullptr) {   // record: 0 no stage events, 1 moments kernel only, 2 every stage
    const bod_config& g = c->cfg;
    const size_t A = g.A, K = g.K, cap = c->capacity, D = c->Dmax;
    const size_t slots = (size_t)c->tiles * kTileAnchors;
    const int cw = cov_width(g.cov_layout);

    // K2 with the tail: the head stream then carries K1 + scan only, so the next run's K1 starts as soon as
    // this one's logits are consumed (not with the pre-NMS filter: its scratch is shared between lanes)
    const bool k2_tail = (hs != ts) && c->k2_on_tail && !c->prefilter;
    // The tail issued on this lane nlanes runs ago may still be reading the lane's state (slot lists by K2,
    // num_survivors / tile_off by the soft-NMS and fusion kernels): nothing of this run may overwrite it before
    // that tail is done -- also when K2 stays on the head stream (pre-NMS filter), where the scans would otherwise
    // rewrite num_survivors under a running tail.
    if (hs != ts && L.tail_pending) CU(c, cudaStreamWaitEvent(hs, L.tail_done, 0));
    const bool do_head = !gh || (gh->phase & 1), do_tail = !gh || (gh->phase & 2);   // graph capture: one half at a time, or both (3)
    // (the moments kernel's tile scheduler counts tickets from zero: the slab starts zeroed and every launch puts
    // the counter back itself; launches of a context never overlap)
    if (record) CU(c, cudaEventRecord(c->ev[0], hs));
    K1Args k1{};
    k1.lv = lv; k1.counts_in = counts;
    k1.probs_out = c->probs ? c->probs + b0 * A * K : nullptr;
    k1.sampled_out = (c->sampled && !counts) ? c->sampled + b0 * A * K : nullptr;
    k1.slot_anchor = L.slot_anchor + b0 * slots; k1.slot_counts = L.slot_counts + b0 * slots * K;
    k1.tile_count = L.tile_count + (size_t)b0 * c->tiles;
    k1.B = nb; k1.N = g.N; k1.A = g.A; k1.K = g.K; k1.tiles = c->tiles;
    k1.num_draws = g.num_draws; k1.seed = g.seed; k1.image_id_base = g.image_id_base + (uint32_t)b0;
#ifdef BOD_DIAGNOSTICS
    k1.debug = c->k1_debug;
#endif
    k1.leave_room = (hs != ts || gh) ? 1 : 0;
    k1.ticket = L.tile_ticket; k1.ticket_base = 0u;
    k1.tl = L.tl;
    k1.clk = L.clk;
    if (do_head) CU(c, launch_k1(k1, hs));
    if (gh && do_head) *gh->k1_out = k1;
    if (record) CU(c, cudaEventRecord(c->ev[1], hs));

    int launches = 3;
    const int32_t* slot_anchor = k1.slot_anchor;      // the slot lists K2 reads: K1's, or the filtered ones
    const float* slot_counts = k1.slot_counts;
    ScanArgs sc{};
    sc.tile_count = k1.tile_count; sc.tile_off = L.tile_off + (size_t)b0 * (c->tiles + 1);
    sc.num_survivors = L.num_survivors + b0; sc.status = c->status;
    sc.B = nb; sc.tiles = c->tiles; sc.capacity = c->capacity;
    sc.tl = L.tl ? L.tl + 4 : nullptr;
    if (c->prefilter && do_head) {
        // scan (dense indexing of the slots) -> filter -> scan again on the new counts; the capacity check
        // belongs to the second scan only
        ScanArgs s0 = sc;
        s0.capacity = 0x7fffffff;
        CU(c, launch_scan(s0, hs));
        PrefilterArgs pf{};
        pf.slot_anchor = k1.slot_anchor; pf.slot_counts = k1.slot_counts; pf.tile_count = k1.tile_count;
        pf.tile_off = sc.tile_off;
        pf.out_anchor = c->pf_anchor + b0 * slots; pf.out_counts = c->pf_counts + b0 * slots * K;
        pf.out_tile_count = c->pf_tile_count + (size_t)b0 * c->tiles;
        pf.key = c->pf_key + b0 * slots; pf.thr_key = c->pf_thr + b0;
        pf.B = nb; pf.A = g.A; pf.K = g.K; pf.tiles = c->tiles; pf.slot_stride = (int)slots;
        pf.dirichlet = g.dirichlet_prior == BOD_DIRICHLET_NON_INFORMATIVE;
        pf.score_threshold = g.score_threshold; pf.top_k = g.pre_nms_top_k;
        CU(c, launch_prefilter(pf, hs));
        sc.tile_count = pf.out_tile_count;
        slot_anchor = pf.out_anchor; slot_counts = pf.out_counts;
        launches += 4;
    }
    // K2 overwrites what the previous tail on this lane (two runs ago) reads
    cudaStream_t k2s = hs;
    // Short moments kernels (small batches: < 0.8 GB of logits per run, ~0.12 ms) leave the head stream idle for a
    // noticeable share of the step while the scan runs between two of them; there the scan (per-lane outputs, read by K2
    // onwards) rides with the tail and the head stream carries K1 back to back (B = 8, K = 8: 69.0 -> 73.8 k images/s).
    // Long ones lose ~1 % that way (the soft-NMS CTAs then find no free SM at the kernel boundary), so they keep the scan.
    static const int scan_tail_env = getenv("BOD_SCAN_TAIL") ? atoi(getenv("BOD_SCAN_TAIL")) : -1;     // experiments
    const bool small_run = scan_tail_env >= 0 ? scan_tail_env != 0 : 4.0 * nb * g.N * A * K < 0.8e9;
    const bool scan_tail = k2_tail && small_run;
    // graph replay: the scan stays with the head (measured: with the scan opening the tail graph the posterior kernel of a
    // small run only finds room when the NEXT moments kernel ends, and every small-batch workload loses 10-25 %);
    // BOD_SCAN_TAIL=1 moves it for experiments
    const bool scan_in_tail_graph = gh && scan_tail_env > 0 && !c->prefilter;
    if (!scan_tail && !scan_in_tail_graph && do_head) {
        CU(c, launch_scan(sc, hs));
        if (record > 1) CU(c, cudaEventRecord(c->ev[2], hs));
    }
    // one graph for the whole run: "the head is done" is an event node between the scan and the posterior kernel
    if (gh && gh->phase == 3) CU(c, cudaEventRecordWithFlags(L.head_done, hs, cudaEventRecordExternal));
    if (!do_tail) return BOD_OK;
    if (scan_in_tail_graph) CU(c, launch_scan(sc, ts, true));
    if (k2_tail) {
        CU(c, cudaEventRecord(L.head_done, hs));
        CU(c, cudaStreamWaitEvent(ts, L.head_done, 0));
        k2s = ts;
        if (scan_tail) {
            CU(c, launch_scan(sc, ts, true));
            if (record > 1) CU(c, cudaEventRecord(c->ev[2], ts));
        }
    }
    K2Args k2{};
    k2.lv = lv; k2.anchors = anchors;
    if (!cw) for (int l = 0; l < k2.lv.n; ++l) k2.lv.cov[l] = nullptr;
    k2.slot_anchor = slot_anchor; k2.slot_counts = slot_counts; k2.tile_off = sc.tile_off;
    k2.num_survivors = sc.num_survivors;
    k2.surv_anchor = L.surv_anchor + b0 * cap; k2.cnt_post = L.cnt_post + b0 * cap * K;
    k2.mu_post = L.mu_post + b0 * cap * 4; k2.sig_post = L.sig_post + b0 * cap * 16;
    k2.score = L.score + b0 * cap; k2.corners = L.corners + b0 * cap; k2.info = L.info + b0 * cap * 2;
    k2.B = nb; k2.N = g.N; k2.A = g.A; k2.K = g.K; k2.tiles = c->tiles; k2.capacity = c->capacity;
    k2.cov_layout = g.cov_layout; k2.use_full_covar = g.use_full_covar;
    k2.dirichlet_prior = g.dirichlet_prior; k2.gaussian_prior = g.gaussian_prior;
    // joint_entropy needs both priors (inference_utils.py:169-170), else falls back to 'score'
    k2.ranking_method = (g.ranking_method == BOD_RANK_JOINT_ENTROPY && g.gaussian_prior != BOD_PRIOR_NONE &&
                         g.dirichlet_prior != BOD_PRIOR_NONE) ? 1 : 0;
    k2.isotropic_variance = g.isotropic_variance; k2.scale_v = g.scale_v; k2.scale_u = g.scale_u;
    k2.anchor_mode = g.anchor_mode; k2.im_h = g.im_h; k2.im_w = g.im_w;
    k2.tl = L.tl ? L.tl + 8 : nullptr;
#ifdef BOD_DIAGNOSTICS
    if (!(c->skip_mask & 1))
#endif
    CU(c, launch_k2(k2, k2s));
    if (gh) *gh->k2_out = k2;
    if (k2.ranking_method == 1) { CU(c, launch_rank_normalise(k2, k2s)); ++launches; }
    if (record > 1) CU(c, cudaEventRecord(c->ev[3], k2s));
    if (hs != ts && !k2_tail) {
        CU(c, cudaEventRecord(L.head_done, hs));
        CU(c, cudaStreamWaitEvent(ts, L.head_done, 0));
    }
    if (record > 1) CU(c, cudaEventRecord(c->ev[6], ts));

    K3Args k3{};
    k3.corners = k2.corners; k3.score = k2.score; k3.num_survivors = sc.num_survivors; k3.surv_anchor = k2.surv_anchor;
    k3.stale = L.stale + b0 * cap; k3.cur = L.cur + b0 * cap; k3.begin = L.begin + b0 * cap;
    k3.pend = L.pend + b0 * cap * kPendStride;
    k3.pw = L.pw + (size_t)b0 * c->pw_rows * c->pstride; k3.pw_rows = c->pw_rows; k3.max_rows = c->k3_rows;
    k3.pstride = c->pstride;
    k3.nms_idx = L.nms_idx + b0 * D; k3.nms_score = L.nms_score + b0 * D; k3.centre_anchor = L.centre_anchor + b0 * D;
    k3.num_dets = L.num_dets + b0;
    k3.B = nb; k3.capacity = c->capacity; k3.Dmax = c->Dmax;
    k3.iou_threshold = g.iou_threshold; k3.soft_nms_sigma = g.soft_nms_sigma;
    k3.threads = c->k3_threads; k3.force_big = c->k3_force_big; k3.psm_max = c->k3_psm_max; k3.seg_cap = c->k3_seg_cap;
    k3.tl = L.tl ? L.tl + 12 : nullptr;
#ifdef BOD_DIAGNOSTICS
    k3.dbg = c->k3_dbg ? c->k3_dbg + (size_t)b0 * 384 : nullptr;
    if (!(c->skip_mask & 2))
#endif
    CU(c, launch_k3(k3, ts));
    if (record > 1) CU(c, cudaEventRecord(c->ev[4], ts));

    K4Args k4{};
    k4.cnt_post = k2.cnt_post; k4.mu_post = k2.mu_post; k4.sig_post = k2.sig_post; k4.num_survivors = sc.num_survivors;
    k4.nms_idx = k3.nms_idx; k4.num_dets = k3.num_dets;
    k4.corners = k2.corners; k4.member = L.member + b0 * D * c->words;
    k4.out_means = L.out_means + b0 * D * 4; k4.out_covs = L.out_covs + b0 * D * 16;
    k4.out_param = L.out_param + b0 * D * K; k4.out_count = L.out_count + b0 * D * K;
    k4.B = nb; k4.K = g.K; k4.capacity = c->capacity; k4.Dmax = c->Dmax; k4.words = c->words;
    k4.calibration = g.cov_calibration; k4.iou_threshold = g.iou_threshold;
    k4.tl = L.tl ? L.tl + 16 : nullptr;
    // the context's (sticky) status word rides in the lane's result block, so a block fetch brings it along
    k4.status_in = c->status; k4.status_out = (b0 + nb == c->cfg.B) ? L.block_status : nullptr;
#ifdef BOD_DIAGNOSTICS
    if (!(c->skip_mask & 4))
#endif
    CU(c, launch_k4(k4, ts));
    if (record > 1) CU(c, cudaEventRecord(c->ev[5], ts));
    if (hs != ts) { CU(c, cudaEventRecord(L.tail_done, ts)); L.tail_pending = true; }
    c->launches += launches + 2;   // + soft-NMS, K4 (membership + fusion)
    return BOD_OK;
}

// wait for every tail in flight (pipelined contexts) before a synchronous entry reuses lane 0
static int drain_tails(bod_ctx* c) {
    if (c->nlanes > 1)
        for (int l = 0; l < c->nlanes; ++l) CU(c, cudaStreamSynchronize(c->lane[l].tail_stream));
    return BOD_OK;
}

static int check_inputs(bod_ctx* c, const float* cls, const float* box, const float* cov, const float* anchors) {
    if (!c) return BOD_ERR_INVALID;
    if (!cls || !box) return fail(c, BOD_ERR_INVALID, "cls and box must not be NULL");
    if (c->cfg.cov_layout != BOD_COV_NONE && !cov) return fail(c, BOD_ERR_INVALID, "cov is NULL but cov_layout != NONE");
    if (c->cfg.anchor_mode == BOD_ANCHORS_TENSOR && !anchors) return fail(c, BOD_ERR_INVALID, "anchors is NULL but anchor_mode = TENSOR");
    return BOD_OK;
}

// Durations of the lane's moments-kernel launches that no report has covered yet (its launch clock: the last
// kClkSlots launches at most), added to *sum_ms / *runs.  The caller has synchronised with the lane's work.
static int fold_launch_clock(bod_ctx* c, Lane& L, double* sum_ms, long long* runs) {
    if (!L.clk) return BOD_OK;
    unsigned long long h[2 + 2 * kClkSlots];
    CU(c, cudaMemcpy(h, L.clk, sizeof h, cudaMemcpyDeviceToHost));
    const unsigned long long n = h[0];
    unsigned long long first = L.clk_read;
    if (n - first > (unsigned long long)kClkSlots) first = n - kClkSlots;
    for (unsigned long long i = first; i < n; ++i) {
        const unsigned long long t0 = h[2 + 2 * (i % kClkSlots)], t1 = h[3 + 2 * (i % kClkSlots)];
        if (t1 > t0) { *sum_ms += (double)(t1 - t0) * 1e-6; ++*runs; }
    }
    L.clk_read = n;
    return BOD_OK;
}

// Capture (first replay of a lane) or re-point (input tensors changed) the lane's two graphs -- head: ticket
// reset, moments kernel, scan; tail: posterior, soft-NMS, fusion -- and launch them on the lane's stream.  The
// event that tells other streams "the head is done" is recorded between the two launches by an ordinary
// cudaEventRecord, so its meaning for later cudaStreamWaitEvent calls is the usual one.
static int capture_half(bod_ctx* c, Lane& L, int phase, const LevelTable& lv, const float* anchors, const float* counts) {
    cudaStream_t ls = L.tail_stream;
    GraphHooks gh{phase, &L.ga1, &L.ga2};
    CU(c, cudaStreamBeginCapture(ls, cudaStreamCaptureModeRelaxed));
    int rc = run_range(c, L, 0, c->cfg.B, lv, anchors, counts, ls, ls, false, &gh);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(ls, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(c, BOD_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    const int slot = phase == 2 ? 1 : 0;
    L.graph[slot] = g;
    // the kernel nodes whose arguments follow the caller's tensors
    size_t n = 0;
    CU(c, cudaGraphGetNodes(g, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    CU(c, cudaGraphGetNodes(g, nodes.data(), &n));
    for (int which = 1; which <= 2; ++which) {
        if (!(phase & which)) continue;
        const void* want = which == 1 ? k1_kernel_func(L.ga1) : k2_kernel_func(L.ga2);
        cudaGraphNode_t found = nullptr;
        for (cudaGraphNode_t nd : nodes) {
            cudaGraphNodeType t;
            CU(c, cudaGraphNodeGetType(nd, &t));
            if (t != cudaGraphNodeTypeKernel) continue;
            cudaKernelNodeParams p;
            CU(c, cudaGraphKernelNodeGetParams(nd, &p));
            if (p.func == want) found = nd;
        }
        if (!found) return fail(c, BOD_ERR_CUDA, "graph capture: %s kernel node not found", which == 1 ? "moments" : "posterior");
        (which == 1 ? L.gk1 : L.gk2) = found;
    }
    CU(c, cudaGraphInstantiate(&L.gexec[slot], g, 0));
    return BOD_OK;
}

static int replay_run(bod_ctx* c, Lane& L, const LevelTable& lv, const float* anchors, const float* counts) {
    cudaStream_t ls = L.tail_stream;
    if (L.gexec[0] && (L.ga1.counts_in == nullptr) != (counts == nullptr)) {   // sampler <-> injected counts: other outputs
        for (int i = 0; i < 2; ++i) {
            if (L.gexec[i]) cudaGraphExecDestroy(L.gexec[i]);
            if (L.graph[i]) cudaGraphDestroy(L.graph[i]);
            L.gexec[i] = nullptr; L.graph[i] = nullptr;
        }
    }
    // BOD_MERGED_GRAPH=1 (experiment): one graph per run, the posterior kernel right behind the scan
    static const bool merged = getenv("BOD_MERGED_GRAPH") && atoi(getenv("BOD_MERGED_GRAPH")) != 0;
    if (!L.gexec[0]) {
        int rc = merged ? capture_half(c, L, 3, lv, anchors, counts) : capture_half(c, L, 1, lv, anchors, counts);
        if (!rc && !merged) rc = capture_half(c, L, 2, lv, anchors, counts);
        if (rc) return rc;
    } else {
        K1Args a1 = L.ga1;
        a1.lv = lv; a1.counts_in = counts;
        a1.seed = c->cfg.seed; a1.image_id_base = c->cfg.image_id_base;
        if (memcmp(&a1, &L.ga1, sizeof a1) != 0) {
            cudaError_t e = k1_graph_update(L.gexec[0], L.gk1, a1);
            if (e != cudaSuccess) return fail(c, BOD_ERR_CUDA, "graph update (moments kernel): %s", cudaGetErrorString(e));
            L.ga1 = a1;
        }
        K2Args a2 = L.ga2;
        a2.lv = lv; a2.anchors = anchors;
        a2.scale_v = c->cfg.scale_v; a2.scale_u = c->cfg.scale_u;
        if (!cov_width(c->cfg.cov_layout)) for (int l = 0; l < a2.lv.n; ++l) a2.lv.cov[l] = nullptr;
        if (memcmp(&a2, &L.ga2, sizeof a2) != 0) {
            cudaError_t e = k2_graph_update(L.gexec[merged ? 0 : 1], L.gk2, a2);
            if (e != cudaSuccess) return fail(c, BOD_ERR_CUDA, "graph update (posterior kernel): %s", cudaGetErrorString(e));
            L.ga2 = a2;
        }
    }
    // the caller's tensors are ready; the previous lane's moments kernel is done (one moments kernel at a time has the GPU)
    const Lane& P = c->lane[(int)((&L - c->lane) + c->nlanes - 1) % c->nlanes];
    CU(c, cudaStreamWaitEvent(ls, c->ev_in, 0));
    CU(c, cudaStreamWaitEvent(ls, P.head_done, 0));
    CU(c, cudaGraphLaunch(L.gexec[0], ls));
    if (!merged) {
        CU(c, cudaEventRecord(L.head_done, ls));
        CU(c, cudaGraphLaunch(L.gexec[1], ls));
    }
    CU(c, cudaEventRecord(L.tail_done, ls));
    L.tail_pending = true;
    c->launches = 5;            // moments, scan, posterior, soft-NMS, fusion (the moments kernel puts its ticket counter back itself)
    return BOD_OK;
}

// issue one whole run (serial on the caller's stream, or head / tail on the context's streams)
static int issue_run(bod_ctx* c, const LevelTable& lv, const float* anchors, const float* counts, void* cuda_stream) {
    if (c->cfg.N < 2) return fail(c, BOD_ERR_INVALID, "bod_run needs N (mc_dropout_samples) >= 2: the sample covariance divides by N-1");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    int rc;
    bool recorded = c->timing;        // stage events are recorded by the stream launches, not by graph replays
    c->launches = 0;
    c->ev = c->evring[c->runs_recorded % bod_ctx::kEvRing];
    if (c->nlanes == 1) {
        c->cur = 0;
        rc = run_range(c, c->lane[0], 0, c->cfg.B, lv, anchors, counts, st, st, c->timing ? 2 : 0);
        if (rc) return rc;
        c->last_stream = st;
    } else {
        // pipelined: the head runs on the context's own stream once the caller's stream has reached this
        // point; the tail floats on the lane's own stream.  The caller's stream only waits for the head (the
        // last reader of the inputs); results are complete at bod_fetch / bod_wait_results.
        c->cur = (c->cur + 1) % c->nlanes;
        Lane& L = c->lane[c->cur];
        CU(c, cudaEventRecord(c->ev_in, st));
        // Graph replay pays once a run is long (B = 32: +4 %); runs of a few images are faster through the streams,
        // where a result fetch behind every run costs nothing (measured at B = 4, 300 steps: streams 47.9 k images/s
        // with a fetch per run, graphs 44.9 k without and 33.9 k with it).  BOD_GRAPHS=2 forces replay.
        const bool long_run = 4.0 * c->cfg.B * c->cfg.N * c->cfg.A * c->cfg.K >= 0.8e9;
        const bool graphs = c->use_graphs && !c->prefilter && c->k2_on_tail && (long_run || c->force_graphs);
        recorded = false;         // pipelined contexts record no stage events (the moments kernel keeps a launch clock)
        if (graphs && L.uses > 0) {
            // Replay: the whole run (moments, scan, posterior, soft-NMS, fusion) is one graph launch on the lane's
            // stream -- a dozen stream calls per run make batches of a few images host-bound.  Stream order puts it
            // behind the lane's previous run; an event node inside the graph puts its moments kernel behind the
            // previous lane's; only the two kernels that read the caller's tensors are re-pointed when those change.
            rc = replay_run(c, L, lv, anchors, counts);
            if (rc) return rc;
        } else {
            // Heads (moments kernel + scan) go one after the other on the context's own stream.  (Measured and rejected:
            // a head stream per lane, so that the next moments kernel's CTAs move in while the previous one's retire --
            // the block scheduler interleaves the two grids instead, every run finishes later and small batches lose
            // 20-25 %.)  With graphs the first run of a lane goes through the streams: one-time kernel attributes and
            // tables are set up there, and its head has to follow the previous lane's, which may have been a replay.
            cudaStream_t hs = c->own_stream;
            // (not with emit_probs: the probability / sampled-count planes are per context, two moments kernels in
            // flight would both write them)
            if (c->hold_inputs && !graphs && c->head_a && c->head_b && !c->probs && !c->sampled)
                hs = (c->head_flip ^= 1) ? c->head_b : c->head_a;
            CU(c, cudaStreamWaitEvent(hs, c->ev_in, 0));
            if (graphs) {
                const Lane& P = c->lane[(c->cur + c->nlanes - 1) % c->nlanes];
                if (P.uses > 0) CU(c, cudaStreamWaitEvent(hs, P.head_done, 0));
            }
            // (pipelined: no stage events -- seven event records per run are a tenth of the host time of a one-image run,
            // and timing events around the moments kernel put ~10 us between two of them; the moments kernel keeps its
            // own launch clock instead, K1Args::clk)
            rc = run_range(c, L, 0, c->cfg.B, lv, anchors, counts, hs, L.tail_stream, 0);
            if (rc) return rc;
        }
        ++L.uses;
        // (held inputs, bod_set_input_hold: the caller's stream is not made to wait for the head, so consecutive runs
        // issued from one stream do not depend on each other through it)
        if (!c->hold_inputs) CU(c, cudaStreamWaitEvent(st, L.head_done, 0));
        c->last_stream = L.tail_stream;
    }
    c->lane[c->cur].ticket = ++c->next_ticket;
    c->lane[c->cur].fetch_pending = false;
    c->lane[c->cur].h_block_status = nullptr;
    if (recorded) ++c->runs_recorded;
    c->last_timed = recorded;
    c->ran = true; c->used_sampler = (counts == nullptr);
    return BOD_OK;
}

extern "C" int bod_run(bod_ctx* c, const float* cls, const float* box, const float* cov, const float* anchors,
                       const float* counts, void* cuda_stream) {
    int rc = check_inputs(c, cls, box, cov, anchors);
    if (rc) return rc;
    if ((reinterpret_cast<uintptr_t>(box) & 15u) || (cov && (reinterpret_cast<uintptr_t>(cov) & 15u)) ||
        (anchors && (reinterpret_cast<uintptr_t>(anchors) & 15u)))
        return fail(c, BOD_ERR_INVALID, "box / cov / anchors must be 16-byte aligned");
    return issue_run(c, levels_concat(c, cls, box, cov), anchors, counts, cuda_stream);
}

extern "C" int bod_run_levels(bod_ctx* c, const float* const* cls, const float* const* box, const float* const* cov,
                              const float* anchors, const float* counts, void* cuda_stream) {
    if (!c) return BOD_ERR_INVALID;
    if (c->cfg.n_levels < 2) return fail(c, BOD_ERR_STATE, "the context was created without n_levels / level_anchors");
    if (!cls || !box) return fail(c, BOD_ERR_INVALID, "cls and box must not be NULL");
    if (c->cfg.cov_layout != BOD_COV_NONE && !cov) return fail(c, BOD_ERR_INVALID, "cov is NULL but cov_layout != NONE");
    if (c->cfg.anchor_mode == BOD_ANCHORS_TENSOR && !anchors) return fail(c, BOD_ERR_INVALID, "anchors is NULL but anchor_mode = TENSOR");
    if (anchors && (reinterpret_cast<uintptr_t>(anchors) & 15u)) return fail(c, BOD_ERR_INVALID, "anchors must be 16-byte aligned");
    for (int l = 0; l < c->cfg.n_levels; ++l) {
        if (!cls[l] || !box[l] || (c->cfg.cov_layout != BOD_COV_NONE && !cov[l]))
            return fail(c, BOD_ERR_INVALID, "level %d: NULL tensor", l);
        if ((reinterpret_cast<uintptr_t>(box[l]) & 15u) || (c->cfg.cov_layout != BOD_COV_NONE && (reinterpret_cast<uintptr_t>(cov[l]) & 15u)))
            return fail(c, BOD_ERR_INVALID, "level %d: box / cov must be 16-byte aligned", l);
    }
    return issue_run(c, levels_split(c, cls, box, c->cfg.cov_layout != BOD_COV_NONE ? cov : nullptr), anchors, counts, cuda_stream);
}

extern "C" int bod_set_sampler_stream(bod_ctx* c, uint64_t seed, uint32_t image_id_base) {
    if (!c) return BOD_ERR_INVALID;
    c->cfg.seed = seed; c->cfg.image_id_base = image_id_base;
    return BOD_OK;
}
extern "C" int bod_set_image_scale(bod_ctx* c, float scale_v, float scale_u) {
    if (!c) return BOD_ERR_INVALID;
    if (!(scale_v > 0.0f) || !(scale_u > 0.0f)) return fail(c, BOD_ERR_INVALID, "scale factors must be positive");
    c->cfg.scale_v = scale_v; c->cfg.scale_u = scale_u;
    return BOD_OK;
}

extern "C" int bod_set_input_hold(bod_ctx* c, int enabled) {
    if (!c) return BOD_ERR_INVALID;
    c->hold_inputs = enabled != 0;
    return BOD_OK;
}

extern "C" int bod_wait_results(bod_ctx* c, void* cuda_stream) {
    if (!c) return BOD_ERR_INVALID;
    if (!c->ran) return fail(c, BOD_ERR_STATE, "no bod_run has been issued on this context");
    CU(c, cudaSetDevice(c->device));
    // every run issued so far (each lane's last one), together with the copies bod_fetch_async put behind them
    if (c->nlanes > 1)
        for (int l = 0; l < c->nlanes; ++l)
            if (c->lane[l].tail_pending)
                CU(c, cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(cuda_stream), c->lane[l].tail_done, 0));
    return BOD_OK;
}

// validation_utils.post_process_predictions (validation_utils.py:10-77) for the batch
extern "C" int bod_validate_run(bod_ctx* c, const float* cls, const float* box, const float* anchors,
                                const bod_val_scaling* scaling, void* cuda_stream) {
    if (!c) return BOD_ERR_INVALID;
    if (!cls || !box || !anchors) return fail(c, BOD_ERR_INVALID, "cls, box and anchors must not be NULL");
    if ((reinterpret_cast<uintptr_t>(box) & 15u) || (reinterpret_cast<uintptr_t>(anchors) & 15u))
        return fail(c, BOD_ERR_INVALID, "box / anchors must be 16-byte aligned");
    if (scaling && (scaling->mode < 0 || scaling->mode > 2)) return fail(c, BOD_ERR_INVALID, "bad scaling mode");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    { int rc0 = drain_tails(c); if (rc0) return rc0; }                  // drain pipelined runs first
    c->cur = 0;
    c->lane[0].ticket = ++c->next_ticket; c->lane[0].fetch_pending = false;
    Lane& L = c->lane[0];
    const bod_config& g = c->cfg;
    c->launches = 0;
    ValArgs v{};
    v.cls = cls; v.box = box; v.anchors = anchors;
    v.slot_anchor = L.slot_anchor; v.slot_counts = L.slot_counts; v.tile_count = L.tile_count;
    v.tile_off = L.tile_off; v.num_survivors = L.num_survivors;
    v.surv_anchor = L.surv_anchor; v.cnt_post = L.cnt_post; v.mu_post = L.mu_post; v.score = L.score; v.corners = L.corners;
    v.nms_idx = L.nms_idx; v.num_dets = L.num_dets;
    v.out_means = L.out_means; v.out_covs = L.out_covs; v.out_param = L.out_param; v.out_count = L.out_count;
    v.B = g.B; v.A = g.A; v.K = g.K; v.tiles = c->tiles; v.capacity = c->capacity; v.Dmax = c->Dmax;
    v.scale_mode = scaling ? scaling->mode : BOD_VAL_SCALE_NONE;
    for (int i = 0; i < 4; ++i) v.shift[i] = scaling ? scaling->shift[i] : 0.0f;
    v.norm_h = scaling ? scaling->norm_h : 1.0f; v.norm_w = scaling ? scaling->norm_w : 1.0f;
    v.scale_h = scaling ? scaling->scale_h : 1.0f; v.scale_w = scaling ? scaling->scale_w : 1.0f;
    CU(c, launch_val_filter(v, st));
    ScanArgs sc{};
    sc.tile_count = L.tile_count; sc.tile_off = L.tile_off; sc.num_survivors = L.num_survivors; sc.status = c->status;
    sc.B = g.B; sc.tiles = c->tiles; sc.capacity = c->capacity;
    CU(c, launch_scan(sc, st));
    CU(c, launch_val_survivors(v, st));
    K3Args k3{};
    k3.corners = L.corners; k3.score = L.score; k3.num_survivors = L.num_survivors; k3.surv_anchor = L.surv_anchor;
    k3.stale = L.stale; k3.cur = L.cur; k3.begin = L.begin; k3.pend = L.pend;
    k3.pw = L.pw; k3.pw_rows = c->pw_rows; k3.max_rows = c->k3_rows;
    k3.pstride = c->pstride;
    k3.nms_idx = L.nms_idx; k3.nms_score = L.nms_score; k3.centre_anchor = L.centre_anchor;
    k3.num_dets = L.num_dets;
    k3.B = g.B; k3.capacity = c->capacity; k3.Dmax = c->Dmax;
    k3.iou_threshold = g.iou_threshold; k3.soft_nms_sigma = g.soft_nms_sigma;
    k3.threads = c->k3_threads; k3.force_big = c->k3_force_big; k3.psm_max = c->k3_psm_max; k3.seg_cap = c->k3_seg_cap;
    CU(c, launch_k3(k3, st));
    CU(c, launch_val_gather(v, st));
    c->launches = 5;
    c->last_timed = false;
    c->last_stream = st; c->ran = true; c->used_sampler = false;
    return BOD_OK;
}

static int sync_and_status(bod_ctx* c) {
    if (!c->ran) return fail(c, BOD_ERR_STATE, "no bod_run has been issued on this context");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->last_stream));
    int32_t status = 0;
    CU(c, cudaMemcpy(&status, c->status, 4, cudaMemcpyDeviceToHost));
    if (status & 1) {
        CU(c, cudaMemset(c->status, 0, 4));        // sticky until reported once
        return fail(c, BOD_ERR_OVERFLOW, "an image produced more survivors than max_survivors=%d", c->capacity);
    }
    return BOD_OK;
}

static int copy_results(bod_ctx* c, const Lane& L, bod_host_results* out, cudaStream_t st) {
    const size_t B = c->cfg.B, D = c->Dmax, K = c->cfg.K;
#define D2H(dst, src, bytes) if (out->dst) CU(c, cudaMemcpyAsync(out->dst, L.src, (bytes), cudaMemcpyDeviceToHost, st))
    D2H(num_dets, num_dets, B * 4);
    D2H(num_survivors, num_survivors, B * 4);
    D2H(means, out_means, B * D * 16);
    D2H(covs, out_covs, B * D * 64);
    D2H(cat_param, out_param, B * D * K * 4);
    D2H(cat_count, out_count, B * D * K * 4);
    D2H(nms_indices, nms_idx, B * D * 4);
    D2H(centre_anchor_idx, centre_anchor, B * D * 4);
    D2H(centre_scores, nms_score, B * D * 4);
#undef D2H
    return BOD_OK;
}

extern "C" int bod_fetch(bod_ctx* c, bod_host_results* out) {
    if (!c || !out) return BOD_ERR_INVALID;
    int rc = sync_and_status(c);
    if (rc) return rc;
    rc = copy_results(c, c->lane[c->cur], out, c->last_stream);
    if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->last_stream));
    return BOD_OK;
}

extern "C" int bod_device_results_of(bod_ctx* c, bod_device_results* out) {
    if (!c || !out) return BOD_ERR_INVALID;
    const Lane& L = c->lane[c->cur];          // the lane of the last issued run
    out->num_dets = L.num_dets; out->num_survivors = L.num_survivors;
    out->means = L.out_means; out->covs = L.out_covs; out->cat_param = L.out_param; out->cat_count = L.out_count;
    out->nms_indices = L.nms_idx; out->centre_anchor_idx = L.centre_anchor; out->centre_scores = L.nms_score;
    return BOD_OK;
}

// ---- streaming retrieval: every run's results, by ticket ----
static Lane* lane_of_ticket(bod_ctx* c, long long ticket) {
    if (ticket <= 0) return nullptr;
    for (int l = 0; l < c->nlanes; ++l) if (c->lane[l].ticket == ticket) return &c->lane[l];
    return nullptr;
}
static cudaStream_t stream_of_lane(bod_ctx* c, const Lane& L) { return c->nlanes > 1 ? L.tail_stream : c->last_stream; }

extern "C" int64_t bod_last_ticket(const bod_ctx* c) { return c ? (int64_t)c->next_ticket : 0; }

extern "C" int bod_fetch_async(bod_ctx* c, int64_t ticket, bod_host_results* out) {
    if (!c || !out) return BOD_ERR_INVALID;
    Lane* L = lane_of_ticket(c, ticket);
    if (!L) return fail(c, BOD_ERR_STATE, "the results of run %lld are gone: its lane has been reused (pipeline_depth = %d)",
                        (long long)ticket, c->nlanes);
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = stream_of_lane(c, *L);
    int rc = copy_results(c, *L, out, st);
    if (rc) return rc;
    if (L->h_status) CU(c, cudaMemcpyAsync(L->h_status, c->status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(c, cudaEventRecord(L->fetch_done, st));
    L->fetch_pending = true;
    L->h_block_status = nullptr;
    // the lane's next run must not overwrite what these copies still have to read
    if (c->nlanes > 1 && L->tail_pending) CU(c, cudaEventRecord(L->tail_done, st));
    return BOD_OK;
}

extern "C" int bod_result_block_layout(const bod_ctx* c, int64_t offsets[10], int64_t* bytes) {
    if (!c || !offsets || !bytes) return BOD_ERR_INVALID;
    for (int i = 0; i < 10; ++i) offsets[i] = c->block_off[i];
    *bytes = c->block_bytes;
    return BOD_OK;
}

extern "C" int bod_fetch_block_async(bod_ctx* c, int64_t ticket, void* host_block) {
    if (!c || !host_block) return BOD_ERR_INVALID;
    Lane* L = lane_of_ticket(c, ticket);
    if (!L) return fail(c, BOD_ERR_STATE, "the results of run %lld are gone: its lane has been reused (pipeline_depth = %d)",
                        (long long)ticket, c->nlanes);
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = stream_of_lane(c, *L);
    CU(c, cudaMemcpyAsync(host_block, L->block, (size_t)c->block_bytes, cudaMemcpyDeviceToHost, st));
    CU(c, cudaEventRecord(L->fetch_done, st));
    L->fetch_pending = true;
    L->h_block_status = reinterpret_cast<const int32_t*>(static_cast<const unsigned char*>(host_block) + c->block_off[9]);
    if (c->nlanes > 1 && L->tail_pending) CU(c, cudaEventRecord(L->tail_done, st));
    return BOD_OK;
}

extern "C" int bod_ticket_wait(bod_ctx* c, int64_t ticket) {
    if (!c) return BOD_ERR_INVALID;
    Lane* L = lane_of_ticket(c, ticket);
    if (!L) return fail(c, BOD_ERR_STATE, "run %lld is not in flight any more (pipeline_depth = %d)", (long long)ticket, c->nlanes);
    CU(c, cudaSetDevice(c->device));
    if (L->fetch_pending) {
        CU(c, cudaEventSynchronize(L->fetch_done));
        const int32_t seen = L->h_block_status ? *L->h_block_status : (L->h_status ? *L->h_status : 0);
        if (seen & 1) {
            CU(c, cudaMemset(c->status, 0, 4));        // sticky until reported once
            if (L->h_status) *L->h_status = 0;
            return fail(c, BOD_ERR_OVERFLOW, "an image produced more survivors than max_survivors=%d", c->capacity);
        }
        return BOD_OK;
    }
    CU(c, cudaStreamSynchronize(stream_of_lane(c, *L)));
    int32_t status = 0;
    CU(c, cudaMemcpy(&status, c->status, 4, cudaMemcpyDeviceToHost));
    if (status & 1) {
        CU(c, cudaMemset(c->status, 0, 4));
        return fail(c, BOD_ERR_OVERFLOW, "an image produced more survivors than max_survivors=%d", c->capacity);
    }
    return BOD_OK;
}

extern "C" int bod_device_results_at(bod_ctx* c, int64_t ticket, bod_device_results* out) {
    if (!c || !out) return BOD_ERR_INVALID;
    Lane* Lp = lane_of_ticket(c, ticket);
    if (!Lp) return fail(c, BOD_ERR_STATE, "the results of run %lld are gone: its lane has been reused", (long long)ticket);
    const Lane& L = *Lp;
    out->num_dets = L.num_dets; out->num_survivors = L.num_survivors;
    out->means = L.out_means; out->covs = L.out_covs; out->cat_param = L.out_param; out->cat_count = L.out_count;
    out->nms_indices = L.nms_idx; out->centre_anchor_idx = L.centre_anchor; out->centre_scores = L.nms_score;
    return BOD_OK;
}

extern "C" void* bod_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void bod_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int bod_fetch_survivors(bod_ctx* c, int32_t b, bod_host_survivors* out) {
    if (!c || !out || b < 0 || b >= c->cfg.B) return BOD_ERR_INVALID;
    int rc = sync_and_status(c);
    if (rc) return rc;
    int32_t S = 0;
    const Lane& L = c->lane[c->cur];
    CU(c, cudaMemcpy(&S, L.num_survivors + b, 4, cudaMemcpyDeviceToHost));
    out->count = S;
    if (S > out->capacity) return fail(c, BOD_ERR_INVALID, "bod_fetch_survivors: capacity %d < S %d", out->capacity, S);
    const size_t cap = c->capacity, K = c->cfg.K, s = (size_t)S;
    if (S == 0) return BOD_OK;
#define D2H(dst, src, off, bytes) if (out->dst) CU(c, cudaMemcpy(out->dst, L.src + (off), (bytes), cudaMemcpyDeviceToHost))
    D2H(anchor_idx, surv_anchor, b * cap, s * 4);
    D2H(counts, cnt_post, b * cap * K, s * K * 4);
    D2H(means, mu_post, b * cap * 4, s * 16);
    D2H(covs, sig_post, b * cap * 16, s * 64);
    D2H(scores, score, b * cap, s * 4);
    D2H(corners, corners, b * cap, s * 16);
#undef D2H
    return BOD_OK;
}

extern "C" int bod_fetch_members(bod_ctx* c, int32_t b, uint32_t* mask, int32_t words_per_row) {
    if (!c || !mask || b < 0 || b >= c->cfg.B) return BOD_ERR_INVALID;
    int rc = sync_and_status(c);
    if (rc) return rc;
    int32_t S = 0, D = 0;
    const Lane& L = c->lane[c->cur];
    CU(c, cudaMemcpy(&S, L.num_survivors + b, 4, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(&D, L.num_dets + b, 4, cudaMemcpyDeviceToHost));
    const int nw = (S + 31) / 32;
    if (words_per_row < nw) return fail(c, BOD_ERR_INVALID, "bod_fetch_members: words_per_row %d < %d", words_per_row, nw);
    if (D == 0 || nw == 0) return BOD_OK;
    CU(c, cudaMemcpy2D(mask, (size_t)words_per_row * 4, L.member + (size_t)b * c->Dmax * c->words, (size_t)c->words * 4,
                       (size_t)nw * 4, D, cudaMemcpyDeviceToHost));
    return BOD_OK;
}

extern "C" int bod_fetch_probs(bod_ctx* c, int32_t b, float* probs) {
    if (!c || !probs || b < 0 || b >= c->cfg.B) return BOD_ERR_INVALID;
    if (!c->probs) return fail(c, BOD_ERR_STATE, "context was created without emit_probs");
    int rc = sync_and_status(c);
    if (rc) return rc;
    const size_t n = (size_t)c->cfg.A * c->cfg.K;
    CU(c, cudaMemcpy(probs, c->probs + b * n, n * 4, cudaMemcpyDeviceToHost));
    return BOD_OK;
}

extern "C" int bod_fetch_sampled_counts(bod_ctx* c, int32_t b, float* counts) {
    if (!c || !counts || b < 0 || b >= c->cfg.B) return BOD_ERR_INVALID;
    if (!c->sampled) return fail(c, BOD_ERR_STATE, "context was created without emit_probs");
    if (!c->used_sampler) return fail(c, BOD_ERR_STATE, "last run injected counts; nothing was sampled");
    int rc = sync_and_status(c);
    if (rc) return rc;
    const size_t n = (size_t)c->cfg.A * c->cfg.K;
    CU(c, cudaMemcpy(counts, c->sampled + b * n, n * 4, cudaMemcpyDeviceToHost));
    return BOD_OK;
}

#ifdef BOD_DIAGNOSTICS
// diagnostic builds only (not part of the public header): the timeline stamps of every lane's last run,
// out[lane][kernel: moments, scan, posterior, soft-NMS, fusion][first CTA start, last CTA start, end, -] in ns
extern "C" int bod_debug_timeline(bod_ctx* c, unsigned long long* out) {
    if (!c || !out || !c->lane[0].tl) return BOD_ERR_STATE;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    for (int l = 0; l < c->nlanes; ++l)
        CU(c, cudaMemcpy(out + 20 * l, c->lane[l].tl, 20 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return BOD_OK;
}
// diagnostic builds only (not part of the public header): per-image, per-warp soft-NMS phase cycle counters
extern "C" int bod_debug_k3_counters(bod_ctx* c, long long* out) {
    if (!c || !out || !c->k3_dbg) return BOD_ERR_STATE;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    CU(c, cudaMemcpy(out, c->k3_dbg, ((size_t)c->cfg.B * 384 + 8) * sizeof(long long), cudaMemcpyDeviceToHost));
    return BOD_OK;
}
#endif

extern "C" int bod_synchronize(bod_ctx* c) {
    if (!c) return BOD_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    return BOD_OK;
}

extern "C" int bod_last_stage_ms(bod_ctx* c, float ms[6]) {
    if (!c || !ms) return BOD_ERR_INVALID;
    int rc = sync_and_status(c);
    if (rc) return rc;
    if (c->nlanes > 1) {                                     // pipelined contexts time the moments kernel only: its launch clock
        for (int i = 0; i < 6; ++i) ms[i] = 0.0f;
        Lane& L = c->lane[c->cur];
        if (!L.clk) return fail(c, BOD_ERR_STATE, "the last run kept no launch clock");
        unsigned long long h[2 + 2 * kClkSlots];
        CU(c, cudaMemcpy(h, L.clk, sizeof h, cudaMemcpyDeviceToHost));
        if (h[0] == 0) return fail(c, BOD_ERR_STATE, "the last run kept no launch clock");
        const unsigned long long i = (h[0] - 1) % kClkSlots;
        if (h[3 + 2 * i] > h[2 + 2 * i]) ms[0] = (float)((double)(h[3 + 2 * i] - h[2 + 2 * i]) * 1e-6);
        return BOD_OK;
    }
    if (!c->last_timed || c->runs_recorded == 0) return fail(c, BOD_ERR_STATE, "the last run recorded no stage events");
    cudaEvent_t* ev = c->evring[(c->runs_recorded - 1) % bod_ctx::kEvRing];
    for (int i = 0; i < 5; ++i) CU(c, cudaEventElapsedTime(&ms[i], ev[i == 3 ? 6 : i], ev[i + 1]));
    CU(c, cudaEventElapsedTime(&ms[5], ev[0], ev[5]));
    return BOD_OK;
}

extern "C" int bod_set_stage_timing(bod_ctx* c, int enabled) {
    if (!c) return BOD_ERR_INVALID;
    c->timing = enabled != 0;
    return BOD_OK;
}

extern "C" int bod_stage_ms_accum(bod_ctx* c, float sum_ms[6], int32_t* runs) {
    if (!c || !sum_ms || !runs) return BOD_ERR_INVALID;
    for (int i = 0; i < 6; ++i) sum_ms[i] = 0.0f;
    *runs = 0;
    if (c->runs_recorded == c->runs_reported && !c->ran) return BOD_OK;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->last_stream));
    { int rc0 = drain_tails(c); if (rc0) return rc0; }      // earlier runs' tails live on other streams
    if (c->nlanes > 1) {
        // pipelined contexts time the moments kernel only (its launch clock, see K1Args::clk: the last 64 launches of
        // every lane at most); the other stages read as zero there
        double sum = 0.0;
        long long n = 0;
        for (int l = 0; l < c->nlanes; ++l) { int rc0 = fold_launch_clock(c, c->lane[l], &sum, &n); if (rc0) return rc0; }
        sum_ms[0] = (float)sum; *runs = (int32_t)n;
        c->runs_reported = c->runs_recorded;
        return BOD_OK;
    }
    long long first = c->runs_reported;
    if (c->runs_recorded - first > bod_ctx::kEvRing) first = c->runs_recorded - bod_ctx::kEvRing;
    for (long long r = first; r < c->runs_recorded; ++r) {
        cudaEvent_t* ev = c->evring[r % bod_ctx::kEvRing];
        float ms = 0.0f;
        for (int i = 0; i < 5; ++i) { CU(c, cudaEventElapsedTime(&ms, ev[i == 3 ? 6 : i], ev[i + 1])); sum_ms[i] += ms; }
        CU(c, cudaEventElapsedTime(&ms, ev[0], ev[5])); sum_ms[5] += ms;
        ++*runs;
    }
    c->runs_reported = c->runs_recorded;
    return BOD_OK;
}

// Device-clock durations of the moments-kernel launches since the previous call (any context; the last 64 launches of
// every lane at most): what bod_stage_ms_accum reports as stage 0 of a pipelined context, available beside the CUDA
// events of a serial one (tests compare the two).
extern "C" int bod_moments_clock_accum(bod_ctx* c, double* sum_ms, int32_t* runs) {
    if (!c || !sum_ms || !runs) return BOD_ERR_INVALID;
    *sum_ms = 0.0; *runs = 0;
    if (!c->ran) return BOD_OK;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->last_stream));
    { int rc0 = drain_tails(c); if (rc0) return rc0; }
    long long n = 0;
    for (int l = 0; l < c->nlanes; ++l) { int rc0 = fold_launch_clock(c, c->lane[l], sum_ms, &n); if (rc0) return rc0; }
    *runs = (int32_t)n;
    return BOD_OK;
}

// ---------------------------------------------------------------------------
// host-buffer entry.  Every anchor's class logits must be inspected, so `cls`
// is staged host->device in image chunks on a copy stream while the previous
// chunk computes.  The box deltas and covariance rows are only needed for the
// S survivors (~2% of the anchors): when `box` / `cov` live in pinned (page-locked,
// device-mapped) host memory, K2 gathers those rows in place over PCIe instead
// of copying the whole [B,N,A,4] and [B,N,A,16] tensors (64% of the input bytes).
// Pageable buffers fall back to plain copies.
// ---------------------------------------------------------------------------
static const float* mapped_device_pointer(const float* host) {
    if (!host) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeHost && at.devicePointer) return static_cast<const float*>(at.devicePointer);
    if (at.type == cudaMemoryTypeManaged) return host;
    return nullptr;
}

extern "C" int bod_run_host(bod_ctx* c, const float* cls, const float* box, const float* cov, const float* anchors,
                            const float* counts, bod_host_results* out) {
    int rc = check_inputs(c, cls, box, cov, anchors);
    if (rc) return rc;
    if (!out) return fail(c, BOD_ERR_INVALID, "out is NULL");
    CU(c, cudaSetDevice(c->device));
    const bod_config& g = c->cfg;
    const size_t B = g.B, N = g.N, A = g.A, K = g.K;
    const size_t cw = cov_width(g.cov_layout);
    // box / cov: read in place when the caller's buffers are device-mapped and 16-byte aligned
    const float* box_m = c->host_copy_all ? nullptr : mapped_device_pointer(box);
    const float* cov_m = (cw && !c->host_copy_all) ? mapped_device_pointer(cov) : nullptr;
    if (box_m && (reinterpret_cast<uintptr_t>(box_m) & 15u)) box_m = nullptr;
    if (cov_m && (reinterpret_cast<uintptr_t>(cov_m) & 15u)) cov_m = nullptr;
    if (!c->in_cls) {
        CU(c, cudaMalloc(&c->in_cls, B * N * A * K * 4));
        CU(c, cudaMalloc(&c->in_anchors, A * 16));
        CU(c, cudaMalloc(&c->in_counts, B * A * K * 4));
    }
    if (!box_m && !c->in_box) CU(c, cudaMalloc(&c->in_box, B * N * A * 16));
    if (cw && !cov_m && !c->in_cov) CU(c, cudaMalloc(&c->in_cov, B * N * A * cw * 4));
    cudaStream_t cs = c->copy_stream, st = c->own_stream;
    { int rc0 = drain_tails(c); if (rc0) return rc0; }                  // drain pipelined runs; this entry is synchronous
    c->cur = 0;
    c->lane[0].ticket = ++c->next_ticket; c->lane[0].fetch_pending = false;
    Lane& L = c->lane[0];
    c->launches = 0;
    c->last_timed = false;
    c->h2d_copied = 0; c->h2d_mapped_rows = 0; c->d2h_copied = 0;
    if (anchors) { CU(c, cudaMemcpyAsync(c->in_anchors, anchors, A * 16, cudaMemcpyHostToDevice, cs)); c->h2d_copied += A * 16; }
    // chunks of images: small enough to overlap copy and compute, large enough to fill the GPU
    const int chunk = (B >= 8) ? (int)((B + 3) / 4) : (int)B;
    int nev = 0;
    for (size_t b0 = 0; b0 < B; b0 += chunk) {
        const size_t nb = (b0 + chunk <= B) ? chunk : B - b0;
        CU(c, cudaMemcpyAsync(c->in_cls + b0 * N * A * K, cls + b0 * N * A * K, nb * N * A * K * 4, cudaMemcpyHostToDevice, cs));
        c->h2d_copied += nb * N * A * K * 4;
        if (counts) {
            CU(c, cudaMemcpyAsync(c->in_counts + b0 * A * K, counts + b0 * A * K, nb * A * K * 4, cudaMemcpyHostToDevice, cs));
            c->h2d_copied += nb * A * K * 4;
        }
        if (!box_m) {
            CU(c, cudaMemcpyAsync(c->in_box + b0 * N * A * 4, box + b0 * N * A * 4, nb * N * A * 16, cudaMemcpyHostToDevice, cs));
            c->h2d_copied += nb * N * A * 16;
        }
        if (cw && !cov_m) {
            CU(c, cudaMemcpyAsync(c->in_cov + b0 * N * A * cw, cov + b0 * N * A * cw, nb * N * A * cw * 4, cudaMemcpyHostToDevice, cs));
            c->h2d_copied += nb * N * A * cw * 4;
        }
        cudaEvent_t ev = c->ev_copy[nev++ & 3];
        CU(c, cudaEventRecord(ev, cs));
        CU(c, cudaStreamWaitEvent(st, ev, 0));
        rc = run_range(c, L, (int)b0, (int)nb,
                       levels_concat(c, c->in_cls + b0 * N * A * K,
                                     box_m ? box_m + b0 * N * A * 4 : c->in_box + b0 * N * A * 4,
                                     cw ? (cov_m ? cov_m + b0 * N * A * cw : c->in_cov + b0 * N * A * cw) : nullptr),
                       anchors ? c->in_anchors : nullptr, counts ? c->in_counts + b0 * A * K : nullptr, st, st, false);
        if (rc) return rc;
    }
    c->last_stream = st; c->ran = true; c->used_sampler = (counts == nullptr);
    rc = copy_results(c, L, out, st);
    if (rc) return rc;
    CU(c, cudaStreamSynchronize(st));
    int32_t status = 0;
    CU(c, cudaMemcpy(&status, c->status, 4, cudaMemcpyDeviceToHost));
    if (status & 1) {
        CU(c, cudaMemset(c->status, 0, 4));
        return fail(c, BOD_ERR_OVERFLOW, "an image produced more survivors than max_survivors=%d", c->capacity);
    }
    // rows read in place from mapped host memory: N * (16 [+ 4*cw]) bytes per survivor
    if (box_m || cov_m) {
        std::vector<int32_t> ns(B);
        CU(c, cudaMemcpy(ns.data(), L.num_survivors, B * 4, cudaMemcpyDeviceToHost));
        int64_t S = 0;
        for (size_t b = 0; b < B; ++b) S += ns[b];
        c->h2d_mapped_rows = S * (int64_t)N * ((box_m ? 16 : 0) + (cov_m ? (int64_t)cw * 4 : 0));
    }
    {
        const size_t D = c->Dmax;
        c->d2h_copied = (int64_t)(B * 4 * 2 + B * D * (16 + 64 + 2 * K * 4 + 12));
    }
    return BOD_OK;
}

extern "C" int bod_last_host_traffic(const bod_ctx* c, int64_t* h2d_copied, int64_t* h2d_gathered, int64_t* d2h) {
    if (!c) return BOD_ERR_INVALID;
    if (h2d_copied) *h2d_copied = c->h2d_copied;
    if (h2d_gathered) *h2d_gathered = c->h2d_mapped_rows;
    if (d2h) *d2h = c->d2h_copied;
    return BOD_OK;
}

// ---------------------------------------------------------------------------
// bayes_od_clustering on its own (inference_utils.py:285-364), one image
// ---------------------------------------------------------------------------
extern "C" int bod_cluster_host(bod_ctx* c, int32_t S, const float* counts, const float* means, const float* covs,
                                int32_t D, const int32_t* centres, const float* affinity, float affinity_threshold,
                                bod_host_results* out) {
    if (!c || !out) return BOD_ERR_INVALID;
    if (S < 0 || D < 0 || (S > 0 && (!counts || !means || !covs)) || (D > 0 && (!centres || !affinity)))
        return fail(c, BOD_ERR_INVALID, "bod_cluster_host: NULL input");
    if (S > c->capacity) return fail(c, BOD_ERR_INVALID, "bod_cluster_host: S=%d exceeds the context capacity %d", S, c->capacity);
    if (D > c->Dmax) return fail(c, BOD_ERR_INVALID, "bod_cluster_host: D=%d exceeds max_output_size %d", D, c->Dmax);
    for (int d = 0; d < D; ++d)
        if (centres[d] < 0 || centres[d] >= S) return fail(c, BOD_ERR_INVALID, "bod_cluster_host: centre index out of range");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = c->own_stream;
    { int rc0 = drain_tails(c); if (rc0) return rc0; }
    c->cur = 0;
    c->lane[0].ticket = ++c->next_ticket; c->lane[0].fetch_pending = false;
    Lane& L = c->lane[0];
    const size_t K = c->cfg.K, Dm = c->Dmax;
    c->launches = 0;
    int rc = BOD_OK;
    // membership bits from the caller's affinity matrix: affinity[s, centre] > thr (:316)
    std::vector<uint32_t> mask((size_t)(D > 0 ? D : 1) * c->words, 0u);
    for (int d = 0; d < D; ++d)
        for (int s = 0; s < S; ++s)
            if (affinity[(size_t)s * S + centres[d]] > affinity_threshold) mask[(size_t)d * c->words + (s >> 5)] |= 1u << (s & 31);
    if (S > 0) {
        CU(c, cudaMemcpyAsync(L.cnt_post, counts, (size_t)S * K * 4, cudaMemcpyHostToDevice, st));
        CU(c, cudaMemcpyAsync(L.mu_post, means, (size_t)S * 16, cudaMemcpyHostToDevice, st));
        CU(c, cudaMemcpyAsync(L.sig_post, covs, (size_t)S * 64, cudaMemcpyHostToDevice, st));
    }
    if (D > 0) {
        CU(c, cudaMemcpyAsync(L.nms_idx, centres, (size_t)D * 4, cudaMemcpyHostToDevice, st));
        CU(c, cudaMemcpyAsync(L.member, mask.data(), (size_t)D * c->words * 4, cudaMemcpyHostToDevice, st));
    }
    CU(c, cudaMemcpyAsync(L.num_survivors, &S, 4, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(L.num_dets, &D, 4, cudaMemcpyHostToDevice, st));
    CU(c, cudaStreamSynchronize(st));          // `mask`, S, D are stack/heap temporaries
    K4Args k4{};
    k4.cnt_post = L.cnt_post; k4.mu_post = L.mu_post; k4.sig_post = L.sig_post; k4.num_survivors = L.num_survivors;
    k4.nms_idx = L.nms_idx; k4.num_dets = L.num_dets; k4.corners = nullptr; k4.member = L.member;
    k4.out_means = L.out_means; k4.out_covs = L.out_covs; k4.out_param = L.out_param; k4.out_count = L.out_count;
    k4.B = 1; k4.K = (int)K; k4.capacity = c->capacity; k4.Dmax = (int)Dm; k4.words = c->words;
    k4.calibration = c->cfg.cov_calibration; k4.iou_threshold = c->cfg.iou_threshold;
    CU(c, launch_k4(k4, st));
    c->launches = 1;
    c->last_stream = st; c->ran = true;
    // B=1 blocks
    bod_host_results o = *out;
    const size_t keepB = c->cfg.B;
    c->cfg.B = 1;
    rc = copy_results(c, L, &o, st);
    c->cfg.B = (int32_t)keepB;
    if (rc) return rc;
    CU(c, cudaStreamSynchronize(st));
    return BOD_OK;
}

extern "C" int bod_generate_anchors(int32_t im_h, int32_t im_w, float* anchors_dev, void* cuda_stream) {
    if (im_h < 1 || im_w < 1) return BOD_ERR_INVALID;
    const int A = count_anchors(im_h, im_w);
    if (!anchors_dev) return A;
    if (launch_generate_anchors(im_h, im_w, anchors_dev, reinterpret_cast<cudaStream_t>(cuda_stream)) != cudaSuccess)
        return BOD_ERR_CUDA;
    return A;
}
