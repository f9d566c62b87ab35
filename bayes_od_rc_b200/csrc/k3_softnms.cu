// k3_softnms.cu — stage K3: exact emulation of TF's soft-NMS centre selection
// plus the cluster-membership bitmasks.  Compiled with -fmad=false.
//
// Reference lines replaced:
//   inference_utils.py:207-212  tf.image.non_max_suppression_with_scores(boxes, scores,
//                               max_output_size, iou_threshold, soft_nms_sigma)
//                               = TF's NonMaxSuppressionV5 CPU kernel (a device->host->device
//                               round trip in the reference graph)
//   inference_utils.py:214-215  box_utils.bbox_iou_vuvu(corners, corners)  [S,S]
//   inference_utils.py:316      affinity_matrix[:, centre] > threshold
// Only the D centre columns of the S x S matrix are ever read by the reference
// (:316), so only those are evaluated here and only as bits.
//
// How the sequential priority-queue loop of the TF kernel is reproduced exactly.
// TF pops the best candidate, multiplies its score by exp(scale*iou^2) for every
// box selected since the candidate's last pop (newest first), and either selects
// it (score unchanged) or pushes it back.  Let t_i be candidate i's score as of
// its last queue update ("stale"), and u_i the score it would have if it were
// popped now (t_i times the weights of the boxes selected since, newest first).
// u_i <= t_i (scores are >= 0), so the candidate TF selects next is
// x = argmax_i (u_i, -i); on the way TF pops, updates and re-pushes exactly the
// candidates whose stale key (t_i, -i) exceeds (u_x, -x).  One ROUND per selected
// box therefore needs one block-wide arg-max plus one pass in which every
// candidate tests the new box; the multiplication order inside each candidate
// (newest selected first within an update epoch) is kept by recomputing u_i from
// t_i over the candidate's pending-selection bitmask.  Weights equal to exactly
// 1.0f (IoU 0, the overwhelmingly common case) leave a score bit-identical, so
// boxes that do not overlap the new centre need no arithmetic at all.
//
// Fast kernel (S <= kFastS): corners and current scores live in shared memory;
// a round is (A) one pass of cheap overlap tests that also folds the arg-max of
// the untouched candidates, (B) a dense pass over the compacted list of
// overlapping candidates (IoUs, exp, membership bits).  Two block barriers per
// round.  One CTA per image (images are independent); B CTAs run concurrently.
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

constexpr int kK3Threads = 512;
constexpr int kFastS = 8192;          // candidates the shared-memory kernel holds
constexpr int kListMax = 3072;        // compacted overlap list (uint16 indices)

BOD_DEVINL unsigned long long make_key(float score, int idx) {
    return ((unsigned long long)float_key(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}
BOD_DEVINL float key_score(unsigned long long k) {
    const uint32_t u = (uint32_t)(k >> 32);
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
BOD_DEVINL int key_index(unsigned long long k) { return (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull)); }

// soft-NMS weight of TF: exp(scale * sim * sim), scale = -0.5 / sigma; hard mode: 1 or 0
BOD_DEVINL float nms_weight(float sim, float scale, bool is_soft, float thr) {
    const float w = exp_cr(scale * sim * sim);
    return (is_soft || sim <= thr) ? w : 0.0f;
}

BOD_DEVINL unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o > v ? o : v;
    }
    return v;
}

// ---------------------------------------------------------------------------
// generic kernel: all candidate state in global memory (any S)
// ---------------------------------------------------------------------------
__device__ void k3_generic(const K3Args& a, int b, unsigned long long (*warp_best)[32], float4* sel_box) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int S = a.num_survivors[b];
    const int Dmax = a.Dmax;
    const float4* corners = a.corners + (size_t)b * a.capacity;
    const float* score = a.score + (size_t)b * a.capacity;
    float* stale = a.stale + (size_t)b * a.capacity;
    float* cur = a.cur + (size_t)b * a.capacity;
    int32_t* begin = a.begin + (size_t)b * a.capacity;
    uint32_t* pend = a.pend + (size_t)b * a.capacity * kPendStride;
    uint32_t* member = a.member + (size_t)b * Dmax * a.words;
    const bool is_soft = a.soft_nms_sigma > 0.0f;
    const float scale = is_soft ? -0.5f / a.soft_nms_sigma : 0.0f;
    const float thr = a.iou_threshold;

    unsigned long long best = 0ull;
    for (int s = tid; s < S; s += blockDim.x) {
        const float sc = score[s];
        stale[s] = sc;
        begin[s] = 0;
#pragma unroll
        for (int w = 0; w < kMaskWords; ++w) pend[(size_t)s * kPendStride + w] = 0u;
        const bool in_queue = sc > -INFINITY;        // scores_data[i] > score_threshold (-inf): NaN stays out
        cur[s] = in_queue ? sc : -INFINITY;          // -inf is never enqueued => usable as "not in queue"
        if (in_queue) { const unsigned long long k = make_key(sc, s); best = k > best ? k : best; }
    }

    int r = 0;
    for (; r < Dmax; ++r) {
        best = warp_max_u64(best);
        if (lane == 0) warp_best[r & 1][warp] = best;
        __syncthreads();
        unsigned long long kx = (lane < nwarps) ? warp_best[r & 1][lane] : 0ull;
        kx = warp_max_u64(kx);
        if (kx == 0ull) break;                        // queue empty
        const int x = key_index(kx);
        const float4 bx = corners[x];
        if (tid == 0) {
            sel_box[r] = bx;
            a.nms_idx[(size_t)b * Dmax + r] = x;
            a.nms_score[(size_t)b * Dmax + r] = key_score(kx);
            a.centre_anchor[(size_t)b * Dmax + r] = a.surv_anchor[(size_t)b * a.capacity + x];
        }
        __syncthreads();

        best = 0ull;
        const int S32 = (S + 31) & ~31;
        for (int s = tid; s < S32; s += blockDim.x) {
            bool mem = false;
            if (s < S) {
                const float4 bs = corners[s];
                float u = cur[s];
                const bool in_queue = (u > -INFINITY) && (s != x);
                if (s == x) cur[s] = -INFINITY;
                const float xI1 = fmaxf(bs.y, bx.y), yI1 = fmaxf(bs.x, bx.x);
                const float xI2 = fminf(bs.w, bx.w), yI2 = fminf(bs.z, bx.z);
                const bool wellformed = (bs.x <= bs.z) && (bs.y <= bs.w) && (bx.x <= bx.z) && (bx.y <= bx.w);
                const bool maybe = !wellformed || (((xI2 - xI1) + 1.0f > 0.0f) && ((yI2 - yI1) + 1.0f > 0.0f));
                if (maybe) mem = repo_iou(bs, bx) > thr;                         // :316, strict >
                if (in_queue) {
                    float t = stale[s];
                    if (u != t && make_key(t, s) > kx) { t = u; stale[s] = t; begin[s] = r; }   // popped before x
                    if (maybe) {
                        const float sim = tf_iou(bs, bx);
                        const float w = nms_weight(sim, scale, is_soft, thr);
                        if (w != 1.0f) {
                            uint32_t* pm = pend + (size_t)s * kPendStride;
                            pm[r >> 5] |= 1u << (r & 31);
                            const int bg = begin[s];
                            float v = t;
                            for (int j = r; j >= bg; --j) {     // pending selections, newest first
                                if (!((pm[j >> 5] >> (j & 31)) & 1u)) continue;
                                const float sj = (j == r) ? sim : tf_iou(bs, sel_box[j]);
                                v = v * nms_weight(sj, scale, is_soft, thr);
                            }
                            u = v;
                            if (!is_soft && w == 0.0f) u = -INFINITY;   // hard-NMS: removed for good
                            cur[s] = u;
                        }
                    }
                    if (u > -INFINITY) { const unsigned long long k = make_key(u, s); best = k > best ? k : best; }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, mem);
            if (lane == 0) member[(size_t)r * a.words + (s >> 5)] = bal;
        }
    }
    if (tid == 0) a.num_dets[b] = r;
    for (int d = r + tid; d < Dmax; d += blockDim.x) {       // padding rows
        a.nms_idx[(size_t)b * Dmax + d] = -1;
        a.centre_anchor[(size_t)b * Dmax + d] = -1;
        a.nms_score[(size_t)b * Dmax + d] = 0.0f;
    }
}

// ---------------------------------------------------------------------------
// fast kernel: all per-candidate state that a round touches lives in shared memory
// ---------------------------------------------------------------------------
constexpr int kWCache = kPendStride; // cached non-unit weights per candidate (global, L2 resident)

struct K3Smem {
    unsigned long long best[2];
    unsigned long long warp_best[2][32];
    int list_n[2];
    float4 sel_box[kMaxOut];
};

__global__ void __launch_bounds__(kK3Threads, 1)
k3_softnms_kernel(K3Args a, int smem_S) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ K3Smem sm;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int S = a.num_survivors[b];
    if (S > smem_S) { k3_generic(a, b, sm.warp_best, sm.sel_box); return; }

    const int W = smem_S / 32;                                           // words per bit row
    float4* corn = reinterpret_cast<float4*>(dyn);                       // [smem_S] corners
    float* ucur = reinterpret_cast<float*>(corn + smem_S);               // [smem_S] up-to-date score, -inf = not queued
    float* stl = ucur + smem_S;                                          // [smem_S] score as of the last queue update
    uint32_t* dirty = reinterpret_cast<uint32_t*>(stl + smem_S);         // [W] u != stale
    uint32_t* mrow = dirty + W;                                          // [2][W] membership row being built
    uint16_t* list = reinterpret_cast<uint16_t*>(mrow + 2 * W);          // [kListMax] overlapping candidates
    uint8_t* beg = reinterpret_cast<uint8_t*>(list + kListMax);          // [smem_S] suppress_begin_index
    uint8_t* nw = beg + smem_S;                                          // [smem_S] cached weights since `beg`

    const int Dmax = a.Dmax;
    const float4* corners = a.corners + (size_t)b * a.capacity;
    const float* score = a.score + (size_t)b * a.capacity;
    // weight cache: weights of the selections in [beg, now) with a non-unit weight, oldest first
    float* wcache = reinterpret_cast<float*>(a.pend) + (size_t)b * a.capacity * kPendStride;
    uint32_t* member = a.member + (size_t)b * Dmax * a.words;
    const bool is_soft = a.soft_nms_sigma > 0.0f;
    const float scale = is_soft ? -0.5f / a.soft_nms_sigma : 0.0f;
    const float thr = a.iou_threshold;
    const int S32 = (S + 31) & ~31, nwords = S32 >> 5;

    // ---- load ----
    if (tid == 0) { sm.best[0] = sm.best[1] = 0ull; sm.list_n[0] = sm.list_n[1] = 0; }
    for (int w = tid; w < nwords; w += kK3Threads) { dirty[w] = 0u; mrow[w] = 0u; mrow[W + w] = 0u; }
    unsigned long long best = 0ull;
    for (int s = tid; s < S; s += kK3Threads) {
        corn[s] = corners[s];
        const float sc = score[s];
        const bool in_queue = sc > -INFINITY;         // scores_data[i] > score_threshold (-inf); NaN stays out
        ucur[s] = in_queue ? sc : -INFINITY;
        stl[s] = sc;
        beg[s] = 0; nw[s] = 0;
        if (in_queue) { const unsigned long long k = make_key(sc, s); best = k > best ? k : best; }
    }
    __syncthreads();                                   // sm.best initialised before the atomics below
    best = warp_max_u64(best);
    if (lane == 0 && best) atomicMax(&sm.best[0], best);
    __syncthreads();

    // the slow part of a round, for one candidate s that (possibly) overlaps the new centre
    auto process = [&](int s, int r, const float4 bx, uint32_t* mr, unsigned long long* next_best) {
        const float4 bs = corn[s];
        const float u0 = ucur[s];
        const int n = nw[s];
        // prefetch the cached weights while the IoUs and the exp are computed
        float wc[kWCache] = {};
        const float4* wp = reinterpret_cast<const float4*>(wcache + (size_t)s * kWCache);
        if (u0 > -INFINITY && n > 0) {
#pragma unroll
            for (int q = 0; q < kWCache / 4; ++q) {
                const float4 v = (4 * q < n) ? __ldcg(wp + q) : make_float4(1.f, 1.f, 1.f, 1.f);
                wc[4 * q] = v.x; wc[4 * q + 1] = v.y; wc[4 * q + 2] = v.z; wc[4 * q + 3] = v.w;
            }
        }
        if (repo_iou(bs, bx) > thr) atomicOr(&mr[s >> 5], 1u << (s & 31));       // :316, strict >
        if (!(u0 > -INFINITY)) return;                                            // selected earlier / never queued
        const float sim = tf_iou(bs, bx);
        const float w = nms_weight(sim, scale, is_soft, thr);
        float u = u0;
        if (w != 1.0f) {
            // u = stale * w * (cached weights, newest first)
            float v = stl[s] * w;
            if (n < kWCache) {
#pragma unroll
                for (int q = kWCache - 1; q >= 0; --q) if (q < n) v = v * wc[q];
                wcache[(size_t)s * kWCache + n] = w;
                nw[s] = (uint8_t)(n + 1);
            } else {
                // cache full (more than kWCache overlapping centres since the last update): recompute
                // every weight from the selected boxes; slot kWCache-1.. are not stored any more
                v = stl[s];
                for (int j = r; j >= (int)beg[s]; --j) {
                    const float sj = (j == r) ? sim : tf_iou(bs, sm.sel_box[j]);
                    const float wj = nms_weight(sj, scale, is_soft, thr);
                    if (wj != 1.0f) v = v * wj;
                }
                nw[s] = (uint8_t)kWCache;                                          // stays in overflow mode
            }
            u = (!is_soft && w == 0.0f) ? -INFINITY : v;                           // hard-NMS: removed for good
            ucur[s] = u;
            atomicOr(&dirty[s >> 5], 1u << (s & 31));
        }
        if (u > -INFINITY) atomicMax(next_best, make_key(u, s));
    };

    int r = 0;
    for (; r < Dmax; ++r) {
        const int cur_buf = r & 1, nxt_buf = cur_buf ^ 1;
        const unsigned long long kx = sm.best[cur_buf];
        if (kx == 0ull) break;                                                     // queue empty
        const int x = key_index(kx);
        const float4 bx = corn[x];
        uint32_t* mr = mrow + cur_buf * W;
        if (tid == 0) {
            sm.sel_box[r] = bx;
            sm.list_n[nxt_buf] = 0;        // last round's list: everyone finished reading it before the barrier
            a.nms_idx[(size_t)b * Dmax + r] = x;
            a.nms_score[(size_t)b * Dmax + r] = key_score(kx);
            a.centre_anchor[(size_t)b * Dmax + r] = a.surv_anchor[(size_t)b * a.capacity + x];
        }
        // flush the membership row of the previous round, then clear it for round r+1
        if (r > 0) {
            uint32_t* pr = mrow + nxt_buf * W;
            for (int w = tid; w < nwords; w += kK3Threads) { member[(size_t)(r - 1) * a.words + w] = pr[w]; pr[w] = 0u; }
        }
        const bool bx_ok = (bx.x <= bx.z) && (bx.y <= bx.w);

        // ---- pass A: overlap tests, commits, arg-max of the untouched candidates ----
        best = 0ull;
#pragma unroll 2
        for (int s = tid; s < S32; s += kK3Threads) {
            bool maybe = false, inline_it = false;
            uint32_t clear = 0u;
            if (s < S) {
                const float4 bs = corn[s];
                const float u = ucur[s];
                const bool in_queue = (u > -INFINITY) && (s != x);
                if (s == x) ucur[s] = -INFINITY;
                // no overlap even with the +1 pixel convention => TF IoU = 0 (weight exactly 1) and
                // repo IoU <= 0: nothing to do for this candidate in this round
                const float xI1 = fmaxf(bs.y, bx.y), yI1 = fmaxf(bs.x, bx.x);
                const float xI2 = fminf(bs.w, bx.w), yI2 = fminf(bs.z, bx.z);
                const bool wellformed = bx_ok && (bs.x <= bs.z) && (bs.y <= bs.w);
                maybe = !wellformed || (((xI2 - xI1) + 1.0f > 0.0f) && ((yI2 - yI1) + 1.0f > 0.0f));
                if (in_queue && ((dirty[s >> 5] >> (s & 31)) & 1u)) {
                    if (make_key(stl[s], s) > kx) {                                // TF popped it before x: update is final
                        stl[s] = u; beg[s] = (uint8_t)r; nw[s] = 0; clear = 1u;
                    }
                }
                if (in_queue && !maybe) { const unsigned long long k = make_key(u, s); best = k > best ? k : best; }
            }
            // this warp owns word s>>5 of `dirty` during pass A
            const unsigned clr = __ballot_sync(0xffffffffu, clear);
            const unsigned bal = __ballot_sync(0xffffffffu, maybe);
            if (clr && lane == 0) dirty[s >> 5] &= ~clr;
            if (bal) {
                int base = 0;
                const int leader = __ffs(bal) - 1;
                if (lane == leader) base = atomicAdd(&sm.list_n[cur_buf], __popc(bal));
                base = __shfl_sync(0xffffffffu, base, leader);
                const int pos = base + __popc(bal & ((1u << lane) - 1u));
                if (maybe) { if (pos < kListMax) list[pos] = (uint16_t)s; else inline_it = true; }
                if (__any_sync(0xffffffffu, inline_it)) {                          // list overflow: handle in place
                    __syncwarp();
                    if (inline_it) process(s, r, bx, mr, &sm.best[nxt_buf]);
                }
            }
        }
        best = warp_max_u64(best);
        if (lane == 0 && best) atomicMax(&sm.best[nxt_buf], best);
        __syncthreads();

        // ---- pass B: the compacted overlapping candidates ----
        if (tid == 0) sm.best[cur_buf] = 0ull;      // every thread has read kx; refilled from the next round's pass A on
        const int n = min(sm.list_n[cur_buf], kListMax);
        for (int e = tid; e < n; e += kK3Threads) process((int)list[e], r, bx, mr, &sm.best[nxt_buf]);
        __syncthreads();
    }
    // flush the last membership row
    if (r > 0) {
        const uint32_t* pr = mrow + ((r - 1) & 1) * W;
        for (int w = tid; w < nwords; w += kK3Threads) member[(size_t)(r - 1) * a.words + w] = pr[w];
    }
    if (tid == 0) a.num_dets[b] = r;
    for (int d = r + tid; d < Dmax; d += kK3Threads) {       // padding rows
        a.nms_idx[(size_t)b * Dmax + d] = -1;
        a.centre_anchor[(size_t)b * Dmax + d] = -1;
        a.nms_score[(size_t)b * Dmax + d] = 0.0f;
    }
}

cudaError_t launch_k3(const K3Args& a, cudaStream_t st) {
    int smem_S = a.capacity < kFastS ? a.capacity : kFastS;
    smem_S = (smem_S + 31) & ~31;
    const size_t smem = (size_t)smem_S * (16 + 4 + 4 + 1 + 1) + (size_t)(smem_S / 32) * 12 + (size_t)kListMax * 2;
    cudaError_t e = cudaFuncSetAttribute(k3_softnms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k3_softnms_kernel<<<a.B, kK3Threads, smem, st>>>(a, smem_S);
    return cudaGetLastError();
}

}  // namespace bod
