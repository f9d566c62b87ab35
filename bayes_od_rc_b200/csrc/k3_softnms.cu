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
constexpr int kFastS = 7680;          // candidates the shared-memory kernel holds

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
    // two REDUX ops: max of the score halves, then max of the (inverted) index halves among the winners
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, (hi == mh) ? lo : 0u);
    return ((unsigned long long)mh << 32) | (unsigned long long)ml;
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
// fast kernel: everything a round touches for every candidate lives in shared
// memory; the pending-weight lists live in global memory (L2 resident).
//
// (1) Commits are LAZY.  The eager formulation above must, every round, look at
// each candidate with pending weights to see whether TF would have popped it
// before x (stale key > kx) and, if so, fold its pending weights into the stale
// score.  Because the selection keys kx_0 > kx_1 > ... are decreasing, that
// history can be replayed exactly the next time the candidate is touched: with
// pending selections j_0 < j_1 < ..., the first pop happens at the smallest round
// rr in (j_0, now] with key(stale) > kx_rr and folds every pending j < rr
// (newest first), and so on.
// (2) Selections are BATCHED.  With every score up to date, walk the candidates in
// key order y_1 > y_2 > ...  y_1 is the next centre.  A later y_q whose weight
// against every centre accepted so far is exactly 1 keeps its score while every
// other score can only drop, so it is the next centre too, provided no skipped
// candidate (one that does overlap an accepted centre) can still outrank it; a
// skipped candidate's new score is at most score * weight * (1 + 1e-5).  A round
// therefore selects up to kBatch centres at once, bit-identically to one by one.
// A round is: (C) block-wide top-kTop of the current scores + acceptance,
// (A) one lean geometric overlap test of every candidate against the batch,
// compacted into per-warp list segments, (B) IoU / exp / replay for the
// overlapping candidates only.  No atomics on the critical path.
// ---------------------------------------------------------------------------
constexpr int kK3Warps = kK3Threads / 32;
constexpr int kSegCap = 256;          // overlap-list entries per warp and round
constexpr int kTop = 8;               // candidates examined per round
constexpr int kBatch = 8;             // centres selected per round at most

struct K3Smem {
    unsigned long long warp_best[2][32];         // generic kernel scratch
    unsigned long long top_w[kK3Warps][kTop];    // per-warp top keys of a round
    unsigned long long sel_key[kMaxOut];         // key (score, -index) of every selected centre
    float4 sel_box[kMaxOut];
    int seg_n[kK3Warps];                         // entries in each warp's list segment
    int batch_n;                                 // centres selected in this round
    int malformed;
};

struct K3Const {                                  // kernel-lifetime constants of the slow path (lives in shared memory)
    const float4* corn; float* ucur; float* stl; uint8_t* npend;
    float* pw; uint8_t* pj; int pstride;          // pending (weight, selection) lists, [S][pstride]
    const unsigned long long* sel_key; const float4* sel_box;
    float scale, thr; int is_soft;
};

// exp(y) rounded to binary32 for the soft-NMS argument range: y = -k/64 + r, table of exp(-k/64) in
// binary64 and a degree-6 Taylor polynomial in r (|r| <= 1/128, error < 2e-17): the binary64 value is
// within ~1e-16 of exp(y), so its binary32 rounding equals the correctly rounded one except with
// probability ~1e-8 per evaluation (same caveat as exp_cr).
__constant__ double c_exp_tab[129];
BOD_DEVINL float exp_neg_cr(float y) {
    if (!(y <= 0.0f && y >= -2.0f)) return exp_cr(y);
    const double yd = (double)y;
    const int k = __double2int_rn(yd * -64.0);
    const double r = fma((double)k, 0.015625, yd);
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return (float)(c_exp_tab[k] * p);
}
BOD_DEVINL float nms_weight_fast(float sim, float scale, bool is_soft, float thr) {
    if (sim == 0.0f) return 1.0f;                                   // exp(+-0) = 1 exactly (0 <= thr: never hard-suppressed)
    const float w = exp_neg_cr(scale * sim * sim);
    return (is_soft || sim <= thr) ? w : 0.0f;
}

// first round rr in [lo, hi] with sel_key[rr] < ks (keys strictly decrease; sel_key[hi] < ks is known)
BOD_DEVINL int first_pop_round(const unsigned long long* sel_key, unsigned long long ks, int lo, int hi) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ks > sel_key[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// The slow part of a round for one QUEUED candidate s whose box intersects the batch centres whose bits
// are set in `mask` (centre k of the batch is selection r0 + k).
__device__ __noinline__ void k3_process(const K3Const* C, const int r0, const uint32_t mask, const int s) {
    const float4 bs = C->corn[s];
    float u = C->ucur[s];
    const float scale = C->scale, thr = C->thr;
    const bool is_soft = C->is_soft != 0;
    int n = C->npend[s];
    float* wrow = C->pw + (size_t)s * C->pstride;
    uint8_t* jrow = C->pj + (size_t)s * C->pstride;
    // issue the loads of the newest pending weights first: their latency overlaps the IoU / exp arithmetic
    const float4* wrow4 = reinterpret_cast<const float4*>(wrow);
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
    const int blk = (n - 1) >> 2;
    float4 cur = (n > 0) ? wrow4[blk] : one4;
    float4 nxt = (blk > 0) ? wrow4[blk - 1] : one4;
    float st = C->stl[s];
    bool changed = false, folded = false, first = true;

    for (uint32_t rem = mask; rem; rem &= rem - 1) {
        const int r = r0 + __ffs(rem) - 1;
        const float sim = tf_iou(bs, C->sel_box[r]);
        const float w = nms_weight_fast(sim, scale, is_soft, thr);
        if (w == 1.0f) continue;                                                   // untouched by this centre
        changed = true;
        if (!is_soft && w == 0.0f) { u = -INFINITY; break; }                       // hard-NMS: removed for good
        const unsigned long long kxr = C->sel_key[r];
        if (n > 0 && make_key(st, s) > kxr) {
            // rare: TF popped this candidate at least once since its list was last touched: replay
            int i0 = 0;
            while (i0 < n) {
                const unsigned long long ks = make_key(st, s);
                if (!(ks > kxr)) break;                                            // keys decrease: no further pop
                const int rr = first_pop_round(C->sel_key, ks, (int)jrow[i0] + 1, r);
                int i1 = i0;
                while (i1 < n && (int)jrow[i1] < rr) ++i1;
                float v = st;
                for (int i = i1 - 1; i >= i0; --i) v = v * wrow[i];                // newest first
                st = v; i0 = i1;
            }
            if (i0 > 0) {                                                          // drop the folded entries
                for (int i = 0; i < n - i0; ++i) { wrow[i] = wrow[i0 + i]; jrow[i] = jrow[i0 + i]; }
                n -= i0; folded = true;
            }
            first = false;
        }
        // u = stale * w(x_r) * (pending weights, newest first)
        float v = st * w;
        if (first) {                                                               // weights prefetched in blocks of 4
            int b4 = blk;
            while (b4 >= 0) {
                const float4 nn = (b4 >= 2) ? wrow4[b4 - 2] : one4;
                const int top = n - 1 - 4 * b4;                                    // highest valid lane of this block (0..3)
                if (top >= 3) v = v * cur.w;
                if (top >= 2) v = v * cur.z;
                if (top >= 1) v = v * cur.y;
                v = v * cur.x;
                cur = nxt; nxt = nn; --b4;
            }
            first = false;
        } else {
            for (int i = n - 1; i >= 0; --i) v = v * wrow[i];
        }
        u = v;
        wrow[n] = w; jrow[n] = (uint8_t)r;
        ++n;
    }
    if (changed) {
        C->ucur[s] = u;
        C->npend[s] = (uint8_t)n;
        if (folded) C->stl[s] = st;
    }
}

// insert key into the descending list t[0..kTop)
BOD_DEVINL void top_insert(unsigned long long (&t)[kTop], unsigned long long key) {
    if (key > t[kTop - 1]) {
        t[kTop - 1] = key;
#pragma unroll
        for (int q = kTop - 1; q > 0; --q)
            if (t[q] > t[q - 1]) { const unsigned long long x = t[q]; t[q] = t[q - 1]; t[q - 1] = x; }
    }
}
// pop the heads of the lanes' sorted lists kTop times: every lane ends with the warp's top-kTop, descending
BOD_DEVINL void warp_top_merge(unsigned long long (&t)[kTop]) {
    unsigned long long out[kTop];
#pragma unroll
    for (int q = 0; q < kTop; ++q) {
        const unsigned long long m = warp_max_u64(t[0]);
        out[q] = m;
        if (m != 0ull && t[0] == m) {                  // keys are unique: exactly one lane pops
#pragma unroll
            for (int i = 0; i < kTop - 1; ++i) t[i] = t[i + 1];
            t[kTop - 1] = 0ull;
        }
    }
#pragma unroll
    for (int q = 0; q < kTop; ++q) t[q] = out[q];
}

__global__ void __launch_bounds__(kK3Threads, 1)
k3_softnms_kernel(K3Args a, int smem_S) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ K3Smem sm;
    __shared__ K3Const kc;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.num_survivors[b];
    if (S > smem_S) { k3_generic(a, b, sm.warp_best, sm.sel_box); return; }

    float4* corn = reinterpret_cast<float4*>(dyn);                       // [smem_S] corners
    float* ucur = reinterpret_cast<float*>(corn + smem_S);               // [smem_S] up-to-date score, -inf = not queued
    float* stl = ucur + smem_S;                                          // [smem_S] score as of the last fold
    uint32_t* list = reinterpret_cast<uint32_t*>(stl + smem_S);          // [kK3Warps][kSegCap] candidate | batch mask << 16
    uint8_t* npend = reinterpret_cast<uint8_t*>(list + kK3Warps * kSegCap);   // [smem_S] pending entries per candidate

    const int Dmax = a.Dmax;
    const float4* corners = a.corners + (size_t)b * a.capacity;
    const float* score = a.score + (size_t)b * a.capacity;
    const bool is_soft = a.soft_nms_sigma > 0.0f;
    const float scale = is_soft ? -0.5f / a.soft_nms_sigma : 0.0f;
    const float thr = a.iou_threshold;
    const int S32 = (S + 31) & ~31;

    if (tid == 0) {
        kc.corn = corn; kc.ucur = ucur; kc.stl = stl; kc.npend = npend;
        kc.pstride = a.pstride;
        kc.pw = a.pw + (size_t)b * a.fastS * a.pstride; kc.pj = a.pj + (size_t)b * a.fastS * a.pstride;
        kc.sel_key = sm.sel_key; kc.sel_box = sm.sel_box;
        kc.scale = scale; kc.thr = thr; kc.is_soft = is_soft ? 1 : 0;
        sm.malformed = 0;
    }
    __syncthreads();

    // ---- load; membership rows start out empty ----
    for (int s = tid; s < S; s += kK3Threads) {
        const float4 c = corners[s];
        corn[s] = c;
        if (!((c.x <= c.z) && (c.y <= c.w))) sm.malformed = 1;          // needs the canonicalising IoU path every round
        const float sc = score[s];
        ucur[s] = (sc > -INFINITY) ? sc : -INFINITY;                     // scores_data[i] > score_threshold (-inf); NaN stays out
        stl[s] = sc;
        npend[s] = 0;
    }
    __syncthreads();
    const bool all_maybe = sm.malformed != 0;

    long long tA = 0, tB = 0, tC = 0, tL = 0, t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    int r = 0, rounds = 0;
    while (r < Dmax) {
        if (a.dbg && tid == 0) t0 = clock64();
        // ---- pass C: block-wide top-kTop of the up-to-date scores ----
        unsigned long long top[kTop];
#pragma unroll
        for (int q = 0; q < kTop; ++q) top[q] = 0ull;
#pragma unroll 4
        for (int s = tid; s < S; s += kK3Threads) {
            const float u = ucur[s];
            if (u > -INFINITY) top_insert(top, make_key(u, s));
        }
        warp_top_merge(top);
        if (lane < kTop) {
            unsigned long long v = top[0];
#pragma unroll
            for (int q = 1; q < kTop; ++q) v = (lane == q) ? top[q] : v;
            sm.top_w[warp][lane] = v;
        }
        __syncthreads();
        if (warp == 0) {
            // merge the kK3Warps x kTop warp results, then accept centres in key order (see header)
            const unsigned long long* src = &sm.top_w[0][0];
            unsigned long long t2[kTop];
#pragma unroll
            for (int q = 0; q < kTop; ++q) t2[q] = 0ull;
#pragma unroll
            for (int i = 0; i < kK3Warps * kTop / 32; ++i) top_insert(t2, src[lane + 32 * i]);
            warp_top_merge(t2);
            // lane p < kTop holds candidate p
            unsigned long long myk = t2[0];
#pragma unroll
            for (int q = 1; q < kTop; ++q) myk = (lane == q) ? t2[q] : myk;
            if (lane >= kTop) myk = 0ull;
            // pairwise weights among the examined candidates: pair (q, i), i < q, on lane q*(q-1)/2 + i
            float wp = 1.0f;
            {
                int q = 1, base = 0;
                while (base + q <= lane) { base += q; ++q; }             // lane -> (q, i)
                const int i = lane - base;
                if (q < kTop) {
                    unsigned long long kq = t2[0], ki = t2[0];
#pragma unroll
                    for (int z = 1; z < kTop; ++z) { kq = (q == z) ? t2[z] : kq; ki = (i == z) ? t2[z] : ki; }
                    if (kq != 0ull && ki != 0ull)
                        wp = nms_weight_fast(tf_iou(corn[key_index(kq)], corn[key_index(ki)]), scale, is_soft, thr);
                }
            }
            static_assert(kTop * (kTop - 1) / 2 <= 32, "one lane per candidate pair");
            const uint32_t nonunit = __ballot_sync(0xffffffffu, wp != 1.0f);
            int m = 0;
            uint32_t acc = 0u;                                           // accepted candidates (bit q)
            if (t2[0] != 0ull) {
                acc = 1u; m = 1;
                float ub_max = -INFINITY;                                // best score a skipped candidate can still reach
#pragma unroll
                for (int q = 1; q < kTop; ++q) {
                    if (t2[q] == 0ull || r + m >= Dmax || m >= kBatch) break;
                    const float sq = key_score(t2[q]);
                    const int base = q * (q - 1) / 2;
                    const uint32_t hit = (nonunit >> base) & acc & ((1u << q) - 1u);   // accepted centres it overlaps
                    if (hit == 0u) {
                        if (!(sq > ub_max)) break;                       // a skipped candidate might still outrank it
                        acc |= 1u << q; ++m;
                    } else {
                        const float w = __shfl_sync(0xffffffffu, wp, base + __ffs(hit) - 1);
                        ub_max = fmaxf(ub_max, sq * w * 1.00001f);
                    }
                }
            }
            // accepted candidate q becomes selection r + (number of accepted before it)
            if (lane < kTop && ((acc >> lane) & 1u)) {
                const int pos = r + __popc(acc & ((1u << lane) - 1u));
                const int x = key_index(myk);
                sm.sel_box[pos] = corn[x];
                sm.sel_key[pos] = myk;
                ucur[x] = -INFINITY;                                     // leaves the queue
                a.nms_idx[(size_t)b * Dmax + pos] = x;
                a.nms_score[(size_t)b * Dmax + pos] = key_score(myk);
                a.centre_anchor[(size_t)b * Dmax + pos] = a.surv_anchor[(size_t)b * a.capacity + x];
            }
            if (lane == 0) sm.batch_n = m;
        }
        __syncthreads();
        const int m = sm.batch_n;
        if (m == 0) break;                                                     // queue empty
        if (a.dbg && tid == 0) t1 = clock64();

        // ---- pass A: geometric overlap of every candidate with the batch centres ----
        float4 bxs[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; ++q) bxs[q] = sm.sel_box[r + (q < m ? q : 0)];
        int cnt = 0;                                                           // entries in this warp's segment
        uint32_t* seg = list + warp * kSegCap;
#pragma unroll 2
        for (int s = tid; s < S32; s += kK3Threads) {
            uint32_t mask = 0u;
            if (s < S && ucur[s] > -INFINITY) {
                const float4 bs = corn[s];
                // TF's intersection area is max(dy,0)*max(dx,0): no positive intersection => IoU = 0 and the
                // weight is exactly 1, the centre does nothing to this candidate
#pragma unroll
                for (int q = 0; q < kBatch; ++q) {
                    if (q < m) {
                        const float dx = fminf(bs.w, bxs[q].w) - fmaxf(bs.y, bxs[q].y);
                        const float dy = fminf(bs.z, bxs[q].z) - fmaxf(bs.x, bxs[q].x);
                        if ((dx > 0.0f && dy > 0.0f) || all_maybe) mask |= 1u << q;
                    }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, mask != 0u);
            if (mask) {
                const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                if (pos < kSegCap) seg[pos] = (uint32_t)s | (mask << 16);
                else k3_process(&kc, r, mask, s);                              // segment overflow: handle in place
            }
            cnt += __popc(bal);
        }
        if (lane == 0) sm.seg_n[warp] = min(cnt, kSegCap);
        __syncthreads();
        if (a.dbg && tid == 0) t2 = clock64();

        // ---- pass B: the compacted overlapping candidates (entry e -> segment, slot) ----
        int total = 0;
#pragma unroll
        for (int w = 0; w < kK3Warps; ++w) total += sm.seg_n[w];
        for (int base = 0; base < total; base += kK3Threads) {
            int e = base + tid, sgw = -1, slot = 0;
#pragma unroll
            for (int w = 0; w < kK3Warps; ++w) {
                const int c = sm.seg_n[w];
                if (sgw < 0 && e >= 0 && e < c) { sgw = w; slot = e; }
                e -= c;
            }
            if (sgw >= 0) { const uint32_t ent = list[sgw * kSegCap + slot]; k3_process(&kc, r, ent >> 16, (int)(ent & 0xFFFFu)); }
        }
        __syncthreads();
        if (a.dbg && tid == 0) { t3 = clock64(); tC += t1 - t0; tA += t2 - t1; tB += t3 - t2; tL += total; }
        r += m;
        ++rounds;
    }
    if (a.dbg && tid == 0) {
        a.dbg[b * 8 + 0] = tA; a.dbg[b * 8 + 1] = tB; a.dbg[b * 8 + 2] = tL; a.dbg[b * 8 + 3] = rounds; a.dbg[b * 8 + 4] = S;
        a.dbg[b * 8 + 5] = tC; a.dbg[b * 8 + 6] = r;
    }
    if (tid == 0) a.num_dets[b] = r;
    for (int d = r + tid; d < Dmax; d += kK3Threads) {       // padding rows
        a.nms_idx[(size_t)b * Dmax + d] = -1;
        a.centre_anchor[(size_t)b * Dmax + d] = -1;
        a.nms_score[(size_t)b * Dmax + d] = 0.0f;
    }
}

// ---------------------------------------------------------------------------
// cluster membership (inference_utils.py:214-215 + :316): bit s of row d <=>
// bbox_iou_vuvu(survivor s, centre d) > threshold, for the D centre columns only.
// One CTA per (centre, image); fully parallel, off the sequential path.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_membership_kernel(K3Args a) {
    const int d = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    if (d >= a.num_dets[b]) return;
    const int S = a.num_survivors[b];
    const float4* corners = a.corners + (size_t)b * a.capacity;
    uint32_t* row = a.member + ((size_t)b * a.Dmax + d) * a.words;
    const float4 bx = corners[a.nms_idx[(size_t)b * a.Dmax + d]];
    const bool bx_ok = (bx.x <= bx.z) && (bx.y <= bx.w);
    const float thr = a.iou_threshold;
    const int S32 = (S + 31) & ~31;
    for (int s = tid; s < S32; s += 256) {
        bool mem = false;
        if (s < S) {
            const float4 bs = corners[s];
            // no overlap even with the +1 pixel convention ((hi - lo) + 1 > 0 <=> hi - lo > -1 in binary32)
            // => intersection 0 => IoU <= 0 <= threshold
            const float dx = fminf(bs.w, bx.w) - fmaxf(bs.y, bx.y);
            const float dy = fminf(bs.z, bx.z) - fmaxf(bs.x, bx.x);
            const bool wellformed = bx_ok && (bs.x <= bs.z) && (bs.y <= bs.w);
            if (!wellformed || (dx > -1.0f && dy > -1.0f)) mem = repo_iou(bs, bx) > thr;   // strict >
        }
        const unsigned bal = __ballot_sync(0xffffffffu, mem);
        if (lane == 0) row[s >> 5] = bal;
    }
}

int k3_fast_capacity(int capacity) {
    int s = capacity < kFastS ? capacity : kFastS;
    return (s + 31) & ~31;
}

static bool g_exp_tab_ready[64] = {false};

cudaError_t launch_k3(const K3Args& a, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !g_exp_tab_ready[dev]) {
        double tab[129];
        for (int k = 0; k <= 128; ++k) tab[k] = exp(-(double)k / 64.0);
        cudaError_t e0 = cudaMemcpyToSymbol(c_exp_tab, tab, sizeof tab);
        if (e0 != cudaSuccess) return e0;
        g_exp_tab_ready[dev] = true;
    }
    const int smem_S = a.fastS;
    const size_t smem = (size_t)smem_S * (16 + 4 + 4 + 1) + (size_t)kK3Warps * kSegCap * 4;
    cudaError_t e = cudaFuncSetAttribute(k3_softnms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k3_softnms_kernel<<<a.B, kK3Threads, smem, st>>>(a, smem_S);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k3_membership_kernel<<<dim3(a.Dmax, a.B), 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bod
