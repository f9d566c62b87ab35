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
// candidates whose stale key (t_i, -i) exceeds (u_x, -x).  Weights equal to exactly
// 1.0f (IoU 0, the overwhelmingly common case) leave a score bit-identical, so
// boxes that do not overlap a new centre need no arithmetic at all.
//
// Three kernels-in-one, chosen per image by its survivor count S:
//   S <= shared-memory pool (~7.4 k): the fast path below, every per-candidate array in shared memory;
//   S <= 65 535: the same code with the per-candidate arrays in (L2-resident) global scratch rows;
//   beyond: k3_generic, the literal one-round-per-selection formulation (list entries hold 16-bit indices).
// One CTA per image (images are independent); B CTAs run concurrently.
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

constexpr int kK3Threads = 1024;
constexpr int kFastS = 7424;          // candidates the shared-memory kernel holds

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
// fast kernel: everything a round touches lives in shared memory.
//
// (1) Selections are BATCHED.  With every score up to date, walk the candidates in
// key order y_1 > y_2 > ...  y_1 is the next centre.  A later y_q whose weight
// against every centre accepted so far is exactly 1 keeps its score while every
// other score can only drop, so it is the next centre too, provided no skipped
// candidate (one that does overlap an accepted centre) can still outrank it; a
// skipped candidate's new score is at most score * weight * (1 + 5e-5).  A round
// therefore selects up to kBatch centres at once, bit-identically to one by one.
// (2) Update epochs are resolved when a candidate is touched.  Each candidate
// keeps the weights it has collected since TF last popped it ("pending", oldest
// first), its score as of that pop (stale) and its up-to-date score
// u = stale * w_newest * ... * w_oldest.  TF pops a candidate with pending
// weights right before selection j iff key(stale) > key(selection j) -- the
// selection keys decrease -- and then folds every pending weight into the stale
// score, newest first.  A candidate overlapping the round's batch walks the
// <= kBatch selections of the batch in order; pops that happened in rounds that
// did not touch it fold the same (whole) pending list and are caught by the
// first comparison of its next walk.
// (3) The block-wide top-kTop of the NEXT round is folded into the same passes:
// every thread tracks the best two keys it has seen (untouched candidates in
// pass A, updated ones in pass B); warps merge by popping heads (REDUX) into a
// list of kTop1 keys each.  A list is cut where a thread runs out of tracked
// keys (its third best is unknown), and the acceptance walk stops at the largest
// such cut -- fewer centres in that round, never a wrong one.
// (4) The expensive arithmetic runs one (candidate, centre) PAIR per thread.
// A round is: acceptance (a team of four warps behind a named barrier: rank the
// 128 listed keys by counting, one pair of examined candidates per thread, then
// warp 0 walks them) | pass A: one lean overlap test of every survivor against
// the batch (a superset filter: the centre grown by 2 px), overlapping pairs
// compacted into the warp's own list segment | pass B1: IoU + exp for every
// listed pair, and the pair's cluster-membership bit (bbox_iou_vuvu > threshold,
// inference_utils.py:316) | pass B2: the epoch walk of every listed candidate
// over its precomputed weights.  A survivor is scanned by the same thread every
// round, so B1 / B2 work on the warp's own segment without a block barrier: two
// block barriers per round, no global memory on the critical path: the first
// psm pending weights of every candidate live in shared memory (psm is chosen
// per image from its survivor count), the rest spill to global rows.
// ---------------------------------------------------------------------------
constexpr int kK3Warps = kK3Threads / 32;
constexpr int kSegCap = 4096 / kK3Warps; // (candidate, centre) pairs listed per warp and round
constexpr int kTop1 = 128 / kK3Warps;  // keys every warp lists per round
constexpr int kTop = 16;              // candidates examined per round
constexpr int kBatch = 16;            // centres selected per round at most
constexpr int kExpTab = 129;
constexpr int kTeamWarps = 4;         // warps that share a round's acceptance work

BOD_DEVINL void team_barrier() {      // named barrier 1: the acceptance team only
    asm volatile("bar.sync 1, %0;" ::"n"(kTeamWarps * 32) : "memory");
}

struct K3Smem {
    unsigned long long warp_best[2][32];         // generic kernel scratch
    unsigned long long top_w[kK3Warps][kTop1];   // per-warp top keys (descending, 0 = none)
    unsigned long long bound_w[kK3Warps];        // every key of the warp that is not in top_w is below this (0: there is none)
    unsigned long long sel_key[kMaxOut];         // key (score, -index) of every selected centre
    float4 sel_box[kMaxOut];
    unsigned long long cand_key[kTop];           // the round's examined candidates ...
    float4 cand_box[kTop];
    float4 batch_ebox[kBatch];                   // the round's centres grown by 2 px: pass A's lean overlap test
    float wpair[kTop][kTop];                     // ... their pairwise soft-NMS weights [q][i], i < q
    uint32_t rowmask[kTop];                      // bit i of row q: wpair[q][i] != 1
    double exp_tab[kExpTab];                     // exp(-k/64)
    int batch_n;                                 // centres selected in this round
    int malformed;
};

struct K3State {                                  // kernel-lifetime constants (registers)
    const float4* corn; float* ucur; float* stl; uint8_t* npend;
    float* pws; int S32, psm;                     // pending weights in shared memory: entry i of candidate s at pws[s*psm+i], i < psm (psm % 4 == 0)
    float* pwg; int pstride;                      // spill rows in global memory: entry i >= psm at pwg[s*pstride+i]
    uint32_t* member; int words;                  // membership rows of this image
    float scale, thr; bool is_soft;
};

// list entry: candidate | centre-of-batch << 16 | first pair of its candidate << 20 | candidate queued << 21 | pairs of the candidate << 22
BOD_DEVINL uint32_t ent_make(int s, int q, bool head, bool queued, int c) {
    return (uint32_t)s | ((uint32_t)q << 16) | ((uint32_t)head << 20) | ((uint32_t)queued << 21) | ((uint32_t)c << 22);
}
BOD_DEVINL int ent_s(uint32_t e) { return (int)(e & 0xFFFFu); }
BOD_DEVINL int ent_q(uint32_t e) { return (int)((e >> 16) & 15u); }
BOD_DEVINL bool ent_head(uint32_t e) { return (e >> 20) & 1u; }
BOD_DEVINL bool ent_queued(uint32_t e) { return (e >> 21) & 1u; }
BOD_DEVINL int ent_pairs(uint32_t e) { return (int)((e >> 22) & 31u); }

// exp(y) rounded to binary32 for the soft-NMS argument range: y = -k/64 + r, table of exp(-k/64) in
// binary64 and a degree-6 Taylor polynomial in r (|r| <= 1/128, error < 2e-17): the binary64 value is
// within ~1e-16 of exp(y), so its binary32 rounding equals the correctly rounded one except with
// probability ~1e-8 per evaluation (same caveat as exp_cr).
__constant__ double c_exp_tab[kExpTab];
BOD_DEVINL float exp_neg_cr(float y, const double* tab) {
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
    return (float)(tab[k] * p);
}
BOD_DEVINL float nms_weight_fast(float sim, float scale, bool is_soft, float thr, const double* tab) {
    if (sim == 0.0f) return 1.0f;                                   // exp(+-0) = 1 exactly (0 <= thr: never hard-suppressed)
    const float w = exp_neg_cr(scale * sim * sim, tab);
    return (is_soft || sim <= thr) ? w : 0.0f;
}

// bbox_iou_vuvu(survivor, centre) > threshold (strict), skipping the division when the boxes cannot
// overlap even with the +1 pixel convention ((hi - lo) + 1 > 0 <=> hi - lo > -1 in binary32)
BOD_DEVINL bool is_member(const float4 bs, const float4 bx, float thr) {
    const float dx = fminf(bs.w, bx.w) - fmaxf(bs.y, bx.y);
    const float dy = fminf(bs.z, bx.z) - fmaxf(bs.x, bx.x);
    const bool wellformed = (bx.x <= bx.z) && (bx.y <= bx.w) && (bs.x <= bs.z) && (bs.y <= bs.w);
    if (!wellformed || (dx > -1.0f && dy > -1.0f)) return repo_iou(bs, bx) > thr;
    return false;
}

// best two keys a thread has seen in a round
struct Top2 {
    unsigned long long a = 0ull, b = 0ull;
    BOD_DEVINL void add(unsigned long long k) {
        if (k > a) { b = a; a = k; } else if (k > b) { b = k; }
    }
};

BOD_DEVINL void pend_put(const K3State& C, int s, int i, float w) {
    if (i < C.psm) C.pws[s * C.psm + i] = w; else C.pwg[(size_t)s * C.pstride + i] = w;
}
// st * (pending weights, newest first).  The shared-memory part of the list is a row of the candidate
// (psm is a multiple of 4): blocks of four weights per LDS.128, the next block in flight while the
// current one is multiplied in.
BOD_DEVINL float pend_product(const K3State& C, int s, int n, float st) {
    float v = st;
    int i = n - 1;
    for (; i >= C.psm; --i) v = v * C.pwg[(size_t)s * C.pstride + i];        // spilled entries (rare)
    if (i < 0) return v;
    const float4* row = reinterpret_cast<const float4*>(C.pws + s * C.psm);
    int blk = i >> 2;
    float4 cur = row[blk];
    const int top = i & 3;
    {
        const float4 nxt = row[blk > 0 ? blk - 1 : 0];
        if (top >= 3) v = v * cur.w;
        if (top >= 2) v = v * cur.z;
        if (top >= 1) v = v * cur.y;
        v = v * cur.x;
        cur = nxt;
    }
    for (--blk; blk >= 0; --blk) {
        const float4 nxt = row[blk > 0 ? blk - 1 : 0];
        v = v * cur.w; v = v * cur.z; v = v * cur.y; v = v * cur.x;
        cur = nxt;
    }
    return v;
}

// Epoch walk of one QUEUED candidate s over the round's batch (selections r0 .. r0+m-1), driven by the
// candidate's listed pairs in ascending centre order: step(q, w) for every batch centre q it overlaps
// (w = its soft-NMS weight against that centre), then finish().  TF pops s right before selection j iff
// it has pending weights and key(stale) > key(selection j); between two of its pairs the pending list
// does not change and the selection keys decrease, so such a pop exists in (q_prev, q] iff
// key(stale) > key(selection q), and wherever it happens it folds the same list.
struct K3Walk {
    float st;
    int n, n_in, last_q;
    bool folded, dead;
    unsigned long long ks;
    BOD_DEVINL void begin(const K3State& C, int s) {
        st = C.stl[s]; n = C.npend[s]; n_in = n; last_q = -1; folded = false; dead = false;
        ks = make_key(st, s);
    }
    BOD_DEVINL void fold(const K3State& C, int s) {        // newest first
        st = pend_product(C, s, n, st);
        n = 0; folded = true;
        ks = make_key(st, s);
    }
    BOD_DEVINL void step(const K3State& C, const K3Smem& sm, int r0, int s, int q, float w) {
        if (dead) return;
        if (n > 0 && ks > sm.sel_key[r0 + q]) fold(C, s);
        last_q = q;
        if (w != 1.0f) {
            if (!C.is_soft && w == 0.0f) { dead = true; return; }           // hard-NMS: removed for good
            pend_put(C, s, n, w);
            ++n;
        }
    }
    // returns the candidate's new up-to-date score (-inf: removed by hard-NMS)
    BOD_DEVINL float finish(const K3State& C, const K3Smem& sm, int r0, int m, int s) {
        if (dead) { C.ucur[s] = -INFINITY; return -INFINITY; }
        if (last_q < m - 1 && n > 0 && ks > sm.sel_key[r0 + m - 1]) fold(C, s);   // popped before a later selection of the batch
        if (!folded && n == n_in) return C.ucur[s];        // every weight was exactly 1: untouched
        const float u = pend_product(C, s, n, st);
        C.ucur[s] = u;
        C.npend[s] = (uint8_t)n;
        if (folded) C.stl[s] = st;
        return u;
    }
};

// A whole candidate in place (only when a warp's list segment is full): membership bits and, if the
// candidate is queued, its walk with the weights computed on the fly.
__device__ __noinline__ float k3_process_inplace(const K3State* Cp, const K3Smem& sm, const int r0, const int m, const uint32_t mask,
                                                 const int s, const bool queued) {
    const K3State C = *Cp;                                  // a shared-memory copy: nothing of the caller's is forced to the stack
    const float4 bs = C.corn[s];
    for (uint32_t rem = mask; rem; rem &= rem - 1) {
        const int q = __ffs(rem) - 1;
        if (is_member(bs, sm.sel_box[r0 + q], C.thr))
            atomicOr(&C.member[(size_t)(r0 + q) * C.words + (s >> 5)], 1u << (s & 31));
    }
    if (!queued) return -INFINITY;
    K3Walk wk;
    wk.begin(C, s);
    for (uint32_t rem = mask; rem; rem &= rem - 1) {
        const int q = __ffs(rem) - 1;
        wk.step(C, sm, r0, s, q, nms_weight_fast(tf_iou(bs, sm.sel_box[r0 + q]), C.scale, C.is_soft, C.thr, sm.exp_tab));
    }
    return wk.finish(C, sm, r0, m, s);
}

// Warp-level merge of the lanes' Top2 pairs: every lane returns with the warp's best keys in out[]
// (descending, 0 = none), cut after the first key whose lane has nothing tracked behind it; `bound`
// is that key (0 when the list was not cut and the warp has no further keys).
BOD_DEVINL void warp_top_merge2(Top2 t, unsigned long long (&out)[kTop1], unsigned long long& bound) {
    unsigned long long cur = t.a, nxt = t.b;
    bool spent = false;                                     // this lane's second key has been promoted already
    bool cut = false;
    bound = 0ull;
#pragma unroll
    for (int q = 0; q < kTop1; ++q) {
        const unsigned long long mx = cut ? 0ull : warp_max_u64(cur);
        out[q] = mx;
        const bool mine = (mx != 0ull) && (cur == mx);      // keys are unique: exactly one lane pops
        const bool exhausted = mine && spent;               // its third best is unknown
        if (mine) { cur = nxt; nxt = 0ull; spent = true; }
        if (!cut && __any_sync(0xffffffffu, exhausted)) { cut = true; bound = mx; }
    }
    if (!cut) bound = out[kTop1 - 1];                       // untracked keys are below the last one listed
}

__global__ void __launch_bounds__(kK3Threads, 1)
k3_softnms_kernel(K3Args a, int smem_S, int pool_bytes) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ K3Smem sm;
    __shared__ K3State kc;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.num_survivors[b];
    // more survivors than the shared-memory pool holds: same algorithm with the per-candidate state in
    // global memory (L2-resident scratch rows of the workspace); beyond 16-bit list entries: the literal kernel
    const bool big = S > smem_S;
    if (big && S > a.max_rows) { k3_generic(a, b, sm.warp_best, sm.sel_box); return; }

    const int S32 = (S + 31) & ~31;
    uint32_t* list = reinterpret_cast<uint32_t*>(dyn);                    // [kK3Warps][kSegCap] pair entries (ent_make)
    float* wl = reinterpret_cast<float*>(list + kK3Warps * kSegCap);      // [kK3Warps][kSegCap] their soft-NMS weights
    float4* corn = reinterpret_cast<float4*>(wl + kK3Warps * kSegCap);    // [S32] corners
    float* ucur = reinterpret_cast<float*>(corn + S32);                   // [S32] up-to-date score, -inf = not queued
    float* stl = ucur + S32;                                              // [S32] score as of the last fold
    // pending weights per candidate held in shared memory: whatever fits behind the fixed arrays
    int psm = S32 > 0 ? (pool_bytes - S32 * 25) / (S32 * 4) : 0;
    psm = psm < 0 ? 0 : (psm > a.pstride ? a.pstride : psm);
    if (a.psm_max >= 0 && psm > a.psm_max) psm = a.psm_max;              // diagnostics: force the global spill rows
    psm &= ~3;                                                            // rows are read four weights at a time
    float* pws = stl + S32;                                               // [S32][psm]
    uint8_t* npend = reinterpret_cast<uint8_t*>(pws + (size_t)psm * S32); // [S32] pending entries per candidate
    if (big) {
        corn = const_cast<float4*>(a.corners + (size_t)b * a.capacity);   // read in place, never written
        ucur = a.cur + (size_t)b * a.capacity;
        stl = a.stale + (size_t)b * a.capacity;
        npend = reinterpret_cast<uint8_t*>(a.begin + (size_t)b * a.capacity);
        psm = 0; pws = nullptr;
    }

    const int Dmax = a.Dmax;
    const float4* corners = a.corners + (size_t)b * a.capacity;
    const float* score = a.score + (size_t)b * a.capacity;
    K3State C;
    C.corn = corn; C.ucur = ucur; C.stl = stl; C.npend = npend;
    C.pws = pws; C.S32 = S32; C.psm = psm;
    C.pwg = a.pw + (size_t)b * a.pw_rows * a.pstride; C.pstride = a.pstride;
    C.member = a.member + (size_t)b * Dmax * a.words; C.words = a.words;
    C.is_soft = a.soft_nms_sigma > 0.0f;
    C.scale = C.is_soft ? -0.5f / a.soft_nms_sigma : 0.0f;
    C.thr = a.iou_threshold;

    if (tid == 0) { sm.malformed = 0; kc = C; }
    for (int k = tid; k < kExpTab; k += kK3Threads) sm.exp_tab[k] = c_exp_tab[k];
    // membership rows start out empty (only the words K4 / bod_fetch_members read)
    {
        const int nw = S32 >> 5;
        for (int d = warp; d < Dmax; d += kK3Warps)
            for (int w = lane; w < nw; w += 32) C.member[(size_t)d * a.words + w] = 0u;
    }
    __syncthreads();

    // ---- load; the first round's top keys come from the initial scores ----
    Top2 t2k;
    for (int s = tid; s < S; s += kK3Threads) {
        const float4 c = corners[s];
        if (!big) corn[s] = c;
        // corners out of order need the canonicalising IoU path for every pair; so do coordinates so large
        // that the 2 px margin of pass A's lean test is not safely above their rounding
        if (!((c.x <= c.z) && (c.y <= c.w)) || !(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))) < 1.0e5f))
            sm.malformed = 1;
        const float sc = score[s];
        const bool queued = sc > -INFINITY;                              // scores_data[i] > score_threshold (-inf); NaN stays out
        ucur[s] = queued ? sc : -INFINITY;
        stl[s] = sc;
        npend[s] = 0;
        if (queued) t2k.add(make_key(sc, s));
    }

    long long tA = 0, tB = 0, tC = 0, tL = 0, t0 = 0, t1 = 0, t2c = 0, t3 = 0, tm = 0, tw = 0, tM = 0, tW = 0, tx1 = 0, tx2 = 0, tR = 0, tP = 0;
    int r = 0, rounds = 0;
    while (r < Dmax) {
        if (a.dbg && tid == 0) t0 = clock64();
        // ---- warp lists of this round's best keys ----
        {
            unsigned long long out[kTop1], bound;
            warp_top_merge2(t2k, out, bound);
            if (lane < kTop1) {
                unsigned long long v = out[0];
#pragma unroll
                for (int q = 1; q < kTop1; ++q) v = (lane == q) ? out[q] : v;
                sm.top_w[warp][lane] = v;
            }
            if (lane == 0) sm.bound_w[warp] = bound;
        }
        if (tid < kTop) { sm.cand_key[tid] = 0ull; sm.rowmask[tid] = 0u; }     // filled by the acceptance team below
        if (a.dbg && tid == 0) tm = clock64();
        __syncthreads();
        if (a.dbg && tid == 0) tw = clock64();
        if (warp < kTeamWarps) {
            // ---- acceptance, by the first four warps (named barrier 1 among them) ----
            // (i) the block's top-kTop keys: thread t ranks entry t of the kK3Warps x kTop1 = 128 listed keys
            // by counting the larger ones; G = the largest cut of any warp's list: below it an untracked key
            // might outrank a listed one, so such entries are dropped (validity is a prefix of the order)
            static_assert(kK3Warps * kTop1 == kTeamWarps * 32, "one listed key per thread of the acceptance team");
            const unsigned long long* flat = &sm.top_w[0][0];
            {
                unsigned long long G = (lane < kK3Warps) ? sm.bound_w[lane] : 0ull;
                G = warp_max_u64(G);
                const unsigned long long key = flat[tid];
                if (key != 0ull && key >= G) {
                    // the listed scores are non-negative floats in practice, so the high words alone decide almost
                    // every comparison: count on 32-bit words, resolve equal scores (ties) on the index words
                    const uint32_t khi = (uint32_t)(key >> 32), klo = (uint32_t)key;
                    const uint4* f4 = reinterpret_cast<const uint4*>(flat);        // (lo, hi, lo, hi) of two keys
                    int r0 = 0, r1 = 0;
#pragma unroll 16
                    for (int j = 0; j < kTeamWarps * 16; ++j) {
                        const uint4 kk = f4[j];
                        r0 += (kk.y > khi) || (kk.y == khi && kk.x > klo);
                        r1 += (kk.w > khi) || (kk.w == khi && kk.z > klo);
                    }
                    const int rank = r0 + r1;
                    if (rank < kTop) { sm.cand_key[rank] = key; sm.cand_box[rank] = corn[key_index(key)]; }
                }
            }
            team_barrier();
            if (a.dbg && tid == 0) tx1 = clock64();
            // (ii) pairwise weights among the examined candidates, one pair (q, i), i < q, per thread.  Most pairs
            // do not intersect at all (weight exactly 1): the division + exp runs only for those that do
            if (tid < kTop * (kTop - 1) / 2) {
                int q = 1, base = 0;
                while (base + q <= tid) { base += q; ++q; }
                const int i = tid - base;
                if (sm.cand_key[q] != 0ull) {                             // then candidate i < q exists too
                    const float4 bq = sm.cand_box[q], bi = sm.cand_box[i];
                    const float dx = fminf(bq.w, bi.w) - fmaxf(bq.y, bi.y);
                    const float dy = fminf(bq.z, bi.z) - fmaxf(bq.x, bi.x);
                    float w = 1.0f;
                    if ((dx > 0.0f && dy > 0.0f) || sm.malformed != 0)
                        w = nms_weight_fast(tf_iou(bq, bi), C.scale, C.is_soft, C.thr, sm.exp_tab);
                    sm.wpair[q][i] = w;
                    if (w != 1.0f) atomicOr(&sm.rowmask[q], 1u << i);
                }
            }
            team_barrier();
            if (a.dbg && tid == 0) tx2 = clock64();
        }
        if (warp == 0) {
            // (iii) accept centres in key order (see header); every lane walks, lane p < kTop owns candidate p
            const unsigned long long myk = (lane < kTop) ? sm.cand_key[lane] : 0ull;
            const int nvalid = __popc(__ballot_sync(0xffffffffu, myk != 0ull));      // candidates are a prefix
            // everything the walk reads is fetched up front; only the accept / skip decisions are sequential
            float sqv[kTop];
            uint32_t rowv[kTop];
#pragma unroll
            for (int q = 0; q < kTop; ++q) { sqv[q] = key_score(sm.cand_key[q]); rowv[q] = sm.rowmask[q]; }
            int m = 0;
            uint32_t acc = 0u;                                           // accepted candidates (bit q)
            if (nvalid > 0) {
                acc = 1u; m = 1;
                float ub_max = -INFINITY;                                // best score a skipped candidate can still reach
#pragma unroll
                for (int q = 1; q < kTop; ++q) {
                    if (q >= nvalid || r + m >= Dmax || m >= kBatch) break;
                    const float sq = sqv[q];
                    const uint32_t hit = rowv[q] & acc;                  // accepted centres it overlaps
                    if (hit == 0u) {
                        if (!(sq > ub_max)) break;                       // a skipped candidate might still outrank it
                        acc |= 1u << q; ++m;
                    } else {
                        const float w = sm.wpair[q][__ffs(hit) - 1];
                        ub_max = fmaxf(ub_max, sq * w * 1.00005f);     // 2n+2 roundings apart, n <= 255
                    }
                }
            }
            // accepted candidate q becomes selection r + (number of accepted before it)
            if (lane < kTop && ((acc >> lane) & 1u)) {
                const int pos = r + __popc(acc & ((1u << lane) - 1u));
                const int x = key_index(myk);
                const float4 cb = sm.cand_box[lane];
                sm.sel_box[pos] = cb;
                sm.batch_ebox[pos - r] = make_float4(cb.x - 2.0f, cb.y - 2.0f, cb.z + 2.0f, cb.w + 2.0f);
                sm.sel_key[pos] = myk;
                ucur[x] = -INFINITY;                                     // leaves the queue
                a.nms_idx[(size_t)b * Dmax + pos] = x;
                a.nms_score[(size_t)b * Dmax + pos] = key_score(myk);
            }
            if (lane == 0) sm.batch_n = m;
        }
        __syncthreads();
        const int m = sm.batch_n;
        if (m == 0) break;                                                     // queue empty
        if (a.dbg && tid == 0) t1 = clock64();
        const bool all_maybe = sm.malformed != 0;

        // ---- pass A: geometric overlap of every survivor with the batch centres (queued or not: selected
        // and removed survivors still belong to clusters) ----
        t2k = Top2();
        int cnt = 0;                                                           // entries in this warp's segment
        uint32_t* seg = list + warp * kSegCap;
        const int seg_cap = (a.seg_cap >= 0 && a.seg_cap < kSegCap) ? a.seg_cap : kSegCap;   // tests shrink it
        for (int s = tid; s < S32; s += kK3Threads) {
            uint32_t mask = 0u;
            bool queued = false;
            if (s < S) {
                const float u = ucur[s];
                queued = u > -INFINITY;
                const float4 bs = corn[s];
                // No overlap even with the +1 pixel convention => not a member, and TF's intersection area
                // max(dy,0)*max(dx,0) is 0 => IoU = 0, weight exactly 1: the centre does nothing to this survivor.
                // The test here only has to be a superset of "(hi - lo) > -1 in both dimensions" (passes B1 / B2
                // decide exactly), so it compares against the centre grown by 2 px: four compares per pair.
                for (int q = 0; q < m; ++q) {
                    const float4 e = sm.batch_ebox[q];
                    if ((bs.z > e.x && bs.x < e.z && bs.w > e.y && bs.y < e.w) || all_maybe) mask |= 1u << q;
                }
                if (mask == 0u && queued) t2k.add(make_key(u, s));
            }
            const int c = __popc(mask);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total != 0) {
                if (cnt + total <= seg_cap) {
                    int pos = cnt + incl - c;
                    bool head = true;
                    for (uint32_t rem = mask; rem; rem &= rem - 1) {
                        seg[pos++] = ent_make(s, __ffs(rem) - 1, head, queued, c);
                        head = false;
                    }
                    cnt += total;
                } else if (mask) {                                             // segment full: handle in place
                    const float u = k3_process_inplace(&kc, sm, r, m, mask, s, queued);
                    if (u > -INFINITY) t2k.add(make_key(u, s));
                }
            }
        }
        __syncwarp();
        if (a.dbg && tid == 0) t2c = clock64();

        // Every survivor is scanned by the same thread in every round (s = tid + k * kK3Threads), so a warp's
        // segment only ever lists survivors whose state this warp owns: passes B1 / B2 run per warp on the
        // warp's own segment and need no block barrier.
        // ---- pass B1: one listed pair per lane: soft-NMS weight + membership bit ----
        float* wseg = wl + warp * kSegCap;
        for (int e = lane; e < cnt; e += 32) {
            const uint32_t ent = seg[e];
            const int s = ent_s(ent), q = ent_q(ent);
            const float4 bs = corn[s], bx = sm.sel_box[r + q];
            wseg[e] = ent_queued(ent) ? nms_weight_fast(tf_iou(bs, bx), C.scale, C.is_soft, C.thr, sm.exp_tab) : 1.0f;
            if (is_member(bs, bx, C.thr)) atomicOr(&C.member[(size_t)(r + q) * C.words + (s >> 5)], 1u << (s & 31));
        }
        __syncwarp();
        // ---- pass B2: one listed candidate per lane (its first pair): the epoch walk ----
        for (int e = lane; e < cnt; e += 32) {
            const uint32_t ent = seg[e];
            if (ent_head(ent) && ent_queued(ent)) {
                const int s = ent_s(ent), c = ent_pairs(ent);
                K3Walk wk;
                wk.begin(C, s);
                for (int j = 0; j < c; ++j) wk.step(C, sm, r, s, ent_q(seg[e + j]), wseg[e + j]);
                const float u = wk.finish(C, sm, r, m, s);
                if (u > -INFINITY) t2k.add(make_key(u, s));
            }
        }
        __syncwarp();
        if (a.dbg && tid == 0) { t3 = clock64(); tC += t1 - tx2; tM += tm - t0; tW += tw - tm; tR += tx1 - tw; tP += tx2 - tx1; tA += t2c - t1; tB += t3 - t2c; tL += cnt; }
        r += m;
        ++rounds;
        // no barrier here: the warp lists of the next round go to sm.top_w, which warp 0 finished reading
        // before the batch barrier; seg_n / list / wl are next written after the top-of-round barrier
    }
    if (a.dbg && tid == 0) {
        a.dbg[b * 8 + 0] = tA; a.dbg[b * 8 + 1] = tB; a.dbg[b * 8 + 2] = tR; a.dbg[b * 8 + 3] = rounds; a.dbg[b * 8 + 4] = tP; (void)tL;
        a.dbg[b * 8 + 5] = tC; a.dbg[b * 8 + 6] = tM; a.dbg[b * 8 + 7] = tW;
    }
    if (tid == 0) a.num_dets[b] = r;
    __syncthreads();                                         // sel_key of the last round
    for (int d = tid; d < r; d += kK3Threads)               // off the rounds' critical path: one gather at the end
        a.centre_anchor[(size_t)b * Dmax + d] = a.surv_anchor[(size_t)b * a.capacity + key_index(sm.sel_key[d])];
    for (int d = r + tid; d < Dmax; d += kK3Threads) {       // padding rows
        a.nms_idx[(size_t)b * Dmax + d] = -1;
        a.centre_anchor[(size_t)b * Dmax + d] = -1;
        a.nms_score[(size_t)b * Dmax + d] = 0.0f;
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
        double tab[kExpTab];
        for (int k = 0; k < kExpTab; ++k) tab[k] = exp(-(double)k / 64.0);
        cudaError_t e0 = cudaMemcpyToSymbol(c_exp_tab, tab, sizeof tab);
        if (e0 != cudaSuccess) return e0;
        g_exp_tab_ready[dev] = true;
    }
    const int smem_S = a.fastS;
    // dynamic shared memory: pair lists + the per-candidate pool (corners, scores, pending weights);
    // as much as one CTA can have next to the static part, so small images keep long pending lists on chip
    const size_t lists = (size_t)kK3Warps * kSegCap * 8;
    const size_t most = (227 * 1024 - sizeof(K3Smem) - 1024 - lists) & ~(size_t)15;
    static_assert((size_t)kFastS * 25 <= 227 * 1024 - sizeof(K3Smem) - 1024 - (size_t)kK3Warps * kSegCap * 8 - 16,
                  "kFastS candidates must fit the shared-memory pool");
    const size_t smem = lists + most;
    cudaError_t e = cudaFuncSetAttribute(k3_softnms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k3_softnms_kernel<<<a.B, kK3Threads, smem, st>>>(a, smem_S, (int)most);
    return cudaGetLastError();
}

}  // namespace bod
