// k3_softnms.cu — stage K3: exact emulation of TF's soft-NMS centre selection.
// Compiled with -fmad=false.
//
// Reference lines replaced:
//   inference_utils.py:207-212  tf.image.non_max_suppression_with_scores(boxes, scores,
//                               max_output_size, iou_threshold, soft_nms_sigma)
//                               = TF's NonMaxSuppressionV5 CPU kernel (a device->host->device
//                               round trip in the reference graph)
// (The cluster-membership test of :214-215 / :316 lives in K4, one warp per centre.)
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
// boxes that do not overlap a new centre need no arithmetic at all.  On the bench
// scenes TF performs 2 000 - 3 000 pops and ~140 000 IoU evaluations per image; the
// formulation below needs ~14 rounds of (one lean overlap test per candidate and
// new centre) + (exact arithmetic for the few hundred pairs that do overlap).
//
// One CTA per image (images are independent); B CTAs run concurrently.  Chosen per
// image by its survivor count S:
//   the per-candidate state fits the shared-memory pool: everything a round touches is on chip;
//   S <= max_rows (65 535): the same code with the per-candidate arrays in (L2-resident) global scratch;
//   beyond: k3_generic, the literal one-round-per-selection formulation.
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

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
        for (int s = tid; s < S; s += blockDim.x) {
            const float4 bs = corners[s];
            float u = cur[s];
            const bool in_queue = (u > -INFINITY) && (s != x);
            if (s == x) cur[s] = -INFINITY;
            if (!in_queue) continue;
            const float xI1 = fmaxf(bs.y, bx.y), yI1 = fmaxf(bs.x, bx.x);
            const float xI2 = fminf(bs.w, bx.w), yI2 = fminf(bs.z, bx.z);
            const bool wellformed = (bs.x <= bs.z) && (bs.y <= bs.w) && (bx.x <= bx.z) && (bx.y <= bx.w);
            const bool maybe = !wellformed || (((xI2 - xI1) + 1.0f > 0.0f) && ((yI2 - yI1) + 1.0f > 0.0f));
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
    if (tid == 0) a.num_dets[b] = r;
    for (int d = r + tid; d < Dmax; d += blockDim.x) {       // padding rows
        a.nms_idx[(size_t)b * Dmax + d] = -1;
        a.centre_anchor[(size_t)b * Dmax + d] = -1;
        a.nms_score[(size_t)b * Dmax + d] = 0.0f;
    }
}

// ---------------------------------------------------------------------------
// round kernel.
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
// (3) The block-wide top list of the NEXT round is folded into the same passes:
// every thread tracks the best two keys it has seen (untouched candidates in
// pass A, updated ones in pass B); warps merge by popping heads (REDUX) into a
// list of kTop1 keys each.  A list is cut where a thread runs out of tracked
// keys (its third best is unknown), and the acceptance walk stops at the largest
// such cut -- fewer centres in that round, never a wrong one.
// (4) Work layout.  Candidate s belongs to warp s % W, lane (s / W) % 32 (W warps
// per CTA): neighbouring survivor indices -- which is what the top of a tie group
// looks like, ties go to the lower index -- land in different warps, so the
// per-warp lists rarely cut the block-wide top list.  A round is:
//   acceptance (four warps behind a named barrier: rank the listed keys by
//   counting, one pair of examined candidates per thread, then warp 0 walks them)
//   | pass A: every thread tests its own queued candidates against the batch
//   centres grown by 2 px (a superset filter: four compares per pair, centres
//   in the outer loop, the thread's candidates in registers), candidates that
//   may overlap are listed (one entry per candidate: who, and a bit per centre)
//   in the warp's own segment
//   | pass B: one listed candidate per lane: IoU + exp for its pairs, then the
//   epoch walk.  A candidate's state is only ever touched by its own warp, so
//   pass B needs no block barrier: two block barriers per round.
// The first psm pending weights of every candidate live in shared memory (psm is
// what fits for the image's survivor count), the rest spill to global rows.
// ---------------------------------------------------------------------------
constexpr int kTop = 32;              // candidates examined per round (one per lane of the acceptance warp)
constexpr int kBatch = 31;            // centres selected per round at most (a bit per centre in a 32-bit mask + the wake bit)
#ifndef BOD_K3_TAU
#define BOD_K3_TAU 0.97f
#endif
constexpr float kTauFrac = BOD_K3_TAU;     // scores below this fraction of the lowest examined score are kept as upper bounds
constexpr int kStateBytes = 25;       // shared memory per candidate besides its pending weights: corners, two scores, a count
constexpr int kSimOld = 32;           // pending weights of an examined candidate the batch loop can multiply in itself
constexpr int kListed = 128;          // keys listed per round by all warps together
constexpr int kExpTab = 129;
constexpr int kPairsCta = 4096;       // (candidate, centre) pairs the warps' list segments hold together

struct K3Smem {
    unsigned long long warp_best[2][32];         // generic kernel scratch
    unsigned long long top_flat[kListed];        // per-warp top keys (descending, 0 = none): warp w at [w * kTop1, (w+1) * kTop1)
    unsigned long long bound_w[32];              // every key of the warp that is not listed is below this (0: there is none)
#ifdef BOD_DIAGNOSTICS
    int bound_kind[32];                          // what bound_w is: 0 the warp's last listed key, 1 a cut (a thread ran out of tracked keys), 2 a lazy bound
#endif
    unsigned long long sel_key[kMaxOut];         // key (score, -index) of every selected centre
    float4 sel_box[kMaxOut];
    unsigned long long cand_key[kTop];           // the round's examined candidates ...
    float4 cand_box[kTop];
    float4 batch_ebox[kBatch + 1];               // the round's centres grown by 2 px: pass A's lean overlap test
    float wpair[kTop][kTop + 1];                 // ... their pairwise soft-NMS weights where they are not exactly 1 (both triangles)
    uint32_t rowmask[kTop];                      // bit i of row q: the weight of the pair (q, i) is not 1
    double exp_tab[kExpTab];                     // exp(-k/64)
    unsigned long long tau[2];                   // the round's laziness threshold (see k3_walk), by round parity
    int batch_n;                                 // centres selected in this round
    int malformed;
    unsigned long long next_key;                 // best listed key that is not examined (0: none)
    int rank_cnt[kListed];                       // listed keys larger than listed key t
    float oldw[kTop][kSimOld + 1];               // the batch loop's copy of an examined candidate's pending weights
    float4 cand_ebox[kTop];                      // the examined candidates' boxes grown by 2 px (the speculative overlap test)
};

struct K3State {                                  // kernel-lifetime constants (registers)
    const float4* corn; float* ucur; float* stl; uint8_t* npend;
    float* pws; int psm;                          // pending weights in shared memory: entry i of state row si at pws[si*psm+i], i < psm
    float* pwg; int pstride;                      // spill rows in global memory: entry i >= psm of survivor s at pwg[s*pstride+i]
    float scale, thr; bool is_soft;
};

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

// best two keys a thread has seen in a round
struct Top2 {
    unsigned long long a = 0ull, b = 0ull;
    BOD_DEVINL void add(unsigned long long k) {
        if (k > a) { b = a; a = k; } else if (k > b) { b = k; }
    }
};

BOD_DEVINL void pend_put(const K3State& C, int si, int s, int i, float w) {
    if (i < C.psm) C.pws[si * C.psm + i] = w; else C.pwg[(size_t)s * C.pstride + i] = w;
}
// entries 4b .. 4b+3 of a pending list (psm and pstride are multiples of 4: a block never straddles)
BOD_DEVINL float4 pend_block(const K3State& C, int si, int s, int b) {
    return (4 * b < C.psm) ? *reinterpret_cast<const float4*>(C.pws + si * C.psm + 4 * b)
                           : *reinterpret_cast<const float4*>(C.pwg + (size_t)s * C.pstride + 4 * b);
}
// v * (entries n-1 .. 0 of the pending list), in that order: four weights per load, the next block in
// flight while the current one is multiplied in
BOD_DEVINL float pend_product(const K3State& C, int si, int s, int n, float v) {
    if (n <= 0) return v;
    int blk = (n - 1) >> 2;
    float4 cur = pend_block(C, si, s, blk);
    const int top = (n - 1) & 3;
    {
        const float4 nxt = pend_block(C, si, s, blk > 0 ? blk - 1 : 0);
        if (top >= 3) v = v * cur.w;
        if (top >= 2) v = v * cur.z;
        if (top >= 1) v = v * cur.y;
        v = v * cur.x;
        cur = nxt;
    }
#pragma unroll 1
    for (--blk; blk >= 0; --blk) {
        const float4 nxt = pend_block(C, si, s, blk > 0 ? blk - 1 : 0);
        v = v * cur.w; v = v * cur.z; v = v * cur.y; v = v * cur.x;
        cur = nxt;
    }
    return v;
}

// list entry of one (candidate, centre) pair: row of the block | owner lane << 3 | centre of the batch << 8 |
// first pair of its candidate << 13 | pairs of the candidate << 14
BOD_DEVINL uint32_t ent_make(int j, int ln, int q, bool head, int c) {
    return (uint32_t)j | ((uint32_t)ln << 3) | ((uint32_t)q << 8) | ((uint32_t)head << 13) | ((uint32_t)c << 14);
}
BOD_DEVINL int ent_row(uint32_t e) { return (int)(e & 7u); }
BOD_DEVINL int ent_lane(uint32_t e) { return (int)((e >> 3) & 31u); }
BOD_DEVINL int ent_q(uint32_t e) { return (int)((e >> 8) & 31u); }
BOD_DEVINL int ent_pairs(uint32_t e) { return (int)((e >> 14) & 63u); }

// Epoch walk of one QUEUED candidate (survivor s, state row si) over a batch of selections whose keys are
// sel_key[0 .. qlast] (the caller passes the batch's slice of sm.sel_key), driven by the candidate's listed
// pairs in ascending centre order: ents[i] names the centre, ws[i] is the exact soft-NMS weight, i < c.
// TF pops s right before selection j iff it has pending weights and key(stale) > key(selection j), and then
// folds every pending weight into the stale score, newest first.  Between two of the candidate's pairs the
// pending list does not change and the selection keys decrease, so such a pop exists in (q_prev, q] iff
// key(stale) > key(selection q), and wherever it happens it folds the same list; pops hidden behind centres
// the candidate does not overlap (weight exactly 1) are caught by the next comparison the same way.
// The pending list is (old entries, in memory) ++ (this batch's non-unit weights, in ws): a product over it,
// newest first, walks ws backwards (unit weights multiply exactly) and then the old entries.
//
// Lazy scores.  A candidate far below the top of the queue needs no score, only a bound that keeps it out of
// the examined candidates: TF itself never looks at it until the queue's front reaches it.  When no pop can be
// due in this batch (key(stale) <= key(last selection); the selection keys decrease) the walk would only append
// the new weights, so that is all that is done, and the score is replaced by an UPPER BOUND (previous score or
// bound x new weights x a rounding allowance), stored negated.  Bounds at or above `tau` (a key well below the
// examined ones) are not allowed: such a candidate takes the exact path -- which never reads the old score --
// and owners list their bounded candidates for it (with no centre at all) once tau has come down to them.
// The pending list, the stale score and the epoch logic are the same on both paths; only how often the
// O(list) product is evaluated differs (without this, candidates TF never pops accumulate lists of dozens of
// weights that are re-multiplied on every touch).
// One loop with one product site (steps 0..c-1: the pairs, c: pops behind the batch's last selections,
// c+1: the up-to-date score): the kernel's round loop has to stay small for the instruction cache.
// Returns the new score (>= 0), the negated bound, or -inf (removed by hard-NMS).
#ifdef BOD_DIAGNOSTICS
__device__ unsigned long long g_k3_cnt[8];     // 0 walks, 1 bounded (lazy) walks, 2 untouched, 3 products, 4 product entries, 5 folds, 6 woken
#endif
#if defined(BOD_DIAGNOSTICS) && (BOD_DIAGNOSTICS + 0) < 2   // timers-only builds: why a round's batch ended, in the otherwise unused slots
#define K3_GB_SLOT (gb_is_next ? 0 : 4 + gb_kind)
#define K3_WHY(i) do { if (lane == 0) atomicAdd(&g_k3_cnt[i], 1ull); } while (0)   // (one atomic per round): [0] a candidate outside the examined
#else                                                        // 32 may come first, [1] pending list too long for the loop, [2] batch / output full, [3] rounds
#define K3_GB_SLOT 0
#define K3_WHY(i)
#endif
#if defined(BOD_DIAGNOSTICS) && (BOD_DIAGNOSTICS + 0) >= 2   // event counters (global atomics: they cost a third of the kernel's time);
#define K3_CNT(i, v) atomicAdd(&g_k3_cnt[i], (unsigned long long)(v))   // -DBOD_DIAGNOSTICS alone keeps the phase timers only
#else
#define K3_CNT(i, v)
#endif
BOD_DEVINL float k3_walk(const K3State& C, const unsigned long long* sel_key, int qlast, unsigned long long tau,
                         int si, int s, const uint32_t* ents, const float* ws, int c) {
    K3_CNT(0, 1);
    float st = C.stl[si];
    int n_old = C.npend[si];              // old entries still pending (0 once folded)
    unsigned long long ks = make_key(st, s);
    const float uprev = C.ucur[si];
    if (!(uprev > -INFINITY)) return -INFINITY;    // removed by an earlier piece of this batch (hard-NMS, piecewise pass B)
    const bool was_exact = __float_as_int(uprev) >= 0;
    if (qlast >= 0 && !(ks > sel_key[qlast])) {
        float ub = fabsf(uprev);
        if (was_exact) ub = ub * 1.00005f;                          // the same factors in another order: <= 2n+2 roundings apart, n <= 255
        int nn = 0;
        bool dead = false;
#pragma unroll 1
        for (int i = 0; i < c; ++i) {
            const float w = ws[i];
            if (w != 1.0f) { ub = ub * w * 1.0000005f; ++nn; dead = dead || (!C.is_soft && w == 0.0f); }
        }
        if (dead) { C.ucur[si] = -INFINITY; return -INFINITY; }     // hard-NMS: removed for good
        if (nn == 0 && was_exact) return uprev;   // every weight was exactly 1: untouched
        if (make_key(ub, s) < tau) {
            K3_CNT(1, 1);
            int base = n_old;
#pragma unroll 1
            for (int i = 0; i < c; ++i) { const float w = ws[i]; if (w != 1.0f) pend_put(C, si, s, base++, w); }
            C.npend[si] = (uint8_t)base;
            const float nb = __int_as_float(__float_as_int(ub) | (int)0x80000000);
            C.ucur[si] = nb;
            return nb;
        }
    }
    if (!was_exact) K3_CNT(6, 1);
    bool folded = false;
    int nf = 0, nn = 0;                   // ws[nf..) belong to the pending list; nn of them are not 1
    int last_q = -1;
    float u = 0.0f;
#pragma unroll 1
    for (int i = 0; i <= c + 1; ++i) {
        bool want;                        // the product over the pending list is needed at this step
        float w = 1.0f;
        int q = qlast;
        if (i < c) {
            w = ws[i];
            if (w == 1.0f) continue;                                 // the centre does nothing to this candidate
            q = ent_q(ents[i]);
            want = (n_old + nn) > 0 && ks > sel_key[q];              // popped (folded) before selection q
        } else if (i == c) {
            want = last_q < qlast && (n_old + nn) > 0 && ks > sel_key[qlast];   // popped before a later selection of the batch
                                                                                 // (qlast = -1: no selections, nothing to check)
        } else {
            if (!folded && nn == 0 && was_exact) return uprev;       // every weight was exactly 1: untouched
            want = true;
        }
        if (want) {
            const int upto = i < c ? i : c;
            float v = st;
#pragma unroll 1
            for (int k = upto - 1; k >= nf; --k) v = v * ws[k];
            v = pend_product(C, si, s, n_old, v);
            K3_CNT(3, 1); K3_CNT(4, n_old + (upto - nf)); if (i <= c) K3_CNT(5, 1);
            if (i <= c) { st = v; n_old = 0; nf = upto; nn = 0; folded = true; ks = make_key(st, s); }
            else u = v;
        }
        if (i < c) {
            last_q = q;
            if (!C.is_soft && w == 0.0f) { C.ucur[si] = -INFINITY; return -INFINITY; }   // hard-NMS: removed for good
            ++nn;
        }
    }
    int base = n_old;
#pragma unroll 1
    for (int k = nf; k < c; ++k) { const float w = ws[k]; if (w != 1.0f) pend_put(C, si, s, base++, w); }
    C.npend[si] = (uint8_t)base;
    C.ucur[si] = u;
    if (folded) C.stl[si] = st;
    return u;
}

// Warp-level merge of the lanes' Top2 pairs: every lane returns with the warp's best keys in out[]
// (descending, 0 = none), cut after the first key whose lane has nothing tracked behind it; `bound`
// is that key (0 when the list was not cut and the warp has no further keys).
template <int kTop1>
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

#ifdef BOD_DIAGNOSTICS
#define K3_T(var) long long var = clock64()
#define K3_ACC(slot, t1, t0) if (lane == 0) dbg_acc[slot] += (t1) - (t0)
#else
#define K3_T(var)
#define K3_ACC(slot, t1, t0)
#endif

// The rounds of one image.  BIG: per-candidate state in the workspace's global rows (more survivors than the
// shared-memory pool holds); otherwise every state pointer is derived from the dynamic shared array, so the
// compiler emits LDS / STS instead of generic accesses.
template <int NT, bool BIG, int kCH>
BOD_DEVINL void k3_rounds(const K3Args& a, const int pool_bytes, K3Smem& sm, const int S) {
    constexpr int W = NT / 32;                              // warps
    constexpr int kTop1 = kListed / W;                      // keys every warp lists per round
    constexpr int kSegCap = kPairsCta / W;                  // pairs per warp segment
    static_assert(kTop1 >= 2 && kTop1 * W == kListed && kSegCap >= 32, "CTA size");
    extern __shared__ __align__(16) unsigned char dyn[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // Candidate rows: thread `tid` owns survivor s = k*NT + lane*W + warp of row k; its state row (shared
    // memory) is si = k*NT + tid.  More survivors than the pool holds: state in the workspace's global
    // rows, indexed by s, corners read in place.
    const int nrows = (S + NT - 1) / NT;
    const int SP = nrows * NT;
    uint32_t* list = reinterpret_cast<uint32_t*>(dyn);                    // [W][kSegCap] pair entries (ent_make)
    float* wl = reinterpret_cast<float*>(list + kPairsCta);               // [W][kSegCap] their soft-NMS weights
    uint16_t* hl = reinterpret_cast<uint16_t*>(wl + kPairsCta);           // [W][kSegCap] first pair of every listed candidate
    constexpr bool big = BIG;
    float4* corn = reinterpret_cast<float4*>(hl + kPairsCta);             // [SP] corners
    float* ucur = reinterpret_cast<float*>(corn + SP);                    // [SP] up-to-date score, -inf = not queued
    float* stl = ucur + SP;                                               // [SP] score as of the last fold
    // pending weights per candidate held in shared memory: whatever fits behind the fixed arrays
    int psm = (SP > 0 && !big) ? (pool_bytes - SP * kStateBytes) / (SP * 4) : 0;
    psm = psm < 0 ? 0 : (psm > a.pstride ? a.pstride : psm);
    if (a.psm_max >= 0 && psm > a.psm_max) psm = a.psm_max;              // tests: force the global spill rows
    psm &= ~3;                                                            // rows are read four weights at a time
    float* pws = stl + SP;                                                // [SP][psm]
    uint8_t* npend = reinterpret_cast<uint8_t*>(pws + (size_t)psm * SP);  // [SP] pending entries per candidate
    if constexpr (BIG) {
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
    C.pws = pws; C.psm = psm;
    C.pwg = a.pw + (size_t)b * a.pw_rows * a.pstride; C.pstride = a.pstride;
    C.is_soft = a.soft_nms_sigma > 0.0f;
    C.scale = C.is_soft ? -0.5f / a.soft_nms_sigma : 0.0f;
    C.thr = a.iou_threshold;
    // state row of survivor s (used where a key names the candidate)
    auto row_of = [&](int s) -> int {
        if (big) return s;
        const int r = s & (NT - 1);
        return (s - r) + (r % W) * 32 + r / W;
    };

    if (tid == 0) sm.malformed = 0;
    for (int k = tid; k < kExpTab; k += NT) sm.exp_tab[k] = c_exp_tab[k];
    __syncthreads();

    // ---- load; the first round's top keys come from the initial scores ----
    Top2 t2k;
    unsigned long long ixmax = 0ull;                        // largest key(upper bound) this thread has seen in the round
    for (int k = 0; k < nrows; ++k) {
        const int s = k * NT + lane * W + warp;
        if (s >= S) continue;
        const int si = big ? s : k * NT + tid;
        const float4 c = corners[s];
        if (!big) corn[si] = c;
        // corners out of order need the canonicalising IoU path for every pair; so do coordinates so large
        // that the 2 px margin of pass A's lean test is not safely above their rounding
        if (!((c.x <= c.z) && (c.y <= c.w)) || !(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))) < 1.0e5f))
            sm.malformed = 1;
        const float sc = score[s];
        const bool queued = sc > -INFINITY;                              // scores_data[i] > score_threshold (-inf); NaN stays out
        ucur[si] = queued ? sc : -INFINITY;
        stl[si] = sc;
        npend[si] = 0;
        if (queued) t2k.add(make_key(sc, s));
    }

#ifdef BOD_DIAGNOSTICS
    // per warp (lane 0): 0 merge, 1 wait at barrier 1, 2 rank, 3 pairwise, 4 walk, 5 wait at barrier 2, 6 pass A, 7 pass B,
    // 8 listed candidates, 9 spilled weights read, 10 rounds, 11 psm
    long long dbg_acc[12] = {0};
#endif
    int r = 0, rnd = 0;
    while (r < Dmax) {
        K3_T(c0);
        // ---- warp lists of this round's best keys ----
        {
            unsigned long long out[kTop1], bound;
            warp_top_merge2<kTop1>(t2k, out, bound);
            if (lane < kTop1) {
                unsigned long long v = out[0];
#pragma unroll
                for (int q = 1; q < kTop1; ++q) v = (lane == q) ? out[q] : v;
                sm.top_flat[warp * kTop1 + lane] = v;
            }
            // a candidate whose score is only an upper bound is never examined, and nothing below its bound is
            const unsigned long long ixw = warp_max_u64(ixmax);
            if (lane == 0) sm.bound_w[warp] = bound > ixw ? bound : ixw;
#ifdef BOD_DIAGNOSTICS
            if (lane == 0) sm.bound_kind[warp] = ixw > bound ? 2 : (bound == out[kTop1 - 1] ? 0 : 1);
#endif
        }
        if (tid < kTop) { sm.cand_key[tid] = 0ull; sm.rowmask[tid] = 0u; }     // filled by the acceptance phases below
        if (tid < kListed) sm.rank_cnt[tid] = 0;
        if (tid == 0) { sm.tau[rnd & 1] = 0ull; sm.next_key = 0ull; }
        K3_T(c1);
        __syncthreads();
        K3_T(c2);
        K3_ACC(0, c1, c0); K3_ACC(1, c2, c1);
        // ---- acceptance (every warp helps with the ranking and the pairwise table; warp 0 then runs the batch) ----
        // (i) the block's top-kTop keys: listed key t is ranked by counting the larger ones, NT / kListed threads
        // per key, each over its share of the list.  G = the largest cut of any warp's list: below it an
        // untracked key might outrank a listed one, so such entries are dropped (validity is a prefix of the order)
        {
            constexpr int kParts = NT / kListed;                         // threads per listed key (warp-uniform part)
            constexpr int kShare = kListed / kParts;                     // keys each of them compares against
            static_assert(kParts >= 1 && kParts * kListed == NT && kShare % 2 == 0, "ranking layout");
            const int t = tid & (kListed - 1), part = tid / kListed;
            const unsigned long long G = warp_max_u64((lane < W) ? sm.bound_w[lane] : 0ull);
            const unsigned long long key = sm.top_flat[t];
            if (key != 0ull && key >= G) {
                const ulonglong2* f2 = reinterpret_cast<const ulonglong2*>(sm.top_flat + part * kShare);
                int cnt = 0;
#pragma unroll 4
                for (int j = 0; j < kShare / 2; ++j) {
                    const ulonglong2 kk = f2[j];
                    cnt += (int)(kk.x > key) + (int)(kk.y > key);
                }
                if (kParts > 1) { if (cnt) atomicAdd(&sm.rank_cnt[t], cnt); } else sm.rank_cnt[t] = cnt;
            }
            __syncthreads();
            if (part == 0 && key != 0ull && key >= G) {
                const int rank = sm.rank_cnt[t];
                if (rank < kTop) {
                    const float4 cb = corn[row_of(key_index(key))];
                    sm.cand_key[rank] = key; sm.cand_box[rank] = cb;
                    sm.cand_ebox[rank] = make_float4(cb.x - 2.0f, cb.y - 2.0f, cb.z + 2.0f, cb.w + 2.0f);
                } else if (rank == kTop) sm.next_key = key;               // the best candidate that is not examined
            }
            __syncthreads();
        }
        K3_T(c3);
        K3_ACC(2, c3, c2);
        // (ii) pairwise weights among the examined candidates, one pair (q, i), i < q, per thread.  Most pairs do
        // not intersect at all (weight exactly 1): the division + exp runs only for those that do
        for (int p = tid; p < kTop * (kTop - 1) / 2; p += NT) {
            int q = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
            while (q * (q - 1) / 2 > p) --q;
            while ((q + 1) * q / 2 <= p) ++q;
            const int i = p - q * (q - 1) / 2;
            if (sm.cand_key[q] != 0ull) {                                 // then candidate i < q exists too
                const float4 bq = sm.cand_box[q], bi = sm.cand_box[i];
                const float dx = fminf(bq.w, bi.w) - fmaxf(bq.y, bi.y);
                const float dy = fminf(bq.z, bi.z) - fmaxf(bq.x, bi.x);
                if ((dx > 0.0f && dy > 0.0f) || sm.malformed != 0) {
                    const float w = nms_weight_fast(tf_iou(bq, bi), C.scale, C.is_soft, C.thr, sm.exp_tab);
                    if (w != 1.0f) {
                        sm.wpair[q][i] = w; sm.wpair[i][q] = w;
                        atomicOr(&sm.rowmask[q], 1u << i); atomicOr(&sm.rowmask[i], 1u << q);
                    }
                }
            }
        }
        __syncthreads();
        K3_T(c4);
        K3_ACC(3, c4, c3);
        K3_T(c5);
        if (warp == 0) {
            // (iii) the batch: TF's loop restricted to the examined candidates, one SELECTION per iteration, every lane
            // (= examined candidate) in parallel.  A lane keeps TF's queue entry of its candidate: the stale score st, the
            // weights pending since its last pop (earlier rounds' in sm.oldw, this batch's in registers, newest first),
            // and u = st x pending weights, newest first: the score TF computes when it pops the candidate now.  The
            // next centre is the lane with the largest key(u): TF pops, updates and re-pushes candidates until the
            // front of the queue is a candidate whose score does not change, and that is the largest current score.  On
            // the way it pops exactly the candidates with pending weights whose stale key exceeds the new centre's key
            // (k3_walk's rule): those lanes fold (st = u, nothing pending).  Then the new centre's weight becomes
            // pending for the lanes it overlaps.  Every candidate outside the examined set has a current score below Gb
            // (the listed keys' cut, the bounds of lazily scored candidates, the best listed key that was not
            // examined) and scores only fall, so while the winner's key is not below Gb it is TF's next selection, bit
            // for bit; the batch ends where a candidate outside might come first.  Nothing is stored for examined
            // candidates that are not selected: the passes below walk them like any other.
            static_assert(kTop == 32, "one examined candidate per lane");
            const unsigned long long myk = sm.cand_key[lane];
            const int nvalid = __popc(__ballot_sync(0xffffffffu, myk != 0ull));      // candidates are a prefix
            unsigned long long Gb = warp_max_u64((lane < W) ? sm.bound_w[lane] : 0ull);
            const bool any_bound = Gb != 0ull;
#if defined(BOD_DIAGNOSTICS) && (BOD_DIAGNOSTICS + 0) < 2
            const bool gb_is_next = sm.next_key > Gb;         // the bound is the best listed key that was not examined; else slot 4 + kind of the binding warp bound
            int gb_kind = 0;
            { const unsigned bw = __ballot_sync(0xffffffffu, lane < W && sm.bound_w[lane] == Gb && Gb != 0ull);
              if (bw) gb_kind = sm.bound_kind[__ffs(bw) - 1]; }
#endif
            { const unsigned long long nk = sm.next_key; Gb = nk > Gb ? nk : Gb; }
            const int x = key_index(myk);
            const int si = (myk != 0ull) ? row_of(x) : 0;
            float st = 0.0f;
            int nold = 0;
            if (myk != 0ull) { st = stl[si]; nold = npend[si]; }
            const uint32_t nz = sm.rowmask[lane];                        // examined candidates whose weight against this one is not 1
            bool old = nold > 0;                                         // weights of earlier rounds are pending
            const bool long_old = nold > kSimOld;                        // ... more than this loop can multiply in itself
            if (old && !long_old) {
#pragma unroll 4
                for (int i = 0; i < nold; ++i)
                    sm.oldw[lane][i] = (i < C.psm) ? C.pws[si * C.psm + i] : C.pwg[(size_t)x * C.pstride + i];
            }
            float u = key_score(myk);                                    // examined candidates have exact scores
            uint32_t hu = (myk != 0ull) ? float_key(u) : 0u;             // key of u, high word (never 0 for a live lane)
            uint32_t hs = float_key(st);                                 // key of st, high word
            const uint32_t lo = (uint32_t)myk;                           // low word of both: ~index
            float q0 = 1.0f, q1 = 1.0f, q2 = 1.0f, q3 = 1.0f;            // this batch's pending weights, q0 the newest
            int nq = 0;
            const uint32_t Gbh = (uint32_t)(Gb >> 32), Gbl = (uint32_t)Gb;
            int m = 0;
            float lowest = INFINITY;                                     // lowest score selected
#pragma unroll 1
            while (m < kBatch && r + m < Dmax) {
                const uint32_t mh = __reduce_max_sync(0xffffffffu, hu);
                if (mh == 0u || mh < Gbh) { K3_WHY(K3_GB_SLOT); break; }
                uint32_t bal = __ballot_sync(0xffffffffu, hu == mh);
                if (bal & (bal - 1u)) {                                  // equal scores: the lower index comes first
                    const uint32_t ml = __reduce_max_sync(0xffffffffu, hu == mh ? lo : 0u);
                    bal = __ballot_sync(0xffffffffu, hu == mh && lo == ml);
                }
                const int L = __ffs(bal) - 1;
                const uint32_t ll = __shfl_sync(0xffffffffu, lo, L);
                if (mh == Gbh && ll < Gbl) { K3_WHY(K3_GB_SLOT); break; }
                K3_CNT(2, lane == 0);
                bool stop = false;
                if (lane == L) {                                         // selection r + m
                    const int pos = r + m;
                    sm.sel_box[pos] = sm.cand_box[lane];
                    sm.batch_ebox[m] = sm.cand_ebox[lane];
                    sm.sel_key[pos] = ((unsigned long long)mh << 32) | lo;
                    ucur[si] = -INFINITY;                                // leaves the queue
                    a.nms_idx[(size_t)b * Dmax + pos] = x;
                    a.nms_score[(size_t)b * Dmax + pos] = u;
                    hu = 0u;
                } else if (hu != 0u) {
                    // popped before this selection: pending weights and a stale key above the new centre's
                    if ((nq > 0 || old) && (hs > mh || (hs == mh && lo > ll))) { st = u; hs = hu; nq = 0; old = false; }
                    if ((nz >> L) & 1u) {                                // the new centre's weight is pending now
                        const float w = sm.wpair[lane][L];
                        if (!C.is_soft && w == 0.0f) hu = 0u;            // hard-NMS: removed for good (the passes below see it too)
                        else if (nq == 4 || (old && long_old)) stop = true;   // more than this loop keeps: the batch ends here
                        else {
                            q3 = q2; q2 = q1; q1 = q0; q0 = w; ++nq;
                            float v = st * q0;
                            if (nq > 1) v = v * q1;
                            if (nq > 2) v = v * q2;
                            if (nq > 3) v = v * q3;
                            if (old) {
#pragma unroll 4
                                for (int i = nold - 1; i >= 0; --i) v = v * sm.oldw[lane][i];
                            }
                            u = v; hu = float_key(u);
                        }
                    }
                }
                ++m; lowest = fminf(lowest, key_score((unsigned long long)mh << 32));
                if (__any_sync(0xffffffffu, stop)) { K3_WHY(1); break; }
            }
            if (!(m < kBatch && r + m < Dmax)) K3_WHY(2);
            K3_WHY(3);
            if (lane == 0) {
                // laziness threshold of the round's passes: a fixed fraction below the lowest score examined or selected
                if (nvalid > 0) {
                    lowest = fminf(lowest, key_score(sm.cand_key[nvalid - 1]));
                    sm.tau[rnd & 1] = make_key(lowest * kTauFrac, 0x7fffffff);
                }
                // nothing examinable although keys exist (every listed key is below some candidate's upper bound):
                // a refresh round (-1) makes all bounds exact
                sm.batch_n = (m == 0 && any_bound) ? -1 : m;
            }
        }
        K3_T(c6);
        __syncthreads();
        K3_T(c7);
        K3_ACC(4, c6, c5); K3_ACC(5, c7, c6);
        const bool refresh = sm.batch_n < 0;
        const int m = refresh ? 0 : sm.batch_n;
        if (m == 0 && !refresh) break;                                         // queue empty
        const bool all_maybe = sm.malformed != 0;
        const uint32_t full = (1u << m) - 1u;                                  // m <= kBatch = 15
        const uint32_t wake = 1u << m;                                         // pseudo-centre: "make this score exact"
        const unsigned long long tau = refresh ? 0ull : sm.tau[rnd & 1];

        t2k = Top2();
        ixmax = 0ull;
        uint32_t* seg = list + warp * kSegCap;
        float* wseg = wl + warp * kSegCap;
        uint16_t* hseg = hl + warp * kSegCap;
        const int seg_cap = (a.seg_cap >= 32 && a.seg_cap < kSegCap) ? a.seg_cap : kSegCap;   // tests shrink it
        for (int k0 = 0; k0 < nrows; k0 += kCH) {
            // ---- pass A: the thread's queued candidates of this block against the batch centres.
            // No positive intersection => TF's IoU is 0 and the weight exactly 1: the centre does nothing to the
            // candidate.  The test here only has to be a superset of "positive intersection" (pass B decides
            // exactly), so it compares against the centre grown by 2 px: four compares per pair.
            K3_T(a0);
            float4 bx[kCH];
            float uq[kCH];
            uint32_t mk[kCH];
#pragma unroll
            for (int j = 0; j < kCH; ++j) {
                const int k = k0 + j;
                const int s = k * NT + lane * W + warp;
                uq[j] = -INFINITY; mk[j] = 0u;
                bx[j] = make_float4(NAN, NAN, NAN, NAN);                       // compares false against everything
                if (k < nrows && s < S) {
                    const int si = big ? s : k * NT + tid;
                    uq[j] = ucur[si];
                    if (uq[j] > -INFINITY) bx[j] = corn[si];
                }
            }
            for (int q = 0; q < m; ++q) {
                const float4 e = sm.batch_ebox[q];
#pragma unroll
                for (int j = 0; j < kCH; ++j) {
                    const bool hit = (bx[j].z > e.x) & (bx[j].x < e.z) & (bx[j].w > e.y) & (bx[j].y < e.w);
                    mk[j] |= (uint32_t)hit << q;
                }
            }
#pragma unroll
            for (int j = 0; j < kCH; ++j) {
                if (uq[j] > -INFINITY) {
                    if (all_maybe) mk[j] = full;
                    const unsigned long long key = make_key(fabsf(uq[j]), (k0 + j) * NT + lane * W + warp);
                    if (__float_as_int(uq[j]) >= 0) {                          // an exact score
                        if (mk[j] == 0u) t2k.add(key);
                    } else if (key >= tau) mk[j] |= wake;                      // a bound the threshold has come down to
                    else if (mk[j] == 0u) ixmax = key > ixmax ? key : ixmax;
                }
            }
            K3_T(a1);
            K3_ACC(6, a1, a0);

            // ---- pass B on the pairs with a centre in `qm` and a row in [jlo, jhi): list them in the warp's
            // segment (pairs of a candidate adjacent, ascending centre), B1: one pair per lane: IoU + exp;
            // B2: one listed candidate per lane: the epoch walk over its weights.  Returns false (nothing done)
            // when the pairs do not fit the segment.
            auto pass_b = [&](const uint32_t qm, const int jlo, const int jhi, const bool add_keys) -> bool {
                int pc = 0, cc = 0;
#pragma unroll
                for (int j = 0; j < kCH; ++j) {
                    const uint32_t mm = (j >= jlo && j < jhi) ? (mk[j] & qm) : 0u;
                    pc += __popc(mm); cc += (mm != 0u);
                }
                int incl = (pc << 16) | cc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                const int tot = __shfl_sync(0xffffffffu, incl, 31);
                const int npairs = tot >> 16, ncand = tot & 0xffff;
                if (npairs == 0) return true;
                if (npairs > seg_cap) return false;
                {
                    int pos = (incl >> 16) - pc, hpos = (incl & 0xffff) - cc;
#pragma unroll
                    for (int j = 0; j < kCH; ++j) {
                        const uint32_t mm = (j >= jlo && j < jhi) ? (mk[j] & qm) : 0u;
                        if (mm != 0u) {
                            hseg[hpos++] = (uint16_t)pos;
                            const int c = __popc(mm);
                            bool head = true;
                            for (uint32_t rem = mm; rem; rem &= rem - 1) { seg[pos++] = ent_make(j, lane, __ffs(rem) - 1, head, c); head = false; }
                        }
                    }
                }
                __syncwarp();
                for (int e = lane; e < npairs; e += 32) {                      // B1
                    const uint32_t ent = seg[e];
                    const int k = k0 + ent_row(ent), ln = ent_lane(ent);
                    const int si = big ? (k * NT + ln * W + warp) : (k * NT + warp * 32 + ln);
                    const int q = ent_q(ent);
                    wseg[e] = (q < m) ? nms_weight_fast(tf_iou(corn[si], sm.sel_box[r + q]), C.scale, C.is_soft, C.thr, sm.exp_tab)
                                      : 1.0f;                                  // the pseudo-centre of a woken candidate
                }
                __syncwarp();
                const int qlast = 31 - __clz(qm & full);                       // pieces always hold a real centre
                for (int h = lane; h < ncand; h += 32) {                       // B2
                    const int e = hseg[h];
                    const uint32_t ent = seg[e];
                    const int k = k0 + ent_row(ent), ln = ent_lane(ent);
                    const int s = k * NT + ln * W + warp;
                    const int si = big ? s : (k * NT + warp * 32 + ln);
                    const float u = k3_walk(C, sm.sel_key + r, qlast, tau, si, s, seg + e, wseg + e, ent_pairs(ent));
                    if (add_keys && u > -INFINITY) {
                        const unsigned long long key = make_key(fabsf(u), s);
                        if (__float_as_int(u) >= 0) t2k.add(key); else ixmax = key > ixmax ? key : ixmax;
                    }
                }
                __syncwarp();
#ifdef BOD_DIAGNOSTICS
                if (lane == 0) { dbg_acc[8] += ncand; dbg_acc[9] += npairs; }
#endif
                return true;
            };
            // The whole block at once; when its pairs overflow the segment: a batch may be applied in pieces of
            // consecutive centres -- rounds are only a grouping of consecutive selections -- so the centres are
            // halved until a piece fits, and row by row where one centre alone overflows (<= 32 pairs then); the
            // owners pick the final scores up afterwards.
            // (One call site: the round loop has to stay small enough for the instruction cache.)
            {
                bool whole = true, add = true;
                int lo = 0, wd = m, jlo = 0, jhi = kCH;
#pragma unroll 1
                for (;;) {
                    const int hi = (lo + wd < m) ? lo + wd : m;
                    const uint32_t qm = whole ? (full | wake)
                                              : ((((1u << hi) - 1u) & ~((1u << lo) - 1u)) | (m == 0 ? wake : 0u));
                    if (pass_b(qm, jlo, jhi, add)) {
                        if (whole) break;
                        if (jhi < kCH) { jlo = jhi; jhi = jlo + 1; continue; }   // next row of this piece
                        lo = hi; jlo = 0; jhi = kCH;
                        if (lo >= m) break;
                    } else {
                        add = false;
                        K3_CNT(7, 1);
                        if (whole) { whole = false; wd = (m + 1) >> 1; }       // halves
                        else if (wd > 1) wd = (wd + 1) >> 1;
                        else { jlo = 0; jhi = 1; }                             // this centre row by row
                    }
                }
                if (!add) {
#pragma unroll
                    for (int j = 0; j < kCH; ++j) {
                        if (mk[j] != 0u) {
                            const int k = k0 + j;
                            const int s = k * NT + lane * W + warp;
                            const float u = ucur[big ? s : k * NT + tid];
                            if (u > -INFINITY) {
                                const unsigned long long key = make_key(fabsf(u), s);
                                if (__float_as_int(u) >= 0) t2k.add(key); else ixmax = key > ixmax ? key : ixmax;
                            }
                        }
                    }
                }
            }
            K3_T(a2);
            K3_ACC(7, a2, a1);
        }
        r += m;
        ++rnd;
#ifdef BOD_DIAGNOSTICS
        dbg_acc[10] += 1;
#endif
        // no barrier here: the warp lists of the next round go to sm.top_flat, which the team finished
        // reading before the batch barrier; the segments are warp-private
    }
#ifdef BOD_DIAGNOSTICS
    if (a.dbg) {
        dbg_acc[11] = psm;
        if (b == 0 && tid == 0) for (int i = 0; i < 8; ++i) a.dbg[(size_t)gridDim.x * 384 + i] = (long long)g_k3_cnt[i];
        if (lane == 0) for (int i = 0; i < 12; ++i) a.dbg[((size_t)b * 32 + warp) * 12 + i] = dbg_acc[i];
    }
#endif
    if (tid == 0) a.num_dets[b] = r;
    __syncthreads();                                         // sel_key of the last round
    for (int d = tid; d < r; d += NT)                       // off the rounds' critical path: one gather at the end
        a.centre_anchor[(size_t)b * Dmax + d] = a.surv_anchor[(size_t)b * a.capacity + key_index(sm.sel_key[d])];
    for (int d = r + tid; d < Dmax; d += NT) {               // padding rows
        a.nms_idx[(size_t)b * Dmax + d] = -1;
        a.centre_anchor[(size_t)b * Dmax + d] = -1;
        a.nms_score[(size_t)b * Dmax + d] = 0.0f;
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, 1)
k3_softnms_kernel(K3Args a, int pool_bytes) {
    BOD_TIMELINE(a.tl);
    __shared__ K3Smem sm;
    const int b = blockIdx.x;
    const int S = a.num_survivors[b];
    if (S > a.max_rows) { k3_generic(a, b, sm.warp_best, sm.sel_box); return; }
    const long long SP = (long long)((S + NT - 1) / NT) * NT;
    constexpr int kCH = NT >= 512 ? 4 : 8;                  // candidates per thread handled per pass-A/B block (registers, pairs per list segment)
    if (SP * kStateBytes > (long long)pool_bytes || a.force_big != 0) k3_rounds<NT, true, kCH>(a, pool_bytes, sm, S);
    else k3_rounds<NT, false, kCH>(a, pool_bytes, sm, S);
}

static bool g_exp_tab_ready[64] = {false};

template <int NT>
static cudaError_t launch_k3_nt(const K3Args& a, cudaStream_t st) {
    // dynamic shared memory: candidate lists + the per-candidate pool (corners, scores, pending weights);
    // as much as one CTA can have next to the static part, so small images keep long pending lists on chip
    const size_t lists = (size_t)kPairsCta * 10;
    const size_t most = (227 * 1024 - sizeof(K3Smem) - 1024 - lists) & ~(size_t)15;
    const size_t smem = lists + most;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k3_softnms_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    k3_softnms_kernel<NT><<<a.B, NT, smem, st>>>(a, (int)most);
    return cudaGetLastError();
}

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
    switch (a.threads) {
        case 256: return launch_k3_nt<256>(a, st);
        case 1024: return launch_k3_nt<1024>(a, st);
        default: return launch_k3_nt<512>(a, st);
    }
}

}  // namespace bod
