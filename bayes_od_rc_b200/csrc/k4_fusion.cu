// k4_fusion.cu — stage K4: cluster membership + per-cluster Bayesian fusion.  Compiled with -fmad=false.
//
// Reference lines replaced: inference_utils.py:214-215 (box_utils.bbox_iou_vuvu(corners, corners), [S,S]:
// only the D centre columns are ever read, :316, so only those are evaluated, as bits) and
// bayes_od_clustering, inference_utils.py:285-364
//   :316       members = affinity[:, centre] > threshold          (one bitmask row per centre)
//   :321-324   P_i = inv(Sigma_i);  Sigma_f = inv(sum_i P_i)       (information form)
//   :327-331   mu_f = Sigma_f * sum_i P_i mu_i
//   :334-349   normalised member scores; more than 3 members: keep the 3 with the
//              smallest KL(centre || member) (scipy.stats.entropy + np.argpartition;
//              ties -> lowest survivor index)
//   :351-352   cat_param = mean of the kept scores, cat_count = sum of the kept counts
//   :361       Sigma_f * 70
//
// One warp per (image, centre).  The warp first tests every survivor of the image against its centre
// (32 survivors per step: one coalesced 512-byte read of corners, the ballot is the row's next word; the
// exact IoU with its division only runs for boxes that can overlap at all) and writes the bitmask row
// consumers fetch (bod_fetch_members).  The member bits of a cluster are scattered over
// the survivor index space, so they are first compacted into an ascending index
// list; then 32 members at a time are handled densely: lanes invert the members'
// covariances in parallel, while the sums over members stay sequential in
// ascending survivor order (one lane per matrix entry walks the chunk), which
// keeps the result bit-identical to a sequential host loop while the expensive
// part (the 4x4 inverses, the KL terms) runs 32 wide.
#include <cstdlib>
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

constexpr int kK4Warps = 4;
constexpr int kK4List = 1024;       // member indices buffered per warp (larger clusters are streamed in segments)

// KL(pk || qk) as scipy.stats.entropy computes it: both re-normalised, rel_entr summed.
// pk is passed already normalised (it is the same for every member of a cluster).
template <int K>
BOD_DEVINL float kl_div_norm(const float (&p)[K], const float (&qk_raw)[K]) {
    float sq = 0.0f, acc = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) sq = sq + qk_raw[k];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float pk = p[k], q = qk_raw[k] / sq;
        float t;
        if (pk > 0.0f && q > 0.0f) t = (pk == q) ? 0.0f : pk * log_cr(pk / q);   // log(1) = 0 exactly: p * 0 = 0
        else if (pk == 0.0f && q >= 0.0f) t = 0.0f;
        else t = INFINITY;
        acc = acc + t;
    }
    return acc;
}

// one (image, centre) on one warp; st / ml: the warp's staging rows and member index list in shared memory
template <int K>
BOD_DEVINL void k4_centre(const K4Args& a, const int d, const int b, float (*st)[21], uint32_t* ml) {
    const int lane = threadIdx.x & 31;
    if (d >= a.num_dets[b]) {                     // padding rows of the result blocks read as zero
        const size_t prow = (size_t)b * a.Dmax + d;
        if (lane < 16) a.out_covs[prow * 16 + lane] = 0.0f;
        if (lane < 4) a.out_means[prow * 4 + lane] = 0.0f;
        for (int k = lane; k < K; k += 32) { a.out_param[prow * K + k] = 0.0f; a.out_count[prow * K + k] = 0.0f; }
        return;
    }
    const int S = a.num_survivors[b];
    const int nwords = (S + 31) >> 5;
    uint32_t* row = a.member + ((size_t)b * a.Dmax + d) * a.words;
    const float* cnt = a.cnt_post + (size_t)b * a.capacity * K;
    const float4* mu = reinterpret_cast<const float4*>(a.mu_post) + (size_t)b * a.capacity;
    const float4* sig = reinterpret_cast<const float4*>(a.sig_post) + (size_t)b * a.capacity * 4;
    const int centre = a.nms_idx[(size_t)b * a.Dmax + d];

    // membership row of this centre (:214-215, :316) and the cluster size (the KL ranking is only needed
    // for more than 3 members, :338)
    int m = 0;
    if (a.corners) {
        const float4* corn = a.corners + (size_t)b * a.capacity;
        const float4 bx = corn[centre];
        const float thr = a.iou_threshold;
        for (int w0 = 0; w0 < nwords; w0 += 4) {             // four words per step: the reads are in flight together
            float4 bs[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int s = ((w0 + i) << 5) + lane;
                bs[i] = (s < S) ? corn[s] : make_float4(NAN, NAN, NAN, NAN);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int s = ((w0 + i) << 5) + lane;
                const unsigned bal = __ballot_sync(0xffffffffu, s < S && is_member(bs[i], bx, thr));
                if (w0 + i < nwords) {
                    if (lane == 0) row[w0 + i] = bal;
                    m += __popc(bal);
                }
            }
        }
        __syncwarp();                                        // the row is re-read below by the other lanes
    } else {
        for (int w = lane; w < nwords; w += 32) m += __popc(row[w]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    }
    const bool need_kl = m > 3;

    // centre's raw counts and its normalised, re-normalised score (:339-340 and scipy's pk / sum(pk))
    float cc[K], cp[K];
    {
        const float* c0 = cnt + (size_t)centre * K;
        float ccs = 0.0f, sp = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) { cc[k] = c0[k]; ccs = ccs + cc[k]; }
#pragma unroll
        for (int k = 0; k < K; ++k) { cp[k] = cc[k] / ccs; sp = sp + cp[k]; }
#pragma unroll
        for (int k = 0; k < K; ++k) cp[k] = cp[k] / sp;
    }

    float acc = 0.0f;            // lanes 0..15: precision-sum entry, lanes 16..19: weighted-mean-sum entry
    // running top-3 by (KL, survivor index); replicated on every lane
    float best_kl[3] = {INFINITY, INFINITY, INFINITY};
    int best_s[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    int nbest = 0;

    for (int w0 = 0; w0 < nwords;) {
        // ---- compact the member bits of the next words into an ascending index list ----
        int filled = 0, w = w0;
        for (; w < nwords; w += 32) {
            const int wi = w + lane;
            const uint32_t bits = (wi < nwords) ? row[wi] : 0u;
            int c = __popc(bits), pre = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += y; }
            const int tot = __shfl_sync(0xffffffffu, pre, 31);
            if (filled + tot > kK4List) break;                     // segment full: process what we have first
            int pos = filled + pre - c;
            for (uint32_t rem = bits; rem; rem &= rem - 1) ml[pos++] = (uint32_t)((wi << 5) + __ffs(rem) - 1);
            filled += tot;
        }
        if (w == w0) {
            // a single 32-word stripe holds more members than the buffer (cannot happen with kK4List >= 1024)
            return;
        }
        w0 = w;
        __syncwarp();

        // ---- 32 members at a time ----
        for (int base = 0; base < filled; base += 32) {
            const int nb = min(32, filled - base);
            float kl = INFINITY;
            int s = 0x7fffffff;
            if (lane < nb) {
                s = (int)ml[base + lane];
                float Sg[4][4], P[4][4], x[4], y[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 r = sig[(size_t)s * 4 + i];
                    Sg[i][0] = r.x; Sg[i][1] = r.y; Sg[i][2] = r.z; Sg[i][3] = r.w;
                }
                inv4(Sg, P);                                                       // :321-322
                const float4 mv = mu[s];
                x[0] = mv.x; x[1] = mv.y; x[2] = mv.z; x[3] = mv.w;
                mv4(P, x, y);                                                      // :327-329
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) st[lane][4 * i + j] = P[i][j];
#pragma unroll
                for (int i = 0; i < 4; ++i) st[lane][16 + i] = y[i];
                if (need_kl) {
                    // member's normalised score and its KL from the centre (:335-336, 344)
                    float c[K];
                    const float* cs = cnt + (size_t)s * K;
                    bool same = true;
#pragma unroll
                    for (int k = 0; k < K; ++k) { c[k] = cs[k]; same = same && (c[k] == cc[k]); }
                    if (same) {
                        kl = 0.0f;                     // identical counts: every term is p * log(1) = 0
#pragma unroll
                        for (int k = 0; k < K; ++k) if (!(cp[k] >= 0.0f)) kl = NAN;    // (NaN counts propagate)
                    } else {
                        float su = 0.0f;
#pragma unroll
                        for (int k = 0; k < K; ++k) su = su + c[k];
#pragma unroll
                        for (int k = 0; k < K; ++k) c[k] = c[k] / su;
                        kl = kl_div_norm<K>(cp, c);
                    }
                }
            }
            __syncwarp();
            // sequential (ascending survivor) accumulation, one lane per entry (:324, :330)
            if (lane < 20) {
                for (int e = 0; e < nb; ++e) acc = acc + st[e][lane];
            }
            if (need_kl) {
                // chunk-local three smallest (KL, index), merged into the running three; strict < keeps
                // the earlier member on ties (np.argpartition's tie order is implementation-defined)
                float ckl = kl;
                int cs_ = s;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    // lexicographic arg-min of (ckl, cs_) across the warp; NaN never wins
                    float bk = ckl; int bs = cs_;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
                        const int os = __shfl_xor_sync(0xffffffffu, bs, o);
                        if (ok < bk || (ok == bk && os < bs) || (!(bk == bk) && ok == ok)) { bk = ok; bs = os; }
                    }
                    if (bs == 0x7fffffff) break;                                   // chunk exhausted
                    if (cs_ == bs) { ckl = INFINITY; cs_ = 0x7fffffff; }           // consumed
                    // insert into the running top-3 (ascending KL, ties -> earlier index stays first)
                    int pos = nbest;
                    while (pos > 0 && bk < best_kl[pos - 1]) --pos;
                    if (pos < 3) {
                        for (int q = (nbest < 3 ? nbest : 2); q > pos; --q) { best_kl[q] = best_kl[q - 1]; best_s[q] = best_s[q - 1]; }
                        best_kl[pos] = bk; best_s[pos] = bs;
                        if (nbest < 3) ++nbest;
                    } else break;                                                  // the rest of the chunk is not smaller either
                }
            }
            __syncwarp();
        }
    }

    const size_t orow = (size_t)b * a.Dmax + d;
    float* om = a.out_means + orow * 4;
    float* oc = a.out_covs + orow * 16;
    float* op = a.out_param + orow * K;
    float* on = a.out_count + orow * K;
    if (m == 0) {   // the reference raises on an empty cluster; emit NaNs like the oracle
        if (lane < 16) oc[lane] = NAN;
        if (lane < 4) om[lane] = NAN;
        for (int k = lane; k < K; k += 32) { op[k] = NAN; on[k] = 0.0f; }
        return;
    }

    // gather the 20 sums on every lane, invert, fuse (:324, :331, :361)
    float Ps[4][4], ws[4], Fc[4][4], mf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Ps[i][j] = __shfl_sync(0xffffffffu, acc, 4 * i + j);
#pragma unroll
    for (int i = 0; i < 4; ++i) ws[i] = __shfl_sync(0xffffffffu, acc, 16 + i);
    inv4(Ps, Fc);
    mv4(Fc, ws, mf);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) om[i] = mf[i];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) oc[4 * i + j] = Fc[i][j] * a.calibration;
    }

    // categorical merge (:338-354): all members when m <= 3, else the three picks, ascending
    int pick[3] = {best_s[0], best_s[1], best_s[2]};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i + 1; j < 3; ++j)
            if (pick[j] < pick[i]) { const int t = pick[i]; pick[i] = pick[j]; pick[j] = t; }
    for (int k = lane; k < K; k += 32) {
        float ps = 0.0f, ns = 0.0f;
        int used = 0;
        if (m > 3) {
            for (int q = 0; q < 3; ++q) {
                const float* c = cnt + (size_t)pick[q] * K;
                float su = 0.0f;
                for (int kk = 0; kk < K; ++kk) su = su + c[kk];
                ps = ps + c[k] / su; ns = ns + c[k]; ++used;
            }
        } else {
            // m <= 3: all members; they are still in the index list (a cluster this small is one segment)
            for (int q = 0; q < m; ++q) {
                const float* c = cnt + (size_t)ml[q] * K;
                float su = 0.0f;
                for (int kk = 0; kk < K; ++kk) su = su + c[kk];
                ps = ps + c[k] / su; ns = ns + c[k]; ++used;
            }
        }
        op[k] = ps / (float)used;
        on[k] = ns;
    }
}

// grid = (gx, B): CTA (x, b) takes the centres x*4 + warp, stepping by 4*gx.  One CTA per group of four centres when
// the kernel has the GPU to itself; BOD_K4_GRIDX (experiments) caps gx, which makes the grid small enough to be placed
// at once beside the resident moments-kernel CTAs of a pipelined context.
template <int K>
__global__ void __launch_bounds__(kK4Warps * 32, 6)
k4_fusion_kernel(K4Args a) {
    BOD_TIMELINE(a.tl);
    __shared__ float stage[kK4Warps][32][21];     // per lane: 16 precision entries + 4 weighted-mean entries (+pad)
    __shared__ uint32_t mlist[kK4Warps][kK4List]; // ascending member indices of the segment being processed
    const int warp = threadIdx.x >> 5, b = blockIdx.y;
    if (a.status_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.status_out = *a.status_in;
    for (int d = blockIdx.x * kK4Warps + warp; d < a.Dmax; d += gridDim.x * kK4Warps) {   // warp-uniform
        k4_centre<K>(a, d, b, stage[warp], mlist[warp]);
        __syncwarp();
    }
}

cudaError_t launch_k4(const K4Args& a, cudaStream_t st) {
    int gx = (a.Dmax + kK4Warps - 1) / kK4Warps;
    static const int gx_env = getenv("BOD_K4_GRIDX") ? atoi(getenv("BOD_K4_GRIDX")) : 0;
    if (gx_env >= 1 && gx_env < gx) gx = gx_env;
    dim3 grid(gx, a.B), block(kK4Warps * 32);
    switch (a.K) {
#define BOD_CASE(KK) case KK: k4_fusion_kernel<KK><<<grid, block, 0, st>>>(a); break;
        BOD_CASE(2) BOD_CASE(3) BOD_CASE(4) BOD_CASE(5) BOD_CASE(6) BOD_CASE(7) BOD_CASE(8) BOD_CASE(9)
        BOD_CASE(10) BOD_CASE(11) BOD_CASE(12) BOD_CASE(13) BOD_CASE(16) BOD_CASE(21) BOD_CASE(32)
#undef BOD_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace bod
