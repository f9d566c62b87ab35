// k4_fusion.cu — stage K4: per-cluster Bayesian fusion.  Compiled with -fmad=false.
//
// Reference lines replaced: bayes_od_clustering, inference_utils.py:285-364
//   :316       members = affinity[:, centre] > threshold          (bitmask row from K3)
//   :321-324   P_i = inv(Sigma_i);  Sigma_f = inv(sum_i P_i)       (information form)
//   :327-331   mu_f = Sigma_f * sum_i P_i mu_i
//   :334-349   normalised member scores; more than 3 members: keep the 3 with the
//              smallest KL(centre || member) (scipy.stats.entropy + np.argpartition;
//              ties -> lowest survivor index)
//   :351-352   cat_param = mean of the kept scores, cat_count = sum of the kept counts
//   :361       Sigma_f * 70
//
// One warp per (image, centre).  Lanes invert the covariances of 32 members at a
// time in parallel; the sums over members stay sequential in ascending survivor
// order (one lane per matrix entry walks the member bits), which keeps the result
// bit-identical to a sequential host loop while the expensive part (the 4x4
// inverses) runs 32 wide.
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

constexpr int kK4Warps = 4;

// KL(pk || qk) as scipy.stats.entropy computes it: both re-normalised, rel_entr summed.
template <int K>
BOD_DEVINL float kl_div(const float (&pk_raw)[K], const float (&qk_raw)[K]) {
    float sp = 0.0f, sq = 0.0f, acc = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) { sp = sp + pk_raw[k]; sq = sq + qk_raw[k]; }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float p = pk_raw[k] / sp, q = qk_raw[k] / sq;
        float t;
        if (p > 0.0f && q > 0.0f) t = p * log_cr(p / q);
        else if (p == 0.0f && q >= 0.0f) t = 0.0f;
        else t = INFINITY;
        acc = acc + t;
    }
    return acc;
}

template <int K>
__global__ void __launch_bounds__(kK4Warps * 32)
k4_fusion_kernel(K4Args a) {
    __shared__ float stage[kK4Warps][32][21];     // per lane: 16 precision entries + 4 weighted-mean entries (+pad)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.x * kK4Warps + warp, b = blockIdx.y;
    if (d >= a.Dmax) return;                      // warp-uniform
    if (d >= a.num_dets[b]) {                     // padding rows of the result blocks read as zero
        const size_t prow = (size_t)b * a.Dmax + d;
        if (lane < 16) a.out_covs[prow * 16 + lane] = 0.0f;
        if (lane < 4) a.out_means[prow * 4 + lane] = 0.0f;
        for (int k = lane; k < K; k += 32) { a.out_param[prow * K + k] = 0.0f; a.out_count[prow * K + k] = 0.0f; }
        return;
    }
    const int S = a.num_survivors[b];
    const int nwords = (S + 31) >> 5;
    const uint32_t* row = a.member + ((size_t)b * a.Dmax + d) * a.words;
    const float* cnt = a.cnt_post + (size_t)b * a.capacity * K;
    const float4* mu = reinterpret_cast<const float4*>(a.mu_post) + (size_t)b * a.capacity;
    const float4* sig = reinterpret_cast<const float4*>(a.sig_post) + (size_t)b * a.capacity * 4;
    const int centre = a.nms_idx[(size_t)b * a.Dmax + d];
    float (*st)[21] = stage[warp];

    // centre's normalised score (:339-340)
    float cs[K];
    {
        const float* cc = cnt + (size_t)centre * K;
        float ccs = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) ccs = ccs + cc[k];
#pragma unroll
        for (int k = 0; k < K; ++k) cs[k] = cc[k] / ccs;
    }

    float acc = 0.0f;            // lanes 0..15: precision-sum entry, lanes 16..19: weighted-mean-sum entry
    int m = 0;
    float best_kl[3] = {0.0f, 0.0f, 0.0f};
    int best_s[3] = {-1, -1, -1};
    int nbest = 0;

    for (int w = 0; w < nwords; ++w) {
        const uint32_t bits = row[w];
        if (bits == 0u) continue;                 // warp-uniform
        const int s = (w << 5) + lane;
        float kl = 0.0f;
        if ((bits >> lane) & 1u) {
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
            // member's normalised score and its KL from the centre (:335-336, 344)
            float sc[K];
            const float* c = cnt + (size_t)s * K;
            float su = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) su = su + c[k];
#pragma unroll
            for (int k = 0; k < K; ++k) sc[k] = c[k] / su;
            kl = kl_div<K>(cs, sc);
        }
        __syncwarp();
        // sequential (ascending survivor) accumulation, one lane per entry (:324, :330)
        if (lane < 20) {
            for (uint32_t rem = bits; rem; rem &= rem - 1) acc = acc + st[__ffs(rem) - 1][lane];
        }
        // running top-3 by (KL, survivor index), strict < so that earlier members win ties
        for (uint32_t rem = bits; rem; rem &= rem - 1) {
            const int l = __ffs(rem) - 1;
            const float v = __shfl_sync(0xffffffffu, kl, l);
            const int sv = (w << 5) + l;
            int pos = nbest;
            while (pos > 0 && v < best_kl[pos - 1]) --pos;
            if (pos < 3) {
                for (int q = (nbest < 3 ? nbest : 2); q > pos; --q) { best_kl[q] = best_kl[q - 1]; best_s[q] = best_s[q - 1]; }
                best_kl[pos] = v; best_s[pos] = sv;
                if (nbest < 3) ++nbest;
            }
        }
        m += __popc(bits);
        __syncwarp();
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
            for (int w = 0; w < nwords; ++w)
                for (uint32_t rem = row[w]; rem; rem &= rem - 1) {
                    const float* c = cnt + (size_t)((w << 5) + __ffs(rem) - 1) * K;
                    float su = 0.0f;
                    for (int kk = 0; kk < K; ++kk) su = su + c[kk];
                    ps = ps + c[k] / su; ns = ns + c[k]; ++used;
                }
        }
        op[k] = ps / (float)used;
        on[k] = ns;
    }
}

cudaError_t launch_k4(const K4Args& a, cudaStream_t st) {
    dim3 grid((a.Dmax + kK4Warps - 1) / kK4Warps, a.B), block(kK4Warps * 32);
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
