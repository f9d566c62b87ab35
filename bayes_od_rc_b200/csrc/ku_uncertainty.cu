// ku_uncertainty.cu — uncertainty scoring of fused detections on the GPU (second half):
// the per-detection entropies and the minimum-uncertainty-error (MUE) curve the reference's offline scripts
// compute in Python loops over every detection of the validation set.
//
// Reference lines replaced:
//   src/core/evaluation_utils_2d.py:280-285  compute_gaussian_entropy_np(cov)        (np.linalg.det, np.round(., 5))
//   src/core/evaluation_utils_2d.py:288-290  compute_categorical_entropy_np(params)  (binary32 terms and sum)
//   src/core/evaluation_utils_2d.py:129-212  compute_mu_error(gt, predictions, thresholds): rank the predictions of
//       one category by entropy (ascending, stable), match each to the best-overlapping ground-truth box of its
//       image (first maximum), greedy TP / FP marking per IoU threshold, cumulative sums, the uncertainty error
//       curve 0.5 (TP_total - TP_cum) / max(TP_total, 1) + 0.5 FP_cum / max(FP_total, 1), its minimum and arg-min
//   callers: src/retina_net/offline_eval/{bdd,kitti}/compute_uncertainty_error.py:91-132 through evaluate_u_error (:236-250)
//
// Kernels: U1 one thread per detection (4x4 determinant by LU with partial pivoting, binary32 like numpy on the
// binary32 .npy covariances; K-term entropy);
// U2 one thread per prediction (its image's ground-truth boxes are a contiguous CSR row: best IoU, first maximum);
// U3 one thread per image (the greedy marking only orders predictions of the SAME image, so images are independent:
// each walks its predictions in entropy order); U4 one CTA (two passes over the ranked flags: totals, then the
// running sums and the minimum).  Sorting stays on the host (two stable argsorts); everything else is one call.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/bayesod.h"
#include "bod_common.cuh"

namespace bod {

// ---- U1: entropies ---------------------------------------------------------
__global__ void entropy_kernel(int n, int K, const float* __restrict__ covs, const float* __restrict__ params,
                               double* __restrict__ gauss, float* __restrict__ cat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (covs && gauss) {
        // The covariances are binary32 arrays (the .npy files run_inference.py writes), so np.linalg.det runs LAPACK's
        // single-precision LU and np.round / + 1e-12 / np.log stay in binary32; only the final sum with the binary64
        // constant d/2 + d/2 log(2 pi) is binary64 (NumPy >= 2 promotion rules; under NumPy 1.x the scalar + 1e-12
        // promoted to binary64 one step earlier -- a difference of < 1e-6 in the entropy).
        float a[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[r][c] = covs[(size_t)i * 16 + 4 * r + c];
        // LU with partial pivoting (first maximum), det = sign * prod(diagonal): sgetrf's algorithm
        float det = 1.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int p = j;
            float mx = fabsf(a[j][j]);
#pragma unroll
            for (int r = j + 1; r < 4; ++r) { const float v = fabsf(a[r][j]); if (v > mx) { mx = v; p = r; } }
#pragma unroll
            for (int r = j + 1; r < 4; ++r) {
                if (p == r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) { const float t = a[j][c]; a[j][c] = a[r][c]; a[r][c] = t; }
                    det = -det;
                }
            }
            if (a[j][j] != 0.0f) {
                const float inv = 1.0f / a[j][j];
#pragma unroll
                for (int r = j + 1; r < 4; ++r) {
                    const float f = a[r][j] * inv;
#pragma unroll
                    for (int c = j + 1; c < 4; ++c) a[r][c] = a[r][c] - f * a[j][c];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) det = det * a[j][j];                         // numpy multiplies the diagonal after the factorisation
        const float rounded = rintf(det * 1.0e5f) / 1.0e5f + 1e-12f;             // np.round(det, 5) + 1e-12 (:282)
        const float term = 0.5f * logf(rounded);                                 // :284
        gauss[i] = (2.0 + 2.0 * log(2.0 * 3.141592653589793)) + (double)term;    // dims_constant = 4 / 2 (:281, :283)
    }
    if (params && cat) {
        float s = 0.0f;
        for (int k = 0; k < K; ++k) { const float p = params[(size_t)i * K + k]; s = s + p * logf(p); }
        cat[i] = -s;                                                              // :289
    }
}

// ---- U2: best-overlapping ground-truth box of every prediction -----------------
__global__ void mue_match_kernel(int n, const double* __restrict__ boxes, const int32_t* __restrict__ image,
                                 const int32_t* __restrict__ gt_off, const double* __restrict__ gt,
                                 double* __restrict__ ovmax, int32_t* __restrict__ jmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x1 = boxes[4 * (size_t)i], y1 = boxes[4 * (size_t)i + 1], x2 = boxes[4 * (size_t)i + 2], y2 = boxes[4 * (size_t)i + 3];
    const int im = image[i];
    double best = -INFINITY;
    int jb = -1;
    for (int j = gt_off[im]; j < gt_off[im + 1]; ++j) {                          // :160-177
        const double gx1 = gt[4 * (size_t)j], gy1 = gt[4 * (size_t)j + 1], gx2 = gt[4 * (size_t)j + 2], gy2 = gt[4 * (size_t)j + 3];
        const double iw = fmax(fmin(gx2, x2) - fmax(gx1, x1) + 1.0, 0.0);
        const double ih = fmax(fmin(gy2, y2) - fmax(gy1, y1) + 1.0, 0.0);
        const double inters = iw * ih;
        const double uni = ((x2 - x1 + 1.0) * (y2 - y1 + 1.0) + (gx2 - gx1 + 1.0) * (gy2 - gy1 + 1.0)) - inters;
        const double ov = inters / uni;
        if (ov > best) { best = ov; jb = j; }                                    // np.max / np.argmax: first maximum
    }
    ovmax[i] = best; jmax[i] = jb;
}

// ---- U3: greedy TP / FP marking, one thread per image ---------------------------
// by_image lists, image after image, the RANKS (positions in entropy order) of the image's predictions, ascending
__global__ void mue_greedy_kernel(int n_images, int T, const int32_t* __restrict__ img_off, const int32_t* __restrict__ by_image,
                                  const int32_t* __restrict__ order, const double* __restrict__ ovmax, const int32_t* __restrict__ jmax,
                                  const double* __restrict__ thr, uint8_t* __restrict__ checked /*[G,T]*/, uint8_t* __restrict__ tp /*[n,T]*/) {
    const int im = blockIdx.x * blockDim.x + threadIdx.x;
    if (im >= n_images) return;
    for (int e = img_off[im]; e < img_off[im + 1]; ++e) {
        const int rank = by_image[e];
        const int i = order[rank];
        const double ov = ovmax[i];
        const int j = jmax[i];
        for (int t = 0; t < T; ++t) {                                            // :181-192
            uint8_t is_tp = 0;
            if (ov > thr[t] && checked[(size_t)j * T + t] == 0) { is_tp = 1; checked[(size_t)j * T + t] = 1; }
            tp[(size_t)rank * T + t] = is_tp;
        }
    }
}

// ---- U4: the curve and its minimum, one CTA ------------------------------------
__global__ void __launch_bounds__(1024) mue_curve_kernel(int n, int T, const uint8_t* __restrict__ tp, double* __restrict__ out /*[2 + 2T]*/) {
    __shared__ long long s_tot[32];
    __shared__ long long s_scan[1024];
    __shared__ double s_min[32];
    __shared__ long long s_arg[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double best = INFINITY;
    long long best_at = 0;                                                       // flat index into the [n, T] matrix
    for (int t = 0; t < T; ++t) {
        // totals (:200-201): every prediction is a TP or an FP at a threshold
        long long c = 0;
        for (int i = tid; i < n; i += 1024) c += tp[(size_t)i * T + t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) s_tot[warp] = c;
        __syncthreads();
        long long total_tp = 0;
        for (int w = 0; w < 32; ++w) total_tp += s_tot[w];
        __syncthreads();
        const long long total_fp = (long long)n - total_tp;
        const double dtp = (double)(total_tp > 1 ? total_tp : 1), dfp = (double)(total_fp > 1 ? total_fp : 1);
        // running sums in chunks of 1024 ranks (:203-207)
        long long carry = 0;
        for (int base = 0; base < n; base += 1024) {
            const int i = base + tid;
            const long long v = (i < n) ? tp[(size_t)i * T + t] : 0;
            s_scan[tid] = v;
            __syncthreads();
            for (int o = 1; o < 1024; o <<= 1) {
                const long long add = (tid >= o) ? s_scan[tid - o] : 0;
                __syncthreads();
                s_scan[tid] += add;
                __syncthreads();
            }
            if (i < n) {
                const long long tpc = carry + s_scan[tid];
                const long long fpc = (long long)(i + 1) - tpc;
                const double u = 0.5 * (double)(total_tp - tpc) / dtp + 0.5 * (double)fpc / dfp;
                const long long at = (long long)i * T + t;
                if (u < best || (u == best && at < best_at)) { best = u; best_at = at; }
            }
            carry += s_scan[1023];
            __syncthreads();
        }
        if (tid == 0) { out[2 + 2 * t] = (double)total_tp; out[3 + 2 * t] = (double)total_fp; }
    }
    // np.min / np.argmin over the whole [n, T] matrix: smallest value, first flat index (:209-213)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oa = __shfl_xor_sync(0xffffffffu, best_at, o);
        if (ob < best || (ob == best && oa < best_at)) { best = ob; best_at = oa; }
    }
    if (lane == 0) { s_min[warp] = best; s_arg[warp] = best_at; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 32; ++w)
            if (s_min[w] < best || (s_min[w] == best && s_arg[w] < best_at)) { best = s_min[w]; best_at = s_arg[w]; }
        out[0] = best; out[1] = (double)best_at;
    }
}

}  // namespace bod

using namespace bod;

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <class T> T* as() { return static_cast<T*>(p); }
};
#define UCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_unc_err, sizeof g_unc_err, "%s: %s", #call, cudaGetErrorString(e_)); return BOD_ERR_CUDA; } } while (0)
thread_local char g_unc_err[256] = {0};
}  // namespace

extern "C" const char* bod_uncertainty_last_error(void) { return g_unc_err; }

extern "C" int bod_entropies(int device, int32_t n, int32_t K, const float* covs, const float* cat_params,
                             double* gaussian_out, float* categorical_out) {
    if (n < 0 || (cat_params && K < 1) || (covs && !gaussian_out) || (cat_params && !categorical_out)) return BOD_ERR_INVALID;
    if (n == 0) return BOD_OK;
    UCU(cudaSetDevice(device));
    DevBuf dc, dp, dg, dk;
    if (covs) { UCU(dc.alloc((size_t)n * 64)); UCU(dg.alloc((size_t)n * 8)); UCU(cudaMemcpy(dc.p, covs, (size_t)n * 64, cudaMemcpyHostToDevice)); }
    if (cat_params) { UCU(dp.alloc((size_t)n * K * 4)); UCU(dk.alloc((size_t)n * 4)); UCU(cudaMemcpy(dp.p, cat_params, (size_t)n * K * 4, cudaMemcpyHostToDevice)); }
    entropy_kernel<<<(n + 255) / 256, 256>>>(n, K, covs ? dc.as<float>() : nullptr, cat_params ? dp.as<float>() : nullptr,
                                             covs ? dg.as<double>() : nullptr, cat_params ? dk.as<float>() : nullptr);
    UCU(cudaGetLastError());
    if (covs) UCU(cudaMemcpy(gaussian_out, dg.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (cat_params) UCU(cudaMemcpy(categorical_out, dk.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return BOD_OK;
}

extern "C" int bod_mu_error(int device, int32_t n, const double* pred_boxes, const int32_t* pred_image, const int32_t* order,
                            int32_t n_images, const int32_t* img_off, const int32_t* by_image,
                            const int32_t* gt_off, const double* gt_boxes, int32_t n_thr, const double* thresholds,
                            double* min_u_error, int64_t* argmin_flat, double* totals /*[n_thr][2] or NULL*/) {
    if (n < 0 || n_images < 0 || n_thr < 1 || n_thr > 16 || !min_u_error || !argmin_flat) return BOD_ERR_INVALID;
    if (n > 0 && (!pred_boxes || !pred_image || !order || !img_off || !by_image || !gt_off || !thresholds)) return BOD_ERR_INVALID;
    if (n == 0) { *min_u_error = NAN; *argmin_flat = -1; return BOD_OK; }        // the reference raises on an empty list (np.min)
    const int G = gt_off[n_images];
    if (G > 0 && !gt_boxes) return BOD_ERR_INVALID;
    UCU(cudaSetDevice(device));
    DevBuf db, di, dord, dio, dbi, dgo, dg, dthr, dov, djm, dchk, dtp, dout;
    UCU(db.alloc((size_t)n * 32)); UCU(di.alloc((size_t)n * 4)); UCU(dord.alloc((size_t)n * 4));
    UCU(dio.alloc((size_t)(n_images + 1) * 4)); UCU(dbi.alloc((size_t)n * 4)); UCU(dgo.alloc((size_t)(n_images + 1) * 4));
    UCU(dg.alloc((size_t)G * 32)); UCU(dthr.alloc((size_t)n_thr * 8)); UCU(dov.alloc((size_t)n * 8)); UCU(djm.alloc((size_t)n * 4));
    UCU(dchk.alloc((size_t)(G > 0 ? G : 1) * n_thr)); UCU(dtp.alloc((size_t)n * n_thr)); UCU(dout.alloc((size_t)(2 + 2 * n_thr) * 8));
    UCU(cudaMemcpy(db.p, pred_boxes, (size_t)n * 32, cudaMemcpyHostToDevice));
    UCU(cudaMemcpy(di.p, pred_image, (size_t)n * 4, cudaMemcpyHostToDevice));
    UCU(cudaMemcpy(dord.p, order, (size_t)n * 4, cudaMemcpyHostToDevice));
    UCU(cudaMemcpy(dio.p, img_off, (size_t)(n_images + 1) * 4, cudaMemcpyHostToDevice));
    UCU(cudaMemcpy(dbi.p, by_image, (size_t)n * 4, cudaMemcpyHostToDevice));
    UCU(cudaMemcpy(dgo.p, gt_off, (size_t)(n_images + 1) * 4, cudaMemcpyHostToDevice));
    if (G > 0) UCU(cudaMemcpy(dg.p, gt_boxes, (size_t)G * 32, cudaMemcpyHostToDevice));
    UCU(cudaMemcpy(dthr.p, thresholds, (size_t)n_thr * 8, cudaMemcpyHostToDevice));
    UCU(cudaMemset(dchk.p, 0, (size_t)(G > 0 ? G : 1) * n_thr));
    mue_match_kernel<<<(n + 255) / 256, 256>>>(n, db.as<double>(), di.as<int32_t>(), dgo.as<int32_t>(), dg.as<double>(),
                                               dov.as<double>(), djm.as<int32_t>());
    mue_greedy_kernel<<<(n_images + 127) / 128, 128>>>(n_images, n_thr, dio.as<int32_t>(), dbi.as<int32_t>(), dord.as<int32_t>(),
                                                       dov.as<double>(), djm.as<int32_t>(), dthr.as<double>(), dchk.as<uint8_t>(),
                                                       dtp.as<uint8_t>());
    mue_curve_kernel<<<1, 1024>>>(n, n_thr, dtp.as<uint8_t>(), dout.as<double>());
    UCU(cudaGetLastError());
    std::vector<double> out(2 + 2 * n_thr);
    UCU(cudaMemcpy(out.data(), dout.p, out.size() * 8, cudaMemcpyDeviceToHost));
    *min_u_error = out[0];
    *argmin_flat = (int64_t)out[1];
    if (totals) for (int t = 0; t < 2 * n_thr; ++t) totals[t] = out[2 + t];
    return BOD_OK;
}
