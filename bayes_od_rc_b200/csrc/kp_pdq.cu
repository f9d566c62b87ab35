// kp_pdq.cu — PDQ spatial quality on the GPU: the heaviest offline consumer of the
// covariances this path produces.
//
// Reference (src/retina_net/offline_eval/), one Python process per image from a multiprocessing Pool (pdq.py:76-77):
//   pdq_data_holders.py:92-117    PBoxDetInst.calc_heatmap: product of two Gaussian-corner heat maps
//   pdq_data_holders.py:120-182   find_roi: Mahalanobis window -> region of interest of one corner
//   pdq_data_holders.py:185-247   gen_single_heatmap: bivariate normal CDF over the ROI, replicated to the image,
//                                 minus the probability mass outside the image
//   pdq.py:199-230                _calc_fg_loss / _calc_bg_loss: [H,W,G] x [H,W,D] tensordots of log heat maps
//   pdq.py:423-424                sum of the background loss over the whole image (false-positive spatial quality)
// The reference materialises D dense [H,W] float32 maps per image (3.7 MB each at 720x1280) and contracts them with
// G dense boolean masks.  Here nothing dense exists unless asked for (bod_pdq_heatmaps):
//   P1 pdq_roi_kernel     one CTA per Gaussian corner: the Mahalanobis window scan of find_roi, one thread per window
//                         row (the passing pixels of a row form one run: its ends come from the quadratic and are
//                         settled with the reference's own predicate), block-reduced to the ROI
//   P2 pdq_table_kernel   the CDF of every ROI pixel (binary64 Genz BVND; the Gauss-Legendre nodes depend on the
//                         corner only: P1 computes them once, P2 keeps them in shared memory; along a table row the
//                         quadrature terms follow a two-multiplication recurrence instead of one exp each) + the two
//                         "outside the image" border vectors, as one compact float32 table per corner; every other
//                         pixel of a corner's map is a replica of a table entry
//   P3 pdq_sum_kernel     one work item per (detection, overlapping ground-truth box) and one per detection for the
//                         whole-image term: heat-map pixels are rebuilt from the two tables on the fly, log terms in
//                         binary32 like numpy's, sums in binary64, split over kSplit CTAs with a fixed reduction order
//   P5 pdq_heatmap_kernel dense [D,H,W] maps on request (HBM-write bound: 4*H*W bytes per detection)
// Ground truth is box-shaped (bdd/compute_pdq.py:107-113), so masks are rectangles and a pair that does not overlap
// the detection's support needs no pixel work at all (host: fg = n_fg * log(1e-14), bg = whole-image sum).
//
// Arithmetic: binary64 for the CDF and the Mahalanobis distances, binary32 for the maps; CUDA's erfc/exp/sin are
// within a few ulp of binary64, i.e. the float32 table entries agree with the CPU oracle except for rare 1-ulp
// roundings; parity is by tolerance (tests/test_pdq.py), not bit-exact.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <exception>
#include <new>
#include <vector>

#include "../../include/bayesod.h"

namespace {

constexpr float kHeatThresh = 0.0027f;      // pdq_data_holders.py:8
constexpr double kMahThresh = 3.439;        // pdq_data_holders.py:9
constexpr double kSmall = 1e-14;            // pdq_data_holders.py:10, pdq.py:8
constexpr int kThreads = 256;
constexpr int kSplit = 8;                   // at most this many CTAs per work item of P3
constexpr int kSumPixelsPerCta = 8192;      // P3: an item gets one CTA per this many pixels

struct Corner {
    double mean[2];          // (y, x) in the corner's own frame
    double cov[4];           // [[var_y, c], [c', var_x]]
    double c00;              // cdf(-eps, -eps): the mass counted twice when the ROI touches both borders
    long long off;           // float offset of the table in the pool: P [rh*rw], outx [rh], outy [rw]
    int32_t x1, y1, x2, y2;  // ROI, inclusive
    int32_t status;          // 0 ok; -1 the reference raises for this corner; -2 non-positive variance
    int32_t pad;
};

struct Item {                // P3 work item: iterate rows [y_lo, y_hi] x columns [x_lo, x_hi] of detection `det`
    int32_t det, x_lo, x_hi, y_lo, y_hi;
    int32_t fx_end, fy_end;  // a pixel is foreground iff x < fx_end && y < fy_end (INT_MIN: no foreground)
    int32_t nsplit;          // CTAs that share the item's rows (1..kSplit, by its pixel count)
};

struct SumCta { int32_t item, split; };   // P3 launches one CTA per entry

__device__ __forceinline__ double phi(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

__constant__ double cGLX[3][10] = {
    {0.9324695142031522, 0.6612093864662647, 0.2386191860831970},
    {0.9815606342467191, 0.9041172563704750, 0.7699026741943050, 0.5873179542866171, 0.3678314989981802,
     0.1252334085114692},
    {0.9931285991850949, 0.9639719272779138, 0.9122344282513259, 0.8391169718222188, 0.7463319064601508,
     0.6360536807265150, 0.5108670019508271, 0.3737060887154196, 0.2277858511416451, 0.07652652113349733}};
__constant__ double cGLW[3][10] = {
    {0.1713244923791705, 0.3607615730481384, 0.4679139345726904},
    {0.04717533638651177, 0.1069393259953183, 0.1600783285433464, 0.2031674267230659, 0.2334925365383547,
     0.2491470458134029},
    {0.01761400713915212, 0.04060142980038694, 0.06267204833410906, 0.08327674157670475, 0.1019301198172404,
     0.1181945319615184, 0.1316886384491766, 0.1420961093183821, 0.1491729864726037, 0.1527533871307259}};
__constant__ int cGLN[3] = {3, 6, 10};

// Per-corner quadrature of the |r| < 0.925 branch: nodes do not depend on the pixel.
struct Quad {
    double sn[20], inv[20], w[20];   // sin(asin(r) (1 +- x_i) / 2), 1 / (1 - sn^2), weights
    double scale;                    // asin(r) / (4 pi)
    double r;
    int n;                           // 0: r == 0 or the high-correlation branch
    int pad;
};

// One node per calling thread (t in [0, 20)); thread 0 also writes the scalars.
__device__ void quad_init_node(Quad& q, double r, int t) {
    const bool mid = fabs(r) < 0.925 && fabs(r) > 0;
    const int ng = fabs(r) < 0.3 ? 0 : fabs(r) < 0.75 ? 1 : 2;
    const int n = mid ? 2 * cGLN[ng] : 0;
    if (t == 0) { q.r = r; q.n = n; q.pad = 0; q.scale = mid ? asin(r) / (4 * 3.14159265358979323846) : 0.0; }
    if (t < n) {
        const int i = t >> 1, is = (t & 1) ? 1 : -1;                     // same node order as the serial loops of the oracle
        const double sn = sin(asin(r) * (is * cGLX[ng][i] + 1) / 2);
        q.sn[t] = sn;
        q.inv[t] = 1.0 / (1 - sn * sn);
        q.w[t] = cGLW[ng][i];
    } else if (t < 20) {                                                 // unused slots: P2 copies the whole struct
        q.sn[t] = 0.0; q.inv[t] = 0.0; q.w[t] = 0.0;
    }
}

__device__ void quad_init(Quad& q, double r) {
    for (int t = 0; t < 20; ++t) quad_init_node(q, r, t);
}

// P(X > dh, Y > dk), |r| >= 0.925 (Genz 2004): same expansion as oracle/pdq_oracle.c, written for the device
__device__ __noinline__ double bvnd_high(double dh, double dk, double r) {
    const double twopi = 6.283185307179586;
    double h = dh, k = dk, hk = h * k, bvn = 0.0;
    if (r < 0) { k = -k; hk = -hk; }
    if (fabs(r) < 1) {
        const double as = (1 - r) * (1 + r);
        double a = sqrt(as);
        const double bs = (h - k) * (h - k), c = (4 - hk) / 8, d = (12 - hk) / 16;
        double asr = -(bs / as + hk) / 2;
        if (asr > -100) bvn = a * exp(asr) * (1 - c * (bs - as) * (1 - d * bs / 5) / 3 + c * d * as * as / 5);
        if (-hk < 100) {
            const double b = sqrt(bs);
            bvn -= exp(-hk / 2) * sqrt(twopi) * phi(-b / a) * b * (1 - c * bs * (1 - d * bs / 5) / 3);
        }
        a /= 2;
        for (int i = 0; i < 10; ++i)
            for (int is = -1; is <= 1; is += 2) {
                double xs = a * (is * cGLX[2][i] + 1);
                xs *= xs;
                const double rs = sqrt(1 - xs);
                asr = -(bs / xs + hk) / 2;
                if (asr > -100)
                    bvn += a * cGLW[2][i] * exp(asr) *
                           (exp(-hk * xs / (2 * (1 + rs) * (1 + rs))) / rs - (1 + c * xs * (1 + d * xs)));
            }
        bvn = -bvn / twopi;
    }
    if (r > 0) {
        bvn += phi(-fmax(h, k));
    } else {
        bvn = -bvn;
        if (k > h) bvn += h < 0 ? phi(k) - phi(h) : phi(-h) - phi(-k);
    }
    return bvn;
}

// P(X <= h, Y <= k) for the standard bivariate normal with correlation q.r
__device__ double bvn_cdf(const Quad& q, double h, double k) {
    double bvn;
    if (fabs(q.r) < 0.925) {
        bvn = 0;
        if (q.n) {
            const double hk = h * k, hs = (h * h + k * k) / 2;      // BVND(-h, -k): hk and hs are unchanged
            for (int i = 0; i < q.n; ++i) bvn += q.w[i] * exp((q.sn[i] * hk - hs) * q.inv[i]);
            bvn *= q.scale;
        }
        bvn += phi(h) * phi(k);
    } else {
        bvn = bvnd_high(-h, -k, q.r);
    }
    return bvn < 0 ? 0 : bvn > 1 ? 1 : bvn;
}

__device__ __forceinline__ double corner_cdf(const Corner& c, const Quad& q, double sy, double sx, double y, double x) {
    return bvn_cdf(q, (y - c.mean[0]) / sy, (x - c.mean[1]) / sx);
}

__device__ __forceinline__ int trunc_int(double v) { return (int)v; }

// ---------------------------------------------------------------------------------------------------------------
// P1: find_roi.  One CTA per corner (corner = 2 * detection + {0: top-left, 1: bottom-right in the flipped frame}).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pdq_roi_kernel(const int32_t* __restrict__ boxes, const double* __restrict__ covs,
                                                           int H, int W, Corner* __restrict__ corners, Quad* __restrict__ quads) {
    const int ci = blockIdx.x, d = ci >> 1, which = ci & 1;
    __shared__ Corner c;
    __shared__ int s_box[4], s_win[6];       // bbox x1 y1 x2 y2; window: minx miny nx ny dmx dmy
    __shared__ int s_mode;                   // 0 scan, 1 done (singular / error)
    if (threadIdx.x == 0) {
        const int32_t* b = boxes + 4 * d;
        const double* cv = covs + 8 * d + 4 * which;
        // calc_heatmap :96-103: covariances in (y, x) order, the second one transposed; corner means
        if (which == 0) { c.mean[0] = b[1]; c.mean[1] = b[0]; c.cov[0] = cv[3]; c.cov[1] = cv[2]; c.cov[2] = cv[1]; c.cov[3] = cv[0]; }
        else { c.mean[0] = H - (b[3] + 1); c.mean[1] = W - (b[2] + 1); c.cov[0] = cv[3]; c.cov[1] = cv[1]; c.cov[2] = cv[2]; c.cov[3] = cv[0]; }
        c.c00 = 0; c.off = 0; c.status = 0; c.pad = 0;
        c.x1 = c.y1 = c.x2 = c.y2 = 0;
        s_mode = 0;
        if (!(c.cov[0] > 0) || !(c.cov[3] > 0)) { c.status = -2; s_mode = 1; }
        else {
            const double stdy = sqrt(c.cov[0]), stdx = sqrt(c.cov[3]);
            const int minx = trunc_int(fmax(c.mean[1] - stdx * 5, 0.0)), miny = trunc_int(fmax(c.mean[0] - stdy * 5, 0.0));
            const int maxx = trunc_int(fmin(c.mean[1] + stdx * 5, (double)(W - 1)));
            const int maxy = trunc_int(fmin(c.mean[0] + stdy * 5, (double)(H - 1)));
            const double det = c.cov[0] * c.cov[3] - c.cov[1] * c.cov[2];
            if (fabs(det) < 1e-8) {                                          // :141-142
                c.x1 = minx; c.y1 = miny; c.x2 = max(0, maxx); c.y2 = max(0, maxy);
                s_mode = 1;
            } else {
                const int ny = max(maxy + 1 - miny, 1), nx = max(maxx + 1 - minx, 1);
                const int dmy = max(min(trunc_int(c.mean[0] - miny), H - 1), 0), dmx = max(min(trunc_int(c.mean[1] - minx), W - 1), 0);
                s_win[0] = minx; s_win[1] = miny; s_win[2] = nx; s_win[3] = ny; s_win[4] = dmx; s_win[5] = dmy;
                if (dmy >= ny || dmx >= nx) { c.status = -1; s_mode = 1; }
                s_box[0] = nx; s_box[1] = ny; s_box[2] = -1; s_box[3] = -1;
            }
        }
    }
    __syncthreads();
    if (s_mode == 0) {
        const int minx = s_win[0], miny = s_win[1], nx = s_win[2], ny = s_win[3], dmx = s_win[4], dmy = s_win[5];
        const bool shy = dmy > 0 && dmy < H - 1, shx = dmx > 0 && dmx < W - 1;        // :163-166
        const double det = c.cov[0] * c.cov[3] - c.cov[1] * c.cov[2];
        const double v0 = c.cov[3] / det, v1 = -c.cov[1] / det, v2 = -c.cov[2] / det, v3 = c.cov[0] / det;
        int bx1 = nx, by1 = ny, bx2 = -1, by2 = -1;
        // the reference's predicate for window pixel (y, x), evaluated exactly as its dense scan does
        auto inside = [&](int y, int x) {
            const int sy = (shy && y < dmy) ? y + 1 : y, sx = (shx && x < dmx) ? x + 1 : x;
            const double dy = (double)(sy + miny) - c.mean[0], dx = (double)(sx + minx) - c.mean[1];
            const double m = sqrt(dy * (v0 * dy + v1 * dx) + dx * (v2 * dy + v3 * dx));
            return m <= kMahThresh || (y == dmy && x == dmx);
        };
        if (det > 0 && v3 > 0) {
            // Positive definite: the pixels of a row that pass form one run (the ellipse is convex and the column shift is
            // monotone), so a row costs a handful of predicate evaluations: the ends of the run are estimated from the
            // quadratic in dx and settled by the exact predicate on the columns next to them; one thread per row.
            for (int y = threadIdx.x; y < ny; y += kThreads) {
                const int sy = (shy && y < dmy) ? y + 1 : y;
                const double dy = (double)(sy + miny) - c.mean[0];
                const double bq = (v1 + v2) * dy, cq = v0 * dy * dy - kMahThresh * kMahThresh;
                const double disc = bq * bq - 4 * v3 * cq, sq = disc > 0 ? sqrt(disc) : 0.0;
                const double off = c.mean[1] - (double)minx;                       // evaluation point s = dx + off
                const double lo = fmin(fmax((-bq - sq) / (2 * v3) + off, -4.0), (double)nx + 4.0);
                const double hi = fmin(fmax((-bq + sq) / (2 * v3) + off, -4.0), (double)nx + 4.0);
                const int s_lo = (int)ceil(lo), s_hi = (int)floor(hi);
                // column whose evaluation point is s: columns x < dmx evaluate x + 1 when the shift applies
                const int e_lo = min(max((shx && s_lo <= dmx) ? s_lo - 1 : s_lo, 0), nx - 1);    // clipped to the window: a run
                const int e_hi = min(max((shx && s_hi < dmx) ? s_hi - 1 : s_hi, 0), nx - 1);     // that leaves it ends at its edge
                int rmin = -1, rmax = -1;
                for (int x = max(e_lo - 2, 0); x <= min(e_lo + 2, nx - 1); ++x) if (inside(y, x)) { rmin = x; break; }
                for (int x = min(e_hi + 2, nx - 1); x >= max(e_hi - 2, 0); --x) if (inside(y, x)) { rmax = x; break; }
                if (y == dmy) {                                                    // the forced pixel joins the run's extent
                    rmin = rmin < 0 ? dmx : min(rmin, dmx);
                    rmax = max(rmax, dmx);
                }
                if (rmin >= 0 || rmax >= 0) {
                    if (rmin < 0) rmin = rmax;
                    if (rmax < 0) rmax = rmin;
                    bx1 = min(bx1, rmin); bx2 = max(bx2, rmax); by1 = min(by1, y); by2 = max(by2, y);
                }
            }
        } else {
            const int total = nx * ny;                                                // <= H * W
            for (int i = threadIdx.x; i < total; i += kThreads) {
                const int y = i / nx, x = i - y * nx;
                if (inside(y, x)) { bx1 = min(bx1, x); bx2 = max(bx2, x); by1 = min(by1, y); by2 = max(by2, y); }
            }
        }
        if (bx2 >= 0) { atomicMin(&s_box[0], bx1); atomicMin(&s_box[1], by1); atomicMax(&s_box[2], bx2); atomicMax(&s_box[3], by2); }
        __syncthreads();
        if (threadIdx.x == 0) {
            c.x1 = max(0, s_box[0] + minx); c.y1 = max(0, s_box[1] + miny);
            c.x2 = max(0, s_box[2] + minx); c.y2 = max(0, s_box[3] + miny);
        }
    }
    if (threadIdx.x == 0) {
        if (c.status == 0 && (c.x2 > W - 1 || c.y2 > H - 1 || c.x1 > c.x2 || c.y1 > c.y2)) c.status = -1;
        corners[ci] = c;
    }
    // the corner's quadrature nodes, one per thread (P2 only loads them)
    if (threadIdx.x < 20 && c.status != -2) quad_init_node(quads[ci], c.cov[1] / (sqrt(c.cov[0]) * sqrt(c.cov[3])), threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------
// P2: CDF tables.  Entries of a corner: P [rh*rw] | outx [rh] | outy [rw] | c00.
// Work unit = one thread: a run of kRun consecutive columns of one table row, or one border entry; one CTA per chunk
// of kThreads units (the host lays the chunks out after reading the ROIs back).
// Along a row only k = (x - mean_x) / sigma_x changes, in equal steps dk, and every quadrature term of the
// |r| < 0.925 branch is exp(f(k)) with f quadratic in k:  f(k) = inv_i (sn_i h k - h^2/2 - k^2/2).  So a run needs two
// exp per node (the term E and its step ratio Q at the first column) and then two multiplications per node and column:
//   E(k + dk) = E(k) Q(k),   Q(k + dk) = Q(k) G,   G = exp(-inv_i dk^2)  (per corner and node)
// instead of one exp per node and column (5x fewer FP64 instructions; the relative error grows by ~2e-16 per step,
// far below the float32 rounding of the table).  Runs whose arguments leave the range where E cannot underflow, steps
// larger than 2 sigma, and the high-correlation branch evaluate every column directly.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRun = 16;

struct Chunk { int32_t corner, first; };    // first work unit of the chunk

__host__ __device__ inline int runs_per_row(int rw) { return (rw + kRun - 1) / kRun; }

__global__ void __launch_bounds__(kThreads, 3) pdq_table_kernel(Corner* __restrict__ corners, const Quad* __restrict__ quads,
                                                             const Chunk* __restrict__ chunks, float* __restrict__ pool) {
    __shared__ Corner c;
    __shared__ Quad q;
    __shared__ double sG[20];
    const Chunk ch = chunks[blockIdx.x];
    if (threadIdx.x == 0) c = corners[ch.corner];
    for (int i = threadIdx.x; i < (int)(sizeof(Quad) / sizeof(double)); i += kThreads)
        reinterpret_cast<double*>(&q)[i] = reinterpret_cast<const double*>(quads + ch.corner)[i];
    __syncthreads();
    const double sy = sqrt(c.cov[0]), sx = sqrt(c.cov[3]), dk = 1.0 / sx;
    if (threadIdx.x < 20) sG[threadIdx.x] = threadIdx.x < q.n ? exp(-q.inv[threadIdx.x] * dk * dk) : 0.0;
    __syncthreads();
    const int rh = c.y2 - c.y1 + 1, rw = c.x2 - c.x1 + 1, rpr = runs_per_row(rw);
    const int np = rh * rw, nruns = rh * rpr, total = nruns + rh + rw + 1;
    const int u = ch.first + (int)threadIdx.x;
    if (u >= total) return;
    float* tab = pool + c.off;
    if (u < nruns) {                                                         // :199-207
        const int ry = u / rpr, x0 = (u - ry * rpr) * kRun, n = min(kRun, rw - x0);
        const double h = (((double)(c.y1 + ry + 1) - kSmall) - c.mean[0]) / sy;
        const double k0 = (((double)(c.x1 + x0 + 1) - kSmall) - c.mean[1]) / sx;
        const double k1 = (((double)(c.x1 + x0 + n) - kSmall) - c.mean[1]) / sx;
        float* out = tab + (long long)ry * rw + x0;
        const bool stepped = q.n > 0 && dk <= 2.0 && fabs(h) <= 7.0 && fabs(k0) <= 7.0 && fabs(k1) <= 7.0;
        if (stepped) {
            double acc[kRun];
#pragma unroll
            for (int j = 0; j < kRun; ++j) acc[j] = 0.0;
            const double hh = h * h;
            for (int i = 0; i < q.n; ++i) {
                const double inv = q.inv[i], a = q.sn[i] * h, w = q.w[i], G = sG[i];
                double E = exp((a * k0 - (hh + k0 * k0) / 2) * inv);         // the term of the direct evaluation at k0
                double Q = exp(((a - k0) * dk - dk * dk / 2) * inv);
#pragma unroll
                for (int j = 0; j < kRun; ++j) { acc[j] += w * E; E *= Q; Q *= G; }
            }
            const double ph = phi(h);
#pragma unroll
            for (int j = 0; j < kRun; ++j)
                if (j < n) {
                    const double k = (((double)(c.x1 + x0 + j + 1) - kSmall) - c.mean[1]) / sx;
                    const double v = acc[j] * q.scale + ph * phi(k);
                    out[j] = (float)(v < 0 ? 0 : v > 1 ? 1 : v);
                }
        } else {
            for (int j = 0; j < n; ++j)
                out[j] = (float)bvn_cdf(q, h, (((double)(c.x1 + x0 + j + 1) - kSmall) - c.mean[1]) / sx);
        }
    } else if (u < nruns + rh) {                                             // :217-223 (used when x1 == 0)
        const int ry = u - nruns;
        tab[np + ry] = c.x1 == 0 ? (float)corner_cdf(c, q, sy, sx, (double)(c.y1 + ry + 1) - kSmall, 0.0 - kSmall) : 0.f;
    } else if (u < nruns + rh + rw) {                                        // :230-235 (used when y1 == 0)
        const int rx = u - nruns - rh;
        tab[np + rh + rx] = c.y1 == 0 ? (float)corner_cdf(c, q, sy, sx, 0.0 - kSmall, (double)(c.x1 + rx + 1) - kSmall) : 0.f;
    } else {                                                                 // :240-241
        corners[ch.corner].c00 = (c.x1 == 0 && c.y1 == 0) ? corner_cdf(c, q, sy, sx, 0.0 - kSmall, 0.0 - kSmall) : 0.0;
    }
}

// gen_single_heatmap's value at (y, x) of the corner's own frame, rebuilt from its table
__device__ __forceinline__ float corner_value(const Corner& c, const float* __restrict__ pool, int y, int x) {
    if (y < c.y1 || x < c.x1) return 0.f;
    const int rh = c.y2 - c.y1 + 1, rw = c.x2 - c.x1 + 1;
    const int ry = min(y, c.y2) - c.y1, rx = min(x, c.x2) - c.x1;
    const float* tab = pool + c.off;
    float v = (y > c.y2 && x > c.x2) ? 1.0f : __ldg(tab + (long long)ry * rw + rx);             // :207-213
    if (c.x1 == 0) v -= __ldg(tab + (long long)rh * rw + ry);                                   // :227
    if (c.y1 == 0) v -= __ldg(tab + (long long)rh * rw + rh + rx);                              // :236
    if (c.x1 == 0 && c.y1 == 0) v = (float)((double)v + c.c00);                                 // :241
    return v < kHeatThresh ? 0.f : v;                                                           // :243
}

__device__ __forceinline__ float heat_value(const Corner& c1, const Corner& c2, const float* __restrict__ pool, int H, int W,
                                            int y, int x) {
    float v = corner_value(c1, pool, y, x) * corner_value(c2, pool, H - 1 - y, W - 1 - x);      // :106-109
    v = v > 1.f ? 1.f : v;                                                                      // :113
    return v < kHeatThresh ? 0.f : v;                                                           // :115
}

// ---------------------------------------------------------------------------------------------------------------
// P3: loss sums of one work item.  Rows are interleaved over (the item's 1..kSplit CTAs) x (8 warps); a warp walks one row with
// everything that depends on the row only (table row, border term, "below the ROI" flag of both corners) hoisted;
// the stretch of a row between the two corners' ROI columns is one value and is added as count x term, and the rows
// between the two corners' ROI rows are all equal: one is summed, weighted by their number.
// partials[cta * 2 + {fg, bg}]: the CTAs of an item are consecutive, the host adds them in order
// ---------------------------------------------------------------------------------------------------------------
struct RowView {                 // one row of gen_single_heatmap's map, in the corner's own frame
    const float* row;            // table row (clamped to the ROI's last row)
    const float* outy;           // border vector over columns (used when y1 == 0)
    float outx;                  // border term of this row (0 unless x1 == 0)
    bool below, zero;            // row below the ROI; row above the ROI (all zeros)
};

__device__ __forceinline__ RowView row_view(const Corner& c, const float* __restrict__ pool, int y) {
    RowView r;
    const int rh = c.y2 - c.y1 + 1, rw = c.x2 - c.x1 + 1;
    const int ry = min(y, c.y2) - c.y1;
    const float* tab = pool + c.off;
    r.zero = y < c.y1;
    r.below = y > c.y2;
    r.row = tab + (long long)max(ry, 0) * rw;
    r.outy = tab + (long long)rh * rw + rh;
    r.outx = (c.x1 == 0 && !r.zero) ? __ldg(tab + (long long)rh * rw + ry) : 0.f;
    return r;
}

__device__ __forceinline__ float row_value(const Corner& c, const RowView& r, int x) {   // == corner_value(c, pool, y, x)
    if (x < c.x1) return 0.f;
    const int rx = min(x, c.x2) - c.x1;
    float v = (r.below && x > c.x2) ? 1.0f : __ldg(r.row + rx);
    v -= r.outx;
    if (c.y1 == 0) v -= __ldg(r.outy + rx);
    if (c.x1 == 0 && c.y1 == 0) v = (float)((double)v + c.c00);
    return v < kHeatThresh ? 0.f : v;
}

__global__ void __launch_bounds__(kThreads) pdq_sum_kernel(const Corner* __restrict__ corners, const float* __restrict__ pool,
                                                           const Item* __restrict__ items, const SumCta* __restrict__ ctas,
                                                           int H, int W, double* __restrict__ partials) {
    __shared__ Corner c1, c2;
    __shared__ Item it;
    __shared__ double red[2][kThreads / 32];
    const SumCta me = ctas[blockIdx.x];
    if (threadIdx.x == 0) {
        it = items[me.item];
        c1 = corners[2 * it.det];
        c2 = corners[2 * it.det + 1];
    }
    __syncthreads();
    const float eps = (float)kSmall;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double fg = 0, bg = 0;
    // Rows below the first corner's ROI and above the second corner's (mirrored) ROI see both corner maps in their
    // row-replicated regime: every row of the band [ya, yb] is the same function of x.  Its first row is summed once and
    // weighted by the number of band rows (foreground: those above fy_end); the other band rows are skipped.
    const int ya = max(it.y_lo, c1.y2 + 1), yb = min(it.y_hi, H - 2 - c2.y2);
    const double log_eps = (double)logf(eps);
    for (int y = it.y_lo + me.split + warp * it.nsplit; y <= it.y_hi; y += it.nsplit * (kThreads / 32)) {
        const bool band = y >= ya && y <= yb;
        if (band && y != ya) continue;
        const int w_bg = band ? yb - ya + 1 : 1;
        const int w_fg = band ? max(min(yb, it.fy_end - 1) - ya + 1, 0) : (y < it.fy_end ? 1 : 0);
        const RowView r1 = row_view(c1, pool, y), r2 = row_view(c2, pool, H - 1 - y);
        if (r1.zero || r2.zero) {                                            // the whole row of the product is zero
            if (w_fg && lane == 0) fg += (double)w_fg * (double)max(min(it.x_hi, it.fx_end - 1) - it.x_lo + 1, 0) * log_eps;
            continue;
        }
        // Columns right of the first corner's ROI and left of the second corner's (mirrored) ROI see both corner maps
        // in their column-replicated regime: the product is one value for the whole segment [xa, xb] of this row.
        const int xa = max(it.x_lo, c1.x2 + 1), xb = min(it.x_hi, W - 2 - c2.x2);
        const bool mid = xa <= xb;
        const int n_left = mid ? xa - it.x_lo : it.x_hi - it.x_lo + 1, n_rest = mid ? n_left + (it.x_hi - xb) : n_left;
        double rf = 0, rb = 0;                                               // this lane's share of the row sums
        for (int t = lane; t < n_rest; t += 32) {
            const int x = t < n_left ? it.x_lo + t : xb + 1 + (t - n_left);
            float h = row_value(c1, r1, x) * row_value(c2, r2, W - 1 - x);   // :106-109
            h = h > 1.f ? 1.f : h;                                           // :113
            h = h < kHeatThresh ? 0.f : h;                                   // :115
            if (w_fg && x < it.fx_end) rf += (double)logf(h + eps);          // pdq.py:222-225
            if (h > 0.f) rb += (double)logf((1.f - h) + eps);                // pdq.py:207-210
        }
        if (mid && lane == 0) {
            float h = row_value(c1, r1, xa) * row_value(c2, r2, W - 1 - xa);
            h = h > 1.f ? 1.f : h;
            h = h < kHeatThresh ? 0.f : h;
            const int n_mid = xb - xa + 1, n_fg = max(min(xb, it.fx_end - 1) - xa + 1, 0);
            if (w_fg) rf += (double)n_fg * (double)logf(h + eps);
            if (h > 0.f) rb += (double)n_mid * (double)logf((1.f - h) + eps);
        }
        fg += (double)w_fg * rf;
        bg += (double)w_bg * rb;
    }
    for (int o = 16; o; o >>= 1) { fg += __shfl_xor_sync(0xffffffffu, fg, o); bg += __shfl_xor_sync(0xffffffffu, bg, o); }
    if (lane == 0) { red[0][warp] = fg; red[1][warp] = bg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double f = 0, b = 0;
        for (int i = 0; i < kThreads / 32; ++i) { f += red[0][i]; b += red[1][i]; }
        partials[(size_t)blockIdx.x * 2 + 0] = f;
        partials[(size_t)blockIdx.x * 2 + 1] = b;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// P5: dense heat maps [D, H, W]: HBM-write bound (4*H*W bytes per detection, most of them zeros).
// grid (ceil(H / kMapRows), D) — row tiles fastest, so that concurrently running CTAs write neighbouring rows of one
// map (5.7 -> 6.1 TB/s against the transposed grid); at most 65535 detections per launch.  A CTA owns kMapRows full rows, one float4 (four pixels of one row) per thread and
// step, streaming stores; rows outside the support rectangle are written without touching the tables.
// Needs W % 4 == 0; pdq_heatmap_scalar_kernel below covers other widths.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMapRows = 8;

__global__ void __launch_bounds__(kThreads) pdq_heatmap_kernel(const Corner* __restrict__ corners, const float* __restrict__ pool,
                                                               int H, int W, float* __restrict__ out) {
    __shared__ Corner c1, c2;
    if (threadIdx.x == 0) { c1 = corners[2 * blockIdx.y]; c2 = corners[2 * blockIdx.y + 1]; }
    __syncthreads();
    // support rectangle of the product: everything outside is exactly zero
    const int sy1 = c1.y1, sx1 = c1.x1, sy2 = H - 1 - c2.y1, sx2 = W - 1 - c2.x1;
    const int w4 = W >> 2, y0 = blockIdx.x * kMapRows, rows = min(kMapRows, H - y0);
    float4* o = reinterpret_cast<float4*>(out + ((size_t)blockIdx.y * H + y0) * W);
    for (int i = threadIdx.x; i < rows * w4; i += kThreads) {
        const int r = i / w4, x = (i - r * w4) << 2, y = y0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= sy1 && y <= sy2 && x + 3 >= sx1 && x <= sx2) {
            v.x = heat_value(c1, c2, pool, H, W, y, x);
            v.y = heat_value(c1, c2, pool, H, W, y, x + 1);
            v.z = heat_value(c1, c2, pool, H, W, y, x + 2);
            v.w = heat_value(c1, c2, pool, H, W, y, x + 3);
        }
        __stcs(o + i, v);
    }
}

__global__ void __launch_bounds__(kThreads) pdq_heatmap_scalar_kernel(const Corner* __restrict__ corners,
                                                                      const float* __restrict__ pool, int H, int W,
                                                                      float* __restrict__ out) {
    __shared__ Corner c1, c2;
    if (threadIdx.x == 0) { c1 = corners[2 * blockIdx.x]; c2 = corners[2 * blockIdx.x + 1]; }
    __syncthreads();
    const long long hw = (long long)H * W;
    for (long long p = (long long)blockIdx.y * kThreads + threadIdx.x; p < hw; p += (long long)gridDim.y * kThreads) {
        const int y = (int)(p / W), x = (int)(p - (long long)y * W);
        out[(size_t)blockIdx.x * hw + p] = heat_value(c1, c2, pool, H, W, y, x);
    }
}

__global__ void pdq_bvn_probe_kernel(int n, const double* __restrict__ h, const double* __restrict__ k, const double* __restrict__ r,
                                     double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Quad q;
    quad_init(q, r[i]);
    out[i] = bvn_cdf(q, h[i], k[i]);
}

// std::vector storage in page-locked host memory: the small per-call lists (ROIs, work lists, partial sums) cross
// PCIe with real asynchronous copies instead of staged ones; capacity persists across calls.
template <class T> struct PinnedAlloc {
    using value_type = T;
    PinnedAlloc() = default;
    template <class U> PinnedAlloc(const PinnedAlloc<U>&) {}
    T* allocate(size_t n) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, n * sizeof(T), cudaHostAllocDefault) != cudaSuccess) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) { cudaFreeHost(p); }
    template <class U> bool operator==(const PinnedAlloc<U>&) const { return true; }
    template <class U> bool operator!=(const PinnedAlloc<U>&) const { return false; }
};
template <class T> using PinnedVec = std::vector<T, PinnedAlloc<T>>;

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        const size_t want = std::max(n, cap * 2);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct bod_pdq_ctx {
    int device = 0, H = 0, W = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    DevBuf<int32_t> d_boxes;
    DevBuf<double> d_covs, d_partials, d_probe;
    DevBuf<Corner> d_corners;
    DevBuf<Quad> d_quads;
    DevBuf<Chunk> d_chunks;
    PinnedVec<Chunk> chunks;
    DevBuf<Item> d_items;
    DevBuf<SumCta> d_sum_ctas;
    PinnedVec<SumCta> sum_ctas;
    DevBuf<float> pool, d_maps;
    PinnedVec<Corner> corners;
    PinnedVec<Item> items;
    PinnedVec<double> partials;
    float ms[3] = {0, 0, 0};
    int64_t table_floats = 0, launches = 0;
    char err[256] = {0};
};

namespace {

char g_create_err[256] = {0};

int fail(bod_pdq_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(c->err, sizeof c->err, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(c->err, sizeof c->err, "%s", what);
    return code;
}

#define PDQ_CUDA(call, what)                                              \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return fail(ctx, BOD_ERR_CUDA, what, e_);  \
    } while (0)

// Upload the detections, run P1, bring the ROIs back, lay the tables out in the pool, run P2.
int build_tables(bod_pdq_ctx* ctx, int32_t D, const int32_t* boxes, const double* covs) {
    PDQ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    PDQ_CUDA(ctx->d_boxes.reserve((size_t)D * 4), "alloc boxes");
    PDQ_CUDA(ctx->d_covs.reserve((size_t)D * 8), "alloc covs");
    PDQ_CUDA(ctx->d_corners.reserve((size_t)D * 2), "alloc corners");
    PDQ_CUDA(ctx->d_quads.reserve((size_t)D * 2), "alloc quadrature nodes");
    PDQ_CUDA(cudaMemcpyAsync(ctx->d_boxes.p, boxes, sizeof(int32_t) * 4 * D, cudaMemcpyHostToDevice, ctx->stream), "H2D boxes");
    PDQ_CUDA(cudaMemcpyAsync(ctx->d_covs.p, covs, sizeof(double) * 8 * D, cudaMemcpyHostToDevice, ctx->stream), "H2D covs");
    PDQ_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream), "event");
    pdq_roi_kernel<<<2 * D, kThreads, 0, ctx->stream>>>(ctx->d_boxes.p, ctx->d_covs.p, ctx->H, ctx->W, ctx->d_corners.p, ctx->d_quads.p);
    PDQ_CUDA(cudaGetLastError(), "pdq_roi_kernel");
    PDQ_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream), "event");
    ctx->corners.resize((size_t)D * 2);
    PDQ_CUDA(cudaMemcpyAsync(ctx->corners.data(), ctx->d_corners.p, sizeof(Corner) * 2 * D, cudaMemcpyDeviceToHost, ctx->stream), "D2H corners");
    PDQ_CUDA(cudaStreamSynchronize(ctx->stream), "sync after P1");
    long long off = 0;
    ctx->chunks.clear();
    for (size_t i = 0; i < ctx->corners.size(); ++i) {
        Corner& c = ctx->corners[i];
        if (c.status != 0) {
            snprintf(ctx->err, sizeof ctx->err,
                     c.status == -2 ? "detection %zu: corner covariance has a non-positive variance"
                                    : "detection %zu: the reference's find_roi / gen_single_heatmap raises for this corner "
                                      "(mean outside its own candidate window or ROI outside the image)", i / 2);
            return BOD_ERR_INVALID;
        }
        c.off = off;
        const long long rh = c.y2 - c.y1 + 1, rw = c.x2 - c.x1 + 1, entries = rh * rw + rh + rw + 1;
        off += (entries + 3) & ~3LL;
        const long long units = rh * runs_per_row((int)rw) + rh + rw + 1;
        for (long long first = 0; first < units; first += kThreads) ctx->chunks.push_back(Chunk{(int32_t)i, (int32_t)first});
    }
    ctx->table_floats = off;
    PDQ_CUDA(ctx->pool.reserve((size_t)std::max<long long>(off, 4)), "alloc table pool");
    PDQ_CUDA(cudaMemcpyAsync(ctx->d_corners.p, ctx->corners.data(), sizeof(Corner) * 2 * D, cudaMemcpyHostToDevice, ctx->stream), "H2D corners");
    PDQ_CUDA(ctx->d_chunks.reserve(ctx->chunks.size()), "alloc chunks");
    PDQ_CUDA(cudaMemcpyAsync(ctx->d_chunks.p, ctx->chunks.data(), sizeof(Chunk) * ctx->chunks.size(), cudaMemcpyHostToDevice, ctx->stream), "H2D chunks");
    pdq_table_kernel<<<(unsigned)ctx->chunks.size(), kThreads, 0, ctx->stream>>>(ctx->d_corners.p, ctx->d_quads.p, ctx->d_chunks.p, ctx->pool.p);
    PDQ_CUDA(cudaGetLastError(), "pdq_table_kernel");
    PDQ_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream), "event");
    ctx->launches += 2;
    return BOD_OK;
}

int finish_timing(bod_pdq_ctx* ctx) {
    PDQ_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream), "event");
    PDQ_CUDA(cudaStreamSynchronize(ctx->stream), "sync");
    for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&ctx->ms[i], ctx->ev[i], ctx->ev[i + 1]);
    return BOD_OK;
}

}  // namespace

extern "C" int bod_pdq_create(bod_pdq_ctx** out, int device, int32_t im_h, int32_t im_w) {
    if (!out || im_h < 1 || im_w < 1 || (long long)im_h * im_w > (1LL << 30)) {       // pixel indices are 32-bit
        snprintf(g_create_err, sizeof g_create_err, "bod_pdq_create: bad argument (image size must be 1 .. 2^30 pixels)");
        return BOD_ERR_INVALID;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) {
        snprintf(g_create_err, sizeof g_create_err, "bod_pdq_create: no CUDA device %d (%s); there is no CPU fallback", device,
                 e != cudaSuccess ? cudaGetErrorString(e) : "out of range");
        return BOD_ERR_CUDA;
    }
    auto* ctx = new bod_pdq_ctx;
    ctx->device = device; ctx->H = im_h; ctx->W = im_w;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e != cudaSuccess) {
        snprintf(g_create_err, sizeof g_create_err, "bod_pdq_create: %s", cudaGetErrorString(e));
        delete ctx;
        return BOD_ERR_CUDA;
    }
    *out = ctx;
    return BOD_OK;
}

extern "C" void bod_pdq_destroy(bod_pdq_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->d_boxes.release(); ctx->d_covs.release(); ctx->d_partials.release(); ctx->d_probe.release();
    ctx->d_corners.release(); ctx->d_quads.release(); ctx->d_chunks.release(); ctx->d_items.release(); ctx->d_sum_ctas.release(); ctx->pool.release(); ctx->d_maps.release();
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* bod_pdq_last_error(const bod_pdq_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int bod_pdq_last_ms(const bod_pdq_ctx* ctx, float ms[3], int64_t* table_floats, int64_t* launches) {
    if (!ctx || !ms) return BOD_ERR_INVALID;
    for (int i = 0; i < 3; ++i) ms[i] = ctx->ms[i];
    if (table_floats) *table_floats = ctx->table_floats;
    if (launches) *launches = ctx->launches;
    return BOD_OK;
}

static int heatmaps_impl(bod_pdq_ctx* ctx, int32_t D, const int32_t* boxes, const double* covs, float* out, int32_t out_on_device) {
    if (D < 0 || (D && (!boxes || !covs || !out))) return fail(ctx, BOD_ERR_INVALID, "bod_pdq_heatmaps: bad argument");
    ctx->launches = 0;
    if (D == 0) return BOD_OK;
    int rc = build_tables(ctx, D, boxes, covs);
    if (rc != BOD_OK) return rc;
    const size_t hw = (size_t)ctx->H * ctx->W;
    float* dst = out;
    if (!out_on_device) {
        PDQ_CUDA(ctx->d_maps.reserve(hw * D), "alloc dense maps");
        dst = ctx->d_maps.p;
    }
    if (ctx->W % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0)
        for (int32_t d0 = 0; d0 < D; d0 += 65535)
            pdq_heatmap_kernel<<<dim3((unsigned)((ctx->H + kMapRows - 1) / kMapRows), (unsigned)std::min(D - d0, 65535)), kThreads, 0,
                                 ctx->stream>>>(ctx->d_corners.p + 2 * (size_t)d0, ctx->pool.p, ctx->H, ctx->W, dst + (size_t)d0 * hw);
    else
        pdq_heatmap_scalar_kernel<<<dim3(D, (unsigned)std::min<size_t>((hw + kThreads - 1) / kThreads, 65535)), kThreads, 0, ctx->stream>>>(
            ctx->d_corners.p, ctx->pool.p, ctx->H, ctx->W, dst);
    PDQ_CUDA(cudaGetLastError(), "pdq_heatmap_kernel");
    ctx->launches += 1;
    rc = finish_timing(ctx);
    if (rc != BOD_OK) return rc;
    if (!out_on_device) PDQ_CUDA(cudaMemcpy(out, dst, sizeof(float) * hw * D, cudaMemcpyDeviceToHost), "D2H maps");
    return BOD_OK;
}

// No C++ exception may cross the C ABI: host-side allocation failures (pinned or pageable) become BOD_ERR_NOMEM.
extern "C" int bod_pdq_heatmaps(bod_pdq_ctx* ctx, int32_t D, const int32_t* boxes, const double* covs, float* out,
                                int32_t out_on_device) {
    if (!ctx) return BOD_ERR_INVALID;
    try {
        return heatmaps_impl(ctx, D, boxes, covs, out, out_on_device);
    } catch (const std::exception&) {
        cudaStreamSynchronize(ctx->stream);
        return fail(ctx, BOD_ERR_NOMEM, "bod_pdq_heatmaps: host allocation failed");
    }
}

static int losses_impl(bod_pdq_ctx* ctx, int32_t n_images, const int32_t* det_offsets, const int32_t* boxes, const double* covs,
                       const int32_t* gt_offsets, const int32_t* gt_boxes, double* fg_loss, double* bg_loss, double* bg_total) {
    if (n_images < 0 || (n_images && (!det_offsets || !gt_offsets))) return fail(ctx, BOD_ERR_INVALID, "bod_pdq_losses: bad argument");
    ctx->launches = 0;
    if (n_images == 0) return BOD_OK;
    const int32_t D = det_offsets[n_images], G = gt_offsets[n_images];
    for (int b = 0; b < n_images; ++b)
        if (det_offsets[b] > det_offsets[b + 1] || gt_offsets[b] > gt_offsets[b + 1] || det_offsets[0] != 0 || gt_offsets[0] != 0)
            return fail(ctx, BOD_ERR_INVALID, "bod_pdq_losses: offsets must start at 0 and be non-decreasing");
    if ((D && (!boxes || !covs || !bg_total)) || (G && !gt_boxes) || (D && G && (!fg_loss || !bg_loss)))
        return fail(ctx, BOD_ERR_INVALID, "bod_pdq_losses: null array");
    if (D == 0) return BOD_OK;
    int rc = build_tables(ctx, D, boxes, covs);
    if (rc != BOD_OK) return rc;
    const int H = ctx->H, W = ctx->W;
    // work list: per detection the whole-image term, then one item per ground-truth box that meets its support
    struct Slot { long long out; int item; };                 // out < 0: bg_total of detection ~out
    std::vector<Slot> slots;
    ctx->items.clear();
    size_t pair_base = 0;
    const float log_eps = logf((float)kSmall);
    for (int b = 0; b < n_images; ++b) {
        const int d0 = det_offsets[b], nd = det_offsets[b + 1] - d0, g0 = gt_offsets[b], ng = gt_offsets[b + 1] - g0;
        for (int d = 0; d < nd; ++d) {
            const Corner& c1 = ctx->corners[2 * (size_t)(d0 + d)];
            const Corner& c2 = ctx->corners[2 * (size_t)(d0 + d) + 1];
            const int sx1 = c1.x1, sy1 = c1.y1, sx2 = W - 1 - c2.x1, sy2 = H - 1 - c2.y1;      // support of the product
            ctx->items.push_back(Item{d0 + d, sx1, sx2, sy1, sy2, INT32_MIN, INT32_MIN, 1});
            slots.push_back(Slot{~(long long)(d0 + d), (int)ctx->items.size() - 1});
            for (int g = 0; g < ng; ++g) {
                const int32_t* gb = gt_boxes + 4 * (size_t)(g0 + g);
                const int gx1 = std::max(gb[0], 0), gy1 = std::max(gb[1], 0), gx2 = std::min(gb[2], W), gy2 = std::min(gb[3], H);
                const long long o = (long long)(pair_base + (size_t)g * nd + d);
                // inclusive box (background exclusion, pdq.py:162-165) clipped to the support
                const int x_lo = std::max(gx1, sx1), x_hi = std::min(std::min(gx2, W - 1), sx2);
                const int y_lo = std::max(gy1, sy1), y_hi = std::min(std::min(gy2, H - 1), sy2);
                // foreground pixels outside the support contribute log(0 + 1e-14) each; inside, the kernel sums them
                const long long n_fg = (long long)std::max(gx2 - gx1, 0) * std::max(gy2 - gy1, 0);
                const long long n_in = (long long)std::max(std::min(gx2 - 1, x_hi) - x_lo + 1, 0) *
                                       std::max(std::min(gy2 - 1, y_hi) - y_lo + 1, 0);
                fg_loss[o] = (double)(n_fg - ((x_lo <= x_hi && y_lo <= y_hi) ? n_in : 0)) * (double)log_eps;
                bg_loss[o] = 0.0;
                if (x_lo <= x_hi && y_lo <= y_hi) {
                    ctx->items.push_back(Item{d0 + d, x_lo, x_hi, y_lo, y_hi, gx2, gy2, 1});
                    slots.push_back(Slot{o, (int)ctx->items.size() - 1});
                }
            }
        }
        pair_base += (size_t)ng * nd;
    }
    const size_t ni = ctx->items.size();
    ctx->sum_ctas.clear();
    for (size_t i = 0; i < ni; ++i) {
        Item& it = ctx->items[i];
        const long long px = (long long)std::max(it.x_hi - it.x_lo + 1, 0) * std::max(it.y_hi - it.y_lo + 1, 0);
        it.nsplit = (int32_t)std::min<long long>(kSplit, std::max<long long>(1, (px + kSumPixelsPerCta - 1) / kSumPixelsPerCta));
        for (int32_t sp = 0; sp < it.nsplit; ++sp) ctx->sum_ctas.push_back(SumCta{(int32_t)i, sp});
    }
    PDQ_CUDA(ctx->d_sum_ctas.reserve(ctx->sum_ctas.size()), "alloc CTA list");
    PDQ_CUDA(cudaMemcpyAsync(ctx->d_sum_ctas.p, ctx->sum_ctas.data(), sizeof(SumCta) * ctx->sum_ctas.size(), cudaMemcpyHostToDevice, ctx->stream), "H2D CTA list");
    PDQ_CUDA(ctx->d_items.reserve(ni), "alloc items");
    const size_t nc = ctx->sum_ctas.size();
    PDQ_CUDA(ctx->d_partials.reserve(nc * 2), "alloc partials");
    PDQ_CUDA(cudaMemcpyAsync(ctx->d_items.p, ctx->items.data(), sizeof(Item) * ni, cudaMemcpyHostToDevice, ctx->stream), "H2D items");
    pdq_sum_kernel<<<(unsigned)ctx->sum_ctas.size(), kThreads, 0, ctx->stream>>>(ctx->d_corners.p, ctx->pool.p, ctx->d_items.p,
                                                                                 ctx->d_sum_ctas.p, H, W, ctx->d_partials.p);
    PDQ_CUDA(cudaGetLastError(), "pdq_sum_kernel");
    ctx->launches += 1;
    ctx->partials.resize(nc * 2);
    PDQ_CUDA(cudaMemcpyAsync(ctx->partials.data(), ctx->d_partials.p, sizeof(double) * nc * 2, cudaMemcpyDeviceToHost, ctx->stream), "D2H partials");
    rc = finish_timing(ctx);
    if (rc != BOD_OK) return rc;
    // fixed-order reduction of the kSplit partial sums
    std::vector<double> fg_item(ni), bg_item(ni);
    for (size_t i = 0, c = 0; i < ni; ++i) {
        double f = 0, b = 0;
        for (int s = 0; s < ctx->items[i].nsplit; ++s, ++c) { f += ctx->partials[c * 2]; b += ctx->partials[c * 2 + 1]; }
        fg_item[i] = f; bg_item[i] = b;
    }
    for (const Slot& s : slots)
        if (s.out < 0) bg_total[~s.out] = bg_item[(size_t)s.item];
    for (const Slot& s : slots)
        if (s.out >= 0) { fg_loss[s.out] += fg_item[(size_t)s.item]; bg_loss[s.out] = -bg_item[(size_t)s.item]; }
    // bg_loss = whole-image sum - the part inside the inclusive box
    pair_base = 0;
    for (int b = 0; b < n_images; ++b) {
        const int d0 = det_offsets[b], nd = det_offsets[b + 1] - d0, ng = gt_offsets[b + 1] - gt_offsets[b];
        for (int g = 0; g < ng; ++g)
            for (int d = 0; d < nd; ++d) bg_loss[pair_base + (size_t)g * nd + d] += bg_total[d0 + d];
        pair_base += (size_t)ng * nd;
    }
    return BOD_OK;
}

extern "C" int bod_pdq_losses(bod_pdq_ctx* ctx, int32_t n_images, const int32_t* det_offsets, const int32_t* boxes,
                              const double* covs, const int32_t* gt_offsets, const int32_t* gt_boxes, double* fg_loss,
                              double* bg_loss, double* bg_total) {
    if (!ctx) return BOD_ERR_INVALID;
    try {
        return losses_impl(ctx, n_images, det_offsets, boxes, covs, gt_offsets, gt_boxes, fg_loss, bg_loss, bg_total);
    } catch (const std::exception&) {
        cudaStreamSynchronize(ctx->stream);
        return fail(ctx, BOD_ERR_NOMEM, "bod_pdq_losses: host allocation failed");
    }
}

extern "C" int bod_pdq_bvn_cdf(bod_pdq_ctx* ctx, int32_t n, const double* h, const double* k, const double* r, double* out) {
    if (!ctx) return BOD_ERR_INVALID;
    if (n < 0 || (n && (!h || !k || !r || !out))) return fail(ctx, BOD_ERR_INVALID, "bod_pdq_bvn_cdf: bad argument");
    if (n == 0) return BOD_OK;
    PDQ_CUDA(cudaSetDevice(ctx->device), "cudaSetDevice");
    PDQ_CUDA(ctx->d_probe.reserve((size_t)n * 4), "alloc probe");
    double* d = ctx->d_probe.p;
    PDQ_CUDA(cudaMemcpyAsync(d, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream), "H2D");
    PDQ_CUDA(cudaMemcpyAsync(d + n, k, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream), "H2D");
    PDQ_CUDA(cudaMemcpyAsync(d + 2 * (size_t)n, r, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream), "H2D");
    pdq_bvn_probe_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n);
    PDQ_CUDA(cudaGetLastError(), "pdq_bvn_probe_kernel");
    PDQ_CUDA(cudaMemcpyAsync(out, d + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
    PDQ_CUDA(cudaStreamSynchronize(ctx->stream), "sync");
    return BOD_OK;
}
