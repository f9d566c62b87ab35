/* pdq_oracle.c — CPU restatement of the PDQ spatial-quality path.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as bayesod_oracle.c): only tests/, bench.py's CPU
 * legs and __graft_entry__.smoke() may load it; the product never does.
 *
 * What it follows (reference repo, src/retina_net/offline_eval/):
 *   pdq_data_holders.py:92-117    PBoxDetInst.calc_heatmap      -> orc_pdq_heatmap
 *   pdq_data_holders.py:120-182   find_roi                      -> orc_pdq_find_roi
 *   pdq_data_holders.py:185-247   gen_single_heatmap            -> orc_pdq_single_heatmap
 *   pdq_data_holders.py:250-260   generate_bounding_box_from_mask (inside find_roi)
 *   pdq.py:199-230                _calc_bg_loss / _calc_fg_loss / _safe_log -> orc_pdq_losses
 *   pdq.py:160-165                _vectorize_img_gts background masks       -> orc_pdq_losses
 *   bdd/compute_pdq.py:107-113    box-shaped ground-truth masks             -> orc_pdq_losses
 *
 * Third-party arithmetic that is not under /root/reference: scipy.stats.multivariate_normal.cdf
 * (scipy unpinned, requirements.txt:7).  For two dimensions every scipy since the Fortran
 * mvndst.f evaluates it deterministically to ~1e-16 (checked here against scipy 1.18.1, which
 * IS installed in the authoring container: tests/golden/make_pdq_golden.py executes the
 * reference's pdq_data_holders.py verbatim, so this oracle is pinned by those goldens).
 * orc_pdq_bvn_cdf restates the published algorithm behind it: A. Genz, "Numerical computation
 * of rectangular bivariate and trivariate normal and t probabilities", Statistics and
 * Computing 14 (2004) — the Drezner-Wesolowsky integral with 6/12/20-point Gauss-Legendre
 * rules and the |r| >= 0.925 expansion.
 *
 * Arithmetic: binary64 for the CDF, the Mahalanobis distances and the loss sums; binary32
 * where numpy computes in float32 (the heat maps and their element-wise logs).  In-place
 * `float32_array += float64_scalar` follows numpy >= 2 (add in binary64, round once).
 * Parity with the CUDA path is by tolerance (stated in tests/test_pdq.py), not bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HEATMAP_THRESH 0.0027f          /* pdq_data_holders.py:8  */
#define MAH_DIST_THRESH 3.439           /* pdq_data_holders.py:9  */
#define SMALL_VAL 1e-14                 /* pdq_data_holders.py:10, pdq.py:8 */

double orc_pdq_phi(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* Gauss-Legendre abscissae / weights (half rules) of orders 6, 12, 20 */
static const double GL_X[3][10] = {
    {0.9324695142031522, 0.6612093864662647, 0.2386191860831970},
    {0.9815606342467191, 0.9041172563704750, 0.7699026741943050, 0.5873179542866171, 0.3678314989981802,
     0.1252334085114692},
    {0.9931285991850949, 0.9639719272779138, 0.9122344282513259, 0.8391169718222188, 0.7463319064601508,
     0.6360536807265150, 0.5108670019508271, 0.3737060887154196, 0.2277858511416451, 0.07652652113349733}};
static const double GL_W[3][10] = {
    {0.1713244923791705, 0.3607615730481384, 0.4679139345726904},
    {0.04717533638651177, 0.1069393259953183, 0.1600783285433464, 0.2031674267230659, 0.2334925365383547,
     0.2491470458134029},
    {0.01761400713915212, 0.04060142980038694, 0.06267204833410906, 0.08327674157670475, 0.1019301198172404,
     0.1181945319615184, 0.1316886384491766, 0.1420961093183821, 0.1491729864726037, 0.1527533871307259}};
static const int GL_N[3] = {3, 6, 10};

/* P(X > dh, Y > dk) for a standard bivariate normal with correlation r (Genz 2004, BVND) */
static double bvnd(double dh, double dk, double r) {
    const double twopi = 6.283185307179586;
    const int ng = fabs(r) < 0.3 ? 0 : fabs(r) < 0.75 ? 1 : 2;
    const int lg = GL_N[ng];
    double h = dh, k = dk, hk = h * k, bvn = 0.0;
    if (fabs(r) < 0.925) {
        if (fabs(r) > 0) {
            const double hs = (h * h + k * k) / 2, asr = asin(r);
            for (int i = 0; i < lg; ++i)
                for (int is = -1; is <= 1; is += 2) {
                    const double sn = sin(asr * (is * GL_X[ng][i] + 1) / 2);
                    bvn += GL_W[ng][i] * exp((sn * hk - hs) / (1 - sn * sn));
                }
            bvn = bvn * asr / (2 * twopi);
        }
        bvn += orc_pdq_phi(-h) * orc_pdq_phi(-k);
    } else {
        if (r < 0) { k = -k; hk = -hk; }
        if (fabs(r) < 1) {
            const double as = (1 - r) * (1 + r);
            double a = sqrt(as);
            const double bs = (h - k) * (h - k), c = (4 - hk) / 8, d = (12 - hk) / 16;
            double asr = -(bs / as + hk) / 2;
            if (asr > -100) bvn = a * exp(asr) * (1 - c * (bs - as) * (1 - d * bs / 5) / 3 + c * d * as * as / 5);
            if (-hk < 100) {
                const double b = sqrt(bs);
                bvn -= exp(-hk / 2) * sqrt(twopi) * orc_pdq_phi(-b / a) * b * (1 - c * bs * (1 - d * bs / 5) / 3);
            }
            a /= 2;
            for (int i = 0; i < lg; ++i)
                for (int is = -1; is <= 1; is += 2) {
                    double xs = a * (is * GL_X[ng][i] + 1);
                    xs *= xs;
                    const double rs = sqrt(1 - xs);
                    asr = -(bs / xs + hk) / 2;
                    if (asr > -100)
                        bvn += a * GL_W[ng][i] * exp(asr) *
                               (exp(-hk * xs / (2 * (1 + rs) * (1 + rs))) / rs - (1 + c * xs * (1 + d * xs)));
                }
            bvn = -bvn / twopi;
        }
        if (r > 0) {
            bvn += orc_pdq_phi(-fmax(h, k));
        } else {
            bvn = -bvn;
            if (k > h) {
                if (h < 0) bvn += orc_pdq_phi(k) - orc_pdq_phi(h);
                else bvn += orc_pdq_phi(-h) - orc_pdq_phi(-k);
            }
        }
    }
    return bvn < 0 ? 0 : bvn > 1 ? 1 : bvn;
}

/* P(X <= h, Y <= k), standardised */
double orc_pdq_bvn_cdf(double h, double k, double r) { return bvnd(-h, -k, r); }

/* multivariate_normal(mean, cov).cdf([y, x]) with mean = (my, mx), cov = [[vy, c], [c, vx]] */
static double gauss_cdf(const double mean[2], const double cov[4], double y, double x) {
    const double sy = sqrt(cov[0]), sx = sqrt(cov[3]);
    return orc_pdq_bvn_cdf((y - mean[0]) / sy, (x - mean[1]) / sx, cov[1] / (sy * sx));
}

static int trunc_int(double v) { return (int)v; }                    /* Python int(): toward zero */
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* find_roi (pdq_data_holders.py:120-182).  roi = [x1, y1, x2, y2] inclusive.  Returns 0, or -1 where the
 * reference itself raises (the pixel of the mean falls outside the candidate window). */
int orc_pdq_find_roi(int H, int W, const double mean[2], const double cov[4], int roi[4]) {
    const double stdy = sqrt(cov[0]), stdx = sqrt(cov[3]);                                   /* :132-133 */
    const int minx = trunc_int(fmax(mean[1] - stdx * 5, 0)), miny = trunc_int(fmax(mean[0] - stdy * 5, 0));
    const int maxx = trunc_int(fmin(mean[1] + stdx * 5, W - 1)), maxy = trunc_int(fmin(mean[0] + stdy * 5, H - 1));
    const double det = cov[0] * cov[3] - cov[1] * cov[2];
    if (fabs(det) < 1e-8) {                                                                  /* :141-142 */
        roi[0] = minx; roi[1] = miny; roi[2] = imax(0, maxx); roi[3] = imax(0, maxy);
        return 0;
    }
    const int ny = imax(maxy + 1 - miny, 1), nx = imax(maxx + 1 - minx, 1);                  /* :145 */
    const double vi[4] = {cov[3] / det, -cov[1] / det, -cov[2] / det, cov[0] / det};         /* np.linalg.inv */
    const int dmy = imax(imin(trunc_int(mean[0] - miny), H - 1), 0);                         /* :161-162 */
    const int dmx = imax(imin(trunc_int(mean[1] - minx), W - 1), 0);
    const int shift_y = dmy > 0 && dmy < H - 1, shift_x = dmx > 0 && dmx < W - 1;
    if (dmy >= ny || dmx >= nx) return -1;           /* :164/:166 shape mismatch or :172 IndexError in the reference */
    int x1 = nx, y1 = ny, x2 = -1, y2 = -1;
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            /* rows above / columns left of the mean take the distance of the next row / column (:163-166):
             * the row shift is applied first, then the column shift on the shifted matrix */
            const int sy = (shift_y && y < dmy) ? y + 1 : y, sx = (shift_x && x < dmx) ? x + 1 : x;
            const double dy = (double)(sy + miny) - mean[0], dx = (double)(sx + minx) - mean[1];
            const double m = sqrt(dy * (vi[0] * dy + vi[1] * dx) + dx * (vi[2] * dy + vi[3] * dx));   /* cdist 'mahalanobis' */
            if (m <= MAH_DIST_THRESH || (y == dmy && x == dmx)) {                            /* :170-172 */
                if (x < x1) x1 = x;
                if (x > x2) x2 = x;
                if (y < y1) y1 = y;
                if (y > y2) y2 = y;
            }
        }
    roi[0] = imax(0, x1 + minx); roi[1] = imax(0, y1 + miny);                                /* :175-179 */
    roi[2] = imax(0, x2 + minx); roi[3] = imax(0, y2 + miny);
    return 0;
}

/* gen_single_heatmap (pdq_data_holders.py:185-247): heatmap [H,W] float32 */
int orc_pdq_single_heatmap(int H, int W, const double mean[2], const double cov[4], float* hm) {
    int roi[4];
    if (!(cov[0] > 0) || !(cov[3] > 0)) return -2;
    if (orc_pdq_find_roi(H, W, mean, cov, roi)) return -1;
    const int x1 = roi[0], y1 = roi[1], x2 = roi[2], y2 = roi[3];
    if (x2 > W - 1 || y2 > H - 1 || x1 > x2 || y1 > y2) return -1;
    memset(hm, 0, sizeof(float) * (size_t)H * W);
    for (int y = y1; y <= y2; ++y)                                                           /* :199-207 */
        for (int x = x1; x <= x2; ++x)
            hm[(size_t)y * W + x] = (float)gauss_cdf(mean, cov, (double)(y + 1) - SMALL_VAL, (double)(x + 1) - SMALL_VAL);
    for (int y = y2 + 1; y < H; ++y)                                                         /* :208-209 */
        for (int x = x1; x <= x2; ++x) hm[(size_t)y * W + x] = hm[(size_t)y2 * W + x];
    for (int y = y1; y <= y2; ++y)                                                           /* :210-212 */
        for (int x = x2 + 1; x < W; ++x) hm[(size_t)y * W + x] = hm[(size_t)y * W + x2];
    for (int y = y2 + 1; y < H; ++y)                                                         /* :213 */
        for (int x = x2 + 1; x < W; ++x) hm[(size_t)y * W + x] = 1.0f;
    if (x1 == 0) {                                                                           /* :217-227 */
        float last = 0.f;
        for (int y = y1; y < H; ++y) {
            if (y <= y2) last = (float)gauss_cdf(mean, cov, (double)(y + 1) - SMALL_VAL, 0.0 - SMALL_VAL);
            for (int x = 0; x < W; ++x) hm[(size_t)y * W + x] -= last;
        }
    }
    if (y1 == 0) {                                                                           /* :230-236 */
        float* row = (float*)calloc((size_t)W, sizeof(float));
        if (!row) return -3;
        for (int x = x1; x < W; ++x)
            row[x] = x <= x2 ? (float)gauss_cdf(mean, cov, 0.0 - SMALL_VAL, (double)(x + 1) - SMALL_VAL) : row[x2];
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) hm[(size_t)y * W + x] -= row[x];
        free(row);
    }
    if (x1 == 0 && y1 == 0) {                                                                /* :240-241 */
        const double c = gauss_cdf(mean, cov, 0.0 - SMALL_VAL, 0.0 - SMALL_VAL);
        for (size_t i = 0; i < (size_t)H * W; ++i) hm[i] = (float)((double)hm[i] + c);
    }
    for (size_t i = 0; i < (size_t)H * W; ++i)                                               /* :243 */
        if (hm[i] < HEATMAP_THRESH) hm[i] = 0.f;
    return 0;
}

/* PBoxDetInst.calc_heatmap (pdq_data_holders.py:92-117).  box = [x1, y1, x2, y2] (ints, compute_pdq.py:118-120);
 * covs = the two 2x2 corner covariances [[var_x, c], [c, var_y]] (top-left, bottom-right), row-major. */
int orc_pdq_heatmap(int H, int W, const int32_t box[4], const double covs[8], float* hm) {
    float* p2 = (float*)malloc(sizeof(float) * (size_t)H * W);
    if (!p2) return -3;
    /* :96 flipud(fliplr(cov)): [[var_y, c'], [c, var_x]];  :101 the second one transposed */
    const double c1[4] = {covs[3], covs[2], covs[1], covs[0]};
    const double c2[4] = {covs[7], covs[5], covs[6], covs[4]};
    const double m1[2] = {(double)box[1], (double)box[0]};                                   /* :98-99 */
    const double m2[2] = {(double)(H - (box[3] + 1)), (double)(W - (box[2] + 1))};           /* :100-103 */
    int rc = orc_pdq_single_heatmap(H, W, m1, c1, hm);
    if (!rc) rc = orc_pdq_single_heatmap(H, W, m2, c2, p2);
    if (!rc)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float v = hm[(size_t)y * W + x] * p2[(size_t)(H - 1 - y) * W + (W - 1 - x)];  /* :106-109 */
                if (v > 1.f) v = 1.f;                                                        /* :113 */
                if (v < HEATMAP_THRESH) v = 0.f;                                             /* :115 */
                hm[(size_t)y * W + x] = v;
            }
    free(p2);
    return rc;
}

/* fg / bg loss sums of pdq.py:199-230 for box-shaped ground truth (compute_pdq.py:107-113):
 *   foreground of object g = rows [y1, y2) x columns [x1, x2) of gt_boxes[g] = [x1, y1, x2, y2];
 *   its background = everything outside rows [y1, y2] x columns [x1, x2] (inclusive, pdq.py:162-165).
 * fg[g*D+d] = sum_fg log(h + 1e-14);  bg[g*D+d] = sum_bg log(1 - h + 1e-14) * (h > 0);
 * bg_total[d] = the same sum over the whole image (the false-positive term, pdq.py:423-424).
 * Element-wise terms in binary32 as numpy computes them, sums in binary64. */
int orc_pdq_losses(int H, int W, int G, const int32_t* gt_boxes, int D, const float* heatmaps, double* fg, double* bg,
                   double* bg_total) {
    const float eps = (float)SMALL_VAL;
    for (int d = 0; d < D; ++d) {
        const float* hm = heatmaps + (size_t)d * H * W;
        double tot = 0;
        for (size_t i = 0; i < (size_t)H * W; ++i)
            if (hm[i] > 0.f) tot += (double)logf((1.f - hm[i]) + eps);
        bg_total[d] = tot;
        for (int g = 0; g < G; ++g) {
            const int x1 = imax(gt_boxes[g * 4 + 0], 0), y1 = imax(gt_boxes[g * 4 + 1], 0);
            const int x2 = imin(gt_boxes[g * 4 + 2], W), y2 = imin(gt_boxes[g * 4 + 3], H);
            double f = 0, in = 0;
            for (int y = y1; y < y2; ++y)
                for (int x = x1; x < x2; ++x) f += (double)logf(hm[(size_t)y * W + x] + eps);
            for (int y = y1; y <= imin(y2, H - 1); ++y)
                for (int x = x1; x <= imin(x2, W - 1); ++x) {
                    const float v = hm[(size_t)y * W + x];
                    if (v > 0.f) in += (double)logf((1.f - v) + eps);
                }
            fg[(size_t)g * D + d] = f;
            bg[(size_t)g * D + d] = tot - in;
        }
    }
    return 0;
}
