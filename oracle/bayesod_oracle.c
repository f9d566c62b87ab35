/*
 * bayesod_oracle.c — CPU ORACLE for the BayesOD post-head path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference arms and __graft_entry__.smoke() may load it.
 * The shipped library (bayes_od_rc_b200/csrc) never links or calls it.
 *
 * What it restates (all paths relative to the reference repo asharakeh/bayes-od-rc):
 *   src/retina_net/experiments/inference_utils.py:25-217   bayes_od_inference minus the model call
 *   src/retina_net/experiments/inference_utils.py:220-277  moments / entropy helpers
 *   src/retina_net/experiments/inference_utils.py:285-364  bayes_od_clustering
 *   src/retina_net/anchor_generator/box_utils.py:5-23, 117-146, 171-192
 *   src/retina_net/anchor_generator/fpn_anchor_generator.py:21-59
 * and, because they are third-party code that is NOT under /root/reference:
 *   tensorflow  (unpinned, requirements.txt:10; "tested on TF 2.0", README.md:7)
 *       tf.image.non_max_suppression_with_scores -> NonMaxSuppressionV5 CPU kernel
 *       (tensorflow/core/kernels/image/non_max_suppression_op.cc, as of TF 2.3+:
 *       priority queue on (score, -index), suppress_begin_index, weight
 *       exp(-0.5/sigma * iou^2), no hard suppression when sigma > 0)
 *       tf.linalg.inv -> LU with partial pivoting (restated as LAPACK sgetf2 + sgetrs)
 *   tensorflow-probability (unpinned): tfp.math.fill_triangular index map,
 *       tfp.distributions.Categorical.sample (UNSEEDED in the reference: counts
 *       are injected for parity; the Philox sampler below restates the PRODUCT's
 *       documented sampler so its draws can be checked bit for bit)
 *   scipy.stats.entropy (KL with re-normalisation), numpy.linalg.inv, np.argpartition.
 *
 * PARITY STATUS: the reference ships no tests, golden vectors or fixtures for
 * this path and TensorFlow is not installable here, so the
 * TF-kernel restatements are "parity unpinned" against real TF.  What IS pinned:
 * the tests/golden npz fixtures are produced by executing the reference's OWN source files
 * (inference_utils.py, box_utils.py) over a numpy-backed `tf` shim
 * (tests/golden/make_golden.py); this oracle is checked against them.
 *
 * Arithmetic contract (so a GPU implementation can be bit-identical):
 *   IEEE binary32, round-to-nearest-even, no FMA contraction (-ffp-contract=off),
 *   reductions are sequential in index order, exp/log are the correctly rounded
 *   binary32 values (computed in binary64 and rounded; TF's kernels call
 *   std::exp(float), which glibc guarantees only to 0.502 ULP), 4x4 inverses are
 *   unblocked right-looking LU with partial pivoting (first maximum) followed by
 *   forward/back substitution on the permuted identity.
 * The same file compiled with -DORC_REAL_IS_DOUBLE gives the binary64 twin used
 * to adjudicate ill-conditioned elements.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_REAL_IS_DOUBLE
typedef double R;
#define ORC(name) orc64_##name
static inline R r_exp(R x) { return exp(x); }
static inline R r_log(R x) { return log(x); }
#else
typedef float R;
#define ORC(name) orc32_##name
static inline R r_exp(R x) { return (float)exp((double)x); }
static inline R r_log(R x) { return (float)log((double)x); }
#endif
static inline R r_abs(R x) { return x < 0 ? -x : x; }
static inline R r_max(R a, R b) { return a > b ? a : b; }   /* tf.maximum on finite data */
static inline R r_min(R a, R b) { return a < b ? a : b; }

/* ------------------------------------------------------------------------- */
/* 4x4 linear algebra                                                         */
/* ------------------------------------------------------------------------- */

/* Unblocked LU with partial pivoting, in place (LAPACK sgetf2 semantics:
 * first-maximum pivot, reciprocal scaling of the sub-column, rank-1 update).
 * Returns the permutation sign (+1/-1), or 0 when a pivot is exactly zero. */
static int lu4(R a[4][4], int piv[4]) {
    int sign = 1;
    for (int j = 0; j < 4; ++j) {
        int p = j;
        R mx = r_abs(a[j][j]);
        for (int i = j + 1; i < 4; ++i) {
            R v = r_abs(a[i][j]);
            if (v > mx) { mx = v; p = i; }
        }
        piv[j] = p;
        if (a[p][j] == (R)0) { sign = 0; continue; }
        if (p != j) {
            for (int k = 0; k < 4; ++k) { R t = a[j][k]; a[j][k] = a[p][k]; a[p][k] = t; }
            sign = -sign;
        }
        if (j < 3) {
            R r = (R)1 / a[j][j];
            for (int i = j + 1; i < 4; ++i) a[i][j] = a[i][j] * r;
            for (int i = j + 1; i < 4; ++i)
                for (int k = j + 1; k < 4; ++k)
                    a[i][k] = a[i][k] - a[i][j] * a[j][k];
        }
    }
    return sign;
}

/* inverse = solve(A, I): tf.linalg.inv (inference_utils.py:75,101,120,129) and
 * np.linalg.inv (:321-324) are both gesv-style LU solves against the identity. */
static int inv4(const R* in, R* out) {
    R a[4][4], b[4][4];
    int piv[4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = in[4 * i + j]; b[i][j] = (i == j) ? (R)1 : (R)0; }
    int sign = lu4(a, piv);
    for (int j = 0; j < 4; ++j) {          /* row interchanges on the right-hand side */
        int p = piv[j];
        if (p != j) for (int k = 0; k < 4; ++k) { R t = b[j][k]; b[j][k] = b[p][k]; b[p][k] = t; }
    }
    for (int c = 0; c < 4; ++c) {
        for (int k = 0; k < 4; ++k)        /* L y = b, unit lower */
            for (int i = k + 1; i < 4; ++i) b[i][c] = b[i][c] - b[k][c] * a[i][k];
        for (int k = 3; k >= 0; --k) {     /* U x = y */
            b[k][c] = b[k][c] / a[k][k];
            for (int i = 0; i < k; ++i) b[i][c] = b[i][c] - b[k][c] * a[i][k];
        }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[4 * i + j] = b[i][j];
    return sign != 0;
}

/* det via the same LU (tf.linalg.det, inference_utils.py:258) */
static R det4(const R* in) {
    R a[4][4]; int piv[4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) a[i][j] = in[4 * i + j];
    int sign = lu4(a, piv);
    R d = (R)sign;
    for (int i = 0; i < 4; ++i) d = d * a[i][i];
    return d;
}

/* C = A * B (or A * B^T), dot products sequential in k (tf.matmul / np.matmul) */
static void mm4(const R* A, const R* B, int transpose_b, R* C) {
    R t[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            R s = (R)0;
            for (int k = 0; k < 4; ++k) s = s + A[4 * i + k] * (transpose_b ? B[4 * j + k] : B[4 * k + j]);
            t[4 * i + j] = s;
        }
    memcpy(C, t, sizeof t);
}
static void mv4(const R* A, const R* x, R* y) {
    R t[4];
    for (int i = 0; i < 4; ++i) {
        R s = (R)0;
        for (int k = 0; k < 4; ++k) s = s + A[4 * i + k] * x[k];
        t[i] = s;
    }
    memcpy(y, t, sizeof t);
}

/* exported for unit tests */
int ORC(inv4)(const R* in, R* out) { return inv4(in, out); }
R   ORC(det4)(const R* in) { return det4(in); }

#ifndef ORC_REAL_IS_DOUBLE
/* ------------------------------------------------------------------------- */
/* fpn_anchor_generator.py:21-59 + bdd_dataset_handler.py:161-186             */
/* ------------------------------------------------------------------------- */
/* Anchors are always binary32 (they are an INPUT tensor of the path). Levels
 * 3..7, aspect ratios [[1,1],[1,2],[2,1]], scales [1.0,1.26,1.59]
 * (retinanet_bdd.yaml:55-58).  Returns A; anchors may be NULL to query. */
int orc_generate_anchors(int im_h, int im_w, float* anchors) {
    static const float ratios[3][2] = {{1.0f, 1.0f}, {1.0f, 2.0f}, {2.0f, 1.0f}};
    static const float scales[3] = {1.0f, 1.26f, 1.59f};
    int total = 0;
    for (int level = 3; level <= 7; ++level) {
        float stride = (float)(1 << level);              /* tf.pow(2.0, layer_number) :27 */
        /* tf.range(0, im/stride): ceil(im/stride) positions :28-29 */
        int nu = (int)ceilf((float)im_w / stride);
        int nv = (int)ceilf((float)im_h / stride);
        float side = (float)(1 << (level + 2));          /* :35 */
        float dims[9][2];
        int d = 0;
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s, ++d) {
                if (ratios[r][0] == 1.0f && ratios[r][1] == 1.0f) {       /* :39-41 */
                    dims[d][0] = ratios[r][0] * side * scales[s];
                    dims[d][1] = ratios[r][1] * side * scales[s];
                } else {                                                   /* :43-48 */
                    float sol = sqrtf((side * side) / (ratios[r][0] * ratios[r][1]));
                    dims[d][0] = ratios[r][0] * sol * scales[s];
                    dims[d][1] = ratios[r][1] * sol * scales[s];
                }
            }
        if (anchors) {
            for (int iv = 0; iv < nv; ++iv)                /* meshgrid, u fastest :30-33 */
                for (int iu = 0; iu < nu; ++iu)
                    for (int a = 0; a < 9; ++a) {          /* tf_repeat / tile :53-57 */
                        float* o = anchors + 4 * (size_t)(total + (iv * nu + iu) * 9 + a);
                        o[0] = ((float)iv + 0.5f) * stride;
                        o[1] = ((float)iu + 0.5f) * stride;
                        o[2] = dims[a][0];
                        o[3] = dims[a][1];
                    }
        }
        total += nv * nu * 9;
    }
    return total;
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11) — restates the PRODUCT's sampler spec  */
/* ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Multinomial(T, probs) counts per anchor, standing in for
 * Categorical(probs).sample(30) -> one_hot -> reduce_sum (inference_utils.py:37-46).
 * The reference's draws are unseeded, so only the DISTRIBUTION is specified by
 * it; this restates the product's sampler so its draws can be checked bit for
 * bit.  Spec (all binary32, sequential):
 *   uniforms: for anchor a of global image g, call j = 0,1,.. of Philox4x32-10
 *     with counter (a, g, j, 0x0B0D) and key (seed lo, seed hi) yields 128 bits =
 *     five 23-bit fields (bits [0,23), [23,46), .. of the little-endian word);
 *     the i-th uniform of the anchor is (field_i + 0.5) * 2^-23, i = 5j + f.
 *   cdf[k] = p[0]+..+p[k], total = cdf[K-1]; m = first argmax of p.
 *   dominant-class split: pw = p[m]^T by square-and-multiply (the mean of N softmax
 *   rows sums to 1 within a few ulp, so p[m] is used as the probability of m).
 *   if pw >= 1e-30:  the number of draws NOT landing on m is Binomial(T, 1-p[m]),
 *       sampled by inversion from 0 upward with the first uniform
 *       (f_0 = pw, f_{j+1} = f_j * ((T-j)/(j+1)) * ((total-p[m]) * (1/p[m])));
 *       each of those draws then picks a class k != m with the next uniform u:
 *       first k != m with u*(total-p[m]) < running sum of p over k != m
 *       (last class != m as the fallback); cnt[m] = T - others.
 *   else (flat distribution, a^T underflows): T plain inverse-cdf draws,
 *       class = first k with u*total < cdf[k], else K-1.
 * Both branches draw exactly Multinomial(T, p/total). */
static float philox_uniform(uint32_t anchor, uint32_t image, const uint32_t key[2], uint32_t* w, int* g) {
    const int f = *g % 5;
    if (f == 0) {
        uint32_t ctr[4] = {anchor, image, (uint32_t)(*g / 5), 0x0B0Du};
        orc_philox4x32_10(ctr, key, w);
    }
    const int bit = 23 * f, wi = bit >> 5, sh = bit & 31;
    const uint64_t two = (uint64_t)w[wi] | ((wi + 1 < 4) ? ((uint64_t)w[wi + 1] << 32) : 0);
    const uint32_t field = (uint32_t)(two >> sh) & 0x7FFFFFu;
    ++*g;
    return ((float)field + 0.5f) * 1.1920928955078125e-07f;   /* 2^-23 */
}

void orc_philox_counts(const float* probs, int A, int K, int T, uint64_t seed,
                       uint32_t image_id, float* counts) {
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    float* cdf = (float*)malloc(sizeof(float) * (size_t)K);
    for (int a = 0; a < A; ++a) {
        const float* p = probs + (size_t)a * K;
        float* c = counts + (size_t)a * K;
        float s = 0.0f;
        int m = 0;
        for (int k = 0; k < K; ++k) { s = s + p[k]; cdf[k] = s; c[k] = 0.0f; if (p[k] > p[m]) m = k; }
        const float total = cdf[K - 1];
        const float pm = p[m], rest = total - pm;
        const float odds = rest * (1.0f / pm);
        float pw = 1.0f, base = pm;
        for (int e = T; e; e >>= 1) { if (e & 1) pw = pw * base; base = base * base; }
        uint32_t w[4]; int g = 0;
        if (pw >= 1e-30f) {
            const float u = philox_uniform((uint32_t)a, image_id, key, w, &g);
            int j = 0;
            float cd = pw, f = pw;
            while (u >= cd && j < T) { f = f * ((float)(T - j) / (float)(j + 1)) * odds; ++j; cd = cd + f; }
            c[m] = (float)(T - j);
            const int last = (m == K - 1) ? K - 2 : K - 1;
            for (int i = 0; i < j; ++i) {
                const float x = philox_uniform((uint32_t)a, image_id, key, w, &g) * rest;
                float acc = 0.0f;
                int cls = last;
                for (int k = 0; k < K; ++k) {
                    if (k == m) continue;
                    acc = acc + p[k];
                    if (x < acc) { cls = k; break; }
                }
                c[cls] = c[cls] + 1.0f;
            }
        } else {
            for (int t = 0; t < T; ++t) {
                const float x = philox_uniform((uint32_t)a, image_id, key, w, &g) * total;
                int cls = K - 1;
                for (int k = 0; k < K - 1; ++k) if (x < cdf[k]) { cls = k; break; }
                c[cls] = c[cls] + 1.0f;
            }
        }
    }
    free(cdf);
}

/* category_filter + boolean_mask (inference_utils.py:48-54): keep anchor iff the
 * FIRST maximum of its counts is not the background column K-1; ascending order. */
int orc_filter(const float* counts, int A, int K, int32_t* keep) {
    int S = 0;
    for (int a = 0; a < A; ++a) {
        const float* c = counts + (size_t)a * K;
        int am = 0;
        for (int k = 1; k < K; ++k) if (c[k] > c[am]) am = k;
        if (am != K - 1) keep[S++] = a;
    }
    return S;
}
#endif /* !ORC_REAL_IS_DOUBLE */

/* ------------------------------------------------------------------------- */
/* H2: softmax + mean over MC samples (inference_utils.py:31-32, 38)           */
/* ------------------------------------------------------------------------- */
void ORC(softmax_mean)(const float* cls, int N, int A, int K, R* probs) {
    R* e = (R*)malloc(sizeof(R) * (size_t)K);
    for (int a = 0; a < A; ++a) {
        R* out = probs + (size_t)a * K;
        for (int k = 0; k < K; ++k) out[k] = (R)0;
        for (int n = 0; n < N; ++n) {
            const float* x = cls + ((size_t)n * A + a) * K;
            R m = (R)x[0];
            for (int k = 1; k < K; ++k) m = r_max(m, (R)x[k]);
            R s = (R)0;
            for (int k = 0; k < K; ++k) { e[k] = r_exp((R)x[k] - m); s = s + e[k]; }
            for (int k = 0; k < K; ++k) out[k] = out[k] + e[k] / s;
        }
        for (int k = 0; k < K; ++k) out[k] = out[k] / (R)N;
    }
    free(e);
}

/* ------------------------------------------------------------------------- */
/* H1, H5-H8r, H9 corners: per-survivor posterior (inference_utils.py:28-29,57-205) */
/* ------------------------------------------------------------------------- */
typedef struct orc_params {
    int32_t N, A, K;
    int32_t cov_layout;         /* 0 none, 1 [N,A,4,4], 2 packed [N,A,10] */
    int32_t use_full_covar;
    int32_t dirichlet_prior;
    int32_t gaussian_prior;
    float   isotropic_variance;
    int32_t ranking_method;
    float   scale_v, scale_u;
} orc_params;

/* tfp.math.fill_triangular (retinanet_model.py:110): packed x0..x9 ->
 * lower-triangular 4x4: rows of concat(x[4:], reverse(x)) reshaped 4x4, lower band */
static void fill_triangular10(const float* x, R* m) {
    R full[16];
    for (int i = 0; i < 6; ++i) full[i] = (R)x[4 + i];       /* x[n:]            */
    for (int i = 0; i < 10; ++i) full[6 + i] = (R)x[9 - i];  /* reverse(x)       */
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) m[4 * i + j] = (j <= i) ? full[4 * i + j] : (R)0;
}
void ORC(fill_triangular)(const float* x, R* m) { fill_triangular10(x, m); }

/* Outputs, one row per survivor s (ascending anchor index keep[s]):
 *   cnt_post [S,K]  dirichlit_posterior_count            (:89-94)
 *   mu_post  [S,4]  gaussian_posterior_means             (:131-145, 160)
 *   sig_post [S,16] gaussian_posterior_covs              (:129, 162-167)
 *   score    [S]    ranking_scores                       (:169-202)
 *   corners  [S,4]  vuhw_to_vuvu(posterior means)        (:204; box_utils.py:5-23) */
void ORC(posterior)(const orc_params* P, const float* box, const float* cov,
                    const float* anchors, const float* counts, const int32_t* keep, int S,
                    R* cnt_post, R* mu_post, R* sig_post, R* score, R* corners) {
    const int N = P->N, A = P->A, K = P->K;
    R* g_info = NULL; R* c_info = NULL;
    if (P->ranking_method == 1) { g_info = (R*)malloc(sizeof(R) * (size_t)(S + 1)); c_info = (R*)malloc(sizeof(R) * (size_t)(S + 1)); }
    R* bx = (R*)malloc(sizeof(R) * (size_t)N * 4);
    const R alpha = (R)1 / (R)K;                                       /* :91 */
    for (int s = 0; s < S; ++s) {
        const int a = keep[s];
        const R av = (R)anchors[4 * a + 0], au = (R)anchors[4 * a + 1];
        const R ah = (R)anchors[4 * a + 2], aw = (R)anchors[4 * a + 3];
        /* box_from_anchor_and_target_bnms, box_utils.py:179-187 (decode every sample) */
        for (int n = 0; n < N; ++n) {
            const float* t = box + ((size_t)n * A + a) * 4;
            bx[4 * n + 0] = ah * (R)t[0] / (R)10 + av;
            bx[4 * n + 1] = aw * (R)t[1] / (R)10 + au;
            bx[4 * n + 2] = ah * r_min(r_max(r_exp((R)t[2] / (R)5), (R)1e-4f), (R)1e4f);
            bx[4 * n + 3] = aw * r_min(r_max(r_exp((R)t[3] / (R)5), (R)1e-4f), (R)1e4f);
        }
        /* compute_mean_covariance_tf :233-242 */
        R mu[4], epi[16];
        for (int i = 0; i < 4; ++i) {
            R acc = (R)0;
            for (int n = 0; n < N; ++n) acc = acc + bx[4 * n + i];
            mu[i] = acc / (R)N;
        }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                R acc = (R)0;
                for (int n = 0; n < N; ++n) acc = acc + (bx[4 * n + i] - mu[i]) * (bx[4 * n + j] - mu[j]);
                epi[4 * i + j] = acc / ((R)N - (R)1);
            }
        /* aleatoric :62-84 */
        R al[16];
        for (int i = 0; i < 16; ++i) al[i] = (R)0;
        if (P->cov_layout != 0) {
            R abar[16], tmp[16];
            for (int i = 0; i < 16; ++i) abar[i] = (R)0;
            for (int n = 0; n < N; ++n) {                                 /* reduce_mean axis 0 :64-65 */
                if (P->cov_layout == 1) {
                    const float* c = cov + ((size_t)n * A + a) * 16;
                    for (int i = 0; i < 16; ++i) abar[i] = abar[i] + (R)c[i];
                } else {
                    fill_triangular10(cov + ((size_t)n * A + a) * 10, tmp);
                    for (int i = 0; i < 16; ++i) abar[i] = abar[i] + tmp[i];
                }
            }
            for (int i = 0; i < 16; ++i) abar[i] = abar[i] / (R)N;
            R Dm[4];
            for (int i = 0; i < 4; ++i) Dm[i] = r_exp(abar[5 * i]);     /* :70 */
            if (P->use_full_covar) {                                       /* :74-80 */
                R M[16], L[16], LD[16];
                memcpy(M, abar, sizeof M);
                for (int i = 0; i < 4; ++i) M[5 * i] = (R)1;
                inv4(M, L);
                for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) LD[4 * i + j] = L[4 * i + j] * Dm[j];
                mm4(LD, L, 1, al);
            } else {
                for (int i = 0; i < 4; ++i) al[5 * i] = Dm[i];           /* :71-73, 82 */
            }
        }
        /* mixing :86-87 */
        R lik[16];
        for (int i = 0; i < 16; ++i) lik[i] = ((R)10 * al[i] + (R)1 * epi[i]) / (R)11;
        /* dirichlet :89-97 */
        const float* cn = counts + (size_t)a * K;
        R* cp = cnt_post + (size_t)s * K;
        R csum = (R)0;
        for (int k = 0; k < K; ++k) {
            cp[k] = (P->dirichlet_prior == 1) ? (R)cn[k] + alpha : (R)cn[k];
            csum = csum + cp[k];
        }
        /* gaussian prior :99-145 */
        R mp[4], sp[16];
        if (P->gaussian_prior == 1) {
            R prec_l[16], prec_post[16], w_l[4], inter[4];
            inv4(lik, prec_l);                                             /* :101 */
            const R prec_p = (R)1 / (R)P->isotropic_variance;              /* inv(diag(var)) :120 */
            memcpy(prec_post, prec_l, sizeof prec_post);
            for (int i = 0; i < 4; ++i) prec_post[5 * i] = prec_l[5 * i] + prec_p;   /* :127 */
            inv4(prec_post, sp);                                           /* :129 */
            const R an[4] = {av, au, ah, aw};                              /* prior mean = anchor :122 */
            mv4(prec_l, mu, w_l);                                          /* :137 */
            for (int i = 0; i < 4; ++i) inter[i] = prec_p * an[i] + w_l[i];  /* :132,140 */
            mv4(sp, inter, mp);                                            /* :141 */
        } else {
            memcpy(mp, mu, sizeof mp); memcpy(sp, lik, sizeof sp);         /* :144-145 */
        }
        /* kitti rescale :147-167 (identity when the scales are 1) */
        {
            const R sc[4] = {(R)P->scale_v, (R)P->scale_u, (R)P->scale_v, (R)P->scale_u};
            for (int i = 0; i < 4; ++i) mp[i] = sc[i] * mp[i];
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) sp[4 * i + j] = (sc[i] * sp[4 * i + j]) * sc[j];
        }
        for (int i = 0; i < 4; ++i) mu_post[4 * (size_t)s + i] = mp[i];
        for (int i = 0; i < 16; ++i) sig_post[16 * (size_t)s + i] = sp[i];
        /* ranking :169-202 */
        if (P->ranking_method == 1) {
            /* compute_gaussian_entropy_tf :247-263, compute_categorical_entropy_tf :266-277 */
            const R two_pi_log = r_log((R)(2.0 * 3.141592653589793));
            R hp = (R)2 + (R)2 * two_pi_log + (R)0.5 * r_log(det4(sp));
            R var4[16];
            for (int i = 0; i < 16; ++i) var4[i] = (R)0;
            for (int i = 0; i < 4; ++i) var4[5 * i] = (R)P->isotropic_variance;
            R hprior = (R)2 + (R)2 * two_pi_log + (R)0.5 * r_log(det4(var4));
            g_info[s] = hprior - hp;
            R ent = (R)0, ent0 = (R)0, psum = (R)0;
            for (int k = 0; k < K; ++k) { R p = cp[k] / csum; ent = ent + p * r_log(p); }
            for (int k = 0; k < K; ++k) psum = psum + alpha;
            for (int k = 0; k < K; ++k) { R p = alpha / psum; ent0 = ent0 + p * r_log(p); }
            c_info[s] = (-ent0) - (-ent);
        } else {
            R best = cp[0] / csum;
            for (int k = 1; k < K; ++k) best = r_max(best, cp[k] / csum);
            score[s] = best;
        }
        /* vuhw_to_vuvu box_utils.py:13-21 */
        corners[4 * (size_t)s + 0] = mp[0] - mp[2] / (R)2;
        corners[4 * (size_t)s + 1] = mp[1] - mp[3] / (R)2;
        corners[4 * (size_t)s + 2] = mp[0] + mp[2] / (R)2;
        corners[4 * (size_t)s + 3] = mp[1] + mp[3] / (R)2;
    }
    if (P->ranking_method == 1 && S > 0) {                                 /* :177-200 */
        R gmin = g_info[0], gmax = g_info[0], cmin = c_info[0], cmax = c_info[0];
        for (int s = 1; s < S; ++s) {
            gmin = r_min(gmin, g_info[s]); gmax = r_max(gmax, g_info[s]);
            cmin = r_min(cmin, c_info[s]); cmax = r_max(cmax, c_info[s]);
        }
        for (int s = 0; s < S; ++s) {
            R g = (g_info[s] - gmin) / r_max((R)1, gmax - gmin);
            R c = (c_info[s] - cmin) / r_max((R)0.001f, cmax - cmin);
            score[s] = c + g;
        }
    }
    free(bx); free(g_info); free(c_info);
}

/* ------------------------------------------------------------------------- */
/* H9: NonMaxSuppressionV5 (tf.image.non_max_suppression_with_scores)          */
/* ------------------------------------------------------------------------- */
/* TF kernel IoU: canonicalise corners, area<=0 -> 0, inter / (a_i + a_j - inter) */
static R tf_iou(const R* bi, const R* bj) {
    const R ymin_i = r_min(bi[0], bi[2]), xmin_i = r_min(bi[1], bi[3]);
    const R ymax_i = r_max(bi[0], bi[2]), xmax_i = r_max(bi[1], bi[3]);
    const R ymin_j = r_min(bj[0], bj[2]), xmin_j = r_min(bj[1], bj[3]);
    const R ymax_j = r_max(bj[0], bj[2]), xmax_j = r_max(bj[1], bj[3]);
    const R area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    const R area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= (R)0 || area_j <= (R)0) return (R)0;
    const R iymin = r_max(ymin_i, ymin_j), ixmin = r_max(xmin_i, xmin_j);
    const R iymax = r_min(ymax_i, ymax_j), ixmax = r_min(xmax_i, xmax_j);
    const R inter = r_max(iymax - iymin, (R)0) * r_max(ixmax - ixmin, (R)0);
    return inter / (area_i + area_j - inter);
}
R ORC(tf_iou)(const R* bi, const R* bj) { return tf_iou(bi, bj); }

typedef struct { int32_t box; R score; int32_t begin; } cand_t;
/* std::priority_queue "less": a below b <=> a.score < b.score, or equal scores and a.box > b.box */
static int cand_below(const cand_t* a, const cand_t* b) {
    return (a->score == b->score) ? (a->box > b->box) : (a->score < b->score);
}
static void heap_push(cand_t* h, int* n, cand_t c) {
    int i = (*n)++;
    h[i] = c;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!cand_below(&h[p], &h[i])) break;
        cand_t t = h[p]; h[p] = h[i]; h[i] = t; i = p;
    }
}
static cand_t heap_pop(cand_t* h, int* n) {
    cand_t top = h[0];
    h[0] = h[--(*n)];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < *n && cand_below(&h[m], &h[l])) m = l;
        if (r < *n && cand_below(&h[m], &h[r])) m = r;
        if (m == i) break;
        cand_t t = h[m]; h[m] = h[i]; h[i] = t; i = m;
    }
    return top;
}

/* Returns D; selected[d] = survivor index, sel_scores[d] = score at selection.
 * `stats` (optional, 2 ints) receives {pops, iou evaluations}. */
int ORC(nms_v5)(const R* boxes, const R* scores, int S, int max_output_size,
                float iou_threshold, float score_threshold, float soft_nms_sigma,
                int32_t* selected, R* sel_scores, int64_t* stats) {
    cand_t* heap = (cand_t*)malloc(sizeof(cand_t) * (size_t)(S > 0 ? S : 1));
    int hn = 0, D = 0;
    int64_t pops = 0, evals = 0;
    for (int i = 0; i < S; ++i)
        if (scores[i] > (R)score_threshold) { cand_t c = {i, scores[i], 0}; heap_push(heap, &hn, c); }
    const int is_soft = soft_nms_sigma > 0.0f;
    const R scale = is_soft ? (R)-0.5f / (R)soft_nms_sigma : (R)0;
    while (D < max_output_size && hn > 0) {
        cand_t c = heap_pop(heap, &hn);
        const R original = c.score;
        int hard = 0;
        ++pops;
        for (int j = D - 1; j >= c.begin; --j) {
            const R sim = tf_iou(boxes + 4 * (size_t)c.box, boxes + 4 * (size_t)selected[j]);
            ++evals;
            R w = r_exp(scale * sim * sim);
            if (!(is_soft || sim <= (R)iou_threshold)) w = (R)0;
            c.score = c.score * w;
            if (!is_soft && sim > (R)iou_threshold) { hard = 1; break; }
            if (c.score <= (R)score_threshold) break;
        }
        c.begin = D;
        if (!hard) {
            if (c.score == original) { selected[D] = c.box; sel_scores[D] = c.score; ++D; continue; }
            if (c.score > (R)score_threshold) heap_push(heap, &hn, c);
        }
    }
    free(heap);
    if (stats) { stats[0] = pops; stats[1] = evals; }
    return D;
}

/* ------------------------------------------------------------------------- */
/* H10: box_utils.bbox_iou_vuvu (box_utils.py:117-146), element (i, j)         */
/* ------------------------------------------------------------------------- */
static R repo_iou(const R* b1, const R* b2) {
    const R y11 = b1[0], x11 = b1[1], y12 = b1[2], x12 = b1[3];
    const R y21 = b2[0], x21 = b2[1], y22 = b2[2], x22 = b2[3];
    const R xI1 = r_max(x11, x21), yI1 = r_max(y11, y21);
    const R xI2 = r_min(x12, x22), yI2 = r_min(y12, y22);
    const R inter = r_max((xI2 - xI1) + (R)1, (R)0) * r_max((yI2 - yI1) + (R)1, (R)0);
    const R a1 = ((x11 - x12) + (R)1) * ((y11 - y12) + (R)1);   /* sic: min - max (:140) */
    const R a2 = ((x21 - x22) + (R)1) * ((y21 - y22) + (R)1);
    const R uni = (a1 + a2) - inter;
    return inter / (uni + (R)0.00001f);
}
R ORC(repo_iou)(const R* b1, const R* b2) { return repo_iou(b1, b2); }

/* full [S,S] matrix (small cases only) */
void ORC(iou_matrix)(const R* corners, int S, R* out) {
    for (int i = 0; i < S; ++i)
        for (int j = 0; j < S; ++j) out[(size_t)i * S + j] = repo_iou(corners + 4 * (size_t)i, corners + 4 * (size_t)j);
}

/* membership bitmask: bit s of row d <=> affinity[s, centre_d] > thr (:316) */
void ORC(membership)(const R* corners, int S, const int32_t* centres, int D, float thr,
                     uint32_t* mask, int words_per_row) {
    for (int d = 0; d < D; ++d) {
        uint32_t* row = mask + (size_t)d * words_per_row;
        memset(row, 0, sizeof(uint32_t) * (size_t)words_per_row);
        for (int s = 0; s < S; ++s)
            if (repo_iou(corners + 4 * (size_t)s, corners + 4 * (size_t)centres[d]) > (R)thr)
                row[s >> 5] |= 1u << (s & 31);
    }
}

/* ------------------------------------------------------------------------- */
/* H12: bayes_od_clustering (inference_utils.py:285-364)                       */
/* ------------------------------------------------------------------------- */
/* KL(center || member) as scipy.stats.entropy(pk, qk) computes it: both
 * arguments re-normalised, rel_entr summed over classes. */
static R kl_div(const R* pk_raw, const R* qk_raw, int K) {
    R sp = (R)0, sq = (R)0, acc = (R)0;
    for (int k = 0; k < K; ++k) { sp = sp + pk_raw[k]; sq = sq + qk_raw[k]; }
    for (int k = 0; k < K; ++k) {
        const R p = pk_raw[k] / sp, q = qk_raw[k] / sq;
        R t;
        if (p > (R)0 && q > (R)0) t = p * r_log(p / q);
        else if (p == (R)0 && q >= (R)0) t = (R)0;
        else t = (R)INFINITY;
        acc = acc + t;
    }
    return acc;
}

/* Outputs are padded by the caller; row d <-> centres[d].
 * Membership is taken from `mask` so the binary64 twin can reuse the binary32
 * clusters.  Top-3 selection (:338-349): smallest KL, ties -> lowest survivor
 * index (np.argpartition's tie order is implementation-defined).
 * Returns the number of clusters with an empty membership (reference would
 * raise on those). */
int ORC(clustering)(const R* cnt, const R* mu, const R* sig, int S, int K,
                    const int32_t* centres, int D, const uint32_t* mask, int words_per_row,
                    float calibration,
                    R* out_scores, R* out_means, R* out_covs, R* out_counts, int32_t* out_members,
                    int32_t* out_chosen /* [D,3] picks of the top-3-KL rule, -1 = rule not applied */) {
    int empty = 0;
    R* sc = (R*)malloc(sizeof(R) * (size_t)K);
    R* cs = (R*)malloc(sizeof(R) * (size_t)K);
    for (int d = 0; d < D; ++d) {
        const uint32_t* row = mask + (size_t)d * words_per_row;
        R psum[16], wsum[4];
        for (int i = 0; i < 16; ++i) psum[i] = (R)0;
        for (int i = 0; i < 4; ++i) wsum[i] = (R)0;
        int m = 0;
        for (int s = 0; s < S; ++s) {
            if (!((row[s >> 5] >> (s & 31)) & 1u)) continue;
            R prec[16], w[4];
            inv4(sig + 16 * (size_t)s, prec);                      /* :321-322 */
            mv4(prec, mu + 4 * (size_t)s, w);                      /* :327-329 */
            for (int i = 0; i < 16; ++i) psum[i] = psum[i] + prec[i];   /* np.sum axis 0 :324 */
            for (int i = 0; i < 4; ++i) wsum[i] = wsum[i] + w[i];       /* :330 */
            ++m;
        }
        if (out_members) out_members[d] = m;
        if (out_chosen) { out_chosen[3 * d] = out_chosen[3 * d + 1] = out_chosen[3 * d + 2] = -1; }
        R* oc = out_covs + 16 * (size_t)d; R* om = out_means + 4 * (size_t)d;
        R* os = out_scores + (size_t)d * K; R* on = out_counts + (size_t)d * K;
        if (m == 0) {
            ++empty;
            for (int i = 0; i < 16; ++i) oc[i] = (R)NAN;
            for (int i = 0; i < 4; ++i) om[i] = (R)NAN;
            for (int k = 0; k < K; ++k) { os[k] = (R)NAN; on[k] = (R)0; }
            continue;
        }
        R fc[16];
        inv4(psum, fc);                                             /* :324 */
        mv4(fc, wsum, om);                                          /* :331 */
        for (int i = 0; i < 16; ++i) oc[i] = fc[i] * (R)calibration;   /* :361 */
        /* categorical part :334-354 */
        int chosen[3] = {-1, -1, -1}; int nchosen = 0;
        if (m > 3) {
            const R* cc = cnt + (size_t)centres[d] * K;
            R ccs = (R)0;
            for (int k = 0; k < K; ++k) ccs = ccs + cc[k];
            for (int k = 0; k < K; ++k) cs[k] = cc[k] / ccs;        /* :339-340 */
            R best[3] = {(R)0, (R)0, (R)0};
            for (int s = 0; s < S; ++s) {
                if (!((row[s >> 5] >> (s & 31)) & 1u)) continue;
                const R* c = cnt + (size_t)s * K;
                R su = (R)0;
                for (int k = 0; k < K; ++k) su = su + c[k];
                for (int k = 0; k < K; ++k) sc[k] = c[k] / su;      /* :335-336 */
                const R kl = kl_div(cs, sc, K);                     /* :344 */
                /* insert into the sorted top-3 (strict <: earlier index wins ties) */
                int pos = nchosen;
                while (pos > 0 && kl < best[pos - 1]) --pos;
                if (pos < 3) {
                    for (int q = (nchosen < 3 ? nchosen : 2); q > pos; --q) { best[q] = best[q - 1]; chosen[q] = chosen[q - 1]; }
                    best[pos] = kl; chosen[pos] = s;
                    if (nchosen < 3) ++nchosen;
                }
            }
            /* accumulate the three in ascending survivor order */
            for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j)
                if (chosen[j] < chosen[i]) { int t = chosen[i]; chosen[i] = chosen[j]; chosen[j] = t; }
            if (out_chosen) for (int i = 0; i < 3; ++i) out_chosen[3 * d + i] = chosen[i];
        }
        for (int k = 0; k < K; ++k) { os[k] = (R)0; on[k] = (R)0; }
        int used = 0;
        for (int s = 0; s < S; ++s) {
            if (!((row[s >> 5] >> (s & 31)) & 1u)) continue;
            if (m > 3 && s != chosen[0] && s != chosen[1] && s != chosen[2]) continue;
            const R* c = cnt + (size_t)s * K;
            R su = (R)0;
            for (int k = 0; k < K; ++k) su = su + c[k];
            for (int k = 0; k < K; ++k) { os[k] = os[k] + c[k] / su; on[k] = on[k] + c[k]; }   /* :351-352 */
            ++used;
        }
        for (int k = 0; k < K; ++k) os[k] = os[k] / (R)used;        /* np.mean :351 */
    }
    free(sc); free(cs);
    return empty;
}

#ifndef ORC_REAL_IS_DOUBLE
/* ------------------------------------------------------------------------- */
/* EXTENSIONS: score threshold + pre-NMS top-k on the keep list */
/* ------------------------------------------------------------------------- */
/* Validation post-process: validation_utils.post_process_predictions          */
/* (src/retina_net/experiments/validation_utils.py:10-77), one image, N = 1.   */
/*   :22-26   box_from_anchor_and_target (box_utils.py:149-168) for EVERY anchor */
/*   :28      vuhw_to_vuvu (box_utils.py:5-23)                                   */
/*   :29-30   softmax over the K logits                                          */
/*   :34-43   argmax != K-1 (first maximum) + boolean_mask (ascending order)     */
/*   :45      top score = max probability                                        */
/*   :47-52   the same NonMaxSuppressionV5 call (100, 0.5, sigma 0.5)            */
/*   :54-66   kitti: (corners / [h,w,h,w]) * [H0,W0,H0,W0]; coco: shift first    */
/*   :68-73   gather classes and (scaled) corners of the selected boxes          */
/* scale_mode: 0 none (bdd), 1 kitti, 2 coco.  Outputs: keep [S] anchor indices, */
/* probs [S,K], corners [S,4] (unscaled, what NMS sees), scores [S], then        */
/* sel [D], sel_scores [D], out_classes [D,K], out_corners [D,4].  Returns S;    */
/* *D_out receives D.  Buffers are sized for A survivors / max_output_size rows. */
/* ------------------------------------------------------------------------- */
typedef struct orc_val_scaling {
    int32_t mode;            /* 0 none, 1 kitti, 2 coco */
    float shift[4];          /* coco: IMAGE_PADDING_KEY[0], subtracted from the corners */
    float norm_h, norm_w;    /* normalize_2d_bounding_boxes (box_utils.py:195-205) */
    float scale_h, scale_w;  /* expand_2d_bounding_boxes (box_utils.py:208-220)    */
} orc_val_scaling;

int orc_val_postprocess(const float* cls, const float* box, const float* anchors, int A, int K,
                        int max_output_size, float iou_threshold, float soft_nms_sigma,
                        const orc_val_scaling* sc,
                        int32_t* keep, float* probs, float* corners, float* scores,
                        int32_t* sel, float* sel_scores, float* out_classes, float* out_corners, int32_t* D_out) {
    int S = 0;
    float* e = (float*)malloc(sizeof(float) * (size_t)K);
    for (int a = 0; a < A; ++a) {
        const float* x = cls + (size_t)a * K;
        float m = x[0];
        for (int k = 1; k < K; ++k) m = r_max(m, x[k]);
        float sum = 0.0f;
        for (int k = 0; k < K; ++k) { e[k] = r_exp(x[k] - m); sum = sum + e[k]; }
        int am = 0;
        float best = e[0] / sum;
        for (int k = 0; k < K; ++k) { e[k] = e[k] / sum; if (e[k] > best) { best = e[k]; am = k; } }
        if (am == K - 1) continue;
        keep[S] = a;
        for (int k = 0; k < K; ++k) probs[(size_t)S * K + k] = e[k];
        scores[S] = best;
        const float* an = anchors + (size_t)a * 4;
        const float* t = box + (size_t)a * 4;
        const float v = an[2] * t[0] / 10.0f + an[0];
        const float u = an[3] * t[1] / 10.0f + an[1];
        const float h = an[2] * r_min(r_max(r_exp(t[2] / 5.0f), 1e-4f), 1e4f);
        const float w = an[3] * r_min(r_max(r_exp(t[3] / 5.0f), 1e-4f), 1e4f);
        float* c = corners + (size_t)S * 4;
        c[0] = v - h / 2.0f; c[1] = u - w / 2.0f; c[2] = v + h / 2.0f; c[3] = u + w / 2.0f;
        ++S;
    }
    free(e);
    for (int d = 0; d < max_output_size; ++d) sel[d] = -1;
    const int D = orc32_nms_v5(corners, scores, S, max_output_size, iou_threshold, -INFINITY, soft_nms_sigma, sel, sel_scores, NULL);
    for (int d = 0; d < D; ++d) {
        const int s = sel[d];
        for (int k = 0; k < K; ++k) out_classes[(size_t)d * K + k] = probs[(size_t)s * K + k];
        for (int i = 0; i < 4; ++i) {
            float c = corners[(size_t)s * 4 + i];
            if (sc && sc->mode == 2) c = c - sc->shift[i];
            if (sc && sc->mode != 0) {
                c = c / ((i & 1) ? sc->norm_w : sc->norm_h);
                c = c * ((i & 1) ? sc->scale_w : sc->scale_h);
            }
            out_corners[(size_t)d * 4 + i] = c;
        }
    }
    *D_out = D;
    return S;
}

/* ------------------------------------------------------------------------- */
/* Ranking score from counts alone ('score' ranking): max_k (c_k+alpha)/sum. */
static float count_score(const float* c, int K, int dirichlet) {
    const float alpha = 1.0f / (float)K;
    float sum = 0.0f, best = 0.0f;
    for (int k = 0; k < K; ++k) sum = sum + (dirichlet ? c[k] + alpha : c[k]);
    for (int k = 0; k < K; ++k) {
        float p = (dirichlet ? c[k] + alpha : c[k]) / sum;
        if (k == 0 || p > best) best = p;
    }
    return best;
}
typedef struct { float score; int32_t idx; } sk_t;
static int sk_cmp(const void* a, const void* b) {
    const sk_t* x = (const sk_t*)a; const sk_t* y = (const sk_t*)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
static int i32_cmp(const void* a, const void* b) { int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return x < y ? -1 : x > y; }
/* (1) drop survivors with score <= thr, (2) keep the top_k by (score desc, anchor asc),
 * (3) restore ascending anchor order.  In place; returns the new S. */
int orc_prefilter(const float* counts, int K, int dirichlet, float score_threshold, int top_k,
                  int32_t* keep, int S) {
    sk_t* v = (sk_t*)malloc(sizeof(sk_t) * (size_t)(S > 0 ? S : 1));
    int n = 0;
    for (int s = 0; s < S; ++s) {
        float sc = count_score(counts + (size_t)keep[s] * K, K, dirichlet);
        if (sc > score_threshold) { v[n].score = sc; v[n].idx = keep[s]; ++n; }
    }
    if (top_k > 0 && n > top_k) { qsort(v, (size_t)n, sizeof(sk_t), sk_cmp); n = top_k; }
    for (int i = 0; i < n; ++i) keep[i] = v[i].idx;
    qsort(keep, (size_t)n, sizeof(int32_t), i32_cmp);
    free(v);
    return n;
}

/* ------------------------------------------------------------------------- */
/* Whole path for a batch (CPU baseline; pthreads over images)                 */
/* ------------------------------------------------------------------------- */
typedef struct orc_run_params {
    orc_params p;
    int32_t B;
    int32_t max_output_size; float iou_threshold, soft_nms_sigma;
    float cov_calibration;
    int32_t num_draws; uint64_t seed; uint32_t image_id_base;
    float score_threshold; int32_t pre_nms_top_k;
} orc_run_params;

/* Outputs padded to Dmax = max_output_size per image (same blocks as
 * bod_host_results). counts may be NULL -> Philox sampler. Returns 0. */
typedef struct orc_job {
    const orc_run_params* rp;
    const float *cls, *box, *cov, *anchors, *counts_in;
    int32_t *num_dets, *num_survivors, *nms_indices, *centre_anchor_idx;
    float *means, *covs, *cat_param, *cat_count;
    volatile int32_t* next;   /* shared image ticket */
} orc_job;

static void orc_run_image(const orc_job* J, int b) {
    const orc_run_params* rp = J->rp;
    const orc_params* P = &rp->p;
    const int N = P->N, A = P->A, K = P->K, Dmax = rp->max_output_size;
    const int cw = (P->cov_layout == 1) ? 16 : (P->cov_layout == 2 ? 10 : 0);
    const float* cls_b = J->cls + (size_t)b * N * A * K;
    const float* box_b = J->box + (size_t)b * N * A * 4;
    const float* cov_b = J->cov ? J->cov + (size_t)b * N * A * cw : NULL;
    float* counts = NULL;
    const float* cnt_b;
    if (J->counts_in) cnt_b = J->counts_in + (size_t)b * A * K;
    else {
        float* probs = (float*)malloc(sizeof(float) * (size_t)A * K);
        counts = (float*)malloc(sizeof(float) * (size_t)A * K);
        orc32_softmax_mean(cls_b, N, A, K, probs);
        orc_philox_counts(probs, A, K, rp->num_draws, rp->seed, rp->image_id_base + (uint32_t)b, counts);
        free(probs);
        cnt_b = counts;
    }
    int32_t* keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)A);
    int S = orc_filter(cnt_b, A, K, keep);
    if (rp->pre_nms_top_k > 0 || rp->score_threshold > -INFINITY)
        S = orc_prefilter(cnt_b, K, P->dirichlet_prior == 1, rp->score_threshold, rp->pre_nms_top_k, keep, S);
    J->num_survivors[b] = S;
    const int S1 = S > 0 ? S : 1;
    float* cp = (float*)malloc(sizeof(float) * (size_t)S1 * K);
    float* mu = (float*)malloc(sizeof(float) * (size_t)S1 * 4);
    float* sg = (float*)malloc(sizeof(float) * (size_t)S1 * 16);
    float* sc = (float*)malloc(sizeof(float) * (size_t)S1);
    float* co = (float*)malloc(sizeof(float) * (size_t)S1 * 4);
    orc32_posterior(P, box_b, cov_b, J->anchors, cnt_b, keep, S, cp, mu, sg, sc, co);
    int32_t* sel = J->nms_indices + (size_t)b * Dmax;
    float* ss = (float*)malloc(sizeof(float) * (size_t)(Dmax > 0 ? Dmax : 1));
    for (int d = 0; d < Dmax; ++d) { sel[d] = -1; J->centre_anchor_idx[(size_t)b * Dmax + d] = -1; }
    int D = orc32_nms_v5(co, sc, S, Dmax, rp->iou_threshold, -INFINITY, rp->soft_nms_sigma, sel, ss, NULL);
    J->num_dets[b] = D;
    const int wpr = (S1 + 31) / 32;
    uint32_t* mask = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)wpr * (size_t)(D > 0 ? D : 1));
    orc32_membership(co, S, sel, D, rp->iou_threshold, mask, wpr);
    float* om = J->means + (size_t)b * Dmax * 4;
    float* oc = J->covs + (size_t)b * Dmax * 16;
    float* op = J->cat_param + (size_t)b * Dmax * K;
    float* on = J->cat_count + (size_t)b * Dmax * K;
    memset(om, 0, sizeof(float) * (size_t)Dmax * 4);
    memset(oc, 0, sizeof(float) * (size_t)Dmax * 16);
    memset(op, 0, sizeof(float) * (size_t)Dmax * K);
    memset(on, 0, sizeof(float) * (size_t)Dmax * K);
    orc32_clustering(cp, mu, sg, S, K, sel, D, mask, wpr, rp->cov_calibration, op, om, oc, on, NULL, NULL);
    for (int d = 0; d < D; ++d) J->centre_anchor_idx[(size_t)b * Dmax + d] = keep[sel[d]];
    free(mask); free(ss); free(cp); free(mu); free(sg); free(sc); free(co); free(keep); free(counts);
}

static void* orc_worker(void* arg) {
    const orc_job* J = (const orc_job*)arg;
    for (;;) {
        int b = __sync_fetch_and_add(J->next, 1);
        if (b >= J->rp->B) break;
        orc_run_image(J, b);
    }
    return NULL;
}

/* Outputs padded to Dmax = max_output_size per image (same blocks as
 * bod_host_results). counts_in may be NULL -> Philox sampler.  Images are
 * independent; `nthreads` host threads pull them from a shared ticket. */
int orc_run_batch(const orc_run_params* rp, const float* cls, const float* box, const float* cov,
                  const float* anchors, const float* counts_in,
                  int32_t* num_dets, int32_t* num_survivors, float* means, float* covs,
                  float* cat_param, float* cat_count, int32_t* nms_indices, int32_t* centre_anchor_idx,
                  int nthreads) {
    int32_t next = 0;
    orc_job J = {rp, cls, box, cov, anchors, counts_in, num_dets, num_survivors, nms_indices,
                 centre_anchor_idx, means, covs, cat_param, cat_count, &next};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > rp->B) nthreads = rp->B;
    if (nthreads <= 1) { orc_worker(&J); return 0; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, orc_worker, &J);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
    return 0;
}
#endif /* !ORC_REAL_IS_DOUBLE */
