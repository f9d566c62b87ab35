// bod_common.cuh — shared device helpers of the BayesOD sm_100a kernels.
//
// Arithmetic contract (DESIGN.md §"Arithmetic contract"): everything that decides
// an index (kept anchors, soft-NMS order, cluster membership) or feeds a fused
// output is IEEE binary32, round-to-nearest, NO FMA contraction, reductions
// sequential in index order, exp/log correctly rounded via binary64.  The
// translation units that include this header for such work are compiled with
// -fmad=false; the only place where FMA/approximate math is allowed is the
// softmax of the moments kernel (its output is a tolerance-checked probability).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BOD_DEVINL __device__ __forceinline__

namespace bod {

constexpr int kTileAnchors = 128;   // anchors per moments-kernel tile (and per compaction tile)
constexpr int kMaxK = 64;           // classes + background supported
constexpr int kMaxOut = 256;        // max_output_size supported (selected-mask words = 8)
constexpr int kMaskWords = kMaxOut / 32;
constexpr int kPendStride = 8;      // 32-bit words of soft-NMS scratch per candidate (generic kernel bitmask)

// ---------------------------------------------------------------------------
// correctly rounded binary32 exp / log (via binary64; 1 ulp of binary64 error
// leaves the binary32 rounding unchanged except with probability ~2^-28)
// ---------------------------------------------------------------------------
// ---- diagnostic builds (-DBOD_DIAGNOSTICS): a kernel's place on the device timeline ----
// tl[0] = start of CTA (0,0), tl[1] = start of the CTA placed last, tl[2] = end of the CTA that ended last
// (%globaltimer, ns).  The guard object stamps the end wherever the kernel returns.
#ifdef BOD_DIAGNOSTICS
struct TimelineStamp {
    unsigned long long* tl;
    static __device__ __forceinline__ unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
    __device__ __forceinline__ explicit TimelineStamp(unsigned long long* p) : tl(p) {
        if (tl && threadIdx.x == 0) {
            const unsigned long long t = now();
            if (blockIdx.x == 0 && blockIdx.y == 0) tl[0] = t;
            atomicMax(tl + 1, t);
        }
    }
    __device__ __forceinline__ ~TimelineStamp() { if (tl && threadIdx.x == 0) atomicMax(tl + 2, now()); }
};
#define BOD_TIMELINE(ptr) TimelineStamp timeline_stamp_(ptr)
#else
#define BOD_TIMELINE(ptr)
#endif

BOD_DEVINL float exp_cr(float x) { return (float)exp((double)x); }
BOD_DEVINL float log_cr(float x) { return (float)log((double)x); }

// ---------------------------------------------------------------------------
// 4x4 linear algebra in registers, same operation order as the arithmetic
// contract: right-looking LU, partial pivoting (first maximum), reciprocal
// scaling, then forward/back substitution on the row-permuted identity.
// inference_utils.py:75,101,120,129 (tf.linalg.inv) and :321-324 (np.linalg.inv).
// ---------------------------------------------------------------------------
struct Mat4 { float m[4][4]; };

BOD_DEVINL void lu4(float (&a)[4][4], float (&b)[4][4], int& sign) {
    sign = 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int p = j;
        float mx = fabsf(a[j][j]);
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            const float v = fabsf(a[i][j]);
            if (v > mx) { mx = v; p = i; }
        }
        const bool nz = (mx != 0.0f);
        if (!nz) sign = 0;
        // row interchange j <-> p on A and on the right-hand side
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            const bool sw = nz && (p == i);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float ta = a[j][k], tb = b[j][k];
                a[j][k] = sw ? a[i][k] : ta;  a[i][k] = sw ? ta : a[i][k];
                b[j][k] = sw ? b[i][k] : tb;  b[i][k] = sw ? tb : b[i][k];
            }
            if (sw) sign = -sign;
        }
        if (j < 3 && nz) {
            const float r = 1.0f / a[j][j];
#pragma unroll
            for (int i = j + 1; i < 4; ++i) a[i][j] = a[i][j] * r;
#pragma unroll
            for (int i = j + 1; i < 4; ++i)
#pragma unroll
                for (int k = j + 1; k < 4; ++k) a[i][k] = a[i][k] - a[i][j] * a[j][k];
        }
    }
}

// out = inverse(in)
BOD_DEVINL void inv4(const float (&in)[4][4], float (&out)[4][4]) {
    float a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[i][j] = in[i][j]; out[i][j] = (i == j) ? 1.0f : 0.0f; }
    int sign;
    lu4(a, out, sign);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int i = k + 1; i < 4; ++i) out[i][c] = out[i][c] - out[k][c] * a[i][k];
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            out[k][c] = out[k][c] / a[k][k];
#pragma unroll
            for (int i = 0; i < k; ++i) out[i][c] = out[i][c] - out[k][c] * a[i][k];
        }
    }
}

// determinant through the same LU (tf.linalg.det, inference_utils.py:258)
BOD_DEVINL float det4(const float (&in)[4][4]) {
    float a[4][4], dummy[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[i][j] = in[i][j]; dummy[i][j] = 0.0f; }
    int sign;
    lu4(a, dummy, sign);
    float d = (float)sign;
#pragma unroll
    for (int i = 0; i < 4; ++i) d = d * a[i][i];
    return d;
}

// y = A x, dot products sequential in k
BOD_DEVINL void mv4(const float (&A)[4][4], const float (&x)[4], float (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s = s + A[i][k] * x[k];
        y[i] = s;
    }
}

// ---------------------------------------------------------------------------
// box geometry
// ---------------------------------------------------------------------------
// TF NonMaxSuppression kernel IoU (non_max_suppression_op.cc IOU<T>): corners
// canonicalised, non-positive area -> 0, no +1, no epsilon.
BOD_DEVINL float tf_iou(const float4 bi, const float4 bj) {
    const float ymin_i = fminf(bi.x, bi.z), xmin_i = fminf(bi.y, bi.w);
    const float ymax_i = fmaxf(bi.x, bi.z), xmax_i = fmaxf(bi.y, bi.w);
    const float ymin_j = fminf(bj.x, bj.z), xmin_j = fminf(bj.y, bj.w);
    const float ymax_j = fmaxf(bj.x, bj.z), xmax_j = fmaxf(bj.y, bj.w);
    const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0.0f || area_j <= 0.0f) return 0.0f;
    const float iymin = fmaxf(ymin_i, ymin_j), ixmin = fmaxf(xmin_i, xmin_j);
    const float iymax = fminf(ymax_i, ymax_j), ixmax = fminf(xmax_i, xmax_j);
    const float inter = fmaxf(iymax - iymin, 0.0f) * fmaxf(ixmax - ixmin, 0.0f);
    if (inter == 0.0f) return 0.0f;              // 0 / (positive) = +0: skip the division
    return inter / (area_i + area_j - inter);
}

// box_utils.bbox_iou_vuvu element (box_utils.py:132-146), including its
// (min - max + 1) area terms and the 1e-5 epsilon.
BOD_DEVINL float repo_iou(const float4 b1, const float4 b2) {
    const float xI1 = fmaxf(b1.y, b2.y), yI1 = fmaxf(b1.x, b2.x);
    const float xI2 = fminf(b1.w, b2.w), yI2 = fminf(b1.z, b2.z);
    const float inter = fmaxf((xI2 - xI1) + 1.0f, 0.0f) * fmaxf((yI2 - yI1) + 1.0f, 0.0f);
    const float a1 = ((b1.y - b1.w) + 1.0f) * ((b1.x - b1.z) + 1.0f);
    const float a2 = ((b2.y - b2.w) + 1.0f) * ((b2.x - b2.z) + 1.0f);
    const float uni = (a1 + a2) - inter;
    return inter / (uni + 0.00001f);
}

// bbox_iou_vuvu(survivor, centre) > threshold (strict; inference_utils.py:316), skipping the division when
// the boxes cannot overlap even with the +1 pixel convention ((hi - lo) + 1 > 0 <=> hi - lo > -1 in binary32;
// then inter = 0 and the quotient is +-0 or NaN, never > a threshold >= 0)
BOD_DEVINL bool is_member(const float4 bs, const float4 bx, float thr) {
    const float dx = fminf(bs.w, bx.w) - fmaxf(bs.y, bx.y);
    const float dy = fminf(bs.z, bx.z) - fmaxf(bs.x, bx.x);
    const bool wellformed = (bx.x <= bx.z) && (bx.y <= bx.w) && (bs.x <= bs.z) && (bs.y <= bs.w);
    if (!wellformed || (dx > -1.0f && dy > -1.0f)) return repo_iou(bs, bx) > thr;
    return false;
}

// order-preserving float -> uint32 key (larger float <=> larger key; -inf > 0)
BOD_DEVINL uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11); counter (anchor, image, call, 0x0B0D)
// ---------------------------------------------------------------------------
BOD_DEVINL uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        uint4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k.x;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k.y;
        n.w = (uint32_t)p0;
        c = n;
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

}  // namespace bod
