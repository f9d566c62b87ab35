// kv_validation.cu — the validation post-process on the GPU.  Compiled with -fmad=false.
//
// Reference lines replaced: validation_utils.post_process_predictions
// (src/retina_net/experiments/validation_utils.py:10-77), the deterministic,
// single-sample sibling of bayes_od_inference that run_validation.py:143-144 calls:
//   :22-28   box_from_anchor_and_target (box_utils.py:149-168) + vuhw_to_vuvu (box_utils.py:5-23)
//   :29-30   softmax over the K logits of every anchor
//   :34-43   argmax != K-1 (first maximum) + boolean_mask (ascending anchor order)
//   :45      top score = max probability
//   :47-52   tf.image.non_max_suppression_with_scores(100, 0.5, sigma 0.5)   -> stage K3, unchanged
//   :54-66   kitti / coco rescaling of the corners
//   :68-73   gather of the selected boxes' class vectors and corners
// Stages: V1 (every anchor: softmax, filter, per-tile compaction into the same
// slot lists K1 uses) -> tile scan -> V2 (every survivor: decode, corners, score)
// -> K3 -> V3 (gather into the padded result blocks: classes in cat_param,
// corners (vuvu) in means).
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

// V1: grid (tiles, B), kTileAnchors threads; one thread per anchor
template <int K>
__global__ void __launch_bounds__(kTileAnchors) val_filter_kernel(ValArgs a) {
    __shared__ int warp_count[kTileAnchors / 32];
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int anchor = tile * kTileAnchors + tid;
    const bool valid = anchor < a.A;
    float p[K];
    bool keep = false;
    if (valid) {
        const float* x = a.cls + ((size_t)b * a.A + anchor) * K;
        float m = __ldg(x);
#pragma unroll
        for (int k = 0; k < K; ++k) { p[k] = __ldg(x + k); m = fmaxf(m, p[k]); }
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) { p[k] = exp_cr(p[k] - m); sum = sum + p[k]; }
        int am = 0;
        float best = p[0] / sum;
#pragma unroll
        for (int k = 0; k < K; ++k) { p[k] = p[k] / sum; if (p[k] > best) { best = p[k]; am = k; } }
        keep = am != K - 1;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_count[warp] = __popc(ballot);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kTileAnchors / 32; ++w) {
        const int c = warp_count[w];
        base += (w < warp) ? c : 0;
        total += c;
    }
    if (keep) {
        const int slot = tile * kTileAnchors + base + __popc(ballot & ((1u << lane) - 1u));
        a.slot_anchor[(size_t)b * a.tiles * kTileAnchors + slot] = anchor;
        float* o = a.slot_counts + ((size_t)b * a.tiles * kTileAnchors + slot) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) o[k] = p[k];
    }
    if (tid == 0) a.tile_count[(size_t)b * a.tiles + tile] = total;
}

// V2: grid (ceil(capacity / 128), B); one thread per survivor
template <int K>
__global__ void __launch_bounds__(128) val_survivor_kernel(ValArgs a) {
    const int b = blockIdx.y;
    const int s = blockIdx.x * 128 + threadIdx.x;
    if (s >= a.num_survivors[b]) return;
    const int32_t* off = a.tile_off + (size_t)b * (a.tiles + 1);
    int lo = 0, hi = a.tiles;                                            // off[lo] <= s < off[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= s) lo = mid; else hi = mid; }
    const int slot = lo * kTileAnchors + (s - off[lo]);
    const int anchor = a.slot_anchor[(size_t)b * a.tiles * kTileAnchors + slot];
    const float* pr = a.slot_counts + ((size_t)b * a.tiles * kTileAnchors + slot) * K;
    const size_t row = (size_t)b * a.capacity + s;
    float best = pr[0];
    float* oc = a.cnt_post + row * K;
#pragma unroll
    for (int k = 0; k < K; ++k) { const float v = pr[k]; oc[k] = v; if (v > best) best = v; }
    const float4 an = __ldg(reinterpret_cast<const float4*>(a.anchors) + anchor);
    const float4 t = __ldg(reinterpret_cast<const float4*>(a.box) + (size_t)b * a.A + anchor);
    const float v = an.z * t.x / 10.0f + an.x;                           // box_utils.py:157-158
    const float u = an.w * t.y / 10.0f + an.y;
    const float h = an.z * fminf(fmaxf(exp_cr(t.z / 5.0f), 1e-4f), 1e4f); // :160-163
    const float w = an.w * fminf(fmaxf(exp_cr(t.w / 5.0f), 1e-4f), 1e4f);
    float c[4] = {v - h / 2.0f, u - w / 2.0f, v + h / 2.0f, u + w / 2.0f};   // box_utils.py:13-21
    a.surv_anchor[row] = anchor;
    a.score[row] = best;
    a.corners[row] = make_float4(c[0], c[1], c[2], c[3]);
    // validation_utils.py:54-66
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (a.scale_mode == 2) c[i] = c[i] - a.shift[i];
        if (a.scale_mode != 0) {
            c[i] = c[i] / ((i & 1) ? a.norm_w : a.norm_h);
            c[i] = c[i] * ((i & 1) ? a.scale_w : a.scale_h);
        }
    }
    reinterpret_cast<float4*>(a.mu_post)[row] = make_float4(c[0], c[1], c[2], c[3]);
}

// V3: grid (Dmax, B); gather of the selected rows, zero padding rows
__global__ void __launch_bounds__(64) val_gather_kernel(ValArgs a) {
    const int d = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const size_t orow = (size_t)b * a.Dmax + d;
    const bool live = d < a.num_dets[b];
    const size_t srow = live ? (size_t)b * a.capacity + a.nms_idx[orow] : 0;
    for (int k = tid; k < a.K; k += 64) {
        a.out_param[orow * a.K + k] = live ? a.cnt_post[srow * a.K + k] : 0.0f;
        a.out_count[orow * a.K + k] = 0.0f;
    }
    if (tid < 4) a.out_means[orow * 4 + tid] = live ? a.mu_post[srow * 4 + tid] : 0.0f;
    if (tid < 16) a.out_covs[orow * 16 + tid] = 0.0f;
}

cudaError_t launch_val_filter(const ValArgs& a, cudaStream_t st) {
    dim3 grid(a.tiles, a.B);
    switch (a.K) {
#define BOD_CASE(KK) case KK: val_filter_kernel<KK><<<grid, kTileAnchors, 0, st>>>(a); break;
        BOD_CASE(2) BOD_CASE(3) BOD_CASE(4) BOD_CASE(5) BOD_CASE(6) BOD_CASE(7) BOD_CASE(8) BOD_CASE(9)
        BOD_CASE(10) BOD_CASE(11) BOD_CASE(12) BOD_CASE(13) BOD_CASE(16) BOD_CASE(21) BOD_CASE(32)
#undef BOD_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_val_survivors(const ValArgs& a, cudaStream_t st) {
    dim3 grid((a.capacity + 127) / 128, a.B);
    switch (a.K) {
#define BOD_CASE(KK) case KK: val_survivor_kernel<KK><<<grid, 128, 0, st>>>(a); break;
        BOD_CASE(2) BOD_CASE(3) BOD_CASE(4) BOD_CASE(5) BOD_CASE(6) BOD_CASE(7) BOD_CASE(8) BOD_CASE(9)
        BOD_CASE(10) BOD_CASE(11) BOD_CASE(12) BOD_CASE(13) BOD_CASE(16) BOD_CASE(21) BOD_CASE(32)
#undef BOD_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_val_gather(const ValArgs& a, cudaStream_t st) {
    val_gather_kernel<<<dim3(a.Dmax, a.B), 64, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bod
