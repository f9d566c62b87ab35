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
// u_i <= t_i, so the candidate TF selects next is x = argmax_i (u_i, -i); on the
// way TF pops, updates and re-pushes exactly the candidates whose stale key
// (t_i, -i) exceeds (u_x, -x).  One ROUND per selected box therefore needs one
// block-wide arg-max plus one pass in which every candidate tests the new box:
// all S IoU evaluations of a round run in parallel, and the multiplication order
// inside each candidate (newest selected first within an update epoch) is kept
// by recomputing u_i from t_i over the candidate's pending-selection bitmask.
// Weights equal to exactly 1.0f (IoU 0, the overwhelmingly common case) leave a
// score bit-identical, so non-overlapping boxes need no work at all.
//
// One CTA per image (images are independent); B CTAs run concurrently.
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

constexpr int kK3Threads = 1024;

BOD_DEVINL unsigned long long make_key(float score, int idx) {
    return ((unsigned long long)float_key(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}

// soft-NMS weight of TF: exp(scale * sim * sim), scale = -0.5 / sigma
BOD_DEVINL float soft_weight(float sim, float scale) { return exp_cr(scale * sim * sim); }

__global__ void __launch_bounds__(kK3Threads, 1)
k3_softnms_kernel(K3Args a) {
    __shared__ unsigned long long warp_best[2][32];
    __shared__ float4 sel_box[kMaxOut];      // corners of the selected boxes, selection order
    __shared__ int sel_idx[kMaxOut];

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.num_survivors[b];
    const int Dmax = a.Dmax;
    const float4* corners = a.corners + (size_t)b * a.capacity;
    const float* score = a.score + (size_t)b * a.capacity;
    float* stale = a.stale + (size_t)b * a.capacity;
    float* cur = a.cur + (size_t)b * a.capacity;
    int32_t* begin = a.begin + (size_t)b * a.capacity;
    uint32_t* pend = a.pend + (size_t)b * a.capacity * kMaskWords;
    uint32_t* member = a.member + (size_t)b * Dmax * a.words;
    const bool is_soft = a.soft_nms_sigma > 0.0f;
    const float scale = is_soft ? -0.5f / a.soft_nms_sigma : 0.0f;
    const float thr = a.iou_threshold;

    // init; `cur` < 0 marks a candidate that left the queue
    unsigned long long best = 0ull;
    for (int s = tid; s < S; s += kK3Threads) {
        const float sc = score[s];
        stale[s] = sc;
        begin[s] = 0;
#pragma unroll
        for (int w = 0; w < kMaskWords; ++w) pend[(size_t)s * kMaskWords + w] = 0u;
        const bool in_queue = sc > -INFINITY;        // scores_data[i] > score_threshold (-inf): NaN stays out
        cur[s] = in_queue ? sc : -INFINITY;          // -inf is never enqueued => usable as "not in queue"
        if (in_queue) { const unsigned long long k = make_key(sc, s); best = k > best ? k : best; }
    }

    int r = 0;
    for (; r < Dmax; ++r) {
        // ---- block arg-max of (cur, -index) over the queue ----
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
            best = o > best ? o : best;
        }
        if (lane == 0) warp_best[r & 1][warp] = best;
        __syncthreads();
        unsigned long long kx = warp_best[r & 1][lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, kx, d);
            kx = o > kx ? o : kx;
        }
        if (kx == 0ull) break;                        // queue empty
        const int x = (int)(0xFFFFFFFFu - (uint32_t)(kx & 0xFFFFFFFFull));
        const float4 bx = corners[x];
        if (tid == 0) {
            sel_box[r] = bx; sel_idx[r] = x;
            a.nms_idx[(size_t)b * Dmax + r] = x;
            a.nms_score[(size_t)b * Dmax + r] = cur[x];
            a.centre_anchor[(size_t)b * Dmax + r] = a.surv_anchor[(size_t)b * a.capacity + x];
        }
        // sel_box[r] is read below by other threads: make it visible
        __syncthreads();

        // ---- one pass: commit popped candidates, test the new box, next arg-max ----
        best = 0ull;
        const int S32 = (S + 31) & ~31;
        for (int s = tid; s < S32; s += kK3Threads) {
            bool mem = false;
            if (s < S) {
                const float4 bs = corners[s];
                float u = cur[s];
                const bool in_queue = (u > -INFINITY) && (s != x);
                if (s == x) cur[s] = -INFINITY;
                // quick geometric reject shared by both IoU definitions: no overlap even
                // with the +1 pixel convention => TF IoU = 0 (weight 1) and repo IoU <= 0.
                const float xI1 = fmaxf(bs.y, bx.y), yI1 = fmaxf(bs.x, bx.x);
                const float xI2 = fminf(bs.w, bx.w), yI2 = fminf(bs.z, bx.z);
                const bool wellformed = (bs.x <= bs.z) && (bs.y <= bs.w) && (bx.x <= bx.z) && (bx.y <= bx.w);
                const bool maybe = !wellformed || (((xI2 - xI1) + 1.0f > 0.0f) && ((yI2 - yI1) + 1.0f > 0.0f));
                if (maybe) mem = repo_iou(bs, bx) > thr;                         // :316, strict >
                if (in_queue) {
                    float t = stale[s];
                    if (u != t && make_key(t, s) > kx) {                          // popped before x: update committed
                        t = u; stale[s] = t; begin[s] = r;
                    }
                    if (maybe) {
                        const float sim = tf_iou(bs, bx);
                        float w = soft_weight(sim, scale);
                        if (!(is_soft || sim <= thr)) w = 0.0f;
                        if (w != 1.0f) {
                            uint32_t* pm = pend + (size_t)s * kMaskWords;
                            pm[r >> 5] |= 1u << (r & 31);
                            // recompute u from t over pending selections >= begin, newest first
                            const int bg = begin[s];
                            float v = t;
                            for (int j = r; j >= bg; --j) {
                                if (!((pm[j >> 5] >> (j & 31)) & 1u)) continue;
                                const float4 bj = (j == r) ? bx : sel_box[j];
                                const float sj = (j == r) ? sim : tf_iou(bs, bj);
                                float wj = soft_weight(sj, scale);
                                if (!(is_soft || sj <= thr)) wj = 0.0f;
                                v = v * wj;
                            }
                            u = v;
                            // hard-NMS (sigma == 0): a zero weight removes the candidate for good
                            if (!is_soft && w == 0.0f) u = -INFINITY;
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
}

cudaError_t launch_k3(const K3Args& a, cudaStream_t st) {
    k3_softnms_kernel<<<a.B, kK3Threads, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bod
