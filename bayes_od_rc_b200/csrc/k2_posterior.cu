// k2_posterior.cu — stages K1b (tile scan) and K2 (per-survivor posterior).
// Compiled with -fmad=false: every operation below is a separately rounded
// binary32 operation in the order of the arithmetic contract.
//
// Reference lines replaced:
//   box_utils.py:171-192            box_from_anchor_and_target_bnms (decode EVERY MC sample)
//   inference_utils.py:57-60,220-244  boolean_mask + compute_mean_covariance_tf
//   inference_utils.py:62-87        aleatoric covariance (diag or L D L^T) and 10:1 mixing
//   inference_utils.py:89-97        Dirichlet prior + posterior score
//   inference_utils.py:99-145       isotropic Gaussian prior fusion (prior mean = anchor)
//   inference_utils.py:147-167      KITTI rescale
//   inference_utils.py:169-202      ranking score
//   box_utils.py:5-23               vuhw_to_vuvu
//   fpn_anchor_generator.py:21-59   anchors (anchor_mode = GENERATE)
//
// Only the S survivors are gathered: N*(16 + 64|40) bytes each, i.e. the
// [N,A,4] and [N,A,4,4] tensors are never streamed in full (the reference
// decodes and reduces them for all A anchors before masking).
#include <cstdlib>
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

// ---------------------------------------------------------------------------
// K1b: exclusive scan of the per-tile survivor counts of every image
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_tiles_kernel(ScanArgs a) {
    BOD_TIMELINE(a.tl);
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* cnt = a.tile_count + (size_t)b * a.tiles;
    int32_t* off = a.tile_off + (size_t)b * (a.tiles + 1);
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < a.tiles; base += (int)blockDim.x) {
        const int t = base + tid;
        const int v = (t < a.tiles) ? cnt[t] : 0;
        int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = x + (warp > 0 ? warp_tot[warp - 1] : 0) + carry;
        if (t < a.tiles) off[t] = incl - v;
        __syncthreads();
        if (tid == (int)blockDim.x - 1) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) {
        int S = carry_s;
        off[a.tiles] = S;
        if (S > a.capacity) { atomicOr(a.status, 1); S = a.capacity; }
        a.num_survivors[b] = S;
    }
}

// beside_k1: 256 threads (8 k registers), so that a scan CTA fits beside the resident moments-kernel CTAs of a pipelined
// context when the scan of run i rides on the tail stream while the head stream already streams the logits of run i+1.
cudaError_t launch_scan(const ScanArgs& a, cudaStream_t st, bool beside_k1) {
    scan_tiles_kernel<<<a.B, beside_k1 ? 256 : 1024, 0, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K1c: pre-NMS filter (EXTENSION, BASELINE.json config 5).  The
// reference has no such knobs; the semantics are the oracle's orc_prefilter:
//   (1) drop survivors whose count score max_k (c_k + alpha) / sum is <= score_threshold,
//   (2) if more than top_k remain keep the top_k by (score desc, anchor index asc),
//   (3) keep ascending anchor order.
// The score depends on the counts only, so this runs between K1 and K2, on the
// per-tile slot lists: (a) a 64-bit key per slot, one warp per tile; (b) one CTA per
// image finds the top_k-th largest key with an 8 x 8-bit radix select (keys are
// unique); (c) every tile is re-compacted, one warp per tile, into a second set of
// slot lists.  The tile scan runs before (dense key indexing) and again after (the
// new counts).
// ---------------------------------------------------------------------------
constexpr int kPfThreads = 1024;
constexpr int kPfTileWarps = 8;           // tiles per CTA in the key / compaction kernels

BOD_DEVINL float count_score(const float* c, int K, bool dirichlet) {
    const float alpha = 1.0f / (float)K;
    float sum = 0.0f, best = 0.0f;
    for (int k = 0; k < K; ++k) sum = sum + (dirichlet ? c[k] + alpha : c[k]);
    for (int k = 0; k < K; ++k) {
        const float p = (dirichlet ? c[k] + alpha : c[k]) / sum;
        if (k == 0 || p > best) best = p;
    }
    return best;
}

// (a) keys: (score, -anchor), 0 = dropped by the threshold; dense entry tile_off[t] + j for slot (t, j).
// One warp per tile, grid (tiles / kPfTileWarps, B).
__global__ void __launch_bounds__(kPfTileWarps * 32) prefilter_keys_kernel(PrefilterArgs a) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kPfTileWarps + (threadIdx.x >> 5);
    if (t >= a.tiles) return;
    const int K = a.K;
    const int cnt = a.tile_count[(size_t)b * a.tiles + t];
    const int off = a.tile_off[(size_t)b * (a.tiles + 1) + t];
    const int32_t* sanchor = a.slot_anchor + (size_t)b * a.slot_stride;
    const float* scounts = a.slot_counts + (size_t)b * a.slot_stride * K;
    unsigned long long* key = a.key + (size_t)b * a.slot_stride;
    for (int j = lane; j < cnt; j += 32) {
        const int slot = t * kTileAnchors + j;
        const float sc = count_score(scounts + (size_t)slot * K, K, a.dirichlet != 0);
        unsigned long long k64 = 0ull;
        if (sc > a.score_threshold)
            k64 = ((unsigned long long)float_key(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)sanchor[slot]);
        key[off + j] = k64;
    }
}

// (b) threshold key per image: the top_k-th largest key (keys are unique), or 1 = keep every
// non-dropped slot.  One CTA per image, 8 x 8-bit radix select over the dense keys.
__global__ void __launch_bounds__(kPfThreads) prefilter_select_kernel(PrefilterArgs a) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long sh_prefix;
    __shared__ unsigned int sh_k, sh_total;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const unsigned long long* key = a.key + (size_t)b * a.slot_stride;
    const int S0 = a.tile_off[(size_t)b * (a.tiles + 1) + a.tiles];
    if (tid == 0) sh_total = 0u;
    __syncthreads();
    unsigned int mine = 0u;
    for (int i = tid; i < S0; i += kPfThreads) mine += key[i] != 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0 && mine) atomicAdd(&sh_total, mine);
    __syncthreads();
    const unsigned int n = sh_total;
    unsigned long long thr_key = 1ull;
    if (a.top_k > 0 && n > (unsigned int)a.top_k) {
        if (tid == 0) { sh_prefix = 0ull; sh_k = (unsigned int)a.top_k; }
        for (int pass = 7; pass >= 0; --pass) {
            for (int i = tid; i < 256; i += kPfThreads) hist[i] = 0u;
            __syncthreads();
            const unsigned long long prefix = sh_prefix;
            const int shift = pass * 8;
            const unsigned long long himask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
            for (int i0 = 0; i0 < S0; i0 += kPfThreads) {                 // warp-uniform trip count
                const int i = i0 + tid;
                const unsigned long long k64 = (i < S0) ? key[i] : 0ull;
                const bool in = k64 != 0ull && (k64 & himask) == prefix;
                // scores take few distinct values, so most keys share their digit: one atomic per distinct
                // digit and warp instead of one per key
                const unsigned int digit = in ? ((unsigned int)(k64 >> shift) & 255u) : 256u;
                const unsigned int peers = __match_any_sync(0xffffffffu, digit);
                if (in && (int)(__ffs(peers) - 1) == lane) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
            }
            __syncthreads();
            if (tid == 0) {
                unsigned int need = sh_k, acc = 0u;
                int d = 255;
                for (; d > 0; --d) { if (acc + hist[d] >= need) break; acc += hist[d]; }
                sh_prefix = prefix | ((unsigned long long)d << shift);
                sh_k = need - acc;                          // rank of the wanted key inside digit d
            }
            __syncthreads();
        }
        thr_key = sh_prefix;
    }
    if (tid == 0) a.thr_key[b] = thr_key;
}

// (c) stable compaction of every tile into the second set of slot lists (out of place: rows are copied
// with independent loads), new tile counts.  One warp per tile, grid (tiles / kPfTileWarps, B).
__global__ void __launch_bounds__(kPfTileWarps * 32) prefilter_compact_kernel(PrefilterArgs a) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kPfTileWarps + (threadIdx.x >> 5);
    if (t >= a.tiles) return;
    const int K = a.K;
    const int cnt = a.tile_count[(size_t)b * a.tiles + t];
    const int off = a.tile_off[(size_t)b * (a.tiles + 1) + t];
    const unsigned long long thr_key = a.thr_key[b];
    const int32_t* sanchor = a.slot_anchor + (size_t)b * a.slot_stride;
    const float* scounts = a.slot_counts + (size_t)b * a.slot_stride * K;
    int32_t* oanchor = a.out_anchor + (size_t)b * a.slot_stride;
    float* ocounts = a.out_counts + (size_t)b * a.slot_stride * K;
    const unsigned long long* key = a.key + (size_t)b * a.slot_stride;
    int kept = 0;
    for (int c0 = 0; c0 < cnt; c0 += 32) {
        const int j = c0 + lane;
        const int slot = t * kTileAnchors + j;
        const bool keep = (j < cnt) && key[off + j] >= thr_key;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int dst = t * kTileAnchors + kept + __popc(bal & ((1u << lane) - 1u));
            oanchor[dst] = sanchor[slot];
            const float* src = scounts + (size_t)slot * K;
            float* d = ocounts + (size_t)dst * K;
            for (int k = 0; k < K; ++k) d[k] = src[k];
        }
        kept += __popc(bal);
    }
    if (lane == 0) a.out_tile_count[(size_t)b * a.tiles + t] = kept;
}

cudaError_t launch_prefilter(const PrefilterArgs& a, cudaStream_t st) {
    dim3 grid((a.tiles + kPfTileWarps - 1) / kPfTileWarps, a.B);
    prefilter_keys_kernel<<<grid, kPfTileWarps * 32, 0, st>>>(a);
    prefilter_select_kernel<<<a.B, kPfThreads, 0, st>>>(a);
    prefilter_compact_kernel<<<grid, kPfTileWarps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// anchors: fpn_anchor_generator.py:21-59, levels 3..7 concatenated P3 -> P7
// ---------------------------------------------------------------------------
struct AnchorLevels {
    int first[6];     // first anchor index of level l (l = 0..4 <-> P3..P7), [5] = A
    int nu[5];        // grid width per level
    float dims[5][9][2];
};

static void host_anchor_levels(int im_h, int im_w, AnchorLevels& L) {
    const float ratios[3][2] = {{1.0f, 1.0f}, {1.0f, 2.0f}, {2.0f, 1.0f}};
    const float scales[3] = {1.0f, 1.26f, 1.59f};
    int total = 0;
    for (int l = 0; l < 5; ++l) {
        const int level = l + 3, stride = 1 << level;
        const int nv = (im_h + stride - 1) / stride, nu = (im_w + stride - 1) / stride;
        const float side = (float)(1 << (level + 2));
        int d = 0;
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s, ++d) {
                if (ratios[r][0] == 1.0f && ratios[r][1] == 1.0f) {
                    L.dims[l][d][0] = ratios[r][0] * side * scales[s];
                    L.dims[l][d][1] = ratios[r][1] * side * scales[s];
                } else {
                    volatile float q = (side * side) / (ratios[r][0] * ratios[r][1]);
                    volatile float sol = sqrtf(q);
                    volatile float h = ratios[r][0] * sol, w = ratios[r][1] * sol;
                    L.dims[l][d][0] = h * scales[s];
                    L.dims[l][d][1] = w * scales[s];
                }
            }
        L.first[l] = total;
        L.nu[l] = nu;
        total += nv * nu * 9;
    }
    L.first[5] = total;
}

int count_anchors(int im_h, int im_w) {
    AnchorLevels L;
    host_anchor_levels(im_h, im_w, L);
    return L.first[5];
}

BOD_DEVINL float4 anchor_of(const AnchorLevels& L, int a) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < 5; ++i) l += (a >= L.first[i]) ? 1 : 0;
    const int r = a - L.first[l];
    const int loc = r / 9, k = r - loc * 9;
    const int iv = loc / L.nu[l], iu = loc - iv * L.nu[l];
    const float stride = (float)(8 << l);
    return make_float4(((float)iv + 0.5f) * stride, ((float)iu + 0.5f) * stride, L.dims[l][k][0], L.dims[l][k][1]);
}

__global__ void generate_anchors_kernel(AnchorLevels L, float4* out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < L.first[5]) out[a] = anchor_of(L, a);
}

cudaError_t launch_generate_anchors(int im_h, int im_w, float* anchors, cudaStream_t st) {
    AnchorLevels L;
    host_anchor_levels(im_h, im_w, L);
    const int A = L.first[5];
    generate_anchors_kernel<<<(A + 255) / 256, 256, 0, st>>>(L, reinterpret_cast<float4*>(anchors));
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------
#ifndef BOD_K2_THREADS
#define BOD_K2_THREADS 128
#endif
constexpr int kK2Threads = BOD_K2_THREADS;
#ifndef BOD_K2_PREFETCH
#define BOD_K2_PREFETCH 2           // samples a thread's gathers run ahead of its decode (0: one sample ahead, in registers)
#endif
constexpr int kK2Prefetch = BOD_K2_PREFETCH;
BOD_DEVINL void cp_async16(void* smem_dst, const void* gmem_src) {
#ifndef BOD_K2_CP_CA
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#else   // through L1 (measured slower: B = 32 step 0.483 against 0.471 ms, KITTI 0.852 against 0.819-0.840)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}

// tfp.math.fill_triangular(x0..x9) element (i,j), lower triangle
// (retinanet_model.py:110): rows of concat(x[4:], reverse(x)) reshaped 4x4.
__device__ __constant__ int kPackedOf[16] = {4, -1, -1, -1, 8, 9, -1, -1, 7, 6, 5, -1, 3, 2, 1, 0};

#ifndef BOD_K2_MINBLOCKS
#define BOD_K2_MINBLOCKS 5
#endif
template <int K>
__global__ void __launch_bounds__(kK2Threads, BOD_K2_MINBLOCKS)
k2_posterior_kernel(K2Args a, AnchorLevels L) {
    BOD_TIMELINE(a.tl);
    extern __shared__ float sbox[];          // decoded boxes [N][4][kK2Threads]
    __shared__ int chunk_first[129];         // prefix of 128-survivor chunks over the images of the batch

    const int tid = threadIdx.x;
    // chunk table: image b owns chunks [chunk_first[b], chunk_first[b+1])
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < a.B; ++b) { chunk_first[b] = acc; acc += (a.num_survivors[b] + kK2Threads - 1) / kK2Threads; }
        chunk_first[a.B] = acc;
    }
    __syncthreads();
    const int total_chunks = chunk_first[a.B];
    const int N = a.N;
    const float alpha = 1.0f / (float)K;                                    // :91

    for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        int b = 0;
        while (chunk >= chunk_first[b + 1]) ++b;
        const int s = (chunk - chunk_first[b]) * kK2Threads + tid;
        const int S = a.num_survivors[b];
        __syncthreads();                                                     // sbox reuse across iterations
        if (s >= S) continue;

        // survivor s -> tile t (binary search in the tile offsets) -> slot
        const int32_t* off = a.tile_off + (size_t)b * (a.tiles + 1);
        int lo = 0, hi = a.tiles;                                            // off[lo] <= s < off[hi]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= s) lo = mid; else hi = mid; }
        const int slot = lo * kTileAnchors + (s - off[lo]);
        const int anchor = a.slot_anchor[(size_t)b * a.tiles * kTileAnchors + slot];
        const float* cn = a.slot_counts + ((size_t)b * a.tiles * kTileAnchors + slot) * K;
        // the level that holds this anchor (one tensor per kind: a single level)
        int lvl = 0;
#pragma unroll
        for (int i = 1; i < kMaxLevels; ++i) lvl += (anchor >= a.lv.first_anchor[i]) ? 1 : 0;     // entries past n hold INT_MAX
        const int A_l = a.lv.rows[lvl];                                  // anchors per sample of the level's tensors
        const int local = anchor - a.lv.first_anchor[lvl] + a.lv.row0[lvl];

        const float4 an = (a.anchor_mode == 1) ? anchor_of(L, anchor)
                                               : __ldg(reinterpret_cast<const float4*>(a.anchors) + anchor);
        const float av = an.x, au = an.y, ah = an.z, aw = an.w;

        // H1 decode every sample (box_utils.py:179-187), accumulate the mean (:233).  The rows of the
        // [N,A,4,4] covariance head of the same sample are summed in the same pass (:67) and everything a
        // sample needs is requested one sample ahead: ten independent 16-byte gathers in flight per thread
        // while the decode (two binary64 exps) of the current sample runs.
        const float4* boxp = reinterpret_cast<const float4*>(a.lv.box[lvl]) + (size_t)b * N * A_l + local;
        const bool cov16 = a.cov_layout == 1;
        const float4* covp = cov16 ? reinterpret_cast<const float4*>(a.lv.cov[lvl]) + ((size_t)b * N * A_l + local) * 4 : nullptr;
        float mu[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        float abar[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) abar[i][j] = 0.0f;
#if BOD_K2_PREFETCH > 0
        // The gathers of a thread run kK2Prefetch samples ahead of its decode, as cp.async copies into the thread's own
        // ring slots in shared memory (no registers held by loads in flight: 15 independent 16-byte gathers per thread
        // instead of 5).  A thread only ever reads what its own copies wrote, so cp.async.wait_group orders everything.
        float4* ring = reinterpret_cast<float4*>(sbox + (size_t)N * 4 * kK2Threads);    // [kK2Prefetch][5][kK2Threads]
        auto issue = [&](int n) {
            if (n < N) {
                float4* slot = ring + (size_t)(n % kK2Prefetch) * 5 * kK2Threads + tid;
                cp_async16(slot, boxp + (size_t)n * A_l);
                if (cov16) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) cp_async16(slot + (size_t)(1 + i) * kK2Threads, covp + (size_t)n * A_l * 4 + i);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");                        // one group per sample, empty past N
        };
#pragma unroll
        for (int n = 0; n < kK2Prefetch; ++n) issue(n);
        for (int n = 0; n < N; ++n) {
            asm volatile("cp.async.wait_group %0;" ::"n"(kK2Prefetch - 1) : "memory");   // sample n has landed
            const float4* slot = ring + (size_t)(n % kK2Prefetch) * 5 * kK2Threads + tid;
            const float4 t = slot[0];
            float4 cr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) cr[i] = cov16 ? slot[(size_t)(1 + i) * kK2Threads] : make_float4(0.f, 0.f, 0.f, 0.f);
            issue(n + kK2Prefetch);                                                      // refill the slot just read
#else
        float4 t_nx = __ldg(boxp);
        float4 c_nx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) c_nx[i] = cov16 ? __ldg(covp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n = 0; n < N; ++n) {
            const float4 t = t_nx;
            float4 cr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) cr[i] = c_nx[i];
            if (n + 1 < N) {
                t_nx = __ldg(boxp + (size_t)(n + 1) * A_l);
                if (cov16) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) c_nx[i] = __ldg(covp + (size_t)(n + 1) * A_l * 4 + i);
                }
            }
#endif
            const float v = ah * t.x / 10.0f + av;
            const float u = aw * t.y / 10.0f + au;
            const float h = ah * fminf(fmaxf(exp_cr(t.z / 5.0f), 1e-4f), 1e4f);
            const float w = aw * fminf(fmaxf(exp_cr(t.w / 5.0f), 1e-4f), 1e4f);
            sbox[(n * 4 + 0) * kK2Threads + tid] = v;
            sbox[(n * 4 + 1) * kK2Threads + tid] = u;
            sbox[(n * 4 + 2) * kK2Threads + tid] = h;
            sbox[(n * 4 + 3) * kK2Threads + tid] = w;
            mu[0] = mu[0] + v; mu[1] = mu[1] + u; mu[2] = mu[2] + h; mu[3] = mu[3] + w;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                abar[i][0] = abar[i][0] + cr[i].x; abar[i][1] = abar[i][1] + cr[i].y;
                abar[i][2] = abar[i][2] + cr[i].z; abar[i][3] = abar[i][3] + cr[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) mu[i] = mu[i] / (float)N;
        // sample covariance :236-242
        float epi[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) epi[i][j] = 0.0f;
        for (int n = 0; n < N; ++n) {
            float d[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = sbox[(n * 4 + i) * kK2Threads + tid] - mu[i];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) epi[i][j] = epi[i][j] + d[i] * d[j];
        }
        const float nm1 = (float)N - 1.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) epi[i][j] = epi[i][j] / nm1;

        // aleatoric :62-84
        float al[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) al[i][j] = 0.0f;
        if (a.cov_layout != 0) {
            if (a.cov_layout == 1) {
                // summed with the box samples above
            } else {
                const float* cp = a.lv.cov[lvl] + ((size_t)b * N * A_l + local) * 10;
                for (int n = 0; n < N; ++n) {
                    const float2* q = reinterpret_cast<const float2*>(cp + (size_t)n * A_l * 10);
                    float x[10];
#pragma unroll
                    for (int i = 0; i < 5; ++i) { const float2 v2 = __ldg(q + i); x[2 * i] = v2.x; x[2 * i + 1] = v2.y; }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            // static map: (0,0)=4 (1,0)=8 (1,1)=9 (2,0)=7 (2,1)=6 (2,2)=5 (3,0)=3 (3,1)=2 (3,2)=1 (3,3)=0
                            constexpr int map[16] = {4, -1, -1, -1, 8, 9, -1, -1, 7, 6, 5, -1, 3, 2, 1, 0};
                            const float e = (map[4 * i + j] >= 0) ? x[map[4 * i + j] >= 0 ? map[4 * i + j] : 0] : 0.0f;
                            abar[i][j] = abar[i][j] + e;
                        }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) abar[i][j] = abar[i][j] / (float)N;
            float Dm[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) Dm[i] = exp_cr(abar[i][i]);           // :70
            if (a.use_full_covar) {                                             // :74-80
#pragma unroll
                for (int i = 0; i < 4; ++i) abar[i][i] = 1.0f;
                float Li[4][4], LD[4][4];
                inv4(abar, Li);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) LD[i][j] = Li[i][j] * Dm[j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float acc = 0.0f;
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc = acc + LD[i][k] * Li[j][k];
                        al[i][j] = acc;
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) al[i][i] = Dm[i];                 // :71-73, 82
            }
        }
        // mixing :86-87
        float lik[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) lik[i][j] = (10.0f * al[i][j] + 1.0f * epi[i][j]) / 11.0f;

        // dirichlet :89-97
        float cp[K];
        float csum = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float c = cn[k];
            cp[k] = (a.dirichlet_prior == 1) ? c + alpha : c;
            csum = csum + cp[k];
        }

        // gaussian prior :99-145
        float mp[4], sp[4][4];
        if (a.gaussian_prior == 1) {
            float prec_l[4][4], prec_post[4][4], w_l[4], inter[4];
            inv4(lik, prec_l);                                                  // :101
            const float prec_p = 1.0f / a.isotropic_variance;                   // :120
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) prec_post[i][j] = (i == j) ? prec_l[i][j] + prec_p : prec_l[i][j];   // :127
            inv4(prec_post, sp);                                                // :129
            const float anv[4] = {av, au, ah, aw};                              // :122
            mv4(prec_l, mu, w_l);                                               // :137
#pragma unroll
            for (int i = 0; i < 4; ++i) inter[i] = prec_p * anv[i] + w_l[i];    // :132,140
            mv4(sp, inter, mp);                                                 // :141
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mp[i] = mu[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) sp[i][j] = lik[i][j];
            }
        }
        // kitti rescale :147-167 (identity for scales of 1)
        {
            const float sc[4] = {a.scale_v, a.scale_u, a.scale_v, a.scale_u};
#pragma unroll
            for (int i = 0; i < 4; ++i) mp[i] = sc[i] * mp[i];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sp[i][j] = (sc[i] * sp[i][j]) * sc[j];
        }

        // outputs
        const size_t row = (size_t)b * a.capacity + s;
        a.surv_anchor[row] = anchor;
        float* ocp = a.cnt_post + row * K;
#pragma unroll
        for (int k = 0; k < K; ++k) ocp[k] = cp[k];
        reinterpret_cast<float4*>(a.mu_post)[row] = make_float4(mp[0], mp[1], mp[2], mp[3]);
        float4* osg = reinterpret_cast<float4*>(a.sig_post) + row * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) osg[i] = make_float4(sp[i][0], sp[i][1], sp[i][2], sp[i][3]);

        // ranking :169-202
        if (a.ranking_method == 1) {
            // information gains; normalised over the image by rank_normalise_kernel
            const float two_pi_log = log_cr(6.2831853071795862f);
            const float hp = 2.0f + 2.0f * two_pi_log + 0.5f * log_cr(det4(sp));           // :247-263
            float var4[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) var4[i][j] = (i == j) ? a.isotropic_variance : 0.0f;
            const float hprior = 2.0f + 2.0f * two_pi_log + 0.5f * log_cr(det4(var4));
            float ent = 0.0f, ent0 = 0.0f, psum = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) { const float p = cp[k] / csum; ent = ent + p * log_cr(p); }   // :266-277
#pragma unroll
            for (int k = 0; k < K; ++k) psum = psum + alpha;
#pragma unroll
            for (int k = 0; k < K; ++k) { const float p = alpha / psum; ent0 = ent0 + p * log_cr(p); }
            a.info[row * 2 + 0] = hprior - hp;
            a.info[row * 2 + 1] = (-ent0) - (-ent);
        } else {
            float best = cp[0] / csum;
#pragma unroll
            for (int k = 1; k < K; ++k) best = fmaxf(best, cp[k] / csum);
            a.score[row] = best;
        }
        // vuhw_to_vuvu box_utils.py:13-21
        a.corners[row] = make_float4(mp[0] - mp[2] / 2.0f, mp[1] - mp[3] / 2.0f, mp[0] + mp[2] / 2.0f, mp[1] + mp[3] / 2.0f);
    }
}

// joint_entropy ranking: min/max normalisation over the survivors of an image (:177-200)
__global__ void __launch_bounds__(1024) rank_normalise_kernel(K2Args a) {
    __shared__ float red[4][32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.num_survivors[b];
    const float* info = a.info + (size_t)b * a.capacity * 2;
    float gmin = INFINITY, gmax = -INFINITY, cmin = INFINITY, cmax = -INFINITY;
    for (int s = tid; s < S; s += 1024) {
        const float g = info[2 * s], c = info[2 * s + 1];
        gmin = fminf(gmin, g); gmax = fmaxf(gmax, g); cmin = fminf(cmin, c); cmax = fmaxf(cmax, c);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, d)); gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, d));
        cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, d)); cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
    }
    if (lane == 0) { red[0][warp] = gmin; red[1][warp] = gmax; red[2][warp] = cmin; red[3][warp] = cmax; }
    __syncthreads();
    gmin = red[0][lane]; gmax = red[1][lane]; cmin = red[2][lane]; cmax = red[3][lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, d)); gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, d));
        cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, d)); cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
    }
    const float gden = fmaxf(1.0f, gmax - gmin), cden = fmaxf(0.001f, cmax - cmin);
    for (int s = tid; s < S; s += 1024) {
        const float g = (info[2 * s] - gmin) / gden;
        const float c = (info[2 * s + 1] - cmin) / cden;
        a.score[(size_t)b * a.capacity + s] = c + g;
    }
}

template <int K>
static cudaError_t launch_k2_k(const K2Args& a, const AnchorLevels& L, cudaStream_t st) {
    const size_t smem = (size_t)a.N * 4 * kK2Threads * sizeof(float) + (size_t)kK2Prefetch * 5 * kK2Threads * sizeof(float4);
    cudaError_t e = ensure_dyn_smem((const void*)k2_posterior_kernel<K>, smem);
    if (e != cudaSuccess) return e;
    // persistent grid: as many CTAs as are resident at once (chunks of 128 survivors are taken grid-stride)
    const int sms = sm_count();
    static std::mutex mu;
    static int occ_N[64] = {0}, occ_v[64] = {0};         // resident CTAs per SM for the last N seen on the device
    int dev = 0, per_sm = 4;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 0 && dev < 64 && occ_N[dev] == a.N && occ_v[dev] > 0) per_sm = occ_v[dev];
        else {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_posterior_kernel<K>, kK2Threads, smem) != cudaSuccess || per_sm < 1)
                per_sm = 4;
            if (dev >= 0 && dev < 64) { occ_N[dev] = a.N; occ_v[dev] = per_sm; }
        }
    }
    static const int cap_env = getenv("BOD_K2_CTAS") ? atoi(getenv("BOD_K2_CTAS")) : 0;   // experiment: fewer resident CTAs per SM
    if (cap_env >= 1 && cap_env < per_sm) per_sm = cap_env;
    k2_posterior_kernel<K><<<sms * per_sm, kK2Threads, smem, st>>>(a, L);
    return cudaGetLastError();
}

cudaError_t launch_k2(const K2Args& a, cudaStream_t st) {
    AnchorLevels L = {};
    if (a.anchor_mode == 1) host_anchor_levels(a.im_h, a.im_w, L);
    switch (a.K) {
#define BOD_CASE(KK) case KK: return launch_k2_k<KK>(a, L, st);
        BOD_CASE(2) BOD_CASE(3) BOD_CASE(4) BOD_CASE(5) BOD_CASE(6) BOD_CASE(7) BOD_CASE(8) BOD_CASE(9)
        BOD_CASE(10) BOD_CASE(11) BOD_CASE(12) BOD_CASE(13) BOD_CASE(16) BOD_CASE(21) BOD_CASE(32)
#undef BOD_CASE
        default: return cudaErrorInvalidValue;
    }
}

// ---- CUDA-graph support (see k1_kernel_func) ----
const void* k2_kernel_func(const K2Args& a) {
    switch (a.K) {
#define BOD_CASE(KK) case KK: return (const void*)k2_posterior_kernel<KK>;
        BOD_CASE(2) BOD_CASE(3) BOD_CASE(4) BOD_CASE(5) BOD_CASE(6) BOD_CASE(7) BOD_CASE(8) BOD_CASE(9)
        BOD_CASE(10) BOD_CASE(11) BOD_CASE(12) BOD_CASE(13) BOD_CASE(16) BOD_CASE(21) BOD_CASE(32)
#undef BOD_CASE
        default: return nullptr;
    }
}
cudaError_t k2_graph_update(cudaGraphExec_t exec, cudaGraphNode_t node, const K2Args& a0) {
    K2Args a = a0;
    cudaKernelNodeParams p;
    cudaError_t e = cudaGraphKernelNodeGetParams(node, &p);
    if (e != cudaSuccess) return e;
    if (p.func != k2_kernel_func(a)) return cudaErrorInvalidValue;
    void* args[2] = {&a, p.kernelParams[1]};                            // (K2Args, AnchorLevels)
    p.kernelParams = args;
    return cudaGraphExecKernelNodeSetParams(exec, node, &p);
}

cudaError_t launch_rank_normalise(const K2Args& a, cudaStream_t st) {
    rank_normalise_kernel<<<a.B, 1024, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bod
