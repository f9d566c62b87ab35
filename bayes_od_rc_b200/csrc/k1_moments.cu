// k1_moments.cu — stage K1: per-anchor class moments over the N MC-dropout
// samples, categorical draw counts, non-background filter and stable per-tile
// compaction.
//
// Reference lines replaced (src/retina_net/experiments/inference_utils.py):
//   :31-32,38  softmax over K per sample, mean over N         -> mean probabilities
//   :37-46     Categorical(probs).sample(30) -> one_hot -> sum -> counts [A,K]
//              (injected tensor in parity mode, Philox4x32-10 otherwise)
//   :48-54     argmax(counts) != K-1 (first maximum) + boolean_mask (stable)
//
// This is the only stage that must touch every anchor: it streams the whole
// [B,N,A,K] logits tensor exactly once (4*N*A*K bytes per image), so it IS the
// HBM roofline of the path.  Data movement: the unit of work is a tile of
// kTileAnchors consecutive anchors of one image; for every MC sample the tile's
// [tile,K] slab is a contiguous span of global memory.  Persistent CTAs (six per
// SM, 160 threads each) take tiles from a global ticket counter; one producer
// thread per CTA bulk-copies slab after slab into a ring of shared-memory stages
// with the TMA engine (cp.async.bulk + full/empty mbarriers), running up to NS
// slabs ahead of the CTA's four consumer warps (one thread per anchor of the
// tile), so ~170-210 KB per SM are always in flight, no thread issues a global load
// and rows of any K (8, 11, 4, ...) are consumed conflict-free from shared memory.
#include <cstdlib>
#include <mutex>
#include "bod_common.cuh"
#include "bod_kernels.h"

namespace bod {

// ---------------------------------------------------------------------------
// mbarrier / bulk-copy PTX wrappers (sm_90+; SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------
// fast math of the softmax (tolerance-checked output): MUFU.EX2 / MUFU.RCP
BOD_DEVINL unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
BOD_DEVINL float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
BOD_DEVINL float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Packed binary32 pairs: sm_100 issues one FFMA2 / FADD2 for two values (the softmax loop is issue-bound, not
// pipe-bound: the pipelined step shares its SMs with the posterior / soft-NMS / fusion kernels of earlier runs).
BOD_DEVINL uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
BOD_DEVINL void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
BOD_DEVINL uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
BOD_DEVINL uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// H2 (inference_utils.py:31-32, 38): sum over the MC samples of softmax(logits row).  exp(x_k - c) for ANY common
// shift c gives the same softmax; the shift used first is the row's background logit (the last column: the largest
// entry of almost every anchor, and a few units from the largest one elsewhere), which costs one multiplication
// instead of a max over the row.  Rows whose entries are so far apart that the sum leaves [1e-30, 1e30] (overflow,
// total underflow, NaN) are redone with the row maximum.  Tolerance-checked output (<= 2e-6 absolute vs the oracle).
template <int K> struct SoftmaxSum {
    static constexpr int KP = K / 2;                 // whole pairs; an odd K keeps its last column in `last`
    uint64_t p2[KP > 0 ? KP : 1];
    float last;
    BOD_DEVINL void clear() {
#pragma unroll
        for (int j = 0; j < KP; ++j) p2[j] = 0ull;
        last = 0.0f;
    }
    static BOD_DEVINL float terms(const float (&x)[K], float c, float (&e)[K]) {
        constexpr float L = 1.4426950408889634f;     // exp(x - c') = 2^(x*log2e - c), c = c'*log2e
        const uint64_t L2 = pk2(L, L), nc2 = pk2(-c, -c);
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            float t0, t1;
            upk2(fma2(pk2(x[2 * j], x[2 * j + 1]), L2, nc2), t0, t1);
            e[2 * j] = ex2_approx(t0); e[2 * j + 1] = ex2_approx(t1);
        }
        if (K & 1) e[K - 1] = ex2_approx(__fmaf_rn(x[K - 1], L, -c));
        float s = 0.0f;
        if (KP > 0) {
            uint64_t s2 = pk2(e[0], e[1]);
#pragma unroll
            for (int j = 1; j < KP; ++j) s2 = add2(s2, pk2(e[2 * j], e[2 * j + 1]));
            float s0, s1;
            upk2(s2, s0, s1);
            s = s0 + s1;
        }
        if (K & 1) s += e[K - 1];
        return s;
    }
    BOD_DEVINL void add_row(const float (&x)[K]) {
        float e[K];
        float s = terms(x, x[K - 1] * 1.4426950408889634f, e);
        if (!(s >= 1e-30f && s <= 1e30f)) {
            float m = x[0];
#pragma unroll
            for (int k = 1; k < K; ++k) m = fmaxf(m, x[k]);
            s = terms(x, m * 1.4426950408889634f, e);
        }
        const float inv = rcp_approx(s);
        const uint64_t inv2 = pk2(inv, inv);
#pragma unroll
        for (int j = 0; j < KP; ++j) p2[j] = fma2(pk2(e[2 * j], e[2 * j + 1]), inv2, p2[j]);
        if (K & 1) last = __fmaf_rn(e[K - 1], inv, last);
    }
    BOD_DEVINL void mean(float scale, float (&p)[K]) const {
#pragma unroll
        for (int j = 0; j < KP; ++j) { float a, b; upk2(p2[j], a, b); p[2 * j] = a * scale; p[2 * j + 1] = b * scale; }
        if (K & 1) p[K - 1] = last * scale;
    }
};

// Launch clock (K1Args::clk): CTA (0,0) opens a slot with its start time, every CTA leaves its end time behind.
BOD_DEVINL void launch_clock_begin(unsigned long long* clk_) {
    if (clk_ != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        volatile unsigned long long* clk = clk_;
        const unsigned long long i = clk[0];
        clk[2 + 2 * (i % kClkSlots)] = global_ns();
        clk[3 + 2 * (i % kClkSlots)] = 0ull;
        __threadfence();
        clk[0] = i + 1;
    }
}
BOD_DEVINL void launch_clock_end(unsigned long long* clk_) {
    if (clk_ != nullptr && threadIdx.x == 0) {
        const unsigned long long i = *reinterpret_cast<volatile unsigned long long*>(clk_) - 1ull;   // CTA (0,0) opened the slot long ago
        atomicMax(clk_ + 3 + 2 * (i % kClkSlots), global_ns());
    }
}

BOD_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
BOD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
BOD_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
BOD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep, do not spin
}
BOD_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The same on 32-bit shared-window addresses: the consumers of the pipeline kernel convert their pointers once,
// outside the sample loop (inside it the generic->shared conversion costs ~14 instructions per iteration).
BOD_DEVINL void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
BOD_DEVINL void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
template <int OFF> BOD_DEVINL float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF> BOD_DEVINL void lds_v4(uint32_t addr, float& a, float& b, float& c, float& d) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr), "n"(OFF));
}
// One anchor's K logits of the current stage.  Rows are K floats apart (the slab is a verbatim copy of the [tile, K]
// span of `cls`), so 32-bit loads are conflict-free only for odd K: K = 8 puts the 32 lanes on 4 banks (8-way
// conflict, 64 wavefronts per row — the LSU pipe, not HBM, then bounds the kernel) and K = 4 on 8 banks.  For
// K % 4 == 0 the row is read with 128-bit loads instead (K = 4: conflict-free, K = 8: 16 wavefronts per row).
template <int K, int I = 0, bool V4 = (K % 4 == 0)> struct LoadRow {
    static BOD_DEVINL void run(uint32_t addr, float (&x)[K]) { x[I] = lds_f32<4 * I>(addr); LoadRow<K, I + 1, V4>::run(addr, x); }
};
template <int K, int I> struct LoadRow<K, I, true> {
    static BOD_DEVINL void run(uint32_t addr, float (&x)[K]) {
        lds_v4<4 * I>(addr, x[I], x[I + 1], x[I + 2], x[I + 3]);
        LoadRow<K, I + 4, true>::run(addr, x);
    }
};
template <int K> struct LoadRow<K, K, false> { static BOD_DEVINL void run(uint32_t, float (&)[K]) {} };
template <int K> struct LoadRow<K, K, true> { static BOD_DEVINL void run(uint32_t, float (&)[K]) {} };

BOD_DEVINL void consumer_barrier() {   // named barrier 1: the kTileAnchors consumer threads only
    asm volatile("bar.sync 1, %0;" ::"n"(kTileAnchors) : "memory");
}
BOD_DEVINL void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------
// Multinomial(T, p) counts for one anchor (stands in for the unseeded
// Categorical(probs).sample(30) -> one_hot -> reduce_sum, inference_utils.py:37-46).
// Uniform i of anchor a in global image g = 23-bit field (i % 5) of Philox call
// (i / 5) with counter (a, g, call, 0x0B0D): u = (field + 0.5) * 2^-23.
// Dominant-class split: the number of draws that do NOT land on the most likely
// class m is Binomial(T, 1 - p_m), drawn by inversion from 0 upward (a background
// anchor needs ~2 steps and ONE Philox call instead of 30 draws / 6 calls); each
// of those draws then picks a class != m by inverse cdf.  Flat distributions
// (p_m^T < 1e-30) fall back to T plain inverse-cdf draws.  Every operation is an
// explicitly rounded binary32 intrinsic so the CPU restatement is bit-identical.
// ---------------------------------------------------------------------------
// (T - j) / (j + 1) for j < kRatioTab, one slot per number of draws T seen on the device.  A slot is
// written once (before the first launch that uses it) and never rewritten, so contexts with different T
// share the table safely; when the slots run out the kernel divides instead (same bits, slower).
constexpr int kRatioTab = 64;
constexpr int kRatioSlots = 8;
__constant__ float c_binom_ratio[kRatioSlots * kRatioTab];

struct PhiloxStream {
    uint4 w;
    uint32_t anchor, image;
    uint2 key;
    int g;
    BOD_DEVINL float next() {
        const int f = g % 5;
        if (f == 0) w = philox4x32_10(make_uint4(anchor, image, (uint32_t)(g / 5), 0x0B0Du), key);
        uint32_t field;
        switch (f) {
            case 0: field = w.x; break;
            case 1: field = (w.x >> 23) | (w.y << 9); break;
            case 2: field = (w.y >> 14) | (w.z << 18); break;
            case 3: field = w.z >> 5; break;
            default: field = (w.z >> 28) | (w.w << 4); break;
        }
        ++g;
        return __fmul_rn(__fadd_rn((float)(field & 0x7FFFFFu), 0.5f), 1.1920928955078125e-07f);
    }
};

// Returns false when the anchor is known to be background-dominant without the
// per-class counts (only possible with `need_counts` false): the dominant class
// is the background column and it received more than half of the draws, so the
// first-maximum argmax (inference_utils.py:48-51) is the background and the
// anchor is dropped; its counts are never read again.  Otherwise cnt[] holds the
// full multinomial counts.
template <int K>
BOD_DEVINL bool philox_counts(const float (&p)[K], uint32_t anchor, uint32_t image, uint2 key, int T, int rslot,
                              bool need_counts, float (&cnt)[K]) {
    float cdf[K];
    float s = 0.0f, pm = p[0];
    int m = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s = __fadd_rn(s, p[k]); cdf[k] = s; cnt[k] = 0.0f;
        if (k > 0 && p[k] > pm) { pm = p[k]; m = k; }
    }
    const float total = cdf[K - 1];
    const float rest = __fsub_rn(total, pm);
    const float odds = __fmul_rn(rest, __frcp_rn(pm));
    float pw = 1.0f, base = pm;
    for (int e = T; e; e >>= 1) { if (e & 1) pw = __fmul_rn(pw, base); base = __fmul_rn(base, base); }
    PhiloxStream rng{make_uint4(0, 0, 0, 0), anchor, image, key, 0};
    if (pw >= 1e-30f) {
        const float u = rng.next();
        int j = 0;
        float cd = pw, f = pw;
        while (u >= cd && j < T) {
            const float ratio = (j < kRatioTab && rslot >= 0) ? c_binom_ratio[rslot * kRatioTab + j]
                                                              : __fdiv_rn((float)(T - j), (float)(j + 1));
            f = __fmul_rn(__fmul_rn(f, ratio), odds);
            ++j;
            cd = __fadd_rn(cd, f);
        }
        if (!need_counts && m == K - 1 && 2 * j < T) return false;     // background keeps a strict majority
        // running sums of p over the classes != m (adding 0 for m is exact)
        float acc[K];
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) { t = __fadd_rn(t, (k == m) ? 0.0f : p[k]); acc[k] = t; }
        const int last = (m == K - 1) ? K - 2 : K - 1;
        for (int i = 0; i < j; ++i) {
            const float x = __fmul_rn(rng.next(), rest);
            int c = last;
#pragma unroll
            for (int k = K - 1; k >= 0; --k) c = (k != m && x < acc[k]) ? k : c;
#pragma unroll
            for (int k = 0; k < K; ++k) cnt[k] = __fadd_rn(cnt[k], (c == k) ? 1.0f : 0.0f);
        }
        const float cm = (float)(T - j);
#pragma unroll
        for (int k = 0; k < K; ++k) cnt[k] = (k == m) ? cm : cnt[k];
    } else {
        for (int t = 0; t < T; ++t) {
            const float x = __fmul_rn(rng.next(), total);
            int c = K - 1;
#pragma unroll
            for (int k = K - 2; k >= 0; --k) c = (x < cdf[k]) ? k : c;
#pragma unroll
            for (int k = 0; k < K; ++k) cnt[k] = __fadd_rn(cnt[k], (c == k) ? 1.0f : 0.0f);
        }
    }
    return true;
}

// tile of the per-image grid -> level, first local anchor of the tile, anchors of the level
struct TileRef { int level, a0, A_l, anchor0, rows, row; };   // rows: anchors per sample of the level's tensor; row: the tile's first row in it
BOD_DEVINL TileRef tile_ref(const LevelTable& lv, int tile) {
    TileRef r;
    if (lv.n == 1) {                                     // one tensor per kind (uniform branch)
        r.level = 0; r.a0 = tile * kTileAnchors; r.A_l = lv.count[0]; r.anchor0 = r.a0;
        r.rows = lv.rows[0]; r.row = lv.row0[0] + r.a0;
        return r;
    }
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i) l += (tile >= lv.first_tile[i]) ? 1 : 0;     // entries past n hold INT_MAX
    r.level = l;
    r.a0 = (tile - lv.first_tile[l]) * kTileAnchors;
    r.A_l = lv.count[l];
    r.anchor0 = lv.first_anchor[l] + r.a0;
    r.rows = lv.rows[l];
    r.row = lv.row0[l] + r.a0;
    return r;
}

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
// grid = (tiles, B), block = kTileAnchors threads.  The N samples of a tile are
// consumed in chunks of NC samples through a two-stage shared-memory ring
// (stage = NC slabs of [tile,K] floats, one mbarrier per stage), so any N fits
// and two chunks are always in flight per CTA (x2 resident CTAs per SM).
// USE_BULK: slabs arrive through cp.async.bulk (needs 16-byte aligned spans);
// otherwise a cooperative coalesced copy fills the same shared layout.
template <int K, bool USE_BULK>
__global__ void __launch_bounds__(kTileAnchors, 2)
k1_moments_kernel(K1Args a, int NC) {
    launch_clock_begin(a.clk);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar[2];
    __shared__ int warp_count[kTileAnchors / 32];

    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const TileRef tr = tile_ref(a.lv, tile);
    const int a0 = tile * kTileAnchors;                      // first slot of the tile
    const int rows = min(kTileAnchors, tr.A_l - tr.a0);      // anchors in this tile
    const int N = a.N;
    const int nchunks = (N + NC - 1) / NC;
    constexpr size_t slab_stride = (size_t)kTileAnchors * K;  // floats per smem slab
    const size_t stage_stride = slab_stride * NC;
    float* ring = reinterpret_cast<float*>(smem_raw);
    const float* src0 = a.lv.cls[tr.level] + ((size_t)b * N * tr.rows + tr.row) * K;   // sample 0 of this tile
    const uint32_t slab_bytes = (uint32_t)rows * K * 4u;

    auto issue = [&](int c) {   // one thread: bulk-copy chunk c into stage c&1
        const int n0 = c * NC, n1 = min(N, n0 + NC);
        uint64_t* br = &bar[c & 1];
        mbar_expect_tx(br, slab_bytes * (uint32_t)(n1 - n0));
        for (int n = n0; n < n1; ++n)
            bulk_g2s(ring + (c & 1) * stage_stride + (size_t)(n - n0) * slab_stride, src0 + (size_t)n * tr.rows * K,
                     slab_bytes, br);
    };

    if (USE_BULK) {
        if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
        if (tid == 0) { issue(0); if (nchunks > 1) issue(1); }
    }

    // counts to inject (parity mode) are fetched while the slabs are in flight
    const int anchor = tr.anchor0 + tid;
    const bool valid = tid < rows;
    float cnt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) cnt[k] = 0.0f;
    if (a.counts_in != nullptr && valid) {
        const float* c = a.counts_in + ((size_t)b * a.A + anchor) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) cnt[k] = __ldg(c + k);
    }

    // H2: softmax per sample, mean over samples (fast-math allowed here)
    float p[K];
    SoftmaxSum<K> acc;
    acc.clear();
    for (int c = 0; c < nchunks; ++c) {
        const int n0 = c * NC, n1 = min(N, n0 + NC);
        const float* stage = ring + (USE_BULK ? (c & 1) : 0) * stage_stride;
        if (USE_BULK) {
            mbar_wait(&bar[c & 1], (uint32_t)((c >> 1) & 1));
        } else {
            __syncthreads();
            const int nflt = rows * K;
            for (int n = n0; n < n1; ++n) {
                const float* src = src0 + (size_t)n * tr.rows * K;
                float* dst = ring + (size_t)(n - n0) * slab_stride;
                for (int e = tid; e < nflt; e += kTileAnchors) dst[e] = __ldg(src + e);
            }
            __syncthreads();
        }
        if (valid) {
            for (int n = n0; n < n1; ++n) {
                const float* row = stage + (size_t)(n - n0) * slab_stride + tid * K;
                float x[K];
#pragma unroll
                for (int k = 0; k < K; ++k) x[k] = row[k];
                acc.add_row(x);
            }
        }
        if (USE_BULK && c + 2 < nchunks) {
            __syncthreads();                       // every thread is done reading stage c&1
            if (tid == 0) issue(c + 2);
        }
    }
    acc.mean(1.0f / (float)N, p);
    if (valid) {
        if (a.probs_out != nullptr) {
            float* o = a.probs_out + ((size_t)b * a.A + anchor) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) o[k] = p[k];
        }
    }

    // H3: categorical draw counts
    bool maybe_fg = true;
    if (a.counts_in == nullptr && valid) {
        maybe_fg = philox_counts<K>(p, (uint32_t)anchor, a.image_id_base + (uint32_t)b,
                                        make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)), a.num_draws, a.ratio_slot,
                                        a.sampled_out != nullptr, cnt);
        if (a.sampled_out != nullptr) {
            float* o = a.sampled_out + ((size_t)b * a.A + anchor) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) o[k] = cnt[k];
        }
    }

    // H4: first-maximum argmax != background, stable compaction inside the tile
    bool keep = false;
    if (valid && maybe_fg) {
        int am = 0;
        float best = cnt[0];
#pragma unroll
        for (int k = 1; k < K; ++k) if (cnt[k] > best) { best = cnt[k]; am = k; }
        keep = (am != K - 1);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int lane = tid & 31, warp = tid >> 5;
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
        const int slot = a0 + base + __popc(ballot & ((1u << lane) - 1u));   // per-tile slot region
        a.slot_anchor[(size_t)b * a.tiles * kTileAnchors + slot] = anchor;
        float* o = a.slot_counts + ((size_t)b * a.tiles * kTileAnchors + slot) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) o[k] = cnt[k];
    }
    if (tid == 0) a.tile_count[(size_t)b * a.tiles + tile] = total;
    launch_clock_end(a.clk);
}

// ---------------------------------------------------------------------------
// persistent, warp-specialised pipeline (the production path)
// ---------------------------------------------------------------------------
constexpr int kMaxStages = 32;
constexpr int kConsumerWarps = kTileAnchors / 32;

constexpr int kPipeCtasPerSM = 6;   // resident CTAs per SM: their finalise phases overlap each other's streaming

// NSU > 0: the ring has exactly NSU stages and N is a multiple of NSU, so every tile starts at stage 0 and a round
// of NSU samples is unrolled with every address a constant offset (no ring bookkeeping in the sample loop);
// NSU = 0: any ring depth NS, any N.
template <int K, int NSU>
// Register cap of the pipeline kernel: 65536 / (MINBLOCKS * 160) -> 56 per thread.  Six CTAs of 56
// registers leave ~11.7 k registers per SM, enough for one fusion (K4) CTA to run beside them in a
// pipelined context; at 64 the kernel is ~1 % faster alone and the pipelined step ~2 % slower (measured).
#ifndef BOD_K1_MINBLOCKS
#define BOD_K1_MINBLOCKS 7
#endif
__global__ void __launch_bounds__(kTileAnchors + 32, BOD_K1_MINBLOCKS)
k1_moments_pipe_kernel(K1Args a, int NS) {
    BOD_TIMELINE(a.tl);
    launch_clock_begin(a.clk);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[2 * kMaxStages];         // one array: empty_bar = full_bar + a compile-time offset
    uint64_t* const full_bar = bars;
    uint64_t* const empty_bar = bars + kMaxStages;
    __shared__ int stage_tile[kMaxStages];            // tile whose first slab sits in the stage (-1: no more work)
    __shared__ int warp_count[2][kConsumerWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N, tiles = a.tiles;
    const int total_tiles = a.B * tiles;
    constexpr size_t slab_stride = (size_t)kTileAnchors * K;      // floats per stage
    float* ring = reinterpret_cast<float*>(smem_raw);

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kConsumerWarps); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ---- producer: one thread feeds the ring, up to NS slabs ahead of the consumers.  Tiles come
        // from a global ticket counter, so CTAs that start late (SMs busy with another stream's
        // kernels) simply take fewer tiles ----
        if (lane == 0) {
            int stage = 0, round = 0;
            for (;;) {
                const uint32_t t = atomicAdd(a.ticket, 1u) - a.ticket_base;
                // every CTA draws exactly one ticket past the last tile, so the launch draws total_tiles + gridDim.x
                // tickets in all: whoever draws the last one puts the counter back for the next launch (launches of
                // a context never overlap), which saves a memset between two launches
                if (t == (uint32_t)total_tiles + gridDim.x - 1u) *a.ticket = a.ticket_base;
                if (round > 0) mbar_wait(&empty_bar[stage], (uint32_t)((round - 1) & 1));
                if (t >= (uint32_t)total_tiles) {                 // no work left: tell the consumers
                    stage_tile[stage] = -1;
                    mbar_arrive(&full_bar[stage]);
                    break;
                }
                stage_tile[stage] = (int)t;
                const int b = (int)t / tiles, tile = (int)t - b * tiles;
                const TileRef tr = tile_ref(a.lv, tile);
                const int rows = min(kTileAnchors, tr.A_l - tr.a0);
                const uint32_t bytes = (uint32_t)rows * K * 4u;
                const float* src = a.lv.cls[tr.level] + ((size_t)b * N * tr.rows + tr.row) * K;
                for (int n = 0; n < N; ++n) {
                    if (n > 0 && round > 0) mbar_wait(&empty_bar[stage], (uint32_t)((round - 1) & 1));
                    mbar_expect_tx(&full_bar[stage], bytes);
                    bulk_g2s(ring + stage * slab_stride, src + (size_t)n * tr.rows * K, bytes, &full_bar[stage]);
                    if (++stage == NS) { stage = 0; ++round; }
                }
            }
        }
        return;
    }

    // ---- consumers: one thread per anchor of the tile ----
    // Shared-window addresses of the ring walk in registers: this thread's row and the full barrier of the current
    // stage (its empty barrier sits a constant behind it); one compare per sample wraps both.
    const uint32_t full0 = smem_u32(&full_bar[0]);
    constexpr uint32_t empty_minus_full = kMaxStages * 8;
    const uint32_t row0 = smem_u32(ring) + (uint32_t)tid * (uint32_t)(K * 4);
    constexpr uint32_t kSlabBytes = (uint32_t)(slab_stride * 4);
    const uint32_t bar_end = full0 + (uint32_t)NS * 8u;
    uint32_t row_addr = row0, bar = full0;
    uint32_t phase = 0;
    const float invN = 1.0f / (float)N;
    for (int tcount = 0;; ++tcount) {
        mbar_wait_a(bar, phase);                             // first slab of the next tile, or the end marker
        const int t = stage_tile[NSU > 0 ? 0 : (bar - full0) >> 3];
        if (t < 0) break;

        // H2: softmax per sample, mean over samples (fast-math allowed here).  Threads past the last anchor of a
        // level's last tile run along on whatever their shared-memory rows hold (nothing of theirs is stored).
        // Nothing but the ring state and the sums is live across this loop (the tile's coordinates are worked out
        // behind it): at 56 registers per thread anything else is spilled or recomputed in every iteration.
        SoftmaxSum<K> acc;
        acc.clear();
        if constexpr (NSU > 0) {
#pragma unroll 1
            for (int r = N; r > 0; r -= NSU) {
#pragma unroll
                for (int s = 0; s < NSU; ++s) {
                    // (the first stage of a tile was waited for above: asking again returns at once, its phase cannot
                    // advance before this warp has released the stage)
                    mbar_wait_a(full0 + 8u * s, phase);
                    float x[K];
                    LoadRow<K>::run(row0 + kSlabBytes * s, x);
                    acc.add_row(x);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(full0 + empty_minus_full + 8u * s);
                }
                phase ^= 1u;
            }
        } else {
#pragma unroll 1
            for (int n = N; n > 0; --n) {
                if (n < N) mbar_wait_a(bar, phase);
#ifdef BOD_DIAGNOSTICS
                if (a.debug < 2)
#endif
                {
                    float x[K];
                    LoadRow<K>::run(row_addr, x);
                    acc.add_row(x);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(bar + empty_minus_full);    // this warp is done with the stage
                row_addr += kSlabBytes; bar += 8;
                if (bar == bar_end) { bar = full0; row_addr = row0; phase ^= 1u; }
            }
        }

        const int b = t / tiles, tile = t - b * tiles;
        const TileRef tr = tile_ref(a.lv, tile);
        const int a0 = tile * kTileAnchors;                  // first slot of the tile
        const int rows = min(kTileAnchors, tr.A_l - tr.a0);
        const int anchor = tr.anchor0 + tid;
        const bool valid = tid < rows;

        float cnt[K];
#pragma unroll
        for (int k = 0; k < K; ++k) cnt[k] = 0.0f;
        if (a.counts_in != nullptr && valid) {              // parity mode: counts to inject
            const float* c = a.counts_in + ((size_t)b * a.A + anchor) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) cnt[k] = __ldg(c + k);
        }
        float p[K];
        acc.mean(invN, p);
        if (valid && a.probs_out != nullptr) {
            float* o = a.probs_out + ((size_t)b * a.A + anchor) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) o[k] = p[k];
        }

#ifdef BOD_DIAGNOSTICS
        if (a.debug >= 1) {                                  // diagnostic builds: data movement (+ softmax) only
            if (valid && p[0] == 12345.0f) a.slot_anchor[0] = 1;
            continue;
        }
#endif

        // H3: categorical draw counts
        bool maybe_fg = true;
        if (a.counts_in == nullptr && valid) {
            maybe_fg = philox_counts<K>(p, (uint32_t)anchor, a.image_id_base + (uint32_t)b,
                                        make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)), a.num_draws, a.ratio_slot,
                                        a.sampled_out != nullptr, cnt);
            if (a.sampled_out != nullptr) {
                float* o = a.sampled_out + ((size_t)b * a.A + anchor) * K;
#pragma unroll
                for (int k = 0; k < K; ++k) o[k] = cnt[k];
            }
        }

        // H4: first-maximum argmax != background, stable compaction inside the tile
        bool keep = false;
        if (valid && maybe_fg) {
            int am = 0;
            float best = cnt[0];
#pragma unroll
            for (int k = 1; k < K; ++k) if (cnt[k] > best) { best = cnt[k]; am = k; }
            keep = (am != K - 1);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        int* wc = warp_count[tcount & 1];                    // double-buffered: one barrier per tile
        if (lane == 0) wc[warp] = __popc(ballot);
        consumer_barrier();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) {
            const int c = wc[w];
            base += (w < warp) ? c : 0;
            total += c;
        }
        if (keep) {
            const int slot = a0 + base + __popc(ballot & ((1u << lane) - 1u));   // per-tile slot region
            a.slot_anchor[(size_t)b * tiles * kTileAnchors + slot] = anchor;
            float* o = a.slot_counts + ((size_t)b * tiles * kTileAnchors + slot) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) o[k] = cnt[k];
        }
        if (tid == 0) a.tile_count[(size_t)b * tiles + tile] = total;
    }
    launch_clock_end(a.clk);
}

static bool k1_aligned(const K1Args& a) {
    // bulk copies need 16-byte aligned sources and sizes for every (level, image, sample, tile)
    if ((kTileAnchors * a.K) % 4 != 0) return false;
    for (int l = 0; l < a.lv.n; ++l)
        if ((reinterpret_cast<uintptr_t>(a.lv.cls[l]) & 15u) != 0 || ((size_t)a.lv.rows[l] * a.K) % 4 != 0 ||
            ((size_t)a.lv.row0[l] * a.K) % 4 != 0)
            return false;
    return true;
}
static int k1_ctas_per_sm() {
    static int v = 0;
    if (v == 0) {
        v = kPipeCtasPerSM;
        if (const char* e = getenv("BOD_K1_CTAS")) { const int x = atoi(e); if (x >= 1 && x <= kPipeCtasPerSM) v = x; }
    }
    return v;
}
static int k1_pipe_ctas(const K1Args& a) {
    int sms = sm_count();
    static const int sms_env = getenv("BOD_K1_SMS") ? atoi(getenv("BOD_K1_SMS")) : 0;                      // experiment: leave SMs free
    if (sms_env >= 1 && sms_env < sms) sms = sms_env;
    int ctas = k1_ctas_per_sm() * sms;
    if (ctas > a.B * a.tiles) ctas = a.B * a.tiles;
    return ctas;
}
uint32_t k1_tickets_per_launch(const K1Args& a) {
    return k1_aligned(a) ? (uint32_t)(a.B * a.tiles + k1_pipe_ctas(a)) : 0u;
}

// The pipeline kernel for (K, ring): rounds of NSU samples unrolled for the class counts and ring depths of the
// BASELINE configurations, the generic ring walk (NSU = 0) otherwise.
template <int K>
static const void* k1_pipe_func(int nsu) {
    if constexpr (K == 4 || K == 8 || K == 11) {
        switch (nsu) {
            case 3: return (const void*)k1_moments_pipe_kernel<K, 3>;
            case 4: return (const void*)k1_moments_pipe_kernel<K, 4>;
            case 5: return (const void*)k1_moments_pipe_kernel<K, 5>;
            case 6: return (const void*)k1_moments_pipe_kernel<K, 6>;
            case 8: return (const void*)k1_moments_pipe_kernel<K, 8>;
            case 10: return (const void*)k1_moments_pipe_kernel<K, 10>;
            default: break;
        }
    }
    return (const void*)k1_moments_pipe_kernel<K, 0>;
}
struct K1Plan { const void* fn; int NS; size_t ring; };
template <int K>
static K1Plan k1_plan(const K1Args& a) {
    const size_t slab = (size_t)kTileAnchors * K * sizeof(float);
    // Ring depth: all the shared memory of the SM when K1 runs alone.  In a pipelined context one posterior CTA
    // (20.5 KB + 1 KB) and one fusion CTA (26.5 KB + 1 KB) have to fit beside the resident K1 CTAs (each: ring +
    // 0.75 KB static + 1 KB reserved) in the SM's 228 KB.
    const unsigned per_cta = a.leave_room ? (228u * 1024u - 50u * 1024u) / k1_ctas_per_sm() - 1792u
                                          : 216u * 1024u / k1_ctas_per_sm() - 1024u;
    int NS = (int)(per_cta / slab);
    // Pipelined contexts: at most five slabs (>= 10 KB) in flight per CTA.  A deeper ring makes the kernel faster alone and
    // the step slower: with K = 4 (2 KB slabs) ten stages stream at 5.1 TB/s alone, but the posterior / soft-NMS kernels
    // of earlier runs then wait on a saturated memory system while they hold their SMs (KITTI shape, B = 64: 1.00 ms per
    // step with ten stages, 0.86 ms with five; measured, round 2).
    if (a.leave_room) { const int cap = (int)(10240 / slab) > 5 ? (int)(10240 / slab) : 5; if (NS > cap) NS = cap; }
    static const int ns_env = getenv("BOD_K1_NS") ? atoi(getenv("BOD_K1_NS")) : 0;                          // experiment: shallower ring
    if (ns_env >= 2 && ns_env < NS) NS = ns_env;
    if (NS > kMaxStages) NS = kMaxStages;
    if (NS < 2) NS = 2;
    int nsu = 0;
    static const bool unroll = !(getenv("BOD_K1_UNROLL") && atoi(getenv("BOD_K1_UNROLL")) == 0);           // experiment: generic walk
    if (unroll && (K == 4 || K == 8 || K == 11)) {
        static const int cand[] = {10, 8, 6, 5, 4, 3};
        for (int c : cand) if (c <= NS && a.N % c == 0) { nsu = c; break; }
    }
    if (nsu > 0) NS = nsu;
    return K1Plan{k1_pipe_func<K>(nsu), NS, (size_t)NS * slab};
}

template <int K>
static cudaError_t launch_k(const K1Args& a, cudaStream_t st) {
    // samples per ring stage: two stages of <= ~50 KB keep two CTAs resident per SM
    const size_t slab = (size_t)kTileAnchors * K * sizeof(float);
    int NC = (int)((50u * 1024u) / slab);
    if (NC < 1) NC = 1;
    if (NC > a.N) NC = a.N;
    const size_t smem = 2 * (size_t)NC * slab;
    // bulk copies need 16-byte aligned sources and sizes for every (image, sample, tile)
    const bool aligned = k1_aligned(a);
    dim3 grid(a.tiles, a.B), block(kTileAnchors);
    cudaError_t e;
    if (aligned) {
        // persistent pipeline: kPipeCtasPerSM CTAs per SM, each with a ring of NS one-sample slabs
        const K1Plan pl = k1_plan<K>(a);
        int NS = pl.NS;
        const int ctas = k1_pipe_ctas(a);
        e = ensure_dyn_smem(pl.fn, pl.ring);
        if (e != cudaSuccess) return e;
        K1Args aa = a;
        void* args[2] = {&aa, &NS};
        return cudaLaunchKernel(pl.fn, dim3(ctas), dim3(kTileAnchors + 32), args, pl.ring, st);
    } else {
        e = ensure_dyn_smem((const void*)k1_moments_kernel<K, false>, smem);
        if (e != cudaSuccess) return e;
        k1_moments_kernel<K, false><<<grid, block, smem, st>>>(a, NC);
    }
    return cudaGetLastError();
}

// slot of the ratio table for T draws on the current device (-1: none left, the kernel divides)
static int ratio_slot_for(int T, cudaError_t* err) {
    static std::mutex mu;
    static int slot_T[64][kRatioSlots];
    static int slot_n[64] = {0};
    *err = cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return -1;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < slot_n[dev]; ++i) if (slot_T[dev][i] == T) return i;
    if (slot_n[dev] == kRatioSlots) return -1;
    float tab[kRatioTab];
    for (int j = 0; j < kRatioTab; ++j) {
        volatile float num = (float)(T - j), den = (float)(j + 1);
        tab[j] = num / den;
    }
    const int slot = slot_n[dev];
    // synchronous copy into a slot no kernel has been told about yet
    *err = cudaMemcpyToSymbol(c_binom_ratio, tab, sizeof tab, (size_t)slot * sizeof tab, cudaMemcpyHostToDevice);
    if (*err != cudaSuccess) return -1;
    slot_T[dev][slot] = T;
    slot_n[dev] = slot + 1;
    return slot;
}

cudaError_t launch_k1(const K1Args& a0, cudaStream_t st) {
    K1Args a = a0;
    cudaError_t e0;
    a.ratio_slot = ratio_slot_for(a.num_draws, &e0);
    if (e0 != cudaSuccess) return e0;
    switch (a.K) {
#define BOD_CASE(KK) case KK: return launch_k<KK>(a, st);
        BOD_CASE(2) BOD_CASE(3) BOD_CASE(4) BOD_CASE(5) BOD_CASE(6) BOD_CASE(7) BOD_CASE(8) BOD_CASE(9)
        BOD_CASE(10) BOD_CASE(11) BOD_CASE(12) BOD_CASE(13) BOD_CASE(16) BOD_CASE(21) BOD_CASE(32)
#undef BOD_CASE
        default: return cudaErrorInvalidValue;
    }
}

// ---- CUDA-graph support: which kernel launch_k1 launches for these arguments, and how to point a captured
// launch of it at new arguments (same shapes; other input tensors) ----
template <int K>
static const void* k1_func_k(const K1Args& a) {
    return k1_aligned(a) ? k1_plan<K>(a).fn : (const void*)k1_moments_kernel<K, false>;
}
const void* k1_kernel_func(const K1Args& a) {
    switch (a.K) {
#define BOD_CASE(KK) case KK: return k1_func_k<KK>(a);
        BOD_CASE(2) BOD_CASE(3) BOD_CASE(4) BOD_CASE(5) BOD_CASE(6) BOD_CASE(7) BOD_CASE(8) BOD_CASE(9)
        BOD_CASE(10) BOD_CASE(11) BOD_CASE(12) BOD_CASE(13) BOD_CASE(16) BOD_CASE(21) BOD_CASE(32)
#undef BOD_CASE
        default: return nullptr;
    }
}
cudaError_t k1_graph_update(cudaGraphExec_t exec, cudaGraphNode_t node, const K1Args& a0) {
    K1Args a = a0;
    cudaError_t e;
    a.ratio_slot = ratio_slot_for(a.num_draws, &e);
    if (e != cudaSuccess) return e;
    cudaKernelNodeParams p;
    e = cudaGraphKernelNodeGetParams(node, &p);
    if (e != cudaSuccess) return e;
    if (p.func != k1_kernel_func(a)) return cudaErrorInvalidValue;      // alignment class changed: recapture
    void* args[2] = {&a, p.kernelParams[1]};                            // (K1Args, ring depth / samples per stage)
    p.kernelParams = args;
    return cudaGraphExecKernelNodeSetParams(exec, node, &p);
}

bool k1_supports(int K) {
    switch (K) {
        case 2: case 3: case 4: case 5: case 6: case 7: case 8: case 9: case 10: case 11: case 12: case 13:
        case 16: case 21: case 32: return true;
        default: return false;
    }
}

}  // namespace bod
