// Dense layers of the actor / critic / discriminator on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
// Reference: the nn.Linear stacks of pacer/pacer/learning/amp_network_sept_builder.py:69-111, amp_network_builder.py:81-84
// (SURVEY 8a rows a11-a13), which the reference runs as fp32 cuBLAS GEMMs (mixed_precision: False).
//
// PRECISION.  north_star asks for 1e-3 relative parity with the fp32 reference.  One bf16 pass (8-bit mantissa) gives
// ~4e-3 and TF32 ~7e-4 after three layers - neither clears the bar element-wise - so every fp32 operand is split into
// two bf16 terms  x = hi + lo  (hi = bf16(x), lo = bf16(x - hi), 16 mantissa bits together) and the product is formed as
//      A*W ~= Ah*Wh + Ah*Wl + Al*Wh          (the dropped Al*Wl term is 2^-16 relative)
// i.e. THREE kind::f16 MMAs per k-step into one fp32 TMEM accumulator ("bf16x3").  Measured error vs fp32: ~1e-5 relative.
//
// KERNEL.  Persistent, warp-specialised, one CTA per SM:
//   warp 0 (one lane)  TMA producer: four 128B-swizzled K-major tiles per stage (Ah, Al: 128 x 64; Wh, Wl: BN x 64 bf16),
//                      out-of-range rows / K-tail zero-filled by the TMA unit
//   warp 1 (one lane)  MMA issuer: 4 k-steps x 3 tcgen05.mma (M=128, N=BN, K=16) per stage, tcgen05.commit frees the
//                      stage and, after the last k-block, publishes the accumulator
//   warps 2-5          epilogue: tcgen05.ld 32 lanes x 32 columns -> +bias, ReLU -> fp32 rows and/or the bf16 hi/lo split
//                      that is the NEXT layer's A operand (the intermediate activations never exist as fp32 in HBM)
//   TMEM: two BN-column accumulators, so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "sim.h"

namespace tc {

constexpr int BM = 128, UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;                           // two per TMEM lane quadrant, each takes half of the tile's columns
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;       // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

// Two tile shapes.  Operand rows are K-major and span exactly one swizzle atom along K:
//   BN = 128: BK = 64 (128-byte swizzle), 3 stages x 64 KB      - small-N layers and the heads
//   BN = 256: BK = 32 ( 64-byte swizzle), 4 stages x 48 KB      - the wide layers: 1.33x the flops per smem/L2 byte
template <int BN> struct Tile {
    static constexpr int BK = BN == 128 ? 64 : 32;
    static constexpr int STAGES = BN == 128 ? 3 : 4;
    static constexpr int SWIZZLE_BYTES = BK * 2;           // 128 or 64
    static constexpr int A_BYTES = BM * BK * 2;            // one bf16 tile of A (hi or lo)
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int BIAS_BYTES = 2 * BN * 4;          // bias slice of the tile, double-buffered with the accumulator
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + 2 * BIAS_BYTES;   // + fused-head weights
    static constexpr int TMEM_COLS = 2 * BN;               // power of two: 256 or 512
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// K-major operand tile whose rows are one swizzle atom wide (SW = 128 or 64 bytes): 8-row atoms are 8*SW bytes apart (SBO);
// LBO unused.  Bit layout: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1 (sm_100),
// [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B.
template <int SW> __device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((8 * SW) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(SW == 128 ? 2 : 4) << 61;
    return d;
}

// instruction descriptor: D fp32 ([4,6) = 1), A and B bf16 ([7,10) = [10,13) = 1), both K-major, N>>3 at [17,23), M>>4 at [24,29)
template <int BN> __device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                 " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
// wait for the thread's outstanding tcgen05.ld; the loaded registers are listed as in/out operands so that the compiler
// cannot schedule a use of them above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

struct GemmArgs {
    const float* bias;       // [N] or NULL
    float* y32; long long ldy;                       // fp32 output (optional)
    __nv_bfloat16* y_hi; __nv_bfloat16* y_lo; long long ldy16;   // split output (optional; N % 32 == 0)
    int M, N, K, relu;
    const int* m_dev;        // optional device-side row count (<= M): tiles beyond it are skipped (compacted row sets)
    // fused single-output head (value / logit layers, N_head = 1): head_part[row][n / 64] = sum over the 64-column group of
    // act(y[row][n]) * head_w[n]; the caller adds the groups in order (+ bias).  The layer's own output may then be omitted.
    const float* head_w; float* head_part; int head_ld;
    // split-K (skinny layers whose tile count is far below the SM count): unit u = (tile, split) covers k-blocks
    // [split * kb_per, ...); split s writes its partial sums to y32 + s * split_stride (bias in split 0 only); the consumer adds them
    int k_splits; long long split_stride;
    // tile order: 0 = consecutive work units walk down M (they share a W tile), 1 = they walk across N (they share an A row
    // block).  When A is much larger than the L2 (the 131 072-row post-horizon pass: 1.6 GB) the M-fast order streams it from
    // DRAM once per N tile (6.8 GB read, ncu); N-fast keeps the co-scheduled CTAs on the same rows
    int n_fast;
};

// One 32-column chunk of one accumulator row: +bias, ReLU, then fp32 store and/or bf16 hi/lo split store.
__device__ __forceinline__ float epilogue_chunk(const uint32_t* r, const float* __restrict__ s_bias, const float* __restrict__ s_head,
                                                int nb, int row, bool row_ok, const GemmArgs& g) {
    if (nb >= g.N) return 0.f;                                              // warp-uniform
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + j);     // same address in every lane: broadcast
        float x0 = __uint_as_float(r[j]) + b4.x, x1 = __uint_as_float(r[j + 1]) + b4.y;
        float x2 = __uint_as_float(r[j + 2]) + b4.z, x3 = __uint_as_float(r[j + 3]) + b4.w;
        v[j] = g.relu ? fmaxf(x0, 0.f) : x0; v[j + 1] = g.relu ? fmaxf(x1, 0.f) : x1;
        v[j + 2] = g.relu ? fmaxf(x2, 0.f) : x2; v[j + 3] = g.relu ? fmaxf(x3, 0.f) : x3;
    }
    float hd = 0.f;
    if (g.head_w) {                                                         // head weights are zero beyond N
#pragma unroll
        for (int j = 0; j < 32; ++j) hd += v[j] * s_head[j];
    }
    if (!row_ok) return hd;
    if (g.y32) {
        float* o = g.y32 + (long long)row * g.ldy + nb;
        if (nb + 32 <= g.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nb + j < g.N) o[j] = v[j];
        }
    }
    if (g.y_hi) {                                                           // next layer's operand: x = hi + lo
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            __nv_bfloat16 h0 = __float2bfloat16_rn(v[j]), h1 = __float2bfloat16_rn(v[j + 1]);
            __nv_bfloat16 l0 = __float2bfloat16_rn(v[j] - __bfloat162float(h0));
            __nv_bfloat16 l1 = __float2bfloat16_rn(v[j + 1] - __bfloat162float(h1));
            ph[j / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            pl[j / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        uint4* oh = reinterpret_cast<uint4*>(g.y_hi + (long long)row * g.ldy16 + nb);
        uint4* ol = reinterpret_cast<uint4*>(g.y_lo + (long long)row * g.ldy16 + nb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            oh[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
            ol[j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
        }
    }
    return hd;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_bf16x3_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                     const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl, GemmArgs g) {
    using T = Tile<BN>;
    constexpr int BK = T::BK, SW = T::SWIZZLE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // swizzled tiles need 1024 B alignment
    const uint32_t bars = base + T::STAGES * T::STAGE_BYTES;                // 8-byte mbarriers
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (T::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * T::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * T::STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * T::STAGES + 4);
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
    float* s_bias = reinterpret_cast<float*>(smem_gen + T::STAGES * T::STAGE_BYTES + T::BAR_BYTES);   // [2][BN]
    float* s_head = s_bias + 2 * BN;                                                                 // [2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // prologue: touches nothing a preceding kernel produces, so with programmatic dependent launch it overlaps that kernel's tail
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl));
        for (int s = 0; s < T::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 32 * NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(T::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
    // everything below reads what earlier kernels wrote (operands, the device-side row count): wait for them to finish and flush
    // (a no-op without the launch attribute), then let the next kernel of the stream start ITS prologue
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (g.m_dev) { const int m = __ldg(g.m_dev); g.M = m < g.M ? m : g.M; }     // uniform: every thread reads the same word
    const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
    const int num_kb = (g.K + BK - 1) / BK;
    const int splits = g.k_splits > 1 ? g.k_splits : 1;
    const int kb_per = (num_kb + splits - 1) / splits;
    const int mn_tiles = tiles_m * tiles_n;
    const int num_tiles = mn_tiles * splits;                    // work units: (output tile, K split)

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < num_tiles; u += gridDim.x) {
                const int t = u % mn_tiles, sp = u / mn_tiles;
                const int m0 = (g.n_fast ? t / tiles_n : t % tiles_m) * BM, n0 = (g.n_fast ? t % tiles_n : t / tiles_m) * BN;
                const int kb0 = sp * kb_per, kb1 = min(num_kb, kb0 + kb_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = base + stage * T::STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), T::STAGE_BYTES);
                    tma_load_2d(sa, &map_ah, full_bar(stage), kb * BK, m0);
                    tma_load_2d(sa + T::A_BYTES, &map_al, full_bar(stage), kb * BK, m0);
                    tma_load_2d(sa + 2 * T::A_BYTES, &map_wh, full_bar(stage), kb * BK, n0);
                    tma_load_2d(sa + 2 * T::A_BYTES + T::B_BYTES, &map_wl, full_bar(stage), kb * BK, n0);
                    if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc<BN>();
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < num_tiles; u += gridDim.x) {
                const int kb0 = (u / mn_tiles) * kb_per, kb1 = min(num_kb, kb0 + kb_per);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);                  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * T::STAGE_BYTES;
                    const uint64_t d_ah = make_smem_desc<SW>(sa), d_al = make_smem_desc<SW>(sa + T::A_BYTES);
                    const uint64_t d_wh = make_smem_desc<SW>(sa + 2 * T::A_BYTES);
                    const uint64_t d_wl = make_smem_desc<SW>(sa + 2 * T::A_BYTES + T::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t ko = (uint64_t)((k * UMMA_K * 2) >> 4);   // advance inside the swizzle row
                        umma_bf16(d_tmem, d_al + ko, d_wh + ko, idesc, ((kb - kb0) | k) != 0);   // small terms first
                        umma_bf16(d_tmem, d_ah + ko, d_wl + ko, idesc, 1);
                        umma_bf16(d_tmem, d_ah + ko, d_wh + ko, idesc, 1);
                    }
                    umma_commit(empty_bar(stage));                          // frees the smem stage when the MMAs retire
                    if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));                                // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: warps 2..9; warp w reads TMEM lane quadrant w % 4, column half (w - 2) / 4 ==========
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int et = threadIdx.x - 64;                                    // 0..255
        constexpr int NC = BN / 64;                                         // 32-column chunks per warp (2 or 4)
        int acc = 0; uint32_t acc_phase = 0;
        float* const y32_base = g.y32;
        for (int u = blockIdx.x; u < num_tiles; u += gridDim.x) {
            const int t = u % mn_tiles, sp = u / mn_tiles;
            const int m0 = (g.n_fast ? t / tiles_n : t % tiles_m) * BM, n0 = (g.n_fast ? t % tiles_n : t / tiles_m) * BN;
            if (splits > 1) g.y32 = y32_base + (long long)sp * g.split_stride;          // this split's partial matrix
            // bias slice of this tile -> smem (one element per epilogue thread), visible after the epilogue-only barrier
            if (et < BN) {
                s_bias[acc * BN + et] = (g.bias && sp == 0 && n0 + et < g.N) ? __ldg(g.bias + n0 + et) : 0.f;
                if (g.head_w) s_head[acc * BN + et] = n0 + et < g.N ? __ldg(g.head_w + n0 + et) : 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < g.M;
            const int c0 = half * (BN / 2);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c0;
            const float* sb = s_bias + acc * BN + c0;
            const float* sh = s_head + acc * BN + c0;
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr, ra);
            tmem_ld_wait(ra);
#pragma unroll
            for (int c = 0; c < NC; c += 2) {                               // two register buffers: the next TMEM load is in
                tmem_ld32(taddr + (c + 1) * 32, rb);                        // flight while this chunk is converted and stored
                float hd = epilogue_chunk(ra, sb + c * 32, sh + c * 32, n0 + c0 + c * 32, row, row_ok, g);
                tmem_ld_wait(rb);
                if (c + 2 < NC) tmem_ld32(taddr + (c + 2) * 32, ra);
                hd += epilogue_chunk(rb, sb + (c + 1) * 32, sh + (c + 1) * 32, n0 + c0 + (c + 1) * 32, row, row_ok, g);
                if (c + 2 < NC) tmem_ld_wait(ra);
                const int grp = (n0 + c0 + c * 32) >> 6;                    // the two chunks are one 64-column group
                if (g.head_part && row_ok && grp < g.head_ld) g.head_part[(long long)row * g.head_ld + grp] = hd;
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));                                   // 256 arrivals release the accumulator
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T::TMEM_COLS));
    }
}

// =====================================================================================================================
// CTA-PAIR variant (cta_group::2): two CTAs of a cluster on the two SMs of a TPC compute one 256 x BN tile.  Each CTA stages
// its own 128 rows of A and HALF of the W tile (BN/2 rows); the leader's single thread issues M = 256 MMAs that read A from
// both SMs and W halves from both SMs.  Per SM this halves the W bytes that come through L2 and shared memory - the resource
// that bounds the bf16x3 kernel (three MMAs read six operand tiles per k-step) - so the tensor pipe can run near its peak.
//   barriers: full[s]   leader's, 2 arrivals (both producers) + the bytes of both CTAs' TMA loads
//             empty[s]  one per CTA, released by a multicast tcgen05.commit
//             tfull[a]  one per CTA (multicast commit), tempty[a] leader's, 2 x 256 epilogue threads
// =====================================================================================================================
template <int BN> struct Tile2 {
    static constexpr int BK = 64;
    static constexpr int BH = BN / 2;                      // W rows staged by each CTA
    static constexpr int STAGES = BN == 256 ? 3 : 4;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BH * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 64 KB (BN 256) / 48 KB (BN 128)
    static constexpr int BAR_BYTES = 256;
    static constexpr int BIAS_BYTES = 2 * BN * 4;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + BAR_BYTES + BIAS_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t num_clusters_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {     // acquire at cluster scope
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAITC_DONE;\n\t"
        "bra WAITC_LOOP;\n\t"
        "WAITC_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {      // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
linear_bf16x3_2cta_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                          const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl, GemmArgs g) {
    using T = Tile2<BN>;
    constexpr int BK = T::BK;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + T::STAGES * T::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (T::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * T::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * T::STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * T::STAGES + 4);
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
    float* s_bias = reinterpret_cast<float*>(smem_gen + T::STAGES * T::STAGE_BYTES + T::BAR_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (g.m_dev) { const int m = __ldg(g.m_dev); g.M = m < g.M ? m : g.M; }
    const int tiles_m = (g.M + 2 * BM - 1) / (2 * BM), tiles_n = (g.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (g.K + BK - 1) / BK;
    const int first_tile = (int)cluster_id_x(), tile_step = (int)num_clusters_x();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl));
        for (int s = 0; s < T::STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * 32 * NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(T::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();                                   // barriers of both CTAs initialised before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer (both CTAs) =====================
            int stage = 0; uint32_t phase = 0;
            for (int t = first_tile; t < num_tiles; t += tile_step) {
                const int m0 = (t % tiles_m) * 2 * BM + (int)rank * BM, n0 = (t / tiles_m) * BN + (int)rank * T::BH;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_cluster(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = base + stage * T::STAGE_BYTES;
                    const uint32_t lfull = mapa_u32(full_bar(stage), 0);             // the LEADER's full barrier
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * T::STAGE_BYTES); // bytes of both CTAs land on it
                    else mbar_arrive_cluster(lfull);
                    tma_load_2d_2sm(sa, &map_ah, lfull, kb * BK, m0);
                    tma_load_2d_2sm(sa + T::A_BYTES, &map_al, lfull, kb * BK, m0);
                    tma_load_2d_2sm(sa + 2 * T::A_BYTES, &map_wh, lfull, kb * BK, n0);
                    tma_load_2d_2sm(sa + 2 * T::A_BYTES + T::B_BYTES, &map_wl, lfull, kb * BK, n0);
                    if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ===================== MMA issuer (leader CTA only): M = 256 across the pair =====================
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = first_tile; t < num_tiles; t += tile_step) {
                mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1);          // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_cluster(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * T::STAGE_BYTES;
                    const uint64_t d_ah = make_smem_desc<128>(sa), d_al = make_smem_desc<128>(sa + T::A_BYTES);
                    const uint64_t d_wh = make_smem_desc<128>(sa + 2 * T::A_BYTES);
                    const uint64_t d_wl = make_smem_desc<128>(sa + 2 * T::A_BYTES + T::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t ko = (uint64_t)((k * UMMA_K * 2) >> 4);
                        umma_bf16_2sm(d_tmem, d_al + ko, d_wh + ko, idesc, (kb | k) != 0);
                        umma_bf16_2sm(d_tmem, d_ah + ko, d_wl + ko, idesc, 1);
                        umma_bf16_2sm(d_tmem, d_ah + ko, d_wh + ko, idesc, 1);
                    }
                    umma_commit_2sm(empty_bar(stage));                      // frees the stage in BOTH CTAs
                    if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(tfull_bar(acc));                            // accumulators complete in both CTAs
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (both CTAs; each drains its own 128 rows) =====================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int et = threadIdx.x - 64;
        constexpr int NC = BN / 64;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = first_tile; t < num_tiles; t += tile_step) {
            const int m0 = (t % tiles_m) * 2 * BM + (int)rank * BM, n0 = (t / tiles_m) * BN;
            if (et < BN) s_bias[acc * BN + et] = (g.bias && n0 + et < g.N) ? __ldg(g.bias + n0 + et) : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
            mbar_wait_cluster(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < g.M;
            const int c0 = half * (BN / 2);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c0;
            const float* sb = s_bias + acc * BN + c0;
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr, ra);
            tmem_ld_wait(ra);
#pragma unroll
            for (int c = 0; c < NC; c += 2) {
                tmem_ld32(taddr + (c + 1) * 32, rb);
                epilogue_chunk(ra, sb + c * 32, sb, n0 + c0 + c * 32, row, row_ok, g);      // (no fused head in the pair kernels)
                tmem_ld_wait(rb);
                if (c + 2 < NC) tmem_ld32(taddr + (c + 2) * 32, ra);
                epilogue_chunk(rb, sb + (c + 1) * 32, sb, n0 + c0 + (c + 1) * 32, row, row_ok, g);
                if (c + 2 < NC) tmem_ld_wait(ra);
            }
            tc_fence_before();
            mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0));              // 512 arrivals on the leader's barrier
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                   // nobody leaves while the pair's MMAs / remote arrives may be in flight
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T::TMEM_COLS));
    }
}

// =====================================================================================================================
// LAYER CHAIN: every dense layer of one network pass (task MLP -> actor / critic -> heads, discriminator) in ONE persistent
// kernel (amp_network_sept_builder.py:69-111 evaluated top to bottom without leaving the SMs).
//   * work unit = one 128 x 128 output tile of one layer; the host lists the tiles of all layers in a dependency-respecting
//     order (segments); CTAs claim tickets from a global counter, so the 148 SMs stay busy across layer boundaries and
//     across layers of different depth - the wave quantisation of one launch per layer (4096 rows = 128 / 256 / 512 tiles on
//     148 SMs) and the launch gaps between them disappear
//   * a layer's A operand is the bf16 hi/lo output the previous layer's epilogues wrote to global memory (L2 resident);
//     the producer waits on a per-(layer, 128-row block) counter that the epilogues bump after their stores
//     (st.global -> fence -> atomic;  ld.acquire -> fence.proxy.async -> TMA load)
//   * single-output heads (value / logit): the CTA that completes a row block adds the per-64-column partial sums in
//     column order (same arithmetic as head_reduce_kernel)
//   * deadlock freedom: tickets are claimed in order and every tile only waits on tiles with lower tickets, which are held by
//     CTAs that are resident and never wait on higher tickets
// Tile shape / pipeline = linear_bf16x3_kernel<128>; results are bit-identical to the per-layer launches.
// =====================================================================================================================
constexpr int CHAIN_MAX_LAYERS = 12, CHAIN_MAX_SEGS = 32, CHAIN_RING = 4;

struct ChainLayer {
    GemmArgs g;
    const float* head_bias; float* head_out;
    int dep;          // layer whose output rows are this layer's A operand (-1: operands complete at launch)
    int need;         // tiles that complete one 128-row block of `dep`
    int tiles_n, num_kb;
    int cnt_off;      // this layer's row-block counters in the workspace
};
struct alignas(64) ChainParams {
    CUtensorMap maps[CHAIN_MAX_LAYERS][6];     // A hi, A lo, W hi, W lo (loads); y hi, y lo (stores, 32 x 32 boxes)
    ChainLayer L[CHAIN_MAX_LAYERS];
    int seg_layer[CHAIN_MAX_SEGS], seg_first[CHAIN_MAX_SEGS], seg_end[CHAIN_MAX_SEGS];   // tickets [seg_end[i-1], seg_end[i])
    int n_segs, total, ws_ints;
    int* ws;          // [0] next ticket, [1] CTAs that have left, [2..) row-block counters; all zero between launches
    int dbg;          // profiling experiments (EMLOCO_CHAIN_DBG): 1 = do not issue the output TMA stores, 2 = no staging either
    long long* trace; // optional (profiling): per ticket {cta, layer << 24 | tile, t claimed, t rows ready, t accumulator ready, t stored, t MMA thread free, t first k-block landed} (globaltimer ns)
};

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

constexpr int CHAIN_STAGE_OUT = 4096;          // per epilogue warp: 32 rows x 64 B of hi, then of lo (64B-swizzled TMA boxes)

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// epilogue_chunk of the chain kernel: same arithmetic; the split output leaves through shared memory and two TMA stores per
// 32 x 32 block (full 64-byte row segments) instead of 16-byte pieces in 32 different rows per store instruction - the
// per-thread stores made the epilogue of a 128 x 128 tile take 6.4 us, as long as a 10-k-block main loop (chain trace).
__device__ __forceinline__ float chain_epilogue_chunk(const uint32_t* r, const float* __restrict__ s_bias, const float* __restrict__ s_head,
                                                      int nb, int row, int row0, bool row_ok, const GemmArgs& g, uint8_t* stage,
                                                      uint32_t stage_u32, const CUtensorMap* map_yh, const CUtensorMap* map_yl,
                                                      int lane, bool& stores_in_flight, int dbg = 0) {
    if (nb >= g.N) return 0.f;                                              // warp-uniform
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + j);
        float x0 = __uint_as_float(r[j]) + b4.x, x1 = __uint_as_float(r[j + 1]) + b4.y;
        float x2 = __uint_as_float(r[j + 2]) + b4.z, x3 = __uint_as_float(r[j + 3]) + b4.w;
        v[j] = g.relu ? fmaxf(x0, 0.f) : x0; v[j + 1] = g.relu ? fmaxf(x1, 0.f) : x1;
        v[j + 2] = g.relu ? fmaxf(x2, 0.f) : x2; v[j + 3] = g.relu ? fmaxf(x3, 0.f) : x3;
    }
    float hd = 0.f;
    if (g.head_w) {
#pragma unroll
        for (int j = 0; j < 32; ++j) hd += v[j] * s_head[j];
    }
    if (g.y32 && row_ok) {
        float* o = g.y32 + (long long)row * g.ldy + nb;
        if (nb + 32 <= g.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nb + j < g.N) o[j] = v[j];
        }
    }
    if (g.y_hi && dbg < 2) {                                                // warp-uniform
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            __nv_bfloat16 h0 = __float2bfloat16_rn(v[j]), h1 = __float2bfloat16_rn(v[j + 1]);
            __nv_bfloat16 l0 = __float2bfloat16_rn(v[j] - __bfloat162float(h0));
            __nv_bfloat16 l1 = __float2bfloat16_rn(v[j + 1] - __bfloat162float(h1));
            ph[j / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            pl[j / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        if (stores_in_flight) {                                             // the previous block's stores still read the buffer
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
        const int sw = (lane >> 1) & 3;                                     // 64-byte swizzle: 16-byte unit ^= (row / 2) % 4
        uint4* sh_hi = reinterpret_cast<uint4*>(stage + lane * 64);
        uint4* sh_lo = reinterpret_cast<uint4*>(stage + 2048 + lane * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sh_hi[j ^ sw] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
            sh_lo[j ^ sw] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic smem writes -> async-proxy (TMA) reads
        __syncwarp();
        if (lane == 0 && dbg == 0) {
            tma_store_2d(map_yh, stage_u32, nb, row0);                      // rows / columns outside [M, N] are clipped by the unit
            tma_store_2d(map_yl, stage_u32 + 2048, nb, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        stores_in_flight = true;
    }
    return hd;
}

__global__ void __launch_bounds__(NUM_THREADS, 1) linear_chain_kernel(const __grid_constant__ ChainParams P) {
    using T = Tile<128>;
    constexpr int BN = 128, BK = T::BK, SW = T::SWIZZLE_BYTES;
    // layout: 3 stages (192 KB) | barriers + ticket ring (256 B) | bias / head slices (2 KB) | pad to 3 KB | output staging (8 x 4 KB): 227 KB,
    // which leaves no room for alignment slack - the dynamic window of a kernel without static shared memory starts 1 KB-aligned
    extern __shared__ __align__(1024) uint8_t chain_smem[];
    uint8_t* const smem_raw = chain_smem;
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) __trap();
    const uint32_t bars = base + T::STAGES * T::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (T::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * T::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * T::STAGES + 2 + a); };
    auto sfull_bar = [&](int r) { return bars + 8u * (2 * T::STAGES + 4 + r); };                       // ring entry written
    auto sempty_bar = [&](int r) { return bars + 8u * (2 * T::STAGES + 4 + CHAIN_RING + r); };          // ring entry consumed
    constexpr uint32_t MISC = 8u * (2 * T::STAGES + 4 + 2 * CHAIN_RING);     // tmem slot, ticket ring, flag
    static_assert(MISC + 12 + 8 * CHAIN_RING <= T::BAR_BYTES, "barrier block");
    const uint32_t tmem_slot = bars + MISC;
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* misc = smem_gen + T::STAGES * T::STAGE_BYTES + MISC;
    volatile int* s_sched = reinterpret_cast<volatile int*>(misc + 8);                               // [CHAIN_RING]
    volatile int* s_last = reinterpret_cast<volatile int*>(misc + 8 + 4 * CHAIN_RING);
    volatile int* s_ticket = reinterpret_cast<volatile int*>(misc + 12 + 4 * CHAIN_RING);              // [CHAIN_RING] (trace only)
    float* s_bias = reinterpret_cast<float*>(smem_gen + T::STAGES * T::STAGE_BYTES + T::BAR_BYTES);   // [2][BN]
    float* s_head = s_bias + 2 * BN;                                                                 // [2][BN]
    constexpr uint32_t OUT_OFF = T::STAGES * T::STAGE_BYTES + 3072;                                   // 1 KB-aligned (the swizzle works on address bits)
    static_assert(T::BAR_BYTES + 2 * T::BIAS_BYTES <= 3072, "staging offset");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < T::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 32 * NUM_EPI_WARPS); }
        for (int r = 0; r < CHAIN_RING; ++r) { mbar_init(sfull_bar(r), 1); mbar_init(sempty_bar(r), 2); }   // MMA thread + one epilogue thread
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(T::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(misc);

    if (warp == 0) {
        if (lane == 0) {
            // ===================== scheduler + TMA producer =====================
            // (A separate scheduler warp that claims and dependency-checks tickets ahead of the loads was tried: the CTAs then
            // hoard up to a ring of tiles, the pass ends ragged - 159 instead of 149 us - and the tile period did not improve.)
            int stage = 0; uint32_t phase = 0;
            int slot = 0; uint32_t sphase = 0;
            int t = atomicAdd(P.ws, 1);
            for (;;) {
                int info = -1, layer = 0, tile = 0;
                if (t < P.total) {
                    int s = 0;
                    while (t >= P.seg_end[s]) ++s;
                    layer = P.seg_layer[s];
                    tile = P.seg_first[s] + t - (s ? P.seg_end[s - 1] : 0);
                    info = (layer << 24) | tile;
                }
                mbar_wait(sempty_bar(slot), sphase ^ 1);
                s_sched[slot] = info;
                s_ticket[slot] = t;
                mbar_arrive(sfull_bar(slot));                               // release: the ring entry is visible to the waiters
                long long* tr = (P.trace && info >= 0) ? P.trace + 8ll * t : nullptr;
                if (tr) { tr[0] = blockIdx.x; tr[1] = info; tr[2] = gtimer(); }
                if (++slot == CHAIN_RING) { slot = 0; sphase ^= 1; }
                if (info < 0) break;
                const ChainLayer& Ly = P.L[layer];
                const int m_blk = tile / Ly.tiles_n;
                const int m0 = m_blk * BM, n0 = (tile - m_blk * Ly.tiles_n) * BN;
                if (Ly.dep >= 0) {                                          // rows written by the previous layer's epilogues
                    const int* c = P.ws + P.L[Ly.dep].cnt_off + m_blk;
                    while (ld_acquire_gpu(c) < Ly.need) __nanosleep(32);
                    // (no proxy fence: the rows were written by TMA stores and are read by TMA loads - the same, async, proxy)
                }
                if (tr) tr[3] = gtimer();
                const CUtensorMap* mp = P.maps[layer];
                const int num_kb = Ly.num_kb;
                int t_next = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = base + stage * T::STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), T::STAGE_BYTES);
                    tma_load_2d(sa, mp + 0, full_bar(stage), kb * BK, m0);
                    tma_load_2d(sa + T::A_BYTES, mp + 1, full_bar(stage), kb * BK, m0);
                    tma_load_2d(sa + 2 * T::A_BYTES, mp + 2, full_bar(stage), kb * BK, n0);
                    tma_load_2d(sa + 2 * T::A_BYTES + T::B_BYTES, mp + 3, full_bar(stage), kb * BK, n0);
                    if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
                    // the next ticket is requested three k-blocks before the end: the atomic's round trip (0.7 us) overlaps the last
                    // loads (not earlier: a ticket held for a whole tile is a tile another CTA could have started)
                    if (kb == (num_kb > 3 ? num_kb - 3 : 0)) t_next = atomicAdd(P.ws, 1);
                }
                t = t_next;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc<BN>();
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int slot = 0; uint32_t sphase = 0;
            for (;;) {
                mbar_wait(sfull_bar(slot), sphase);
                const int info = s_sched[slot];
                long long* tr = (P.trace && info >= 0) ? P.trace + 8ll * s_ticket[slot] : nullptr;
                mbar_arrive(sempty_bar(slot));
                if (++slot == CHAIN_RING) { slot = 0; sphase ^= 1; }
                if (info < 0) break;
                const int num_kb = P.L[info >> 24].num_kb;
                // k-steps of the last k-block that hold columns below K: the rest of the box is zero fill and adds nothing
                const int last_steps = (P.L[info >> 24].g.K - (num_kb - 1) * BK + UMMA_K - 1) / UMMA_K;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                if (tr) tr[6] = gtimer();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (tr && kb == 0) tr[7] = gtimer();
                    const uint32_t sa = base + stage * T::STAGE_BYTES;
                    const uint64_t d_ah = make_smem_desc<SW>(sa), d_al = make_smem_desc<SW>(sa + T::A_BYTES);
                    const uint64_t d_wh = make_smem_desc<SW>(sa + 2 * T::A_BYTES);
                    const uint64_t d_wl = make_smem_desc<SW>(sa + 2 * T::A_BYTES + T::B_BYTES);
                    const int steps = kb == num_kb - 1 ? last_steps : BK / UMMA_K;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (k >= steps) break;
                        const uint64_t ko = (uint64_t)((k * UMMA_K * 2) >> 4);
                        umma_bf16(d_tmem, d_al + ko, d_wh + ko, idesc, (kb | k) != 0);      // same order as the per-layer kernel
                        umma_bf16(d_tmem, d_ah + ko, d_wl + ko, idesc, 1);
                        umma_bf16(d_tmem, d_ah + ko, d_wh + ko, idesc, 1);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: warps 2..9 =====================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int et = threadIdx.x - 64;
        constexpr int NC = BN / 64;
        int acc = 0; uint32_t acc_phase = 0;
        int slot = 0; uint32_t sphase = 0;
        // A finished tile is PUBLISHED (stores complete -> fences -> row-block counter) at the top of the next iteration, after the
        // next ring entry and its bias slice have been requested: the completion latency of the TMA stores overlaps those waits.
        // (Not later: the next blocking wait - the accumulator - may depend on this very publication.)
        int pub = -1, pub_stores = 0;                                       // ring word of the tile to publish
        for (;;) {
            mbar_wait(sfull_bar(slot), sphase);
            const int info = s_sched[slot];
            const ChainLayer& Ly = P.L[info < 0 ? 0 : info >> 24];
            const GemmArgs g = Ly.g;
            const int tile = info & 0xffffff;
            long long* tr = (P.trace && et == 0 && info >= 0) ? P.trace + 8ll * s_ticket[slot] : nullptr;
            const int m_blk = tile / Ly.tiles_n;
            const int m0 = m_blk * BM, n0 = (tile - m_blk * Ly.tiles_n) * BN;
            float bias_v = 0.f, head_v = 0.f;
            if (info >= 0 && et < BN && n0 + et < g.N) {
                if (g.bias) bias_v = __ldg(g.bias + n0 + et);
                if (g.head_w) head_v = __ldg(g.head_w + n0 + et);
            }
            if (pub >= 0) {
                const ChainLayer& Lp = P.L[pub >> 24];
                const int pm = (pub & 0xffffff) / Lp.tiles_n;
                if (pub_stores && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the tile's rows are written
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
                if (et == 0) {
                    const int old = atomicAdd(P.ws + Lp.cnt_off + pm, 1);
                    *s_last = old == Lp.tiles_n - 1;
                    __threadfence();
                }
                if (Lp.head_out) {                                          // uniform per tile
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
                    if (*s_last && et < BM && pm * BM + et < Lp.g.M) {      // this CTA completed the row block: add the partials
                        const float* hp = Lp.g.head_part + (long long)(pm * BM + et) * Lp.g.head_ld;
                        float a = 0.f;
                        for (int q = 0; q < Lp.g.head_ld; ++q) a += __ldcg(hp + q);
                        Lp.head_out[pm * BM + et] = a + (Lp.head_bias ? __ldg(Lp.head_bias) : 0.f);
                    }
                }
            }
            if (info < 0) break;                                            // (the ring is not reused after the end marker)
            if (et < BN) {
                s_bias[acc * BN + et] = bias_v;
                if (g.head_w) s_head[acc * BN + et] = head_v;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");       // everyone has read the ring entry
            if (et == 0) mbar_arrive(sempty_bar(slot));
            if (++slot == CHAIN_RING) { slot = 0; sphase ^= 1; }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (tr) tr[4] = gtimer();
            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < g.M;
            const int c0 = half * (BN / 2);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c0;
            const float* sb = s_bias + acc * BN + c0;
            const float* sh = s_head + acc * BN + c0;
            uint32_t ra[32], rb[32];
            uint8_t* const stage = smem_gen + OUT_OFF + (warp - 2) * CHAIN_STAGE_OUT;
            const uint32_t stage_u32 = base + OUT_OFF + (warp - 2) * CHAIN_STAGE_OUT;
            const CUtensorMap* const map_yh = &P.maps[info >> 24][4];
            const int row0 = m0 + quad * 32;
            bool in_flight = false;
            tmem_ld32(taddr, ra);
            tmem_ld_wait(ra);
#pragma unroll
            for (int c = 0; c < NC; c += 2) {
                tmem_ld32(taddr + (c + 1) * 32, rb);
                float hd = chain_epilogue_chunk(ra, sb + c * 32, sh + c * 32, n0 + c0 + c * 32, row, row0, row_ok, g, stage, stage_u32,
                                                map_yh, map_yh + 1, lane, in_flight, P.dbg);
                tmem_ld_wait(rb);
                if (c + 2 < NC) tmem_ld32(taddr + (c + 2) * 32, ra);
                hd += chain_epilogue_chunk(rb, sb + (c + 1) * 32, sh + (c + 1) * 32, n0 + c0 + (c + 1) * 32, row, row0, row_ok, g, stage,
                                           stage_u32, map_yh, map_yh + 1, lane, in_flight, P.dbg);
                if (c + 2 < NC) tmem_ld_wait(ra);
                const int grp = (n0 + c0 + c * 32) >> 6;
                if (g.head_part && row_ok && grp < g.head_ld) g.head_part[(long long)row * g.head_ld + grp] = hd;
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (tr) tr[5] = gtimer();
            pub = info; pub_stores = in_flight;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T::TMEM_COLS));
    } else if (warp == 0) {
        // the last CTA to leave zeroes the workspace for the next launch (nobody reads it any more)
        int left = 0;
        if (lane == 0) { __threadfence(); left = atomicAdd(P.ws + 1, 1); }
        left = __shfl_sync(0xffffffffu, left, 0);
        if (left == (int)gridDim.x - 1)
            for (int i = lane; i < P.ws_ints; i += 32) P.ws[i] = 0;
    }
}

// ---- fp32 -> (optional running-mean-std normalisation, utils/running_mean_std.py:82-84) -> bf16 hi/lo split ----
// Only columns [0,K) are written: the TMA tensor maps carry the exact K, so pad columns of the pitch are never read.
__global__ void split_bf16_kernel(const float* __restrict__ x, long long ldx, long long M, int K, const float* __restrict__ mean,
                                  const float* __restrict__ var, float eps, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long ld16) {
    const int pairs = (K + 1) / 2;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * pairs) return;
    long long r = i / pairs; int k = (int)(i - r * pairs) * 2;
    float v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        float a = 0.f;
        if (k + u < K) {
            a = x[r * ldx + k + u];
            if (mean) { a = (a - mean[k + u]) / sqrtf(var[k + u] + eps); a = fminf(fmaxf(a, -5.0f), 5.0f); }
        }
        v[u] = a;
    }
    __nv_bfloat16 h0 = __float2bfloat16_rn(v[0]), h1 = __float2bfloat16_rn(v[1]);
    __nv_bfloat16 l0 = __float2bfloat16_rn(v[0] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v[1] - __bfloat162float(h1));
    if (k + 1 < K) {   // ld16 and k are even: 4-byte aligned pair store
        *reinterpret_cast<uint32_t*>(hi + r * ld16 + k) = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        *reinterpret_cast<uint32_t*>(lo + r * ld16 + k) = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    } else {
        hi[r * ld16 + k] = h0; lo[r * ld16 + k] = l0;
    }
}

// ---- host side: tensor maps ----
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    });
    return fn;
}

// [rows, K] bf16, row pitch ld elements (ld % 8 == 0, base 16-byte aligned), box = box_rows x bk (one swizzle atom wide),
// zero fill outside [rows, K]
static bool make_map(CUtensorMap* m, const void* ptr, long long rows, int K, long long ld, int box_rows, int bk) {
    EncodeFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_num_sms = 0;

template <int BN>
static cudaError_t launch(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, long long lda, const __nv_bfloat16* w_hi,
                          const __nv_bfloat16* w_lo, long long ldw, const GemmArgs& g, cudaStream_t st) {
    using T = Tile<BN>;
    CUtensorMap mah, mal, mwh, mwl;
    if (!make_map(&mah, a_hi, g.M, g.K, lda, BM, T::BK) || !make_map(&mal, a_lo, g.M, g.K, lda, BM, T::BK) ||
        !make_map(&mwh, w_hi, g.N, g.K, ldw, BN, T::BK) || !make_map(&mwl, w_lo, g.N, g.K, ldw, BN, T::BK))
        return cudaErrorInvalidValue;
    auto kern = linear_bf16x3_kernel<BN>;
    static bool attr_set = false;   // one flag per template instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (!g_num_sms) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN) * (g.k_splits > 1 ? g.k_splits : 1);
    const int grid = tiles < g_num_sms ? tiles : g_num_sms;
    // programmatic dependent launch (EMLOCO_PDL=1, off by default): this kernel's prologue (tensor-map prefetch, barrier init,
    // TMEM allocation) may run while the previous kernel of the stream is still draining; griddepcontrol.wait in the kernel
    // orders everything else.  Measured +0.3 % on the rollout step: with parallel graph branches the early-launched CTAs sit on
    // SMs that kernels of the other branches could have used, which eats most of the prologue overlap.
    static const bool pdl = [] { const char* e = getenv("EMLOCO_PDL"); return e && e[0] == '1'; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = T::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, mah, mal, mwh, mwl, g);
}

template <int BN>
static cudaError_t launch_2cta(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, long long lda, const __nv_bfloat16* w_hi,
                               const __nv_bfloat16* w_lo, long long ldw, const GemmArgs& g, cudaStream_t st) {
    using T = Tile2<BN>;
    CUtensorMap mah, mal, mwh, mwl;
    if (!make_map(&mah, a_hi, g.M, g.K, lda, BM, T::BK) || !make_map(&mal, a_lo, g.M, g.K, lda, BM, T::BK) ||
        !make_map(&mwh, w_hi, g.N, g.K, ldw, T::BH, T::BK) || !make_map(&mwl, w_lo, g.N, g.K, ldw, T::BH, T::BK))
        return cudaErrorInvalidValue;
    auto kern = linear_bf16x3_2cta_kernel<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (!g_num_sms) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int tiles = ((g.M + 2 * BM - 1) / (2 * BM)) * ((g.N + BN - 1) / BN);
    const int pairs = g_num_sms / 2;
    const int clusters = tiles < pairs ? tiles : pairs;
    kern<<<2 * clusters, NUM_THREADS, T::SMEM, st>>>(mah, mal, mwh, mwl, g);
    return cudaGetLastError();
}


long long* g_chain_trace = nullptr;      // emloco_linear_chain_trace: profiling aid, NULL in production

static cudaError_t launch_chain(const emloco_chain_layer* layers, int n_layers, const int* order, int n_segments, int* ws,
                                long long ws_ints, cudaStream_t st, const char** why) {
    using T = Tile<128>;
    auto bad = [&](const char* m) { if (why) *why = m; return cudaErrorInvalidValue; };
    if (!layers || n_layers <= 0 || n_layers > CHAIN_MAX_LAYERS) return bad("emloco_linear_chain: 1..12 layers");
    if (!ws) return bad("emloco_linear_chain: null workspace");
    static ChainParams P;                                   // (host calls are serialised by the Python GIL / one thread per sim)
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int tiles_m[CHAIN_MAX_LAYERS], tiles[CHAIN_MAX_LAYERS];
    int off = 2;
    for (int l = 0; l < n_layers; ++l) {
        const emloco_chain_layer& a = layers[l];
        if (!a.a_hi || !a.a_lo || !a.w_hi || !a.w_lo || a.M <= 0 || a.N <= 0 || a.K <= 0 || a.lda < a.K || a.ldw < a.K || (a.lda & 7) || (a.ldw & 7))
            return bad("emloco_linear_chain: bad operand");
        if (((uintptr_t)a.a_hi | (uintptr_t)a.a_lo | (uintptr_t)a.w_hi | (uintptr_t)a.w_lo) & 15) return bad("emloco_linear_chain: operand pointers must be 16-byte aligned");
        if (!a.d_y32 && !a.y_hi && !a.d_head_w) return bad("emloco_linear_chain: layer without output");
        if (a.d_y32 && a.ldy < a.N) return bad("emloco_linear_chain: ldy < N");
        if ((a.y_hi == nullptr) != (a.y_lo == nullptr)) return bad("emloco_linear_chain: y_hi and y_lo must come together");
        if (a.y_hi && ((a.N & 31) || a.ldy16 < a.N || (a.ldy16 & 7) || (((uintptr_t)a.y_hi | (uintptr_t)a.y_lo) & 15)))
            return bad("emloco_linear_chain: split output needs N % 32 == 0, pitch % 8 == 0, 16-byte aligned pointers");
        if (a.d_head_w && (!a.d_head_part || !a.d_head_out)) return bad("emloco_linear_chain: null head argument");
        if (a.dep >= l || a.dep < -1) return bad("emloco_linear_chain: dep must name an earlier layer");
        if (a.dep >= 0 && (layers[a.dep].M != a.M || !layers[a.dep].y_hi)) return bad("emloco_linear_chain: dep needs the same M and a split output");
        tiles_m[l] = (int)((a.M + BM - 1) / BM);
        ChainLayer& L = P.L[l];
        L.tiles_n = (a.N + 127) / 128; L.num_kb = (a.K + T::BK - 1) / T::BK;
        tiles[l] = tiles_m[l] * L.tiles_n;
        if (tiles[l] >= (1 << 24)) return bad("emloco_linear_chain: too many tiles");
        L.dep = a.dep; L.need = a.dep >= 0 ? P.L[a.dep].tiles_n : 0;
        L.cnt_off = off; off += tiles_m[l];
        L.head_bias = a.d_head_bias; L.head_out = a.d_head_w ? a.d_head_out : nullptr;
        GemmArgs& g = L.g;
        g.bias = a.d_bias; g.y32 = a.d_y32; g.ldy = a.ldy; g.y_hi = (__nv_bfloat16*)a.y_hi; g.y_lo = (__nv_bfloat16*)a.y_lo; g.ldy16 = a.ldy16;
        g.M = (int)a.M; g.N = a.N; g.K = a.K; g.relu = a.relu & 1; g.m_dev = nullptr;
        g.head_w = a.d_head_w; g.head_part = a.d_head_w ? a.d_head_part : nullptr; g.head_ld = (a.N + 63) / 64;
        g.k_splits = 0; g.split_stride = 0; g.n_fast = 1;
        if (!make_map(&P.maps[l][0], a.a_hi, a.M, a.K, a.lda, BM, T::BK) || !make_map(&P.maps[l][1], a.a_lo, a.M, a.K, a.lda, BM, T::BK) ||
            !make_map(&P.maps[l][2], a.w_hi, a.N, a.K, a.ldw, 128, T::BK) || !make_map(&P.maps[l][3], a.w_lo, a.N, a.K, a.ldw, 128, T::BK))
            return bad("emloco_linear_chain: cuTensorMapEncodeTiled failed");
        if (a.y_hi && (!make_map(&P.maps[l][4], a.y_hi, a.M, a.N, a.ldy16, 32, 32) || !make_map(&P.maps[l][5], a.y_lo, a.M, a.N, a.ldy16, 32, 32)))
            return bad("emloco_linear_chain: cuTensorMapEncodeTiled failed (output)");
    }
    if (ws_ints < off) return bad("emloco_linear_chain: workspace too small");
    P.ws = ws; P.ws_ints = off; P.trace = g_chain_trace;
    { const char* e = getenv("EMLOCO_CHAIN_DBG"); P.dbg = e ? atoi(e) : 0; }
    // the ticket order: given segments, or the layers one after the other
    int seg[CHAIN_MAX_SEGS][3];
    if (order) {
        if (n_segments <= 0 || n_segments > CHAIN_MAX_SEGS) return bad("emloco_linear_chain: 1..32 segments");
        for (int i = 0; i < n_segments; ++i) for (int j = 0; j < 3; ++j) seg[i][j] = order[3 * i + j];
    } else {
        n_segments = n_layers;
        for (int l = 0; l < n_layers; ++l) { seg[l][0] = l; seg[l][1] = 0; seg[l][2] = tiles[l]; }
    }
    // every tile exactly once, and after the tiles of its dep that cover its row block (=> no ticket waits on a later one)
    {
        std::vector<int> claimed[CHAIN_MAX_LAYERS];            // per row block: tiles handed out so far
        std::vector<char> seen[CHAIN_MAX_LAYERS];
        for (int l = 0; l < n_layers; ++l) { claimed[l].assign(tiles_m[l], 0); seen[l].assign(tiles[l], 0); }
        int total = 0;
        for (int i = 0; i < n_segments; ++i) {
            const int l = seg[i][0];
            if (l < 0 || l >= n_layers || seg[i][1] < 0 || seg[i][2] <= 0 || seg[i][1] + seg[i][2] > tiles[l]) return bad("emloco_linear_chain: bad segment");
            for (int t = seg[i][1]; t < seg[i][1] + seg[i][2]; ++t) {
                if (seen[l][t]) return bad("emloco_linear_chain: tile listed twice");
                seen[l][t] = 1;
                const int mb = t / P.L[l].tiles_n;
                if (P.L[l].dep >= 0 && claimed[P.L[l].dep][mb] != P.L[P.L[l].dep].tiles_n)
                    return bad("emloco_linear_chain: a tile is ordered before the tiles it depends on");
                ++claimed[l][mb];
            }
            total += seg[i][2];
            P.seg_layer[i] = l; P.seg_first[i] = seg[i][1]; P.seg_end[i] = total;
        }
        int want = 0;
        for (int l = 0; l < n_layers; ++l) want += tiles[l];
        if (total != want) return bad("emloco_linear_chain: the order does not cover every tile");
        P.n_segs = n_segments; P.total = total;
        for (int i = n_segments; i < CHAIN_MAX_SEGS; ++i) { P.seg_layer[i] = 0; P.seg_first[i] = 0; P.seg_end[i] = 0x7fffffff; }
    }
    auto kern = linear_chain_kernel;
    constexpr int CHAIN_SMEM = T::STAGES * T::STAGE_BYTES + 3072 + NUM_EPI_WARPS * CHAIN_STAGE_OUT;      // = 227 KB
    static_assert(CHAIN_SMEM <= 227 * 1024, "shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (!g_num_sms) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = P.total < g_num_sms ? P.total : g_num_sms;
    kern<<<grid, NUM_THREADS, CHAIN_SMEM, st>>>(P);
    return cudaGetLastError();
}

}  // namespace tc

cudaError_t eml_split_bf16(const float* x, long long ldx, long long M, int K, const float* mean, const float* var, float eps,
                           void* hi, void* lo, long long ld16, cudaStream_t st) {
    if (M <= 0 || K <= 0) return cudaSuccess;
    long long n = M * ((K + 1) / 2);
    tc::split_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, ldx, M, K, mean, var, eps, (__nv_bfloat16*)hi,
                                                                        (__nv_bfloat16*)lo, ld16);
    return cudaGetLastError();
}

cudaError_t eml_linear_bf16x3(const void* a_hi, const void* a_lo, long long lda, const void* w_hi, const void* w_lo, long long ldw,
                              const float* bias, long long M, int N, int K, int relu, float* y32, long long ldy, void* y_hi,
                              void* y_lo, long long ldy16, int tile_n, const int* m_dev, const float* head_w, float* head_part,
                              int head_ld, cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    tc::GemmArgs g;
    g.m_dev = m_dev; g.head_w = head_w; g.head_part = head_part; g.head_ld = head_ld;
    g.k_splits = (tile_n >> 12) & 0xf; g.split_stride = (long long)M * ldy;      // bits 12..15 of the tile word: split-K count
    g.n_fast = (double)M * (double)K * 4.0 > 48e6;                               // hi + lo bytes of A beyond ~L2 / 2
    tile_n &= 0xfff;
    if (g.k_splits > 1 && (relu || y_hi || head_w || !y32 || (tile_n & 0x800))) return cudaErrorInvalidValue;
    if (g.k_splits > 1) {                                   // every split must own at least one k-block (both tile shapes)
        for (int bk : {64, 32}) { const int nkb = (K + bk - 1) / bk, per = (nkb + g.k_splits - 1) / g.k_splits; if ((g.k_splits - 1) * per >= nkb) return cudaErrorInvalidValue; }
    }
    if (head_w && (tile_n & 0x800)) return cudaErrorInvalidValue;       // the pair kernels carry no fused head
    g.bias = bias; g.y32 = y32; g.ldy = ldy; g.y_hi = (__nv_bfloat16*)y_hi; g.y_lo = (__nv_bfloat16*)y_lo; g.ldy16 = ldy16;
    g.M = (int)M; g.N = N; g.K = K; g.relu = relu;
    // tile choice: the 128 x 256 tile does 1.33x the flops per operand byte, but needs enough tiles to fill the 148 SMs
    int sms = tc::g_num_sms;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); tc::g_num_sms = sms; }
    const long long tiles256 = ((M + tc::BM - 1) / tc::BM) * ((N + 255) / 256);
    if (tile_n & 0x800) {                                   // CTA-pair kernels: 256 x 256 or 256 x 128 per pair
        if ((tile_n & 0x7ff) == 256)
            return tc::launch_2cta<256>((const __nv_bfloat16*)a_hi, (const __nv_bfloat16*)a_lo, lda, (const __nv_bfloat16*)w_hi,
                                        (const __nv_bfloat16*)w_lo, ldw, g, st);
        return tc::launch_2cta<128>((const __nv_bfloat16*)a_hi, (const __nv_bfloat16*)a_lo, lda, (const __nv_bfloat16*)w_hi,
                                    (const __nv_bfloat16*)w_lo, ldw, g, st);
    }
    const bool wide = tile_n == 256 || (tile_n == 0 && N >= 256 && tiles256 * 8 >= (long long)sms * 7);
    if (wide)
        return tc::launch<256>((const __nv_bfloat16*)a_hi, (const __nv_bfloat16*)a_lo, lda, (const __nv_bfloat16*)w_hi,
                               (const __nv_bfloat16*)w_lo, ldw, g, st);
    return tc::launch<128>((const __nv_bfloat16*)a_hi, (const __nv_bfloat16*)a_lo, lda, (const __nv_bfloat16*)w_hi,
                           (const __nv_bfloat16*)w_lo, ldw, g, st);
}

// fp32-in / fp32-out convenience path behind emloco_linear(use_tensor_cores = 1): splits both operands into stream-ordered
// scratch, then runs the tcgen05 GEMM.  The rollout uses the explicit split / bf16x3 entry points with persistent buffers.
cudaError_t eml_linear_tc(const float* x, long long ldx, const float* w, const float* b, float* y, long long ldy, long long M,
                          int N, int K, const float* mean, const float* var, float eps, int relu, cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    const long long kp = (K + 63) / 64 * 64;
    __nv_bfloat16* scratch = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&scratch, (size_t)(2 * (M + N) * kp) * sizeof(__nv_bfloat16), st);
    if (e != cudaSuccess) return e;
    __nv_bfloat16 *ah = scratch, *al = ah + M * kp, *wh = al + M * kp, *wl = wh + (long long)N * kp;
    if ((e = eml_split_bf16(x, ldx, M, K, mean, var, eps, ah, al, kp, st)) == cudaSuccess &&
        (e = eml_split_bf16(w, K, N, K, nullptr, nullptr, 0.f, wh, wl, kp, st)) == cudaSuccess)
        e = eml_linear_bf16x3(ah, al, kp, wh, wl, kp, b, M, N, K, relu, y, ldy, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, 0, st);
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    return e != cudaSuccess ? e : e2;
}

void eml_linear_chain_trace(long long* buf) { tc::g_chain_trace = buf; }

long long eml_linear_chain_workspace_ints(const emloco_chain_layer* layers, int n_layers) {
    long long n = 2;
    for (int l = 0; l < n_layers; ++l) n += (layers[l].M + tc::BM - 1) / tc::BM;
    return n;
}

cudaError_t eml_linear_chain(const emloco_chain_layer* layers, int n_layers, const int* order, int n_segments, int* ws, long long ws_ints,
                             cudaStream_t st, const char** why) {
    return tc::launch_chain(layers, n_layers, order, n_segments, ws, ws_ints, st, why);
}
