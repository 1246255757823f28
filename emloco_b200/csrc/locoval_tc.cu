// LocoVal scoring on the tensor cores: ValuePoseNet.forward (reference pacer/pacer/learning/value_pose_net.py:105-149) for the
// full variant (13 waypoints + pose + velocity: 100 -> 49 -> 24 -> 1) at large batch - the "1M synthetic 12-step futures"
// configuration of BASELINE.json and the value filter of social-transmotion/evaluate_jta.py:298-302.
//
// The CUDA-core kernel (locoval.cu) is bound by FP32 issue (6 100 MAC per score, 18 TFLOP/s); here the two wide layers run as
// bf16x3 tcgen05 MMAs (same split-precision scheme as linear_tc.cu: x = hi + lo, Al*Wh + Ah*Wl + Ah*Wh into fp32 TMEM), which
// leaves the kernel bound by its 404 B/score of HBM traffic and the per-row feature generation.
//
// CTA = 128 rows (row = TMEM lane), two threads per row, persistent over 128-row tiles, two CTAs per SM.  Activations never
// touch shared memory: both layers take their A operand from TENSOR memory (TS-mode tcgen05.mma), shared memory holds only the
// weights and the staging buffer of the raw rows.
//   1. each thread reads its half of the staged row (waypoints, pose, velocity), applies the heading normalisation (:73-103)
//      and the toe / spine masks (:141-144), and tcgen05.st's its 56 features as bf16 hi / lo pairs into its TMEM lane
//   2. thread 0: 7 k-steps x 3 MMAs (M=128, N=64, K=16, A from TMEM) -> D1; tcgen05.commit -> mbarrier.  The staging buffer
//      is free from here on: the second half-CTA starts the cp.async copies of the NEXT tile (they land under steps 2-5)
//   3. each thread: tcgen05.ld half of its D1 row, + b1, ReLU, split, tcgen05.st the A operand of layer 2 (over the dead
//      layer-1 columns)
//   4. thread 0: 4 k-steps x 3 MMAs (N=32, A from TMEM) -> D2
//   5. first half-CTA: tcgen05.ld the D2 row, + b2, ReLU, dot w3, + b3, sigmoid -> value; second half: waits for its copies
// Weights are split and staged once per CTA (W1 64 x 128, W2 32 x 64, zero padded, K-major, 128B swizzle).
#include <cuda_bf16.h>
#include "sim.h"

namespace lvtc {

constexpr int IN = 100, H1 = 49, H2 = 24;
constexpr int K1 = 112;                     // IN padded to the MMA K step (7 x 16); the tile itself is two 64-wide swizzle atoms
constexpr int N1 = 64, N2 = 32;
constexpr int A_ATOM = 128 * 128;           // 16 KB; the first 4 x 16 KB of shared memory are the staging buffer of the raw rows
constexpr int W1_ATOM = N1 * 128;           // 8 KB
constexpr int OFF_W1_HI = 4 * A_ATOM, OFF_W1_LO = OFF_W1_HI + 2 * W1_ATOM;
constexpr int W2_ATOM = N2 * 128;           // 4 KB
constexpr int OFF_W2_HI = OFF_W1_LO + 2 * W1_ATOM, OFF_W2_LO = OFF_W2_HI + W2_ATOM;
constexpr int OFF_F32 = OFF_W2_LO + W2_ATOM;                // b1[64] b2[32] w3[32] b3[1]
constexpr int OFF_BAR = OFF_F32 + (64 + 32 + 32 + 4) * 4;
constexpr int SMEM_BYTES = OFF_BAR + 16 + 1024;             // + alignment slack  (~107 KB: two CTAs per SM)
constexpr int POSE_PITCH = 19;             // float4 per staged pose row (18 used): 76-float pitch, conflict-free LDS.128
constexpr int ST_TRAJ = 0, ST_POSE = 20480, ST_VEL = ST_POSE + 128 * POSE_PITCH * 16;   // raw-row staging buffer (< 64 KB)
// tensor-memory map (columns; lane = row): layer-1 A operand hi 0..55 / lo 64..119 (K = 112 -> 56 columns of two bf16), D1
// 128..191, D2 192..223; the layer-2 A operand (K = 64 -> 32 columns) reuses the dead layer-1 columns: hi 0..31, lo 64..95
constexpr int TMEM_COLS = 256;
constexpr int TM_A1H = 0, TM_A1L = 64, TM_D1 = 128, TM_D2 = 192, TM_A2H = 0, TM_A2L = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) inside a K-major tile made of 64-element (128 B) wide atoms of `rows` rows, 128B swizzle:
// the 16-byte chunk index is XORed with the row index modulo 8 (what TMA's SWIZZLE_128B does)
__device__ __forceinline__ uint32_t sw128(int rows, int row, int k) {
    const int atom = k >> 6, kk = k & 63;
    return (uint32_t)(atom * rows * 128 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4) | ((kk & 7) << 1)));
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {       // K-major, SWIZZLE_128B, SBO = 1024 (see linear_tc.cu)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// A operand from tensor memory (lane = row, one 32-bit column = two consecutive K elements), B from shared memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tLVW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra LVW_DONE;\n\tbra LVW_LOOP;\n\tLVW_DONE:\n\t}"
                 ::"r"(bar), "r"(parity) : "memory");
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
__device__ __forceinline__ void tmem_wait(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    // packed conversions (one F2FP per pair); a bf16 widened to fp32 is its bits shifted into the high half
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// CTA = 256 threads = 128 rows x 2 halves (thread pair per row: half 0 = warps 0-3 builds features 0..55 and reads D1 columns
// 0..31, half 1 = warps 4-7 features 56..111 and D1 columns 32..63; both halves of a row sit on the same TMEM lanes).  Two CTAs
// per SM = 16 warps: the per-row work is a long dependent chain, so the kernel is latency-bound and lives off warps in flight.
__global__ void __launch_bounds__(256, 2)
locoval_tc_kernel(const float* __restrict__ traj, int stride, float* pose_rw, const float* __restrict__ vel,
                  const float* __restrict__ weights, float* __restrict__ value, long long B, int flags) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);        // 1024-byte aligned tile base
    const uint32_t sbase = smem_u32(sm);
    float* s_b1 = reinterpret_cast<float*>(sm + OFF_F32);
    float* s_b2 = s_b1 + 64; float* s_w3 = s_b2 + 32; float* s_b3 = s_w3 + 32;
    const uint32_t bar = sbase + OFF_BAR;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8);
    const int tid = threadIdx.x, warp = tid >> 5, half = tid >> 7, rt = tid & 127;
    const bool hide_toe = flags & 4, hide_spine = flags & 8, normalize = flags & 16, writeback = flags & 32;

    // ---- one-time: weights -> bf16 hi/lo, K-major swizzled, zero padded; biases; barrier; TMEM ----
    {
        const float* w1 = weights; const float* b1 = w1 + IN * H1;
        const float* w2 = b1 + H1; const float* b2 = w2 + H1 * H2;
        const float* w3 = b2 + H2; const float* b3 = w3 + H2;
        for (int i = tid; i < N1 * 128 / 2; i += 256) {                    // pairs (n, k..k+1) of the 64 x 128 padded W1
            const int n = i / 64, k = (i % 64) * 2;
            float a = (n < H1 && k < IN) ? w1[n * IN + k] : 0.f, b = (n < H1 && k + 1 < IN) ? w1[n * IN + k + 1] : 0.f;
            uint32_t h, l; split2(a, b, h, l);
            const uint32_t off = sw128(N1, n, k);
            *reinterpret_cast<uint32_t*>(sm + OFF_W1_HI + off) = h; *reinterpret_cast<uint32_t*>(sm + OFF_W1_LO + off) = l;
        }
        for (int i = tid; i < N2 * 64 / 2; i += 256) {                     // 32 x 64 padded W2
            const int n = i / 32, k = (i % 32) * 2;
            float a = (n < H2 && k < H1) ? w2[n * H1 + k] : 0.f, b = (n < H2 && k + 1 < H1) ? w2[n * H1 + k + 1] : 0.f;
            uint32_t h, l; split2(a, b, h, l);
            const uint32_t off = sw128(N2, n, k);
            *reinterpret_cast<uint32_t*>(sm + OFF_W2_HI + off) = h; *reinterpret_cast<uint32_t*>(sm + OFF_W2_LO + off) = l;
        }
        if (tid < 64) s_b1[tid] = tid < H1 ? b1[tid] : 0.f;
        if (tid < 32) { s_b2[tid] = tid < H2 ? b2[tid] : 0.f; s_w3[tid] = tid < H2 ? w3[tid] : 0.f; }
        if (tid == 0) {
            s_b3[0] = b3[0];
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    const long long tiles = (B + 127) / 128;
    // the tile's raw rows -> shared memory with coalesced 16-byte cp.async (every piece of the 52 KB in flight at once, no
    // register round trip); issued by half 1.  The staging area aliases the A tile.  Pose rows are placed at a 76-float pitch
    // so that the per-row LDS.128 reads below are bank-conflict free (72 would be 2-way)
    auto stage_tile = [&](long long t) {
        const int i0 = tid - 128;
        const long long r0 = t * 128;
        const int rows = (int)((B - r0) < 128 ? (B - r0) : 128);
        {
            const float* src = traj + r0 * 13 * stride;
            const int nfl = rows * 13 * stride;
            float* dst = reinterpret_cast<float*>(sm + ST_TRAJ);
            for (int i = i0; i < nfl / 4; i += 128)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(reinterpret_cast<float4*>(dst) + i)), "l"(reinterpret_cast<const float4*>(src) + i) : "memory");
            for (int i = (nfl & ~3) + i0; i < nfl; i += 128) dst[i] = src[i];
        }
        {
            const float4* src = reinterpret_cast<const float4*>(pose_rw + r0 * 72);
            for (int i = i0; i < rows * 18; i += 128) {
                const int r = i / 18, c4 = i - r * 18;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + ST_POSE + (uint32_t)(r * POSE_PITCH + c4) * 16u), "l"(src + i) : "memory");
            }
        }
        {
            const float* src = vel + r0 * 2;
            const int nfl = rows * 2;
            float* dst = reinterpret_cast<float*>(sm + ST_VEL);
            for (int i = i0; i < nfl / 4; i += 128)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(reinterpret_cast<float4*>(dst) + i)), "l"(reinterpret_cast<const float4*>(src) + i) : "memory");
            for (int i = (nfl & ~3) + i0; i < nfl; i += 128) dst[i] = src[i];
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (half == 1 && (long long)blockIdx.x < tiles) { stage_tile(blockIdx.x); asm volatile("cp.async.wait_group 0;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // weight tiles are read by the async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(s_tmem);
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t phase = 0;

    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long b = t * 128 + rt;
        const bool ok = b < B;
        // the rows of the tile after next start moving DRAM -> L2 now; their cp.async staging one iteration later hits L2
        if (tid == 128 && t + 2 * (long long)gridDim.x < tiles) {
            const long long r0 = (t + 2 * (long long)gridDim.x) * 128;
            const unsigned rows = (unsigned)((B - r0) < 128 ? (B - r0) : 128);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(traj + r0 * 13 * stride), "r"((rows * 13u * (unsigned)stride * 4u) & ~15u) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pose_rw + r0 * 72), "r"(rows * 288u) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(vel + r0 * 2), "r"((rows * 8u) & ~15u) : "memory");
        }
        // ---- 1b. this thread's 56 features of its row (registers, statically indexed) ----
        float f[56];
        {
            const int row = ok ? rt : (int)(B - 1 - t * 128);              // tail rows recompute the last valid row
            const float* tr = reinterpret_cast<const float*>(sm + ST_TRAJ) + row * 13 * stride;
            float c = 1.f, s = 0.f;
            if (normalize) {                                               // _rotate_normalization (:76-84): theta = atan2(y1, x1)
                const float x1 = tr[stride], y1 = tr[stride + 1];
                const float xe = fabsf(x1) < 1e-10f ? 1e-10f : x1;
                const float rn = rsqrtf(xe * xe + y1 * y1);                // cos / sin of atan2(y, x) without the angle
                c = xe * rn; s = y1 * rn;
            }
            const float4* p4 = reinterpret_cast<const float4*>(sm + ST_POSE) + row * POSE_PITCH;
            if (half == 0) {
                if (stride == 2) {                                         // 8-byte rows: LDS.64 at a 26-word pitch is conflict free
#pragma unroll
                    for (int n = 0; n < 13; ++n) {                         // row-vector times [[c,-s],[s,c]] (:85-100)
                        const float2 w = reinterpret_cast<const float2*>(tr)[n];
                        f[2 * n] = w.x * c + w.y * s; f[2 * n + 1] = w.y * c - w.x * s;
                    }
                } else {
#pragma unroll
                    for (int n = 0; n < 13; ++n) {
                        const float x = tr[n * stride], y = tr[n * stride + 1];
                        f[2 * n] = x * c + y * s; f[2 * n + 1] = y * c - x * s;
                    }
                }
                // pose floats 0..29 = joints 0..9 -> features 26..55
#pragma unroll
                for (int i = 0; i < 7; ++i) { float4 v = p4[i]; f[26 + 4 * i] = v.x; f[27 + 4 * i] = v.y; f[28 + 4 * i] = v.z; f[29 + 4 * i] = v.w; }
                { const float2 v = *reinterpret_cast<const float2*>(p4 + 7); f[54] = v.x; f[55] = v.y; }
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    float xr = f[26 + 3 * j] * c + f[27 + 3 * j] * s, yr = f[27 + 3 * j] * c - f[26 + 3 * j] * s, z = f[28 + 3 * j];
                    if ((hide_toe && (j == 4 || j == 8)) || (hide_spine && j == 9)) { xr = 0.f; yr = 0.f; z = 0.f; }
                    f[26 + 3 * j] = xr; f[27 + 3 * j] = yr; f[28 + 3 * j] = z;
                }
                if (writeback && ok) {                                     // the reference rotates / zeroes init_pose in place (:97,141-144)
                    float2* o2 = reinterpret_cast<float2*>(pose_rw + b * 72);
#pragma unroll
                    for (int i = 0; i < 15; ++i) o2[i] = make_float2(f[26 + 2 * i], f[27 + 2 * i]);
                }
            } else {
                // pose floats 30..71 = joints 10..23 -> features 56..97 (local 0..41); vel -> local 42,43; K padding 44..55
                { const float2 v = *(reinterpret_cast<const float2*>(p4 + 7) + 1); f[0] = v.x; f[1] = v.y; }
#pragma unroll
                for (int i = 0; i < 10; ++i) { float4 v = p4[8 + i]; f[2 + 4 * i] = v.x; f[3 + 4 * i] = v.y; f[4 + 4 * i] = v.z; f[5 + 4 * i] = v.w; }
#pragma unroll
                for (int j = 0; j < 14; ++j) {                             // joint 10 + j
                    float xr = f[3 * j] * c + f[3 * j + 1] * s, yr = f[3 * j + 1] * c - f[3 * j] * s, z = f[3 * j + 2];
                    if (hide_spine && j <= 1) { xr = 0.f; yr = 0.f; z = 0.f; }
                    f[3 * j] = xr; f[3 * j + 1] = yr; f[3 * j + 2] = z;
                }
                if (writeback && ok) {
                    float2* o2 = reinterpret_cast<float2*>(pose_rw + b * 72 + 30);
#pragma unroll
                    for (int i = 0; i < 21; ++i) o2[i] = make_float2(f[2 * i], f[2 * i + 1]);
                }
                const float2 vv = reinterpret_cast<const float2*>(sm + ST_VEL)[row];
                f[42] = vv.x * c + vv.y * s; f[43] = vv.y * c - vv.x * s;
#pragma unroll
                for (int k = 44; k < 56; ++k) f[k] = 0.f;
            }
        }
        // ---- 1c. bf16 hi / lo -> layer-1 A operand in tensor memory: 28 columns (56 features) per thread and term ----
        {
            uint32_t h[28], l[28];
#pragma unroll
            for (int i = 0; i < 28; ++i) split2(f[2 * i], f[2 * i + 1], h[i], l[i]);
            const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            // every store starts at a multiple of its own width: half 0 = 16 + 8 + 4 columns from 0, half 1 = 4 + 16 + 8 from 28
            const uint32_t ch = lane_base + TM_A1H + (uint32_t)(half * 28), cl = lane_base + TM_A1L + (uint32_t)(half * 28);
            if (half == 0) {
                tmem_st16(ch, h); tmem_st8(ch + 16, h + 16); tmem_st4(ch + 24, h + 24);
                tmem_st16(cl, l); tmem_st8(cl + 16, l + 16); tmem_st4(cl + 24, l + 24);
            } else {
                tmem_st4(ch, h); tmem_st16(ch + 4, h + 4); tmem_st8(ch + 20, h + 20);
                tmem_st4(cl, l); tmem_st16(cl + 4, l + 4); tmem_st8(cl + 20, l + 20);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                   // operands in place; everybody has read its staged row
        // ---- 2. layer 1 on the tensor core (A from TMEM); the staging buffer is free: the next tile's rows start to arrive ----
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < K1 / 16; ++j) {
                const uint32_t wo = (uint32_t)((j >> 2) * W1_ATOM + (j & 3) * 32);
                const uint64_t wh = make_desc(sbase + OFF_W1_HI + wo), wl = make_desc(sbase + OFF_W1_LO + wo);
                const uint32_t ah = tmem + TM_A1H + (uint32_t)(j * 8), al = tmem + TM_A1L + (uint32_t)(j * 8);
                umma_ts(tmem + TM_D1, al, wh, idesc1, j != 0);
                umma_ts(tmem + TM_D1, ah, wl, idesc1, 1);
                umma_ts(tmem + TM_D1, ah, wh, idesc1, 1);
            }
            commit(bar);
        }
        if (half == 1 && t + gridDim.x < tiles) stage_tile(t + gridDim.x);   // waited for at the end of the iteration
        mbar_wait(bar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- 3. hidden layer 1: + b1, ReLU, split -> layer-2 A operand in TMEM (K = 64: 32 columns hi + 32 lo; K 49..63 zero) ----
        {
            const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            uint32_t r0[32];
            tmem_ld32(lane_base + TM_D1 + (uint32_t)(half * 32), r0);
            tmem_wait(r0);
            uint32_t h[16], l[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int jl = 2 * i, j = half * 32 + jl;
                const float a = j < H1 ? fmaxf(__uint_as_float(r0[jl]) + s_b1[j], 0.f) : 0.f;
                const float bq = j + 1 < H1 ? fmaxf(__uint_as_float(r0[jl + 1]) + s_b1[j + 1], 0.f) : 0.f;
                split2(a, bq, h[i], l[i]);
            }
            tmem_st16(lane_base + TM_A2H + (uint32_t)(half * 16), h);
            tmem_st16(lane_base + TM_A2L + (uint32_t)(half * 16), l);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- 4. layer 2 on the tensor core (A from TMEM) ----
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ko = (uint32_t)(j * 32);
                const uint64_t wh = make_desc(sbase + OFF_W2_HI + ko), wl = make_desc(sbase + OFF_W2_LO + ko);
                const uint32_t ah = tmem + TM_A2H + (uint32_t)(j * 8), al = tmem + TM_A2L + (uint32_t)(j * 8);
                umma_ts(tmem + TM_D2, al, wh, idesc2, j != 0);
                umma_ts(tmem + TM_D2, ah, wl, idesc2, 1);
                umma_ts(tmem + TM_D2, ah, wh, idesc2, 1);
            }
            commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- 5. half 0: hidden layer 2, output layer, sigmoid; half 1: waits for the staged rows ----
        if (half == 0) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + TM_D2, r);
            tmem_wait(r);
            float z = s_b3[0];
#pragma unroll
            for (int o = 0; o < H2; ++o) z += s_w3[o] * fmaxf(__uint_as_float(r[o]) + s_b2[o], 0.f);
            if (ok) value[b] = 1.0f / (1.0f + __expf(-z));
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();            // D1 / D2 are free again; the next tile's rows are staged
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

}  // namespace lvtc

cudaError_t eml_locoval_forward_tc(const float* traj, int stride, float* pose, const float* vel, const float* w, float* value,
                                   long long B, int flags, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(lvtc::locoval_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lvtc::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (B <= 0) return cudaSuccess;
    const long long tiles = (B + 127) / 128;
    const int grid = (int)(tiles < 148 * 2 ? tiles : 148 * 2);
    lvtc::locoval_tc_kernel<<<grid, 256, lvtc::SMEM_BYTES, st>>>(traj, stride, pose, vel, w, value, B, flags);
    return cudaGetLastError();
}
