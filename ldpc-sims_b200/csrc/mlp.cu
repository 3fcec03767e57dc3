// mlp.cu - the reference's MLP demappers (nn/llr.py:7-73: Linear + tanh chains, fp32) on the
// 5th-generation tensor cores with fp32-equivalent accuracy.
//
// The reference evaluates  y = tanh(x W^T + b)  layer by layer in fp32 (ATen addmm).  fp32 is not a
// tensor-core input format, so every fp32 operand is split EXACTLY into NS binary16 planes
//     v = p0 + p1 (+ p2),   p0 = f16(v), p1 = f16(v - p0), ...            (11 + 11 (+ 11) significant bits)
// and the product is accumulated in fp32 (TMEM) from the plane pairs (i, j) with i + j < NS:
// NS = 2 -> 3 tcgen05.mma per k-step, |v - p0 - p1| <= 2^-23 |v|: fp32-equivalent (the default);
// NS = 3 -> 6 MMAs (beyond fp32); NS = 1 -> plain fp16.  Operands must stay below 65504 in magnitude
// (LLR-scale inputs, trained weights and tanh outputs do).
//
// One kernel per layer:  C[M,N] = act(A[M,K] W[N,K]^T + bias)
// (persistent CTAs, one per SM, walking the [128 x BN] output tiles)
//   warp 0  TMA producer: 3-D tensor maps (k, row, plane), 128-byte swizzle, [128 x 64] A boxes and
//           [BN x 64] W boxes per plane into a shared-memory ring (mbarrier full/empty, ~200 KB deep)
//   warp 1  allocates TMEM, one elected lane issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN,
//           K=16) for every plane pair and k-slice, tcgen05.commit releases the stage / signals the tile
//   warps 2-17 epilogue: tcgen05.ld (32 lanes x 32 columns) -> + bias -> tanhf -> split into the f16
//           planes of the next layer's A operand, transposed through a swizzled shared-memory patch so that
//           every store instruction writes 8 rows x 64 contiguous bytes (or the fp32 result of the last
//           layer); TMEM holds two accumulator sets, so the epilogue of a tile overlaps the loads and MMAs
//           of the next one
// Activations of a chunk of rows ping-pong between two plane buffers (LDPC_MLP_PER_LAYER); where the shape allows, the whole
// chain runs in ONE cooperative launch per chunk instead (chain_kernel below, LDPC_MLP_CHAIN, the default).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ldpc {
namespace mlp {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
#ifndef MLP_EPI_WARPS
#define MLP_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = MLP_EPI_WARPS;
constexpr int THREADS = 32 * (2 + EPI_WARPS);      // warp 0 TMA, warp 1 MMA, then the epilogue warps (EPI_WARPS / 4 per TMEM lane quadrant)

// ---- PTX wrappers --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol error traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (unsigned spin = 0; !ok; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// the same load with an L2 eviction-priority hint (the fixed createpolicy encodings CUTLASS uses: cute/arch/copy_sm90_desc.hpp)
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_global_v4_hint(void *p, const uint4 &v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, binary16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 64 halves (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);      // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tanh(x) = sign(x) (1 - e) / (1 + e) with e = 2^(-2 log2(e) |x|): two MUFU (ex2, rcp) + five ALU instructions, no branch.
// Absolute error <= ~3e-7 over the whole range (the cancellation in 1 - e near 0 costs relative, not absolute, accuracy, and the
// next layer consumes absolute values): the error of the whole chain against float64 is unchanged (2.2e-6 of the output scale,
// profiles/r02_mlp_experiments.md).  libm tanhf is ~25 instructions with both range branches predicated, and the epilogue warps
// are issue/latency-bound: 4.67 -> 4.58 ms per 2^20 rows.  -DMLP_EXP_NOTANH (timing experiments only) removes the activation.
#ifdef MLP_EXP_NOTANH
__device__ __forceinline__ float tanh_act(float x) { return x * 0.01f; }
#else
__device__ __forceinline__ float tanh_act(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(x) * -2.885390082f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return copysignf((1.0f - e) * r, x);
}
#endif

// Accumulators -> activations for 32 columns of one TMEM lane (= output row): leading (+ correction) accumulator, bias, tanh.
// The columns go in two halves, the TMEM read of the second in flight while the first is summed, biased and passed through tanh
// (all epilogue warps start a tile together and queue on the TMEM read port: 3 % of the forward; four quarters were slower);
// the bias is four vector loads per half issued together (a load per element made every add wait for its own load).
template <bool CORR>
__device__ __forceinline__ void acc_to_values(uint32_t t_main, uint32_t corr_off, const float *bias /* of these 32 columns, or null */, bool act, float (&o)[32]) {
    uint32_t v[32], w[32];
    tmem_ld16_issue(t_main, v);
    if (CORR) tmem_ld16_issue(t_main + corr_off, w);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tmem_ld_wait();
        if (half == 0) {
            tmem_ld16_issue(t_main + 16, v + 16);
            if (CORR) tmem_ld16_issue(t_main + corr_off + 16, w + 16);
        }
#pragma unroll
        for (int q = 16 * half; q < 16 * half + 16; ++q) o[q] = CORR ? __fadd_rn(__uint_as_float(v[q]), __uint_as_float(w[q])) : __uint_as_float(v[q]);
        if (bias) {
            float4 b4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) b4[q] = __ldg(reinterpret_cast<const float4 *>(bias + 16 * half) + q);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = 16 * half + 4 * q;
                o[e] = __fadd_rn(o[e], b4[q].x); o[e + 1] = __fadd_rn(o[e + 1], b4[q].y);
                o[e + 2] = __fadd_rn(o[e + 2], b4[q].z); o[e + 3] = __fadd_rn(o[e + 3], b4[q].w);
            }
        }
        if (act) {
#pragma unroll
            for (int q = 16 * half; q < 16 * half + 16; ++q) o[q] = tanh_act(o[q]);
        }
    }
}

// exact split of an fp32 value into binary16 planes
template <int NS>
__device__ __forceinline__ void split_f16(float v, __half (&p)[NS]) {
    float r = v;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        p[i] = __float2half_rn(r);
        r = __fsub_rn(r, __half2float(p[i]));      // exact: p[i] is r rounded to 11 significant bits
    }
}

// ---- fp32 rows -> f16 planes [NS][M][Kp] (zero padding beyond K) ---------------------------------------
// One thread converts 8 consecutive columns of a row: one 16-byte store per plane.
template <int NS>
__global__ void __launch_bounds__(256) split_rows_kernel(const float *x, long long ld, int K, long long M, int Kp,
                                                         __half *planes, long long plane_stride) {
    const int groups = Kp >> 3;                                           // Kp is a multiple of 8
    const long long total = M * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / groups;
        const int k0 = (int)(i - m * groups) << 3;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (k0 + j < K) ? __ldg(x + m * ld + k0 + j) : 0.0f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
                pk[q] = *reinterpret_cast<const uint32_t *>(&h);
                const float2 hf = __half22float2(h);
                v[2 * q] = __fsub_rn(v[2 * q], hf.x);                     // exact residuals
                v[2 * q + 1] = __fsub_rn(v[2 * q + 1], hf.y);
            }
            *reinterpret_cast<uint4 *>(planes + s * plane_stride + m * Kp + k0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// ---- one layer ---------------------------------------------------------------------------------------------
struct LayerArgs {
    int k_blocks;                 // Kp / 64
    int m_valid;                  // rows of this chunk that exist
    int m_rows;                   // rows to compute (m_valid rounded up to the tile)
    int n_total;                  // N of the layer
    int act;                      // 1 = tanh
    const float *bias;            // [N] or null
    __half *out_planes;    // [NS][chunk_rows][N] (next layer's A operand) or null
    long long out_plane_stride;
    float *out_f32;               // [m_valid][N] row-major (last layer) or null
};

template <int NS, int BN>
struct Smem {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = NS * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
    static constexpr int STORE_STAGING = EPI_WARPS * 2048;   // per epilogue warp: 32 rows x 64 B of one output plane
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 128 /* barriers */ + STORE_STAGING;
};

// -DMLP_TRACE: CTA 0 records clock64() at its pipeline events (stage refilled / stage full / epilogue start, end);
// ldpc_mlp_debug_trace() reads them back.  This is how the store transpose above was found.
#ifdef MLP_TRACE
__device__ long long g_trace[3][256];
__device__ int g_trace_n[3];
#define TRACE(role) do { if (blockIdx.x == 0) { int i_ = g_trace_n[role]; if (i_ < 256) { g_trace[role][i_] = clock64(); g_trace_n[role] = i_ + 1; } } } while (0)
#else
#define TRACE(role) do { } while (0)
#endif

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent: gridDim.x CTAs (one per SM) walk the output tiles t = blockIdx.x, + gridDim.x, ... (column
// tile fastest, so the CTAs running side by side share their A tile in L2).  TMEM holds TWO accumulator
// sets, so the epilogue of tile i overlaps the TMA loads and MMAs of tile i + 1.
template <int NS, int BN>
__global__ void __launch_bounds__(THREADS, 1) layer_kernel(const __grid_constant__ CUtensorMap map_a,
                                                            const __grid_constant__ CUtensorMap map_w, const LayerArgs args) {
    using S = Smem<NS, BN>;
    constexpr int STAGES = S::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;         // swizzle-128B tiles need 1024-byte alignment
    const uint32_t bars = base + STAGES * S::STAGE_BYTES;                  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    const uint32_t store_staging = bars + 128u;                           // [EPI_WARPS][2048]
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = args.n_total / BN;
    const int n_tiles = tiles_n * ((args.m_rows + BM - 1) / BM);
    // two fp32 accumulators per set: columns [0, BN) collect the leading plane pair (0, 0), columns [BN, 2 BN)
    // the correction pairs, whose sum is ~2^-11 of the result - the tensor core's truncating fp32 accumulation
    // then costs ~2^-11 less on the correction streams; the epilogue adds the two in fp32 (RN)
    constexpr uint32_t ACC2 = NS > 1 ? BN : 0;
    constexpr uint32_t ACC_COLS = BN + ACC2;
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
    static_assert(2 * ACC_COLS <= 512, "accumulator sets do not fit the tensor memory");

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {                                                   // ===== TMA producer =====
            uint32_t it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
                for (int kb = 0; kb < args.k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    TRACE(0);
#ifdef MLP_EXP_NOLOAD
                    mbar_arrive(full_bar(s));
                    (void)m0; (void)n0;
#else
                    mbar_expect_tx(full_bar(s), S::STAGE_BYTES);
                    const uint32_t st = base + s * S::STAGE_BYTES;
#pragma unroll
                    for (int p = 0; p < NS; ++p) {
                        tma_load_3d(st + p * S::A_BYTES, &map_a, full_bar(s), kb * BK, m0, p);
                        tma_load_3d(st + NS * S::A_BYTES + p * S::B_BYTES, &map_w, full_bar(s), kb * BK, n0, p);
                    }
#endif
                }
            }
        }
    } else if (warp == 1) {                                                // ===== MMA issuer =====
        constexpr uint32_t idesc = (1u << 4) /* D = f32; A = B = f16 (format 0), both K-major */ |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        uint32_t it = 0, ti = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
            const uint32_t acc = ti & 1, acc_ph = (ti >> 1) & 1;
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);                        // the epilogue has drained this accumulator set
            tc_fence_after();
            const uint32_t d_main = tmem_base + acc * ACC_COLS;
            for (int kb = 0; kb < args.k_blocks; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                if (lane == 0) {
                    TRACE(1);
                    const uint32_t st = base + s * S::STAGE_BYTES;
#ifndef MLP_EXP_NOMMA
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if constexpr (NS == 2) {
                            // The W planes of a stage are adjacent [BN x 64] tiles = ONE [2 BN x 64] K-major operand, and the two
                            // accumulators are adjacent column ranges: A0 [W0;W1]^T is a single N = 2 BN instruction that yields the
                            // leading pair (0,0) and the correction pair (0,1) together (one read of A0 from shared memory instead
                            // of two: the operand reads of three N = 128 instructions saturate the 128 B/clk shared-memory port).
                            constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
                            const uint64_t a0 = umma_desc_sw128(st + k * UMMA_K * 2), a1 = umma_desc_sw128(st + S::A_BYTES + k * UMMA_K * 2);
                            const uint64_t b0 = umma_desc_sw128(st + NS * S::A_BYTES + k * UMMA_K * 2);
                            umma_f16(d_main, a0, b0, idesc2, (kb | k) ? 1u : 0u);
                            umma_f16(d_main + ACC2, a1, b0, idesc, 1u);                     // correction pair (1,0)
                        } else {
#pragma unroll
                            for (int i = 0; i < NS; ++i) {
#pragma unroll
                                for (int j = 0; j + i < NS; ++j) {
                                    const uint64_t ad = umma_desc_sw128(st + i * S::A_BYTES + k * UMMA_K * 2);
                                    const uint64_t bd = umma_desc_sw128(st + NS * S::A_BYTES + j * S::B_BYTES + k * UMMA_K * 2);
                                    if (i + j == 0) umma_f16(d_main, ad, bd, idesc, (kb | k) ? 1u : 0u);
                                    else umma_f16(d_main + ACC2, ad, bd, idesc, (kb | k | (i + j - 1) | i) ? 1u : 0u);   // first correction pair: (0, 1)
                                }
                            }
                        }
                    }
#endif
                    tc_commit(empty_bar(s));                               // stage free once these MMAs have read it
                    if (kb == args.k_blocks - 1) tc_commit(tfull_bar(acc)); // accumulator set complete
                }
                __syncwarp();
            }
        }
    } else {                                                               // ===== epilogue (warps 2..9) =====
        const int quad = warp & 3;                                         // TMEM lane quadrant this warp may read
        constexpr int PARTS = EPI_WARPS / 4;                               // column slices per tile (warps past BN / 32 idle)
        constexpr int PART_COLS = (BN / PARTS) < 32 ? 32 : (BN / PARTS);
        const int part = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
            const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
            const uint32_t acc = ti & 1, acc_ph = (ti >> 1) & 1;
            const long long m = (long long)m0 + row;
            mbar_wait(tfull_bar(acc), acc_ph);
            tc_fence_after();
#ifdef MLP_EXP_NOEPI
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            continue;
#endif
            if (warp == 2 && lane == 0) TRACE(2);
            const uint32_t d_main = tmem_base + acc * ACC_COLS + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int c0 = part * PART_COLS; c0 < (part + 1) * PART_COLS && c0 < BN; c0 += 32) {
                float o[32];
                acc_to_values<(NS > 1)>(d_main + (uint32_t)c0, ACC2, args.bias ? args.bias + n0 + c0 : nullptr, args.act != 0, o);
                if (args.out_f32) {
                    if (m < args.m_valid) {
                        float4 *dst = reinterpret_cast<float4 *>(args.out_f32 + m * args.n_total + n0 + c0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                    }
                } else {
                    // exact plane split, two values per conversion (cvt.rn.f16x2.f32).  A thread owns one ROW (TMEM lane):
                    // storing its 64 bytes directly would make every warp store touch 32 lines with 16 bytes each
                    // (measured: 60 % of the epilogue and a 40 % longer tile).  The warp transposes through a swizzled
                    // 2 KB shared-memory patch instead, so each store instruction writes 8 rows x 64 contiguous bytes.
                    const uint32_t stg = store_staging + (uint32_t)(warp - 2) * 2048u;
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        uint32_t pk[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const __half2 h = __floats2half2_rn(o[2 * q], o[2 * q + 1]);
                            pk[q] = *reinterpret_cast<const uint32_t *>(&h);
                            if (s + 1 < NS) {
                                const float2 hf = __half22float2(h);
                                o[2 * q] = __fsub_rn(o[2 * q], hf.x);          // exact residuals
                                o[2 * q + 1] = __fsub_rn(o[2 * q + 1], hf.y);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)                            // chunk q of row `lane`, XOR-swizzled: conflict-free both ways
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 64u + (uint32_t)((q ^ ((lane >> 1) & 3)) << 4)),
                                         "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                        __syncwarp();
                        __half *const plane = args.out_planes + s * args.out_plane_stride + ((long long)m0 + quad * 32) * args.n_total + n0 + c0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int R = 8 * i + (lane >> 2), C = lane & 3;
                            uint4 v;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                         : "r"(stg + (uint32_t)R * 64u + (uint32_t)((C ^ ((R >> 1) & 3)) << 4)) : "memory");
#ifdef MLP_EXP_NOSTORE
                            if (v.x == 0x12345678u && args.m_valid < 0)
#endif
                            *reinterpret_cast<uint4 *>(plane + (long long)R * args.n_total + C * 8) = v;
                        }
                        __syncwarp();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (warp == 2 && lane == 0) TRACE(2);
            if (lane == 0) mbar_arrive(tempty_bar(acc));                   // this warp is done with the accumulator set
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- the whole chain in ONE launch: row blocks flow through the layers while their activations are still in L2 ----------
// The per-layer launches above stream every activation plane through HBM twice (write, then read by the next launch: 10.6 KB
// per row measured, 20x the algorithmic 516 B).  Here GROUP = 4 persistent CTAs (one per SM) own a block of 128 rows for ALL
// layers: CTA r computes one column tile of every layer (a narrower last layer rotates over the CTAs), writes its [128 x BN]
// slice of the activation planes to a small ring in global memory (INFLIGHT row blocks per group, L2 evict_last stores) and a
// signalling thread publishes it with fence + add on a per-(row block, layer) counter; the TMA producer of a tile of the
// next layer acquires the counter (all GROUP slices of the previous layer present), crosses to the async proxy
// (fence.proxy.async) and loads the full-width A operand from L2.  INFLIGHT row blocks are interleaved layer by layer, so the
// epilogue and the cross-SM hand-over of one overlap the MMAs of the others.  Results are bit-identical to the per-layer
// launches (same MMA order per tile, same epilogue arithmetic).  Measured on 2^20 rows (profiles/r02_mlp_experiments.md):
// INFLIGHT = 4 (78 MB ring): 4.3 ms, 3.4 KB of DRAM traffic per row; 3 (58 MB): +5 %, 1.6 KB; 2 (39 MB): +17 %, 1.3 KB -
// of which 0.6 KB is the input plane split, still a separate launch.  Launched cooperatively: the spin waits need every CTA of
// a group resident.
#ifndef MLP_INFLIGHT
#define MLP_INFLIGHT 4
#endif
constexpr int MAX_CHAIN = 6, GROUP = 4, INFLIGHT = MLP_INFLIGHT;

struct ChainLayer {
    int k_blocks, n_total, bn, act;
    const float *bias;
};
struct ChainArgs {
    int n_layers, m_valid, row_blocks, n_groups;
    ChainLayer L[MAX_CHAIN];
    __half *ring[2];              // activation planes [NS][ring_rows][ring_pitch], the leading `width` columns used; layer l reads ring[l & 1]
    long long ring_rows;          // n_groups * INFLIGHT * 128
    int ring_pitch;               // widest layer: one row pitch for every layer, so the slots of layers of different width cannot alias
    float *out_f32;               // [m_valid][N of the last layer]
    unsigned int *flags;          // [n_groups * INFLIGHT][MAX_CHAIN] epilogue-warp arrivals, monotonic within a launch
};
struct ChainMaps {
    CUtensorMap a[MAX_CHAIN], w[MAX_CHAIN];
};

// Tiles of CTA r of group g, in issue order; every role of the CTA walks the same sequence.  Within a generation of INFLIGHT
// row blocks the layers go round robin, except that the (short) layer-0 tiles of row blocks 2, 3, ... are slotted between the
// layer-1 tiles of row blocks 0, 1, ...: a layer-0 tile has a quarter of the MMA work of a hidden tile but a full epilogue.
template <class F>
__device__ __forceinline__ void chain_tiles(const ChainArgs &a, int g, int r, F &&f) {
    const int mine = (a.row_blocks - g + a.n_groups - 1) / a.n_groups;     // row blocks g, g + n_groups, ... of this group
    for (int j0 = 0; j0 < mine; j0 += INFLIGHT) {
        const int cnt = mine - j0 < INFLIGHT ? mine - j0 : INFLIGHT;
        int n0 = 0, n1 = 0, l2 = 2, s2 = 0;                                // next row block of layer 0 / of layer 1; cursor over the later layers
        for (;;) {                                                         // (one call site: the body is inlined once, its counters stay in registers)
            int s, l;
            if (n1 < cnt) {
                if (n0 < cnt && n0 < n1 + 2) { s = n0++; l = 0; } else { s = n1++; l = 1; }
            } else if (l2 < a.n_layers) {
                s = s2; l = l2;
                if (++s2 == cnt) { s2 = 0; ++l2; }
            } else break;
            const int j = j0 + s;
            const int nt = (r + GROUP - (j & (GROUP - 1))) & (GROUP - 1);   // a layer with fewer than GROUP tiles rotates over the CTAs
            if (nt * a.L[l].bn < a.L[l].n_total) f(j, l, g + a.n_groups * j, nt);
        }
    }
}

__device__ __forceinline__ void wait_counter(const unsigned int *p, unsigned int target) {
    unsigned int v = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= target) break;
        if (spin > (1u << 24)) __trap();
        __nanosleep(64);
    }
    asm volatile("fence.proxy.async;" ::: "memory");       // the data was written through the generic proxy, TMA reads through the async proxy
}

constexpr int CHAIN_THREADS = THREADS + 32;          // + the signalling warp

template <int NS>
__global__ void __launch_bounds__(CHAIN_THREADS, 1) chain_kernel(const __grid_constant__ ChainMaps maps, const ChainArgs args) {
    static_assert(NS == 2, "the chain kernel is the fp32-equivalent two-plane path");
    constexpr int A_BYTES = BM * BK * 2, STAGE_MAX = NS * (A_BYTES + 128 * BK * 2), STAGES = 3;
    constexpr uint32_t ACC_COLS = 256, TMEM_COLS = 512;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE_MAX;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    auto stored_bar = [&](uint32_t t) { return bars + 8u * (2 * STAGES + 5 + (t & 3)); };   // the epilogue warps have issued the stores of tile t
    const uint32_t store_staging = bars + 128u;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x / GROUP, r = blockIdx.x % GROUP;
    const int last = args.n_layers - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
        for (uint32_t t = 0; t < 4; ++t) mbar_init(stored_bar(t), EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 2 + EPI_WARPS) {
        // ===== signaller: publishes a tile once its stores are visible device-wide.  The fence waits for the store
        // acknowledgements (microseconds under load), so it lives here and not in the epilogue warps: they release (CTA scope)
        // on an mbarrier and carry on; this thread acquires it, and its device-scope fence is cumulative over their stores.
        if (lane == 0) {
            uint32_t ti = 0;
            chain_tiles(args, g, r, [&](int j, int l, int, int) {
                mbar_wait(stored_bar(ti), (ti >> 2) & 1);
                ++ti;
                __threadfence();
                atomicAdd(args.flags + (g * INFLIGHT + (j % INFLIGHT)) * MAX_CHAIN + l, (unsigned)EPI_WARPS);
            });
        }
    } else if (warp == 0) {
        if (lane == 0) {                                                   // ===== TMA producer =====
            uint32_t it = 0;
            chain_tiles(args, g, r, [&](int j, int l, int rb, int nt) {
                const ChainLayer &L = args.L[l];
                const int slot = g * INFLIGHT + (j % INFLIGHT);
                const unsigned use = (unsigned)(j / INFLIGHT);             // earlier row blocks that went through this slot
                int arow;
                if (l == 0) {
                    arow = rb * BM;
                    // the first epilogue of this row block overwrites ring[1][slot]: the LAST layer of the previous occupant must have read it
                    if (use) wait_counter(args.flags + slot * MAX_CHAIN + last, (unsigned)EPI_WARPS * (unsigned)(args.L[last].n_total / args.L[last].bn) * use);
                } else {
                    arow = slot * BM;
                    wait_counter(args.flags + slot * MAX_CHAIN + (l - 1), (unsigned)EPI_WARPS * (unsigned)(args.L[l - 1].n_total / args.L[l - 1].bn) * (use + 1));
                }
                const uint32_t b_bytes = (uint32_t)L.bn * BK * 2;
                TRACE(0);
                for (int kb = 0; kb < L.k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(full_bar(s), NS * (A_BYTES + b_bytes));
                    const uint32_t st = base + s * STAGE_MAX;
#pragma unroll
                    for (int p = 0; p < NS; ++p) {        // input planes stream through L2 once; ring and weights should stay
                        tma_load_3d_hint(st + p * A_BYTES, &maps.a[l], full_bar(s), kb * BK, arow, p, l == 0 ? L2_EVICT_FIRST : L2_EVICT_LAST);
                        tma_load_3d_hint(st + NS * A_BYTES + p * b_bytes, &maps.w[l], full_bar(s), kb * BK, nt * L.bn, p, L2_EVICT_LAST);
                    }
                }
            });
        }
    } else if (warp == 1) {                                                // ===== MMA issuer =====
        uint32_t it = 0, ti = 0;
        chain_tiles(args, g, r, [&](int, int l, int, int) {
            const ChainLayer &L = args.L[l];
            const uint32_t acc = ti & 1, acc_ph = (ti >> 1) & 1;
            ++ti;
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            tc_fence_after();
            if (lane == 0) TRACE(1);
            const uint32_t d_main = tmem_base + acc * ACC_COLS;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(L.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * L.bn) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < L.k_blocks; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(full_bar(s), (it / STAGES) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t st = base + s * STAGE_MAX;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {                // A0 [W0;W1]^T (N = 2 BN: pairs (0,0) | (0,1)), then A1 W0^T (pair (1,0))
                        const uint64_t a0 = umma_desc_sw128(st + k * UMMA_K * 2), a1 = umma_desc_sw128(st + A_BYTES + k * UMMA_K * 2);
                        const uint64_t b0 = umma_desc_sw128(st + NS * A_BYTES + k * UMMA_K * 2);
                        umma_f16(d_main, a0, b0, idesc2, (kb | k) ? 1u : 0u);
                        umma_f16(d_main + (uint32_t)L.bn, a1, b0, idesc, 1u);
                    }
                    tc_commit(empty_bar(s));
                    if (kb == L.k_blocks - 1) { tc_commit(tfull_bar(acc)); TRACE(1); }
                }
                __syncwarp();
            }
        });
    } else {                                                               // ===== epilogue warps =====
        const int quad = warp & 3, part = (warp - 2) >> 2;                 // TMEM lane quadrant, 32-column slice
        const int row = quad * 32 + lane;
        uint32_t ti = 0;
        chain_tiles(args, g, r, [&](int j, int l, int rb, int nt) {
            const ChainLayer &L = args.L[l];
            const int slot = g * INFLIGHT + (j % INFLIGHT);
            const int n0 = nt * L.bn;
            const uint32_t acc = ti & 1, acc_ph = (ti >> 1) & 1;
            ++ti;
            mbar_wait(tfull_bar(acc), acc_ph);
            tc_fence_after();
            if (warp == 2 && lane == 0) TRACE(2);
            const uint32_t d_main = tmem_base + acc * ACC_COLS + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int c0 = part * 32; c0 < L.bn; c0 += 32 * (EPI_WARPS / 4)) {
                float o[32];
                acc_to_values<true>(d_main + (uint32_t)c0, (uint32_t)L.bn, L.bias ? L.bias + n0 + c0 : nullptr, L.act != 0, o);
                if (l == last) {
                    const long long m = (long long)rb * BM + row;
                    if (m < args.m_valid) {
                        float4 *dst = reinterpret_cast<float4 *>(args.out_f32 + m * L.n_total + n0 + c0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) __stcs(dst + q, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
                    }
                } else {
                    const uint32_t stg = store_staging + (uint32_t)(warp - 2) * 2048u;
#pragma unroll
                    for (int s = 0; s < NS; ++s) {                         // exact plane split + transposed store (see layer_kernel)
                        uint32_t pk[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const __half2 h = __floats2half2_rn(o[2 * q], o[2 * q + 1]);
                            pk[q] = *reinterpret_cast<const uint32_t *>(&h);
                            if (s + 1 < NS) {
                                const float2 hf = __half22float2(h);
                                o[2 * q] = __fsub_rn(o[2 * q], hf.x);
                                o[2 * q + 1] = __fsub_rn(o[2 * q + 1], hf.y);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 64u + (uint32_t)((q ^ ((lane >> 1) & 3)) << 4)),
                                         "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                        __syncwarp();
                        __half *const plane = args.ring[(l + 1) & 1] + ((long long)s * args.ring_rows + (long long)slot * BM + quad * 32) * args.ring_pitch + n0 + c0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int R = 8 * i + (lane >> 2), C = lane & 3;
                            uint4 u;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                                         : "r"(stg + (uint32_t)R * 64u + (uint32_t)((C ^ ((R >> 1) & 3)) << 4)) : "memory");
                            st_global_v4_hint(plane + (long long)R * args.ring_pitch + C * 8, u, L2_EVICT_LAST);   // dirty ring lines stay in L2 until the slot is rewritten
                        }
                        __syncwarp();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();                                                  // the lanes' stores happen before lane 0's (releasing) arrivals
            if (lane == 0) {
                mbar_arrive(tempty_bar(acc));
                mbar_arrive(stored_bar(ti - 1));
                if (warp == 2) TRACE(2);
            }
        });
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- the chain on CTA PAIRS (cta_group::2) -----------------------------------------------------------------------------
// A single-SM tcgen05.mma fetches its shared-memory operands at ~64 B/clk, and the plane-split products need 20 KB per k-slice
// (profiles/r02_mlp_experiments.md): 320 clk against 192 clk of math.  Here two SMs (a cluster of 2) issue ONE M = 256
// instruction: each CTA holds the A planes of its own 128 rows and HALF of the B operand (for A0 [W0;W1]^T: CTA 0 stages the W0
// tile, CTA 1 the W1 tile; for A1 W0^T each stages half of the W0 rows): 14 KB per SM and k-slice.  A group is two pairs and
// carries a block of 256 rows; pair p computes two of the four column tiles of every layer; CTA c of both pairs depends only on
// rows c*128.. of the previous layer (written by CTA c of both pairs).  Only the leader (cluster rank 0) issues MMAs; both CTAs
// run a TMA producer (transaction bytes on the LEADER's full barrier), epilogue warps and a signaller; tcgen05.commit multicasts
// the "stage free" / "accumulator full" arrivals to both CTAs, the peer's epilogue warps arrive remotely on the leader's
// "accumulator empty" barrier.
#ifndef MLP_PAIR_LEAD
#define MLP_PAIR_LEAD 4
#endif
constexpr int INFLIGHT2 = 2;                         // 256-row blocks in flight per group
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;          // shared::cluster address of the same offset in the pair's even CTA

struct Chain2Maps {
    CUtensorMap a[MAX_CHAIN], w[MAX_CHAIN], w2[MAX_CHAIN];   // w: box [64 x BN] of one plane; w2: box [64 x BN/2]
};

__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1, int c2, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {          // arrives on `bar` of BOTH CTAs once the MMAs issued so far are done
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Tiles of CTA (pair p, rank c) of group g: the scheduler of chain_tiles over "virtual row blocks" v = 2 * (256-row block) + u,
// u = which of the pair's two column tiles.
template <class F>
__device__ __forceinline__ void chain2_tiles(const ChainArgs &a, int g, int p, F &&f) {
    const int mine = (a.row_blocks - g + a.n_groups - 1) / a.n_groups;     // 256-row blocks g, g + n_groups, ... of this group
    for (int j0 = 0; j0 < mine; j0 += INFLIGHT2) {
        const int cnt = 2 * (mine - j0 < INFLIGHT2 ? mine - j0 : INFLIGHT2);
        int n0 = 0, n1 = 0, l2 = 2, s2 = 0;
        for (;;) {
            int v, l;
            if (n1 < cnt) {
                if (n0 < cnt && n0 < n1 + MLP_PAIR_LEAD) { v = n0++; l = 0; } else { v = n1++; l = 1; }   // a layer-1 tile needs BOTH layer-0 tiles of its block from both pairs
            } else if (l2 < a.n_layers) {
                v = s2; l = l2;
                if (++s2 == cnt) { s2 = 0; ++l2; }
            } else break;
            const int j = j0 + (v >> 1);
            const int nt = 2 * (v & 1) + ((p ^ j) & 1);                    // a layer with fewer than 4 tiles alternates between the pairs
            if (nt * a.L[l].bn < a.L[l].n_total) f(j, l, g + a.n_groups * j, nt);
        }
    }
}

template <int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CHAIN_THREADS, 1) chain2_kernel(const __grid_constant__ Chain2Maps maps, const ChainArgs args) {
    static_assert(NS == 2, "two-plane path");
    constexpr int A_BYTES = BM * BK * 2, STAGE_MAX = NS * A_BYTES + 128 * BK * 2 + 64 * BK * 2, STAGES = 3;
    constexpr uint32_t ACC_COLS = 256, TMEM_COLS = 512;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE_MAX;
    auto full_bar = [&](int s) { return bars + 8u * s; };                  // leader's copy is the live one
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };   // leader's copy is the live one (2 * EPI_WARPS arrivals)
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    auto stored_bar = [&](uint32_t t) { return bars + 8u * (2 * STAGES + 5 + (t & 3)); };
    const uint32_t store_staging = bars + 128u;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t c;                                                            // rank in the pair: 0 = leader
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(c));
    const int g = blockIdx.x / 4, p = (blockIdx.x >> 1) & 1;
    const int last = args.n_layers - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EPI_WARPS); }
        for (uint32_t t = 0; t < 4; ++t) mbar_init(stored_bar(t), EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                    // both CTAs' barriers initialised before any remote arrival
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    auto flag_of = [&](int slot, int l) { return args.flags + ((slot * 2 + (int)c) * MAX_CHAIN + l); };

    if (warp == 2 + EPI_WARPS) {                                           // ===== signaller =====
        if (lane == 0) {
            uint32_t ti = 0;
            chain2_tiles(args, g, p, [&](int j, int l, int, int) {
                mbar_wait(stored_bar(ti), (ti >> 2) & 1);
                ++ti;
                __threadfence();
                atomicAdd(flag_of(g * INFLIGHT2 + (j % INFLIGHT2), l), (unsigned)EPI_WARPS);
            });
        }
    } else if (warp == 0) {
        if (lane == 0) {                                                   // ===== TMA producer (both CTAs) =====
            uint32_t it = 0;
            chain2_tiles(args, g, p, [&](int j, int l, int rb, int nt) {
                const ChainLayer &L = args.L[l];
                const int slot = g * INFLIGHT2 + (j % INFLIGHT2);
                const unsigned use = (unsigned)(j / INFLIGHT2);
                int arow;
                if (l == 0) {
                    arow = rb * 2 * BM + (int)c * BM;
                    if (use) wait_counter(flag_of(slot, last), (unsigned)EPI_WARPS * (unsigned)(args.L[last].n_total / args.L[last].bn) * use);
                } else {
                    arow = slot * 2 * BM + (int)c * BM;
                    wait_counter(flag_of(slot, l - 1), (unsigned)EPI_WARPS * (unsigned)(args.L[l - 1].n_total / args.L[l - 1].bn) * (use + 1));
                }
                const uint32_t b1_bytes = (uint32_t)L.bn * BK * 2, b2_bytes = b1_bytes / 2;
                const uint32_t stage_bytes = NS * A_BYTES + b1_bytes + b2_bytes;
                TRACE(0);
                for (int kb = 0; kb < L.k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
                    if (c == 0) mbar_expect_tx(full_bar(s), 2 * stage_bytes);     // the bytes of both CTAs land on the leader's barrier
                    const uint32_t st = base + s * STAGE_MAX, fb = full_bar(s) & PEER_MASK;
                    const uint64_t pa = l == 0 ? L2_EVICT_FIRST : L2_EVICT_LAST;
                    tma_load_3d_pair(st, &maps.a[l], fb, kb * BK, arow, 0, pa);
                    tma_load_3d_pair(st + A_BYTES, &maps.a[l], fb, kb * BK, arow, 1, pa);
                    tma_load_3d_pair(st + NS * A_BYTES, &maps.w[l], fb, kb * BK, nt * L.bn, (int)c, L2_EVICT_LAST);                       // W0 tile | W1 tile
                    tma_load_3d_pair(st + NS * A_BYTES + b1_bytes, &maps.w2[l], fb, kb * BK, nt * L.bn + (int)c * (L.bn / 2), 0, L2_EVICT_LAST);   // half of the W0 rows
                }
            });
        }
    } else if (warp == 1) {
        if (c == 0) {                                                      // ===== MMA issuer (leader only) =====
            uint32_t it = 0, ti = 0;
            chain2_tiles(args, g, p, [&](int, int l, int, int) {
                const ChainLayer &L = args.L[l];
                const uint32_t acc = ti & 1, acc_ph = (ti >> 1) & 1;
                ++ti;
                mbar_wait(tempty_bar(acc), acc_ph ^ 1);
                tc_fence_after();
                if (lane == 0) TRACE(1);
                const uint32_t d_main = tmem_base + acc * ACC_COLS;
                const uint32_t idesc = (1u << 4) | ((uint32_t)(L.bn >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
                const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * L.bn) >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
                const uint32_t b1_bytes = (uint32_t)L.bn * BK * 2;
                for (int kb = 0; kb < L.k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(full_bar(s), (it / STAGES) & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t st = base + s * STAGE_MAX;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t a0 = umma_desc_sw128(st + k * UMMA_K * 2), a1 = umma_desc_sw128(st + A_BYTES + k * UMMA_K * 2);
                            const uint64_t b1 = umma_desc_sw128(st + NS * A_BYTES + k * UMMA_K * 2), b2 = umma_desc_sw128(st + NS * A_BYTES + b1_bytes + k * UMMA_K * 2);
                            umma_f16_pair(d_main, a0, b1, idesc2, (kb | k) ? 1u : 0u);          // [256 x 2 BN]: pairs (0,0) | (0,1)
                            umma_f16_pair(d_main + (uint32_t)L.bn, a1, b2, idesc, 1u);            // [256 x BN]: pair (1,0)
                        }
                        tc_commit_pair(empty_bar(s));
                        if (kb == L.k_blocks - 1) { tc_commit_pair(tfull_bar(acc)); TRACE(1); }
                    }
                    __syncwarp();
                }
            });
        }
    } else {                                                               // ===== epilogue warps (both CTAs: their own 128 rows) =====
        const int quad = warp & 3, part = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        uint32_t ti = 0;
        chain2_tiles(args, g, p, [&](int j, int l, int rb, int nt) {
            const ChainLayer &L = args.L[l];
            const int slot = g * INFLIGHT2 + (j % INFLIGHT2);
            const int n0 = nt * L.bn;
            const uint32_t acc = ti & 1, acc_ph = (ti >> 1) & 1;
            ++ti;
            mbar_wait(tfull_bar(acc), acc_ph);
            tc_fence_after();
            if (warp == 2 && lane == 0) TRACE(2);
            const uint32_t d_main = tmem_base + acc * ACC_COLS + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int c0 = part * 32; c0 < L.bn; c0 += 32 * (EPI_WARPS / 4)) {
                float o[32];
                acc_to_values<true>(d_main + (uint32_t)c0, (uint32_t)L.bn, L.bias ? L.bias + n0 + c0 : nullptr, L.act != 0, o);
                if (l == last) {
                    const long long m = ((long long)rb * 2 + c) * BM + row;
                    if (m < args.m_valid) {
                        float4 *dst = reinterpret_cast<float4 *>(args.out_f32 + m * L.n_total + n0 + c0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) __stcs(dst + q, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
                    }
                } else {
                    const uint32_t stg = store_staging + (uint32_t)(warp - 2) * 2048u;
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        uint32_t pk[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const __half2 h = __floats2half2_rn(o[2 * q], o[2 * q + 1]);
                            pk[q] = *reinterpret_cast<const uint32_t *>(&h);
                            if (s + 1 < NS) {
                                const float2 hf = __half22float2(h);
                                o[2 * q] = __fsub_rn(o[2 * q], hf.x);
                                o[2 * q + 1] = __fsub_rn(o[2 * q + 1], hf.y);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 64u + (uint32_t)((q ^ ((lane >> 1) & 3)) << 4)),
                                         "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                        __syncwarp();
                        __half *const plane = args.ring[(l + 1) & 1] + ((long long)s * args.ring_rows + ((long long)slot * 2 + c) * BM + quad * 32) * args.ring_pitch + n0 + c0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int R = 8 * i + (lane >> 2), C = lane & 3;
                            uint4 u;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                                         : "r"(stg + (uint32_t)R * 64u + (uint32_t)((C ^ ((R >> 1) & 3)) << 4)) : "memory");
                            st_global_v4_hint(plane + (long long)R * args.ring_pitch + C * 8, u, L2_EVICT_LAST);
                        }
                        __syncwarp();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_leader(tempty_bar(acc));                       // the MMA issuer waits for the epilogues of BOTH CTAs
                mbar_arrive(stored_bar(ti - 1));
                if (warp == 2) TRACE(2);
            }
        });
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                    // no CTA of the pair leaves while the other may still signal it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// planes [NS][rows][Kp] f16 -> 3-D map (k, row, plane), box [64 x box_rows x 1], 128-byte swizzle
// (pitch = elements between rows, 0 = dense: the activation ring keeps one row pitch for layers of different width)
static int make_map(CUtensorMap *map, void *ptr, int Kp, long long rows, int ns, int box_rows, int pitch = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LDPC_ECUDA; }
    if (pitch == 0) pitch = Kp;
    const cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)ns};
    const cuuint64_t gstr[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)rows * pitch * 2};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    const cuuint32_t est[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ptr, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return LDPC_ECUDA; }
    return LDPC_OK;
}

struct Layer {
    int K, Kp, N, BN, act;
    float *d_bias = nullptr;
    __half *d_w = nullptr;          // [NS][N][Kp]
    CUtensorMap map_w, map_a;              // map_a: this layer's INPUT planes
};

}  // namespace mlp
}  // namespace ldpc

struct ldpc_mlp {
    int ns, device;
    long long chunk = 0;       // rows the activation buffers currently hold (grown on demand up to chunk_max)
    long long chunk_max = 0, chain_chunk_max = 0;
    int maxw = 0;              // widest layer (elements per row)
    std::vector<ldpc::mlp::Layer> layers;
    __half *d_act[2] = {nullptr, nullptr};   // ping-pong activation planes [NS][chunk][maxw] (one launch per layer)
    // single-launch chain (chain_kernel): input planes of a chunk, the L2-resident activation ring, the arrival counters
    int mode = 0;              // LDPC_MLP_AUTO / _PER_LAYER / _CHAIN
    bool chain_ok = false;     // shape and device allow the chain kernel
    int n_groups = 0;
    long long x_rows = 0;
    __half *d_x = nullptr, *d_ring[2] = {nullptr, nullptr};
    unsigned int *d_flags = nullptr;
    size_t ring_bytes = 0;
    ldpc::mlp::ChainMaps cmaps;
    ldpc::mlp::Chain2Maps cmaps2;
    bool pairs_ok = false;     // cluster launch of CTA pairs available too
    std::mutex mu;             // ldpc_mlp_forward grows and reuses the activation buffers: one call at a time per handle
};

using namespace ldpc;
using namespace ldpc::mlp;

template <int NS>
static int launch_split(const float *x, long long ld, int K, long long M, int Kp, __half *planes, long long plane_stride, cudaStream_t s) {
    const long long total = M * (Kp >> 3);
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    split_rows_kernel<NS><<<grid, 256, 0, s>>>(x, ld, K, M, Kp, planes, plane_stride);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

template <int NS, int BN>
static int launch_layer(const Layer &L, const LayerArgs &a, long long rows, cudaStream_t s) {
    auto k = layer_kernel<NS, BN>;
    LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<NS, BN>::TOTAL));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (long long)(L.N / BN) * ((rows + BM - 1) / BM);
    k<<<(unsigned)std::min<long long>(tiles, sms), THREADS, Smem<NS, BN>::TOTAL, s>>>(L.map_a, L.map_w, a);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

template <int NS>
static int launch_layer_ns(const Layer &L, const LayerArgs &a, long long rows, cudaStream_t s) {
    if (L.BN == 128) return launch_layer<NS, 128>(L, a, rows, s);
    return launch_layer<NS, 64>(L, a, rows, s);
}

extern "C" {

#ifdef MLP_TRACE
int ldpc_mlp_debug_trace(long long *out, int *counts, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 3 * 256);
    cudaMemcpyFromSymbol(counts, g_trace_n, sizeof(int) * 3);
    if (reset) { int z[3] = {0, 0, 0}; cudaMemcpyToSymbol(g_trace_n, z, sizeof(z)); }
    return 0;
}
#endif

void ldpc_mlp_destroy(ldpc_mlp_t *h) {
    if (!h) return;
    for (auto &L : h->layers) { cudaFree(L.d_bias); cudaFree(L.d_w); }
    cudaFree(h->d_act[0]); cudaFree(h->d_act[1]);
    cudaFree(h->d_x); cudaFree(h->d_ring[0]); cudaFree(h->d_flags);
    delete h;
}

int ldpc_mlp_create(int n_layers, const int32_t *dims, const float *const *weights, const float *const *biases,
                    const int32_t *activations, int splits, int64_t chunk_rows, ldpc_mlp_t **out) {
    if (!out || n_layers <= 0 || !dims || !weights) { set_error("ldpc_mlp_create: bad arguments"); return LDPC_EINVAL; }
    *out = nullptr;
    if (splits < 1 || splits > 3) { set_error("ldpc_mlp_create: splits must be 1, 2 or 3"); return LDPC_EINVAL; }
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { set_error("no CUDA device (there is no CPU fallback)"); return LDPC_ECUDA; }
    for (int l = 0; l < n_layers; ++l)
        if (dims[l] <= 0 || dims[l + 1] <= 0 || dims[l + 1] % 64 != 0 || !weights[l]) {
            set_error("ldpc_mlp_create: layer %d: output width must be a positive multiple of 64", l);
            return LDPC_EUNSUPPORTED;
        }
    ldpc_mlp *h = new ldpc_mlp();
    h->ns = splits;
    cudaGetDevice(&h->device);
    // default chunks: per-layer launches 592 row tiles (16 tiles per persistent CTA and layer amortise the fill/drain; both
    // activation buffers of a chunk are 310 MB), the single-launch chain 4x that (its only per-chunk buffer is the input planes)
    h->chain_chunk_max = chunk_rows > 0 ? ((chunk_rows + BM - 1) / BM) * BM : 128LL * 148 * 16;
    if (chunk_rows <= 0) chunk_rows = 128 * 148 * 4;
    h->chunk_max = ((chunk_rows + BM - 1) / BM) * BM;
    int maxw = 0;
    h->layers.resize(n_layers);
    for (int l = 0; l < n_layers; ++l) {
        Layer &L = h->layers[l];
        L.K = dims[l]; L.Kp = ((dims[l] + BK - 1) / BK) * BK; L.N = dims[l + 1];
        L.BN = (L.N % 128 == 0) ? 128 : 64;
        L.act = activations ? activations[l] : (l + 1 < n_layers);
        maxw = std::max(maxw, std::max(L.Kp, L.N));
    }
    h->maxw = maxw;
    int rc = LDPC_OK;
    auto fail = [&](int code) { ldpc_mlp_destroy(h); return code; };
    for (int l = 0; l < n_layers; ++l) {
        Layer &L = h->layers[l];
        const size_t wel = (size_t)L.N * L.Kp;
        float *tmp = nullptr;
        if (cudaMalloc(&L.d_w, h->ns * wel * sizeof(__half)) != cudaSuccess || cudaMalloc(&tmp, (size_t)L.N * L.K * sizeof(float)) != cudaSuccess) { cudaFree(tmp); set_error("ldpc_mlp_create: out of device memory"); return fail(LDPC_ENOMEM); }
        cudaMemcpy(tmp, weights[l], (size_t)L.N * L.K * sizeof(float), cudaMemcpyHostToDevice);
        if (h->ns == 1) rc = launch_split<1>(tmp, L.K, L.K, L.N, L.Kp, L.d_w, (long long)wel, 0);
        else if (h->ns == 2) rc = launch_split<2>(tmp, L.K, L.K, L.N, L.Kp, L.d_w, (long long)wel, 0);
        else rc = launch_split<3>(tmp, L.K, L.K, L.N, L.Kp, L.d_w, (long long)wel, 0);
        cudaDeviceSynchronize();
        cudaFree(tmp);
        if (rc) return fail(rc);
        if (biases && biases[l]) {
            if (cudaMalloc(&L.d_bias, L.N * sizeof(float)) != cudaSuccess) { set_error("ldpc_mlp_create: out of device memory"); return fail(LDPC_ENOMEM); }
            cudaMemcpy(L.d_bias, biases[l], L.N * sizeof(float), cudaMemcpyHostToDevice);
        }
        if ((rc = make_map(&L.map_w, L.d_w, L.Kp, L.N, h->ns, L.BN))) return fail(rc);
    }
    if (cudaGetLastError() != cudaSuccess) { set_error("ldpc_mlp_create: CUDA error while uploading the weights"); return fail(LDPC_ECUDA); }
    {
        int coop = 0, sms = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        h->n_groups = sms / GROUP;
        h->chain_ok = coop && h->ns == 2 && n_layers >= 2 && n_layers <= MAX_CHAIN && h->n_groups > 0;
        for (const Layer &L : h->layers) h->chain_ok = h->chain_ok && L.N / L.BN <= GROUP;
        if (h->chain_ok) {                                                // what the cooperative launch will check: one CTA per SM must fit
            constexpr int SMEM1 = 3 * 2 * (BM * BK * 2 + 128 * BK * 2) + 1024 + 128 + EPI_WARPS * 2048;
            int per_sm = 0;
            if (cudaFuncSetAttribute(chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM1) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_kernel<2>, CHAIN_THREADS, SMEM1) != cudaSuccess || per_sm < 1) {
                cudaGetLastError();
                h->chain_ok = false;                                      // AUTO then runs one launch per layer
            }
        }
        if (h->chain_ok) {                                                // and the pair kernel: every cluster of 2 must be resident at once
            constexpr int SMEM2 = 3 * (2 * BM * BK * 2 + 128 * BK * 2 + 64 * BK * 2) + 1024 + 128 + EPI_WARPS * 2048;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)(h->n_groups * GROUP)); cfg.blockDim = dim3(CHAIN_THREADS); cfg.dynamicSmemBytes = SMEM2;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int clusters = 0;
            h->pairs_ok = cudaFuncSetAttribute(chain2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2) == cudaSuccess &&
                          cudaOccupancyMaxActiveClusters(&clusters, chain2_kernel<2>, &cfg) == cudaSuccess && clusters * 2 >= h->n_groups * GROUP;
            if (!h->pairs_ok) cudaGetLastError();
        }
    }
    *out = h;
    return LDPC_OK;
}

// Activation buffers are sized for the batches actually seen (up to chunk_max rows): 4 KB per row at width 512.
static int ensure_activation_buffers(ldpc_mlp *h, long long rows, cudaStream_t s) {
    const long long want = std::min<long long>(h->chunk_max, ((rows + BM - 1) / BM) * BM);
    if (want <= h->chunk) return LDPC_OK;
    if (h->chunk) LDPC_CUDA_TRY(cudaStreamSynchronize(s));                // earlier launches may still read the old buffers
    for (int b = 0; b < 2; ++b) {
        cudaFree(h->d_act[b]);
        h->d_act[b] = nullptr;
        if (cudaMalloc(&h->d_act[b], (size_t)h->ns * want * h->maxw * sizeof(__half)) != cudaSuccess) {
            h->chunk = 0;
            set_error("ldpc_mlp_forward: out of device memory (%lld activation rows)", want);
            return LDPC_ENOMEM;
        }
    }
    h->chunk = want;
    for (size_t l = 0; l < h->layers.size(); ++l) {                       // input planes of layer l live in d_act[l & 1] with row length Kp
        Layer &L = h->layers[l];
        const int rc = make_map(&L.map_a, h->d_act[l & 1], L.Kp, h->chunk, h->ns, BM);
        if (rc) return rc;
    }
    return LDPC_OK;
}

// Single-launch chain: per chunk one plane-split launch of the input rows, one memset of the counters, one cooperative launch.
static int ensure_chain_buffers(ldpc_mlp *h, long long rows, cudaStream_t s) {
    const Layer &L0 = h->layers[0];
    if (!h->d_ring[0]) {
        const long long ring_rows = (long long)h->n_groups * INFLIGHT * BM;
        const size_t ring_elems = (size_t)h->ns * ring_rows * h->maxw;    // one allocation: the two buffers form ONE L2 persistence window
        if (cudaMalloc(&h->d_ring[0], 2 * ring_elems * sizeof(__half)) != cudaSuccess) { h->d_ring[0] = nullptr; set_error("ldpc_mlp_forward: out of device memory (activation ring)"); return LDPC_ENOMEM; }
        h->d_ring[1] = h->d_ring[0] + ring_elems;
        h->ring_bytes = 2 * ring_elems * sizeof(__half);
        if (cudaMalloc(&h->d_flags, (size_t)h->n_groups * INFLIGHT * 2 * MAX_CHAIN * sizeof(unsigned int)) != cudaSuccess) {
            cudaFree(h->d_ring[0]);                                           // all or nothing: the next call starts over
            h->d_ring[0] = h->d_ring[1] = nullptr;
            set_error("ldpc_mlp_forward: out of device memory");
            return LDPC_ENOMEM;
        }
        for (size_t l = 0; l < h->layers.size(); ++l) {
            Layer &L = h->layers[l];
            h->cmaps.w[l] = L.map_w;
            if (l) { const int rc = make_map(&h->cmaps.a[l], h->d_ring[l & 1], L.Kp, ring_rows, h->ns, BM, h->maxw); if (rc) return rc; }
            h->cmaps2.w[l] = L.map_w;
            h->cmaps2.a[l] = h->cmaps.a[l];
            { const int rc = make_map(&h->cmaps2.w2[l], L.d_w, L.Kp, L.N, h->ns, L.BN / 2); if (rc) return rc; }
        }
    }
    const long long want = std::min<long long>(h->chain_chunk_max, ((rows + BM - 1) / BM) * BM);
    if (want > h->x_rows) {
        if (h->x_rows) LDPC_CUDA_TRY(cudaStreamSynchronize(s));
        cudaFree(h->d_x);
        h->d_x = nullptr; h->x_rows = 0;
        // the input planes keep only ceil(K / 8) * 8 columns (65 -> 72, not 128): the TMA box of the last k-block runs past the tensor's
        // inner extent and is zero-filled there, which is what the zero-padded weight planes expect
        const int kx = ((L0.K + 7) / 8) * 8;
        if (cudaMalloc(&h->d_x, (size_t)h->ns * want * kx * sizeof(__half)) != cudaSuccess) { h->d_x = nullptr; set_error("ldpc_mlp_forward: out of device memory (%lld input rows)", want); return LDPC_ENOMEM; }
        h->x_rows = want;
        const int rc = make_map(&h->cmaps.a[0], h->d_x, kx, h->x_rows, h->ns, BM);
        if (rc) return rc;
        h->cmaps2.a[0] = h->cmaps.a[0];
    }
    return LDPC_OK;
}

static int forward_chain(ldpc_mlp *h, const float *x, long long B, float *y, cudaStream_t s, bool pairs) {
    { const int rc = ensure_chain_buffers(h, B, s); if (rc) return rc; }
    const int nl = (int)h->layers.size();
    const Layer &L0 = h->layers[0];
    constexpr int SMEM1 = 3 * 2 * (BM * BK * 2 + 128 * BK * 2) + 1024 + 128 + EPI_WARPS * 2048;
    constexpr int SMEM2 = 3 * (2 * BM * BK * 2 + 128 * BK * 2 + 64 * BK * 2) + 1024 + 128 + EPI_WARPS * 2048;
    if (pairs) LDPC_CUDA_TRY(cudaFuncSetAttribute(chain2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2));
    else LDPC_CUDA_TRY(cudaFuncSetAttribute(chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM1));
    const int rows_per_block = pairs ? 2 * BM : BM;
    for (long long done = 0; done < B; done += h->x_rows) {
        const long long rows = std::min<long long>(h->x_rows, B - done);
        const int kx = ((L0.K + 7) / 8) * 8;
        { const int rc = launch_split<2>(x + done * L0.K, L0.K, L0.K, rows, kx, h->d_x, h->x_rows * kx, s); if (rc) return rc; }
        ChainArgs a;
        memset(&a, 0, sizeof(a));
        a.n_layers = nl; a.m_valid = (int)rows; a.row_blocks = (int)((rows + rows_per_block - 1) / rows_per_block);
        a.n_groups = std::min(h->n_groups, a.row_blocks);
        for (int l = 0; l < nl; ++l) { const Layer &L = h->layers[l]; a.L[l] = ChainLayer{L.Kp / BK, L.N, L.BN, L.act, L.d_bias}; }
        a.ring[0] = h->d_ring[0]; a.ring[1] = h->d_ring[1];
        a.ring_rows = (long long)h->n_groups * INFLIGHT * BM;             // = n_groups * INFLIGHT2 * 256 for the pairs
        a.ring_pitch = h->maxw;
        a.out_f32 = y + done * h->layers[nl - 1].N;
        a.flags = h->d_flags;
        LDPC_CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, (size_t)h->n_groups * INFLIGHT * 2 * MAX_CHAIN * sizeof(unsigned int), s));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(a.n_groups * GROUP)); cfg.blockDim = dim3(CHAIN_THREADS); cfg.stream = s;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;     // every CTA of a group must be resident: the waits spin
        cfg.attrs = at; cfg.numAttrs = 1;
#ifdef MLP_EXP_PERSIST
        at[1].id = cudaLaunchAttributeAccessPolicyWindow;
        at[1].val.accessPolicyWindow.base_ptr = h->d_ring[0];
        at[1].val.accessPolicyWindow.num_bytes = h->ring_bytes;
        at[1].val.accessPolicyWindow.hitRatio = 1.0f;
        at[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.numAttrs = 2;
#endif
        if (pairs) { cfg.dynamicSmemBytes = SMEM2; LDPC_CUDA_TRY(cudaLaunchKernelEx(&cfg, chain2_kernel<2>, h->cmaps2, a)); }
        else { cfg.dynamicSmemBytes = SMEM1; LDPC_CUDA_TRY(cudaLaunchKernelEx(&cfg, chain_kernel<2>, h->cmaps, a)); }
    }
    return LDPC_OK;
}

int ldpc_mlp_forward(ldpc_mlp_t *h, const float *x, int64_t B, float *y, ldpc_stream_t stream) {
    if (!h || (B > 0 && (!x || !y)) || B < 0) { set_error("ldpc_mlp_forward: bad arguments"); return LDPC_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if (B == 0) return LDPC_OK;
    if (reinterpret_cast<uintptr_t>(y) & 15) { set_error("ldpc_mlp_forward: y must be 16-byte aligned"); return LDPC_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);   // host-side serialisation; work of different calls is still ordered per stream by the caller
    if (h->mode >= LDPC_MLP_CHAIN && !h->chain_ok) { set_error("ldpc_mlp_forward: the single-launch chain needs splits = 2, 2..%d layers of at most %d column tiles and cooperative launch", MAX_CHAIN, GROUP); return LDPC_EUNSUPPORTED; }
    if (h->mode == LDPC_MLP_CHAIN_PAIRS && !h->pairs_ok) { set_error("ldpc_mlp_forward: the single-launch chain on CTA pairs needs every cluster of 2 resident at once on this device"); return LDPC_EUNSUPPORTED; }
    // AUTO: the chain on CTA pairs for large batches (4 % faster from ~64 K rows), on single SMs for small ones (the pairs'
    // longer hand-over chain costs 59 against 45 us at the reference's 1 024-row batches, evaluate_quantized_snr.py:150-157)
    if (h->chain_ok && h->mode != LDPC_MLP_PER_LAYER)
        return forward_chain(h, x, B, y, s, h->mode == LDPC_MLP_CHAIN_PAIRS || (h->mode == LDPC_MLP_AUTO && h->pairs_ok && B >= 32768));
    {
        const int rc0 = ensure_activation_buffers(h, B, s);
        if (rc0) return rc0;
    }
    const int nl = (int)h->layers.size();
    const int K0 = h->layers[0].K, NL = h->layers[nl - 1].N;
    for (long long done = 0; done < B; done += h->chunk) {
        const long long rows = std::min<long long>(h->chunk, B - done);
        const Layer &L0 = h->layers[0];
        // the map of layer 0 describes [chunk][Kp0] planes with plane stride chunk * Kp0
        const long long ps0 = h->chunk * L0.Kp;
        int rc;
        if (h->ns == 1) rc = launch_split<1>(x + done * K0, K0, K0, rows, L0.Kp, h->d_act[0], ps0, s);
        else if (h->ns == 2) rc = launch_split<2>(x + done * K0, K0, K0, rows, L0.Kp, h->d_act[0], ps0, s);
        else rc = launch_split<3>(x + done * K0, K0, K0, rows, L0.Kp, h->d_act[0], ps0, s);
        if (rc) return rc;
        for (int l = 0; l < nl; ++l) {
            const Layer &L = h->layers[l];
            LayerArgs a;
            memset(&a, 0, sizeof(a));
            a.k_blocks = L.Kp / BK; a.m_valid = (int)rows; a.m_rows = (int)rows; a.n_total = L.N; a.act = L.act; a.bias = L.d_bias;
            if (l + 1 < nl) { a.out_planes = h->d_act[(l + 1) & 1]; a.out_plane_stride = h->chunk * (long long)L.N; }
            else a.out_f32 = y + done * NL;
            if (h->ns == 1) rc = launch_layer_ns<1>(L, a, rows, s);
            else if (h->ns == 2) rc = launch_layer_ns<2>(L, a, rows, s);
            else rc = launch_layer_ns<3>(L, a, rows, s);
            if (rc) return rc;
        }
    }
    return LDPC_OK;
}

int ldpc_mlp_set_mode(ldpc_mlp_t *h, int mode) {
    if (!h || mode < LDPC_MLP_AUTO || mode > LDPC_MLP_CHAIN_PAIRS) { set_error("ldpc_mlp_set_mode: bad arguments"); return LDPC_EINVAL; }
    if ((mode >= LDPC_MLP_CHAIN && !h->chain_ok) || (mode == LDPC_MLP_CHAIN_PAIRS && !h->pairs_ok)) { set_error("ldpc_mlp_set_mode: this network / device cannot run the single-launch chain"); return LDPC_EUNSUPPORTED; }
    std::lock_guard<std::mutex> lock(h->mu);
    h->mode = mode;
    return LDPC_OK;
}

}  // extern "C"
