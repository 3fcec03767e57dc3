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
// Activations of a chunk of rows ping-pong between two plane buffers that stay L2-resident.
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// libm tanhf (~1 ulp): measured FASTER in this epilogue than an ex2.approx/rcp.approx formulation (twice, on two
// kernel generations) and it keeps the activations within an ulp of the reference's ATen tanh.
__device__ __forceinline__ float tanh_act(float x) { return tanhf(x); }

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
    const int groups = Kp >> 3;                                           // Kp is a multiple of 64
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
                    mbar_expect_tx(full_bar(s), S::STAGE_BYTES);
                    const uint32_t st = base + s * S::STAGE_BYTES;
#pragma unroll
                    for (int p = 0; p < NS; ++p) {
                        tma_load_3d(st + p * S::A_BYTES, &map_a, full_bar(s), kb * BK, m0, p);
                        tma_load_3d(st + NS * S::A_BYTES + p * S::B_BYTES, &map_w, full_bar(s), kb * BK, n0, p);
                    }
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
            if (warp == 2 && lane == 0) TRACE(2);
            const uint32_t d_main = tmem_base + acc * ACC_COLS + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int c0 = part * PART_COLS; c0 < (part + 1) * PART_COLS && c0 < BN; c0 += 32) {
                uint32_t v[32], w[32];
                tmem_ld32(d_main + (uint32_t)c0, v);
                if (NS > 1) tmem_ld32(d_main + ACC2 + (uint32_t)c0, w);
                float o[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float f = __uint_as_float(v[j]);
                    if (NS > 1) f = __fadd_rn(f, __uint_as_float(w[j]));
                    if (args.bias) f = __fadd_rn(f, __ldg(args.bias + n0 + c0 + j));
                    o[j] = args.act ? tanh_act(f) : f;
                }
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
static int make_map(CUtensorMap *map, void *ptr, int Kp, long long rows, int ns, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LDPC_ECUDA; }
    const cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)ns};
    const cuuint64_t gstr[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)rows * Kp * 2};
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
    long long chunk_max = 0;
    int maxw = 0;              // widest layer (elements per row)
    std::vector<ldpc::mlp::Layer> layers;
    __half *d_act[2] = {nullptr, nullptr};   // ping-pong activation planes [NS][chunk][maxw]
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
    if (chunk_rows <= 0) chunk_rows = 128 * 148 * 4;                      // 592 row tiles x 4 column tiles at N = 512: 16 tiles per persistent CTA (fill/drain amortised)
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

int ldpc_mlp_forward(ldpc_mlp_t *h, const float *x, int64_t B, float *y, ldpc_stream_t stream) {
    if (!h || (B > 0 && (!x || !y)) || B < 0) { set_error("ldpc_mlp_forward: bad arguments"); return LDPC_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if (B == 0) return LDPC_OK;
    if (reinterpret_cast<uintptr_t>(y) & 15) { set_error("ldpc_mlp_forward: y must be 16-byte aligned"); return LDPC_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);   // host-side serialisation; work of different calls is still ordered per stream by the caller
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

}  // extern "C"
