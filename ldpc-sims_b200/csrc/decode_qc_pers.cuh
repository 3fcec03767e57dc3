// decode_qc_pers.cuh - the persistent, TMA-fed form of the code-specialised decoder (fp32 and f16x2).
//
// Same plan (qc_plan.cuh), same node arithmetic (node_math*.cuh) and therefore the same bits as decode_qc_kernel; what
// changes is how a CTA gets its data and how the variable phase is ordered:
//   * CTAs are persistent (two per SM): CTA b decodes tiles b, b + gridDim.x, ...
//   * the channel-LLR tile of the NEXT tile is fetched by one elected thread with a bulk copy (cp.async.bulk, completion
//     on an mbarrier -> SASS UBLKCP / SYNCS) into a staging buffer while the current tile decodes (the per-batch
//     host->device tensor of ofdm_functions.py:156 is the tile being staged); threads then take their NB LLRs from shared
//     memory with immediate offsets instead of 64-bit global address arithmetic.  Dtypes other than f32 / unaligned
//     pointers take a cooperative convert-and-stage path through the same buffer.
//   * the variable phase issues the shared-memory loads of the next batch of block columns before it computes and
//     stores the current one (VarBatches, qc_plan.cuh): its pointers are run-time selections, so the compiler cannot
//     reorder a load above a store by itself.
// Schedules measured and REJECTED on B200 (profiles/r02_decoder_schedule_experiments.md): one 512-thread CTA per SM
// with two codeword groups in lock-step anti-phase (CTA-wide barrier per half iteration, one group's check phase
// against the other's variable phase), the same two groups free-running on named barriers, and a generic slot loop.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "epilogue.cuh"
#include "node_math.cuh"
#include "node_math_h2.cuh"
#include "qc_plan.cuh"
#include "qc_var_pipe.cuh"

namespace ldpc {

// ---- mbarrier / bulk-copy wrappers (PTX ISA: mbarrier, cp.async.bulk) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    const uint32_t addr = smem_u32(b);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
    } while (!ok);
}

template <class T> struct cpt_of { static constexpr int value = 1; };
template <> struct cpt_of<__half2> { static constexpr int value = 2; };

template <class Code, int CWT, int CPT>
struct PersLayout {
    static constexpr int Z = Code::Z, N = Code::NB * Z;
    static constexpr int NLOC = kQc<Code>.n_local, NSM = kQc<Code>.n_smem;
    static constexpr int CWG = CWT * CPT;                          // codewords per tile
    static constexpr int THREADS = ((CWT * Z + 31) / 32) * 32;
    static constexpr int MIN_CTAS = THREADS <= 256 ? 2 : 1;
    static constexpr int HARD_STRIDE = (N + 15) & ~15;
    static constexpr size_t MSG_BYTES = (4 * (size_t)NSM * Z * CWT + 15) & ~size_t(15);
    static constexpr size_t STAGE_BYTES = (size_t)CWG * N * 4;    // f32 rows
    static constexpr size_t HARD_BYTES = (size_t)CWG * HARD_STRIDE;
    static constexpr int SCR_INTS = 3 + 2 * CWG;                  // {uncoded, info, -} + frame flags [CWG] + syndrome weights [CWG]
    static constexpr size_t SCR_BYTES = (SCR_INTS * 4 + 15) & ~size_t(15);
    static constexpr size_t TILE_BYTES = MSG_BYTES + STAGE_BYTES + HARD_BYTES + SCR_BYTES;
    static constexpr size_t SMEM = 16 + TILE_BYTES;               // the mbarrier in front
    static_assert(N % 4 == 0, "bulk copies move whole 16-byte units: the f32 row must be a multiple of 16 bytes");
};

// ---- group-scoped tails (the CTA-scoped forms live in epilogue.cuh) ---------------------------------------
__device__ __forceinline__ void pack_hard_group(const uint8_t *hard_s, int hs_stride, int ncw, int n, uint8_t *packed_g, int tid, int nthr) {
    const int nbytes = (n + 7) >> 3;
    const int nfull = n >> 3;                                       // hard_s rows are 16-byte aligned (PersLayout)
    for (int i = tid; i < ncw * nbytes; i += nthr) {
        const int cw = i / nbytes, by = i - cw * nbytes;
        const uint8_t *h = hard_s + cw * hs_stride + by * 8;
        unsigned v = 0;
        if (by < nfull) {
            const unsigned long long y = *reinterpret_cast<const unsigned long long *>(h) & 0x0101010101010101ull;
            v = (unsigned)((y * 0x8040201008040201ull) >> 56);
        } else {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int idx = by * 8 + b;
                v |= (idx < n ? (unsigned)(h[b] & 1) : 0u) << (7 - b);
            }
        }
        packed_g[(long long)cw * nbytes + by] = (uint8_t)v;
    }
}

// accumulate into scr[0] (uncoded bit errors), scr[1] (information-bit errors), scr[3 + cw] (frame error flag);
// flushed one barrier later by count_flush_group
__device__ __forceinline__ void count_accumulate_group(const uint8_t *hard_s, int hs_stride, int ncw, int n, int k_info,
                                                       const uint8_t *ref_packed_g, int *scr, int tid, int nthr) {
    const int nbytes = (n + 7) >> 3;
    int unc = 0, inf = 0;
    for (int i = tid; i < ncw * n; i += nthr) {
        const int cw = i / n, v = i - cw * n;
        const int hv = hard_s[cw * hs_stride + v];
        const int ref = (ref_packed_g[(long long)cw * nbytes + (v >> 3)] >> (7 - (v & 7))) & 1;
        const int hb = hv & 1, ub = (hv >> 1) & 1;
        unc += (ub != ref);
        const int e = (hb != ref);
        inf += e & (v < k_info);
        if (e) scr[3 + cw] = 1;                                    // benign race: everyone writes 1
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unc += __shfl_xor_sync(0xffffffffu, unc, o);
        inf += __shfl_xor_sync(0xffffffffu, inf, o);
    }
    if ((tid & 31) == 0) {
        if (unc) atomicAdd(&scr[0], unc);
        if (inf) atomicAdd(&scr[1], inf);
    }
}
__device__ __forceinline__ void count_flush_group(int ncw, int n, unsigned long long *counters, int *scr) {   // one thread
    int fe = 0;
    for (int c = 0; c < ncw; ++c) { fe += scr[3 + c]; scr[3 + c] = 0; }
    if (scr[0]) atomicAdd(&counters[0], (unsigned long long)scr[0]);
    if (scr[1]) atomicAdd(&counters[1], (unsigned long long)scr[1]);
    if (fe) atomicAdd(&counters[2], (unsigned long long)fe);
    atomicAdd(&counters[3], (unsigned long long)ncw * (unsigned long long)n);
    atomicAdd(&counters[4], (unsigned long long)ncw);
    scr[0] = 0;
    scr[1] = 0;
}

// =========================================================================================================
// Persistent form of decode_qc_kernel (fixed iteration count, channel LLRs from global memory): the same phases in the
// same order, but a CTA walks tiles  blockIdx.x, blockIdx.x + gridDim.x, ...  and the LLR tile of its NEXT tile is in
// flight (bulk copy into `stage`) while the current one decodes.
template <class Code, int CWT, int UPD, class T, int VB>
__global__ void __launch_bounds__((PersLayout<Code, CWT, cpt_of<T>::value>::THREADS), (PersLayout<Code, CWT, cpt_of<T>::value>::MIN_CTAS))
decode_qc_pers_kernel(const DecodeArgs a) {
    constexpr int CPT = cpt_of<T>::value;
    using L = PersLayout<Code, CWT, CPT>;
    constexpr int Z = Code::Z, NB = Code::NB, MB = Code::MB, N = L::N, TG = L::THREADS, CWG = L::CWG;
    static_assert(CPT == 1 || UPD == UPD_MINSUM || UPD == UPD_NMS, "the f16x2 format implements min-sum and normalized min-sum");
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    uint64_t *const mbar = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned char *const tile_s = smem_raw + 16;
    T *const msg_s = reinterpret_cast<T *>(tile_s);
    float *const stage = reinterpret_cast<float *>(tile_s + L::MSG_BYTES);
    uint8_t *const hard_s = tile_s + L::MSG_BYTES + L::STAGE_BYTES;
    int *const scr = reinterpret_cast<int *>(hard_s + L::HARD_BYTES);

    // codewords interleaved by lane: thread = t * CWT + slot; message word (blk, z) of slot sl at (blk * Z + z) * CWT + sl
    const int t = tid / CWT, sl = tid - t * CWT;
    T *const msg = msg_s + ((t < Z) ? tid : 0);
    T *const lo = msg;
    T *const hi = msg + Z * CWT;

    const long long ntiles = (a.B + CWG - 1) / CWG;
    const bool tma_ok = a.llr_dtype == LDPC_F32 && (reinterpret_cast<uintptr_t>(a.llr) & 15u) == 0;
    const int nbytes = (N + 7) >> 3;

    NodeParams np;
    np.clampv = a.clampv; np.param = a.param;
    np.clamp_h = __float2half2_rn(a.clampv); np.alpha_h = __float2half2_rn(a.param);

    auto tile_ncw = [&](long long j) -> int {
        const long long r = a.B - j * CWG;
        return r <= 0 ? 0 : (int)(r < CWG ? r : CWG);
    };
    auto issue_tma = [&](long long j) {                              // one thread
        const int n = tile_ncw(j);
        if (n > 0) {
            const uint32_t bytes = (uint32_t)n * N * 4u;
            mbar_expect_tx(mbar, bytes);
            bulk_g2s(stage, reinterpret_cast<const float *>(a.llr) + j * CWG * (long long)N, bytes, mbar);
        }
    };
    auto fill_sync = [&](long long j) {                              // whole CTA: any dtype, any alignment
        const int n = tile_ncw(j);
        const long long src = j * CWG * (long long)N;
        for (int i = tid; i < n * N; i += TG) stage[i] = load_llr(a.llr, a.llr_dtype, src + i);
    };

    T llr[NB];
    T loc[L::NLOC > 0 ? L::NLOC : 1];

    auto var_phase_first = [&]() {                                  // C->V messages are the zeros every caller passes (ofdm_functions.py:157)
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            if constexpr (D > 0) {
                T in[D], out[D];
                static_for<D>([&](auto kk) { in[decltype(kk)::value] = zero_of(T()); });
                vnode<D, UPD>(in, llr[c], out);
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                    constexpr int slot = kQc<Code>.col_slot[c][k];
                    if constexpr (is_loc) loc[slot] = out[k];
                    else {
                        constexpr int s = kQc<Code>.col_eff[c][k];
                        constexpr int off = (slot * Z - s) * CWT;
                        ((t < s ? hi : lo) + off)[0] = out[k];
                    }
                });
            }
        });
    };
    // shared-memory loads of batch b+1 issued before batch b is computed and stored (VarBatches, qc_plan.cuh)
    auto var_phase = [&]() {
        using VP = VarPipe<Code, CWT, UPD, T, (VB > 0 ? VB : 0)>;
        T inA[VP::VBW], inB[VP::VBW];
        T *pA[VP::VBW], *pB[VP::VBW];
        VP::template load<0>(t, lo, hi, inA, pA);
        VP::template run<0>(t, lo, hi, llr, loc, inA, pA, inB, pB);
    };
    auto check_phase = [&]() {
        static_for<MB>([&](auto rr) {
            constexpr int r = decltype(rr)::value;
            constexpr int D = kQc<Code>.row_deg[r];
            if constexpr (D > 0) {
                T in[D], out[D];
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr bool is_loc = kQc<Code>.row_loc[r][j];
                    constexpr int slot = kQc<Code>.row_slot[r][j];
                    if constexpr (is_loc) in[j] = loc[slot];
                    else in[j] = msg[slot * Z * CWT];
                });
                cnode<D, UPD>(in, np, out);
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr bool is_loc = kQc<Code>.row_loc[r][j];
                    constexpr int slot = kQc<Code>.row_slot[r][j];
                    if constexpr (is_loc) loc[slot] = out[j];
                    else msg[slot * Z * CWT] = out[j];
                });
            }
        });
    };
    // LLR tile (staging buffer, f32 rows) -> registers, in this thread's lane relabelling
    auto load_llr_regs = [&](bool second) {
        const float *row0 = stage + (sl * CPT) * N;
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int rho = kQc<Code>.rho[c];
            int zv = t + rho;
            if (zv >= Z) zv -= Z;
            if constexpr (CPT == 1) llr[c] = row0[c * Z + zv];
            else llr[c] = __floats2half2_rn(sat_llr(row0[c * Z + zv]), second ? sat_llr(row0[N + c * Z + zv]) : 0.0f);
        });
    };
    // marginal, hard decision (hard_s: bit 0 decoded bit, bit 1 channel decision), posterior / probability / byte outputs
    auto marginal_phase = [&](long long cw0, bool second) {
        T tm[NB];
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            T in[D > 0 ? D : 1];
            static_for<D>([&](auto kk) {
                constexpr int k = decltype(kk)::value;
                constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                constexpr int slot = kQc<Code>.col_slot[c][k];
                if constexpr (is_loc) in[k] = loc[slot];
                else {
                    constexpr int s = kQc<Code>.col_eff[c][k];
                    constexpr int off = (slot * Z - s) * CWT;
                    in[k] = ((t < s ? hi : lo) + off)[0];
                }
            });
            tm[c] = mnode<(D > 0 ? D : 1)>(in, D, llr[c]);
        });
        auto tfv = [&](auto cc, int h) -> float {                    // marginal of codeword h of this thread as a float
            constexpr int c = decltype(cc)::value;
            if constexpr (CPT == 1) return tm[c];
            else return h == 0 ? __low2float(tm[c]) : __high2float(tm[c]);
        };
        unsigned hb[CPT], cb[CPT];
        float tmin = CUDART_INF_F;
#pragma unroll
        for (int h = 0; h < CPT; ++h) {
            hb[h] = 0; cb[h] = 0;
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const float v = tfv(cc, h);
                float l;
                if constexpr (CPT == 1) l = llr[c];
                else l = h == 0 ? __low2float(llr[c]) : __high2float(llr[c]);
                hb[h] |= (v < 0.0f ? 1u : 0u) << c;
                cb[h] |= (l > 0.0f ? 1u : 0u) << c;
                tmin = fminf(tmin, fabsf(v));
            });
        }
        if (!(tmin > 1e-5f)) {                                       // tie band (rare): round the way the reference does (node_math.cuh)
#pragma unroll
            for (int h = 0; h < CPT; ++h) {
                hb[h] = 0;
                static_for<NB>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    hb[h] |= (unsigned)hard_bit(tfv(cc, h)) << c;
                });
            }
        }
        const long long gbase = (cw0 + sl * CPT) * N;
        uint8_t *const hrow = hard_s + (sl * CPT) * L::HARD_STRIDE;
        auto stores = [&](auto with_post) {                         // the output pointer is tested ONCE around the store loop
            constexpr bool POST = decltype(with_post)::value;
            float *const post = a.llr_post + gbase;
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int rho = kQc<Code>.rho[c];
                int zv = t + rho;
                if (zv >= Z) zv -= Z;
                const int idx = c * Z + zv;
#pragma unroll
                for (int h = 0; h < CPT; ++h) {
                    if (h == 1 && !second) break;
                    hrow[h * L::HARD_STRIDE + idx] = (uint8_t)(((hb[h] >> c) & 1u) | (((cb[h] >> c) & 1u) << 1));
                    if constexpr (POST) post[h * N + idx] = __fmul_rn(-2.0f, tfv(cc, h));
                }
            });
        };
        if (a.llr_post) stores(std::true_type{});
        else stores(std::false_type{});
        if (a.prob || a.hard) {                                      // byte / probability outputs: cold path
#pragma unroll 1
            for (int c = 0; c < NB; ++c) {
                float tv[CPT];
                int rho = 0;
                static_for<NB>([&](auto cc) {
                    constexpr int c2 = decltype(cc)::value;
                    constexpr int rho2 = kQc<Code>.rho[c2];
                    if (c == c2) {
                        rho = rho2;
#pragma unroll
                        for (int h = 0; h < CPT; ++h) tv[h] = tfv(cc, h);
                    }
                });
                int zv = t + rho;
                if (zv >= Z) zv -= Z;
#pragma unroll
                for (int h = 0; h < CPT; ++h) {
                    if (h == 1 && !second) break;
                    const long long o = gbase + (long long)h * N + c * Z + zv;
                    if (a.prob) a.prob[o] = prob_one(tv[h]);
                    if (a.hard) a.hard[o] = (uint8_t)((hb[h] >> c) & 1u);
                }
            }
        }
    };
    auto syndrome_phase = [&](bool second) {                        // adds this thread's unsatisfied checks
#pragma unroll
        for (int h = 0; h < CPT; ++h) {
            if (h == 1 && !second) break;
            int w = 0;
            const uint8_t *hs = hard_s + (sl * CPT + h) * L::HARD_STRIDE;
            static_for<MB>([&](auto rr) {
                constexpr int r = decltype(rr)::value;
                constexpr int D = kQc<Code>.row_deg[r];
                constexpr int sg = kQc<Code>.sigma[r];
                int zc = t + sg;
                if (zc >= Z) zc -= Z;
                unsigned par = 0;
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr int s = kQc<Code>.row_shift[r][j];
                    constexpr int cbase = kQc<Code>.row_col[r][j] * Z;
                    int zv = zc + s;
                    if (zv >= Z) zv -= Z;
                    par ^= hs[cbase + zv] & 1u;
                });
                w += (int)par;
            });
            if (w) atomicAdd(&scr[3 + CWG + sl * CPT + h], w);
        }
    };

    // ---- set-up: barrier object, scratch, first tile -----------------------------------------------------
    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < L::SCR_INTS; i += TG) scr[i] = 0;
    __syncthreads();
    if (tma_ok) { if (tid == 0) issue_tma(blockIdx.x); }
    else fill_sync(blockIdx.x);
    __syncthreads();

    uint32_t parity = 0;
#pragma unroll 1
    for (long long j = blockIdx.x; j < ntiles; j += gridDim.x) {
        const int ncw = tile_ncw(j);
        const long long cw0 = j * CWG;
        const bool act = (sl * CPT < ncw) && (t < Z);                // this thread holds a codeword of the tile
        const bool sec = sl * CPT + 1 < ncw;                         // (f16x2) the .y half holds one too
        if (tma_ok) { mbar_wait(mbar, parity); parity ^= 1u; }
        if (act) {
            load_llr_regs(sec);
            var_phase_first();
        }
        __syncthreads();                                             // every thread has taken its LLRs: the staging buffer is free
        if (tma_ok && tid == 0) issue_tma(j + gridDim.x);
        if (act) check_phase();
        __syncthreads();
#pragma unroll 1
        for (int it = 1; it < a.iters; ++it) {
            if (act) var_phase();
            __syncthreads();
            if (act) check_phase();
            __syncthreads();
        }
        if (act) marginal_phase(cw0, sec);
        if (!tma_ok) fill_sync(j + gridDim.x);
        __syncthreads();
        // ---- tail: everything that reads the tile's hard decisions from shared memory ------------------------
        if (a.hard_packed) pack_hard_group(hard_s, L::HARD_STRIDE, ncw, N, a.hard_packed + cw0 * nbytes, tid, TG);
        if (a.iters_used && tid < ncw) a.iters_used[cw0 + tid] = a.iters;
        if (a.syndrome || a.counters) {
            if (a.syndrome && act) syndrome_phase(sec);
            if (a.counters) count_accumulate_group(hard_s, L::HARD_STRIDE, ncw, N, a.k_info, a.ref_packed + cw0 * nbytes, scr, tid, TG);
            __syncthreads();
            if (a.syndrome && tid < ncw) { a.syndrome[cw0 + tid] = scr[3 + CWG + tid]; scr[3 + CWG + tid] = 0; }
            if (a.counters && tid == 32) count_flush_group(ncw, N, a.counters, scr);
        }
    }
}

// ---- host launcher -------------------------------------------------------------------------------------
inline int device_sm_count() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <class Code, int CWT, int UPD, class T, int VB>
int launch_qc_pers(const DecodeArgs &a, cudaStream_t s) {
    using L = PersLayout<Code, CWT, cpt_of<T>::value>;
    const long long ntiles = (a.B + L::CWG - 1) / L::CWG;
    const long long slots = (long long)device_sm_count() * L::MIN_CTAS;      // persistent CTAs: as many as are resident at once
    const int grid = (int)(ntiles < slots ? ntiles : slots);
    auto k = decode_qc_pers_kernel<Code, CWT, UPD, T, VB>;
    LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
    k<<<grid, L::THREADS, L::SMEM, s>>>(a);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

}  // namespace ldpc
