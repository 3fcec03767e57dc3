// decode_tiny.cu - register-resident decoder for SMALL codes whose parity-check matrix is known at
// compile time: the reference's default (64,32) code (bp/parity.py:7-47) first.
//
// One THREAD decodes one codeword.  The whole Tanner graph is compiled into the instruction stream
// (every edge index a compile-time constant), so the E edge messages and the n channel LLRs are plain
// registers: no shared memory, no barriers, no index loads.  HBM traffic is the LLR row in and the
// requested outputs out.  This is the (64,32) counterpart of decode_qc.cu and replaces, for this code,
// the dense E x E masked products of the reference (bp/bp_vc.py:19, bp/bp_cv.py:24-42).
// Arithmetic = node_math.cuh in the generic kernel's order (variable-major edges ascending in the check
// index, check-major edges ascending in the variable index), so results are bit-identical to
// decode_generic.cu and to the CPU oracle.
#include <utility>

#include "common.cuh"
#include "node_math.cuh"
#include "qc_plan.cuh"          // static_for

namespace ldpc {

// ---- compile-time codes -------------------------------------------------------------------------------------
struct Peg64x32 {                                        // bp/parity.py:7-47: row r = {r/2, second[r], 32 + r}
    static constexpr int N = 64, M = 32, DC = 3, MAXDV = 2;
    static constexpr int second[32] = {16, 17, 16, 18, 17, 19, 18, 20, 19, 21, 20, 22, 21, 23, 22, 24,
                                       23, 25, 24, 26, 25, 27, 26, 28, 27, 29, 28, 30, 29, 31, 30, 31};
    static constexpr int var_of(int r, int j) { return j == 0 ? r / 2 : (j == 1 ? second[r] : 32 + r); }
};

template <class Code>
struct TinyPlan {
    static constexpr int N = Code::N, M = Code::M, DC = Code::DC, E = M * DC;
    int chk_var[M][DC] = {};
    int dv[N] = {};
    int var_edge[N][Code::MAXDV] = {};                   // check-major edge ids of a variable, ascending check
    int var_base[N] = {};                                // variable-major id of a variable's first edge
    constexpr TinyPlan() {
        for (int r = 0; r < M; ++r)
            for (int j = 0; j < DC; ++j) {
                const int v = Code::var_of(r, j);
                chk_var[r][j] = v;
                var_edge[v][dv[v]++] = r * DC + j;       // rows visited ascending => checks ascending
            }
        for (int v = 1; v < N; ++v) var_base[v] = var_base[v - 1] + dv[v - 1];
    }
};
template <class Code>
inline constexpr TinyPlan<Code> kTiny{};

// marginal t_v of one variable (recomputed where needed: keeping all N of them live next to the messages spills)
template <class Code, int v, bool WT = false>
__device__ __forceinline__ float tiny_marginal(const float (&x)[Code::M * Code::DC], const float (&llr)[Code::N], const DecodeArgs &a) {
    constexpr int D = kTiny<Code>.dv[v];
    float in[D > 0 ? D : 1];
    static_for<D>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        constexpr int e = kTiny<Code>.var_edge[v][k];
        in[k] = x[e];
    });
    if constexpr (WT) {
        constexpr int vb = kTiny<Code>.var_base[v];
        return marginal_t_weighted<(D > 0 ? D : 1)>(in, D, llr[v], __ldg(a.wf_llr + v), a.wf_edge + vb);
    } else {
        return marginal_t<(D > 0 ? D : 1)>(in, D, llr[v]);
    }
}

// hard decision (bit v of hb) - also the convergence test of the early-termination mode
template <class Code, bool WT>
__device__ __forceinline__ void tiny_hard(const float (&x)[Code::M * Code::DC], const float (&llr)[Code::N], const DecodeArgs &a,
                                          unsigned long long &hb) {
    constexpr int N = Code::N;
    float tmin = CUDART_INF_F;
    hb = 0;
    static_for<N>([&](auto vv) {
        constexpr int v = decltype(vv)::value;
        const float t = tiny_marginal<Code, v, WT>(x, llr, a);
        tmin = fminf(tmin, fabsf(t));
        hb |= (unsigned long long)(t < 0.0f ? 1u : 0u) << v;
    });
    if (!(tmin > 1e-5f)) {                               // tie band (rare): round the way the reference does
        hb = 0;
        static_for<N>([&](auto vv) {
            constexpr int v = decltype(vv)::value;
            hb |= (unsigned long long)hard_bit(tiny_marginal<Code, v, WT>(x, llr, a)) << v;
        });
    }
}

template <class Code>
__device__ __forceinline__ int tiny_syndrome(unsigned long long hb) {
    int w = 0;
    static_for<Code::M>([&](auto rr) {
        constexpr int r = decltype(rr)::value;
        unsigned par = 0;
        static_for<Code::DC>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            constexpr int v = kTiny<Code>.chk_var[r][j];
            par ^= (unsigned)(hb >> v) & 1u;
        });
        w += (int)par;
    });
    return w;
}

// ---- kernel ---------------------------------------------------------------------------------------------------
// EE: syndrome-based early termination compiled in (a separate instantiation: the convergence test inside the loop costs
// the fixed-iteration path 25 % through register pressure)
// WT: the reference's trainable weights (bp_vc.py:16-32) read from the tables of ldpc_decode_weighted (broadcast loads:
// every thread of a warp reads the same word)
template <class Code, int UPD, bool EE, bool WT = false>
__global__ void __launch_bounds__(128) decode_tiny_kernel(const DecodeArgs a) {
    constexpr bool IS_SP = (UPD == UPD_SP);
    constexpr int N = Code::N, M = Code::M, DC = Code::DC, E = M * DC, NBY = (N + 7) / 8;
    const long long cw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = cw < a.B;
    const long long row = active ? cw : 0;

    float llr[N], x[E];
    if (a.llr_dtype == LDPC_F32 && (N % 4) == 0) {
        const float4 *src = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(a.llr) + row * N);
        static_for<N / 4>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            const float4 v = __ldg(src + q);
            llr[4 * q] = v.x; llr[4 * q + 1] = v.y; llr[4 * q + 2] = v.z; llr[4 * q + 3] = v.w;
        });
    } else {
        static_for<N>([&](auto vv) {
            constexpr int v = decltype(vv)::value;
            llr[v] = load_llr(a.llr, a.llr_dtype, row * N + v);
        });
    }
    static_for<E>([&](auto ee) { x[decltype(ee)::value] = 0.0f; });      // the zeros every reference caller passes (ofdm_functions.py:157)

    unsigned long long hb = 0;
    auto marginal_and_hard = [&]() { tiny_hard<Code, WT>(x, llr, a, hb); };
    auto syndrome_weight = [&]() { return tiny_syndrome<Code>(hb); };

    int wit = 0;                                         // iteration index of the weight tables
    auto iterate = [&]() {
        // V -> C, in place
        static_for<N>([&](auto vv) {
            constexpr int v = decltype(vv)::value;
            constexpr int D = kTiny<Code>.dv[v];
            if constexpr (D > 0) {
                float in[D], out[D];
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr int e = kTiny<Code>.var_edge[v][k];
                    in[k] = x[e];
                });
                if constexpr (WT) {
                    constexpr int vb = kTiny<Code>.var_base[v];
                    var_node_weighted<D, IS_SP>(in, D, llr[v], __ldg(a.w_llr + (long long)wit * N + v),
                                                a.w_edge + ((long long)wit * E + vb) * a.w_stride, a.w_stride, out);
                }
                else
                    var_node<D, IS_SP>(in, D, llr[v], out);
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr int e = kTiny<Code>.var_edge[v][k];
                    x[e] = out[k];
                });
            }
        });
        // C -> V, in place
        static_for<M>([&](auto rr) {
            constexpr int r = decltype(rr)::value;
            float in[DC], out[DC];
            static_for<DC>([&](auto jj) { constexpr int j = decltype(jj)::value; in[j] = x[r * DC + j]; });
            if constexpr (IS_SP) check_node_sp<DC>(in, DC, a.clampv, out);
            else check_node_ms_ct<DC, UPD>(in, a.clampv, a.param, out);
            static_for<DC>([&](auto jj) { constexpr int j = decltype(jj)::value; x[r * DC + j] = out[j]; });
        });
    };
    int used = a.iters;
    if constexpr (!EE) {
        // fixed iteration count (the reference's schedule, bp/bp.py:46-47)
#pragma unroll 1
        for (int it = 0; it < a.iters; ++it) { wit = it; iterate(); }
        marginal_and_hard();
    } else {
        // syndrome-based early termination (not in the reference, off in parity runs): marginal + hard decision after EVERY
        // iteration; a codeword whose hard decision satisfies every check stops there.  Every phase appears once in the
        // loop (a second copy of the marginal block pushes some instantiations over the inliner's budget and their
        // register arrays into local memory).
        int it = 0;
#pragma unroll 1
        for (;;) {
            if (it < a.iters) { wit = it; iterate(); ++it; }
            marginal_and_hard();
            if (it >= a.iters) break;
            if (syndrome_weight() == 0) { used = it; break; }
        }
    }

    // ---- outputs --------------------------------------------------------------------------------------------
    unsigned long long ub = 0;                           // bit v = uncoded channel decision (llr > 0)
    static_for<N>([&](auto vv) {
        constexpr int v = decltype(vv)::value;
        ub |= (unsigned long long)(llr[v] > 0.0f ? 1u : 0u) << v;
    });
    // MSB-first bytes (numpy.packbits): byte b bit 7-j = bit 8b+j
    auto pack = [&](unsigned long long bits) {
        unsigned long long w = 0;                        // little-endian word whose byte b is the packed byte b
        static_for<NBY>([&](auto bb) {
            constexpr int b = decltype(bb)::value;
            const unsigned byte = __brev((unsigned)((bits >> (8 * b)) & 0xffu)) >> 24;
            w |= (unsigned long long)byte << (8 * b);
        });
        return w;
    };
    const unsigned long long hpk = pack(hb);

    if (active) {
        if (a.llr_post) {
            float4 *dst = reinterpret_cast<float4 *>(a.llr_post + row * N);
            static_for<N / 4>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                dst[q] = make_float4(__fmul_rn(-2.0f, tiny_marginal<Code, 4 * q, WT>(x, llr, a)), __fmul_rn(-2.0f, tiny_marginal<Code, 4 * q + 1, WT>(x, llr, a)),
                                     __fmul_rn(-2.0f, tiny_marginal<Code, 4 * q + 2, WT>(x, llr, a)), __fmul_rn(-2.0f, tiny_marginal<Code, 4 * q + 3, WT>(x, llr, a)));
            });
        }
        if (a.prob) {
            float4 *dst = reinterpret_cast<float4 *>(a.prob + row * N);
#pragma unroll 1
            for (int q = 0; q < N / 4; ++q) {
                float t4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                static_for<N / 4>([&](auto qq) {
                    constexpr int q2 = decltype(qq)::value;
                    if (q == q2) {
                        t4[0] = tiny_marginal<Code, 4 * q2, WT>(x, llr, a); t4[1] = tiny_marginal<Code, 4 * q2 + 1, WT>(x, llr, a);
                        t4[2] = tiny_marginal<Code, 4 * q2 + 2, WT>(x, llr, a); t4[3] = tiny_marginal<Code, 4 * q2 + 3, WT>(x, llr, a);
                    }
                });
                dst[q] = make_float4(prob_one(t4[0]), prob_one(t4[1]), prob_one(t4[2]), prob_one(t4[3]));
            }
        }
        if (a.hard) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(a.hard + row * N);
            static_for<N / 4>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                const unsigned nib = (unsigned)(hb >> (4 * q)) & 0xfu;
                dst[q] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
            });
        }
        if (a.hard_packed) {
            if constexpr (NBY == 8) *reinterpret_cast<unsigned long long *>(a.hard_packed + row * NBY) = hpk;
            else static_for<NBY>([&](auto bb) { constexpr int b = decltype(bb)::value; a.hard_packed[row * NBY + b] = (uint8_t)(hpk >> (8 * b)); });
        }
        if (a.syndrome) a.syndrome[row] = syndrome_weight();
        if (a.iters_used) a.iters_used[row] = used;
    }
    // ---- fused exact link metrics (evaluate_quantized_snr.py:169-188) ----------------------------------------
    if (a.counters) {
        int unc = 0, inf = 0, fe = 0;
        if (active) {
            unsigned long long ref = 0;
            static_for<NBY>([&](auto bb) {
                constexpr int b = decltype(bb)::value;
                ref |= (unsigned long long)__ldg(a.ref_packed + row * NBY + b) << (8 * b);
            });
            unsigned long long kmask = 0;                                  // packed-byte mask of the first k_info bits
            static_for<NBY>([&](auto bb) {
                constexpr int b = decltype(bb)::value;
                const int rem = a.k_info - 8 * b;
                const unsigned m8 = rem >= 8 ? 0xffu : (rem <= 0 ? 0u : ((0xffu << (8 - rem)) & 0xffu));
                kmask |= (unsigned long long)m8 << (8 * b);
            });
            const unsigned long long de = hpk ^ ref;
            unc = __popcll(pack(ub) ^ ref);
            inf = __popcll(de & kmask);
            fe = de != 0;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, active);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unc += __shfl_xor_sync(0xffffffffu, unc, o);
            inf += __shfl_xor_sync(0xffffffffu, inf, o);
            fe += __shfl_xor_sync(0xffffffffu, fe, o);
        }
        if ((threadIdx.x & 31) == 0 && ballot) {
            const int nact = __popc(ballot);
            if (unc) atomicAdd(&a.counters[0], (unsigned long long)unc);
            if (inf) atomicAdd(&a.counters[1], (unsigned long long)inf);
            if (fe) atomicAdd(&a.counters[2], (unsigned long long)fe);
            atomicAdd(&a.counters[3], (unsigned long long)nact * N);
            atomicAdd(&a.counters[4], (unsigned long long)nact);
        }
    }
}

// ---- registry ---------------------------------------------------------------------------------------------------
template <class Code>
static bool tiny_matches(int m, int n, const int32_t *row_ptr, const int32_t *col_idx) {
    if (m != Code::M || n != Code::N) return false;
    for (int r = 0; r < m; ++r) {
        if (row_ptr[r + 1] - row_ptr[r] != Code::DC) return false;
        for (int j = 0; j < Code::DC; ++j)
            if (col_idx[row_ptr[r] + j] != kTiny<Code>.chk_var[r][j]) return false;
    }
    return true;
}

int tiny_lookup(int m, int n, const int32_t *row_ptr, const int32_t *col_idx) {
    if (tiny_matches<Peg64x32>(m, n, row_ptr, col_idx)) return 0;
    return -1;
}

template <class Code>
static int launch_tiny_t(const DecodeArgs &a, cudaStream_t s) {
    const long long grid = (a.B + 127) / 128;
    if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
    void (*k)(const DecodeArgs) = nullptr;
    if (a.w_edge) {                                      // trainable weights: sum-product and min-sum, fixed iteration count
        if (a.update == UPD_SP) k = decode_tiny_kernel<Code, UPD_SP, false, true>;
        else if (a.update == UPD_MINSUM) k = decode_tiny_kernel<Code, UPD_MINSUM, false, true>;
        else return LDPC_EUNSUPPORTED;
        k<<<(int)grid, 128, 0, s>>>(a);
        LDPC_CUDA_TRY(cudaGetLastError());
        return LDPC_OK;
    }
    switch (a.update) {
        case UPD_SP: k = a.early_exit ? decode_tiny_kernel<Code, UPD_SP, true> : decode_tiny_kernel<Code, UPD_SP, false>; break;
        case UPD_MINSUM: k = a.early_exit ? decode_tiny_kernel<Code, UPD_MINSUM, true> : decode_tiny_kernel<Code, UPD_MINSUM, false>; break;
        case UPD_NMS: k = a.early_exit ? decode_tiny_kernel<Code, UPD_NMS, true> : decode_tiny_kernel<Code, UPD_NMS, false>; break;
        default: k = a.early_exit ? decode_tiny_kernel<Code, UPD_OMS, true> : decode_tiny_kernel<Code, UPD_OMS, false>; break;
    }
    k<<<(int)grid, 128, 0, s>>>(a);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int launch_decode_tiny(int tiny_id, const DecodeArgs &a, cudaStream_t s) {
    if (a.B <= 0) return LDPC_OK;
    if (tiny_id == 0) return launch_tiny_t<Peg64x32>(a, s);
    set_error("unknown register-resident specialisation %d", tiny_id);
    return LDPC_EINVAL;
}

}  // namespace ldpc
