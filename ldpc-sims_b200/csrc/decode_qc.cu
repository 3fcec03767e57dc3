// decode_qc.cu - code-specialised belief-propagation decoder for quasi-cyclic codes whose
// prototype matrix is known at compile time (IEEE 802.11n n=1944 R=1/2 Z=81 first).
//
// The parity-check matrix is compiled INTO the instruction stream: every block's shift and
// shared-memory offset is an immediate, every node degree a compile-time loop bound, so the
// inner loops carry no index loads at all (the reference multiplies by dense E x E masks,
// bp/masking.py:12-147, bp/bp_vc.py:19, bp/bp_cv.py:24-42).
//
// Mapping: thread = (codeword cw of the CTA's tile, lane t in [0,Z)); see QcPlan for which
// check / variable of each block row / column a lane computes.  Channel LLRs (NB per thread)
// and the messages of the thread-local blocks stay in REGISTERS for the whole decode; the
// other blocks keep one fp32 slot per edge in shared memory at  blk*Z + tc  (tc = the check's
// thread), so the check phase is a pure linear access and the variable phase reads a rotated
// window ((t - s') mod Z).  HBM traffic is the LLR load and the result store only.
// Per iteration: variable phase (NB unrolled block columns per thread), barrier, check
// phase (MB unrolled block rows per thread), barrier.  Arithmetic = node_math.cuh, so the
// results are bit-identical to the generic kernel and to the CPU oracle's definition.
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "epilogue.cuh"
#include "node_math.cuh"
#include "qc_plan.cuh"

namespace ldpc {

template <class Code, int CW, int UPD>
__global__ void __launch_bounds__((QcLayout<Code, CW>::THREADS)) decode_qc_kernel(const DecodeArgs a) {
    using L = QcLayout<Code, CW>;
    constexpr bool IS_SP = (UPD == UPD_SP);
    constexpr int Z = Code::Z, NB = Code::NB, MB = Code::MB, N = L::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *msg_s = reinterpret_cast<float *>(smem_raw);
    uint8_t *hard_s = reinterpret_cast<uint8_t *>(msg_s + CW * L::MSG_STRIDE);
    int *scratch = reinterpret_cast<int *>(hard_s + CW * L::HARD_STRIDE);       // [4 + CW]

    const int tid = threadIdx.x, T = blockDim.x;
    const long long cw0 = (long long)blockIdx.x * CW;
    const int ncw = (int)min((long long)CW, a.B - cw0);
    const int cw = tid / Z, t = tid - cw * Z;
    const bool active = cw < ncw;            // also false for the padding threads (cw >= CW)
    for (int i = tid; i < 4 + CW; i += T) scratch[i] = 0;

    float *const msg = msg_s + (active ? cw : 0) * L::MSG_STRIDE;
    // rotated window bases: slot (blk, (t - s') mod Z) = (t < s' ? hi : lo)[blk*Z - s']
    float *const lo = msg + t;
    float *const hi = msg + t + Z;

    // ---- channel LLRs of this thread's NB variables live in registers for the whole decode ----------
    float llr[NB];
    float loc[L::NLOC > 0 ? L::NLOC : 1];
    const long long gbase = (cw0 + (active ? cw : 0)) * N;
    if (active) {
        auto load_all = [&](auto ld) {                                    // one uniform dtype branch, then NB straight loads
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int rho = kQc<Code>.rho[c];
                int zv = t + rho;
                if (zv >= Z) zv -= Z;
                llr[c] = ld(gbase + c * Z + zv);
            });
        };
        if (a.llr_dtype == LDPC_F32) load_all([&](long long i) { return __ldg(reinterpret_cast<const float *>(a.llr) + i); });
        else if (a.llr_dtype == LDPC_F64) load_all([&](long long i) { return (float)__ldg(reinterpret_cast<const double *>(a.llr) + i); });
        else load_all([&](long long i) { return __half2float(__ldg(reinterpret_cast<const __half *>(a.llr) + i)); });
    }

    // V -> C for all NB block columns of this thread.  FIRST: the C->V messages are still the
    // zeros every reference caller passes (ofdm_functions.py:157) - nothing is loaded.
    auto var_phase = [&](auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            if constexpr (D > 0) {
                float in[D], out[D];
                float *ptr[D];
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                    constexpr int slot = kQc<Code>.col_slot[c][k];
                    if constexpr (is_loc) {
                        ptr[k] = nullptr;
                        in[k] = FIRST ? 0.0f : loc[slot];
                    } else {
                        constexpr int s = kQc<Code>.col_eff[c][k];
                        constexpr int off = slot * Z - s;
                        ptr[k] = (t < s ? hi : lo) + off;
                        in[k] = FIRST ? 0.0f : *ptr[k];
                    }
                });
                var_node<D, IS_SP>(in, D, llr[c], out);
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                    constexpr int slot = kQc<Code>.col_slot[c][k];
                    if constexpr (is_loc) loc[slot] = out[k];
                    else *ptr[k] = out[k];
                });
            }
        });
    };
    auto check_phase = [&]() {
        static_for<MB>([&](auto rr) {
            constexpr int r = decltype(rr)::value;
            constexpr int D = kQc<Code>.row_deg[r];
            if constexpr (D > 0) {
                float in[D], out[D];
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr bool is_loc = kQc<Code>.row_loc[r][j];
                    constexpr int slot = kQc<Code>.row_slot[r][j];
                    if constexpr (is_loc) in[j] = loc[slot];
                    else in[j] = msg[slot * Z + t];
                });
                if constexpr (IS_SP) check_node_sp<D>(in, D, a.clampv, out);
                else check_node_ms_ct<D, UPD>(in, a.clampv, a.param, out);
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr bool is_loc = kQc<Code>.row_loc[r][j];
                    constexpr int slot = kQc<Code>.row_slot[r][j];
                    if constexpr (is_loc) loc[slot] = out[j];
                    else msg[slot * Z + t] = out[j];
                });
            }
        });
    };

    if (a.iters > 0) {
        if (active) var_phase(std::true_type{});
        __syncthreads();
        if (active) check_phase();
        __syncthreads();
    }
#pragma unroll 1
    for (int it = 1; it < a.iters; ++it) {
        if (active) var_phase(std::false_type{});
        __syncthreads();
        if (active) check_phase();
        __syncthreads();
    }

    // ---- marginal, P(bit=1), hard decision --------------------------------------------------------
    if (active) {
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            constexpr int rho = kQc<Code>.rho[c];
            float in[D > 0 ? D : 1];
            static_for<D>([&](auto kk) {
                constexpr int k = decltype(kk)::value;
                constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                constexpr int slot = kQc<Code>.col_slot[c][k];
                if constexpr (is_loc) in[k] = (a.iters == 0) ? 0.0f : loc[slot];
                else {
                    constexpr int s = kQc<Code>.col_eff[c][k];
                    constexpr int off = slot * Z - s;
                    in[k] = (a.iters == 0) ? 0.0f : ((t < s ? hi : lo) + off)[0];
                }
            });
            const float tm = marginal_t<(D > 0 ? D : 1)>(in, D, llr[c]);
            const uint8_t hb = hard_bit(tm);
            int zv = t + rho;
            if (zv >= Z) zv -= Z;
            hard_s[cw * L::HARD_STRIDE + c * Z + zv] = hb | ((llr[c] > 0.0f) ? 2 : 0);
            const long long o = gbase + c * Z + zv;
            if (a.prob) a.prob[o] = prob_one(tm);
            if (a.llr_post) a.llr_post[o] = __fmul_rn(-2.0f, tm);
            if (a.hard) a.hard[o] = hb;
        });
    }
    __syncthreads();

    // ---- syndrome weight ------------------------------------------------------------------------
    if (a.syndrome) {
        if (active) {
            int w = 0;
            const uint8_t *h = hard_s + cw * L::HARD_STRIDE;
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
                    par ^= h[cbase + zv] & 1u;
                });
                w += (int)par;
            });
            if (w) atomicAdd(&scratch[4 + cw], w);
        }
        __syncthreads();
        for (int i = tid; i < ncw; i += T) a.syndrome[cw0 + i] = scratch[4 + i];
        __syncthreads();
        for (int i = tid; i < CW; i += T) scratch[4 + i] = 0;
    }
    if (a.hard_packed) pack_hard(hard_s, L::HARD_STRIDE, ncw, N, a.hard_packed + cw0 * ((N + 7) >> 3));
    if (a.counters) {
        __syncthreads();
        count_errors(hard_s, L::HARD_STRIDE, ncw, N, a.k_info, a.ref_packed + cw0 * ((N + 7) >> 3), a.counters,
                     scratch + 1);
    }
}

// ---- registry of compiled specialisations ----------------------------------------------------------
template <class Code>
static bool proto_matches(int Z, int mb, int nb, const int16_t *proto) {
    if (Z != Code::Z || mb != Code::MB || nb != Code::NB) return false;
    for (int r = 0; r < mb; ++r)
        for (int c = 0; c < nb; ++c)
            if (proto[r * nb + c] != Code::proto[r][c]) return false;
    return true;
}

void qc_plan_info(int qc_id, int out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    if (qc_id == 0) {
        using L = QcLayout<Wifi1944R12, 3>;
        out[0] = L::NLOC; out[1] = L::NSM; out[2] = L::THREADS; out[3] = 3;
    }
}

int qc_lookup(int Z, int mb, int nb, const int16_t *proto) {
    if (proto_matches<Wifi1944R12>(Z, mb, nb, proto)) return 0;
    return -1;
}

template <class Code, int CW, int UPD>
static int launch_qc_one(const DecodeArgs &a, cudaStream_t s) {
    using L = QcLayout<Code, CW>;
    const long long grid = (a.B + CW - 1) / CW;
    if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
    auto k = decode_qc_kernel<Code, CW, UPD>;
    LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
    k<<<(int)grid, L::THREADS, L::SMEM, s>>>(a);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

template <class Code, int CW>
static int launch_qc_t(const DecodeArgs &a, cudaStream_t s) {
    using L = QcLayout<Code, CW>;
    const long long grid = (a.B + CW - 1) / CW;
    if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
    void (*k)(const DecodeArgs) = nullptr;
    switch (a.update) {
        case UPD_SP: k = decode_qc_kernel<Code, CW, UPD_SP>; break;
        case UPD_MINSUM: k = decode_qc_kernel<Code, CW, UPD_MINSUM>; break;
        case UPD_NMS: k = decode_qc_kernel<Code, CW, UPD_NMS>; break;
        default: k = decode_qc_kernel<Code, CW, UPD_OMS>; break;
    }
    LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
    k<<<(int)grid, L::THREADS, L::SMEM, s>>>(a);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int launch_decode_qc(int qc_id, const DecodeArgs &a, cudaStream_t s) {
    if (a.B <= 0) return LDPC_OK;
    switch (qc_id) {
        case 0: {
            static const int cw = [] { const char *e = getenv("LDPC_QC_CW"); return e ? atoi(e) : 3; }();
            if (a.update == UPD_MINSUM) {
                if (cw == 6) return launch_qc_one<Wifi1944R12, 6, UPD_MINSUM>(a, s);
                if (cw == 1) return launch_qc_one<Wifi1944R12, 1, UPD_MINSUM>(a, s);
            }
            return launch_qc_t<Wifi1944R12, 3>(a, s);
        }
        default: set_error("unknown QC specialisation %d", qc_id); return LDPC_EINVAL;
    }
}

}  // namespace ldpc
