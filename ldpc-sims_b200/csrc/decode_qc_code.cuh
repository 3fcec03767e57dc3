// decode_qc_code.cuh - everything one compiled quasi-cyclic code contributes to the library: the instantiations of
// the code-specialised kernels (fp32, f16x2, single-launch simulator) behind one table of entry points.  Each
// prototype of qc_protos.cuh gets its own translation unit (decode_qc_cNN.cu) so the codes compile in parallel;
// decode_qc.cu holds the registry.
#pragma once
#include "decode_qc_kernel.cuh"
#include "decode_qc_h2_kernel.cuh"
#include "decode_qc_pers.cuh"
#include "qc_protos.cuh"

namespace ldpc {

struct QcCodeEntry {
    const char *name;
    bool (*matches)(int Z, int mb, int nb, const int16_t *proto);
    int (*decode)(const DecodeArgs &, cudaStream_t);
    int (*decode_h2)(const DecodeArgs &, cudaStream_t);                        // f16x2: min-sum / normalized min-sum
    int (*sim_fused)(const DecodeArgs &, const LinkParams &, cudaStream_t);    // LDPC_EUNSUPPORTED -> three-launch chain
    int (*decode_tma)(const DecodeArgs &, cudaStream_t);                       // persistent / bulk-copy form (LDPC_KERNEL_QC_TMA); LDPC_EUNSUPPORTED -> decode
    void (*plan_info)(int out[4]);
};

// FULL: every update rule in the single-launch simulator and every OFDM size (the headline code); otherwise the
// simulator is compiled for min-sum over 64-point OFDM only and other requests take the three-launch chain.
template <class Code, int CW, bool FULL>
struct QcCodeImpl {
    using L = QcLayout<Code, CW>;

    static bool matches(int Z, int mb, int nb, const int16_t *proto) {
        if (Z != Code::Z || mb != Code::MB || nb != Code::NB) return false;
        for (int r = 0; r < mb; ++r)
            for (int c = 0; c < nb; ++c)
                if (proto[r * nb + c] != Code::proto[r][c]) return false;
        return true;
    }
    static void plan_info(int out[4]) { out[0] = L::NLOC; out[1] = L::NSM; out[2] = L::THREADS; out[3] = CW; }

    static int decode(const DecodeArgs &a, cudaStream_t s) {
        if (a.B <= 0) return LDPC_OK;
        const long long grid = (a.B + CW - 1) / CW;
        if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
        void (*k)(const DecodeArgs, const LinkParams) = nullptr;
        if (!a.early_exit) {
            switch (a.update) {
                case UPD_SP: k = decode_qc_kernel<Code, CW, UPD_SP, 0, false>; break;
                case UPD_MINSUM: k = decode_qc_kernel<Code, CW, UPD_MINSUM, 0, false>; break;
                case UPD_NMS: k = decode_qc_kernel<Code, CW, UPD_NMS, 0, false>; break;
                default: k = decode_qc_kernel<Code, CW, UPD_OMS, 0, false>; break;
            }
        } else if (a.update == UPD_MINSUM) {
            k = decode_qc_kernel<Code, CW, UPD_MINSUM, 0, true>;
        } else if constexpr (FULL) {                       // syndrome-based early termination with the other rules: headline code only
            switch (a.update) {
                case UPD_SP: k = decode_qc_kernel<Code, CW, UPD_SP, 0, true>; break;
                case UPD_NMS: k = decode_qc_kernel<Code, CW, UPD_NMS, 0, true>; break;
                default: k = decode_qc_kernel<Code, CW, UPD_OMS, 0, true>; break;
            }
        }
        if (!k) return LDPC_EUNSUPPORTED;                   // the dispatcher takes the generic kernel (same bits)
        LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
        k<<<(int)grid, L::THREADS, L::SMEM, s>>>(a, LinkParams());
        LDPC_CUDA_TRY(cudaGetLastError());
        return LDPC_OK;
    }

    template <int UPD>
    static int launch_h2(const DecodeArgs &a, cudaStream_t s) {
        const size_t smem = L::MSG_BYTES + (size_t)2 * CW * L::HARD_STRIDE + sizeof(int) * (8 + 2 * CW);
        const long long grid = (a.B + 2 * CW - 1) / (2 * CW);
        if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
        auto k = decode_qc_h2_kernel<Code, CW, UPD>;
        LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(int)grid, L::THREADS, smem, s>>>(a);
        LDPC_CUDA_TRY(cudaGetLastError());
        return LDPC_OK;
    }
    static int decode_h2(const DecodeArgs &a, cudaStream_t s) {
        if (a.B <= 0) return LDPC_OK;
        if (a.update == UPD_MINSUM) return launch_h2<UPD_MINSUM>(a, s);
        if (a.update == UPD_NMS) return launch_h2<UPD_NMS>(a, s);
        set_error("the f16x2 kernel implements min-sum and normalized min-sum only");
        return LDPC_EUNSUPPORTED;
    }

    template <int UPD, int SIM>
    static int launch_sim(const DecodeArgs &a, const LinkParams &lp, cudaStream_t s) {
        const long long grid = (a.B + CW - 1) / CW;
        if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
        auto k = decode_qc_kernel<Code, CW, UPD, SIM, false>;
        LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
        k<<<(int)grid, L::THREADS, L::SMEM, s>>>(a, lp);
        LDPC_CUDA_TRY(cudaGetLastError());
        return LDPC_OK;
    }
    static int sim_fused(const DecodeArgs &a, const LinkParams &lp, cudaStream_t s) {
        if constexpr (!kQc<Code>.dual_diagonal) return LDPC_EUNSUPPORTED;   // the in-kernel encoder needs the 802.11n-style parity part
        else return sim_fused_dd(a, lp, s);
    }
    static int sim_fused_dd(const DecodeArgs &a, const LinkParams &lp, cudaStream_t s) {
        if (a.B <= 0) return LDPC_OK;
        if (lp.ofdm_size == 64) {
            if (a.update == UPD_MINSUM) return launch_sim<UPD_MINSUM, 64>(a, lp, s);
            if constexpr (FULL) {
                switch (a.update) {
                    case UPD_SP: return launch_sim<UPD_SP, 64>(a, lp, s);
                    case UPD_NMS: return launch_sim<UPD_NMS, 64>(a, lp, s);
                    default: return launch_sim<UPD_OMS, 64>(a, lp, s);
                }
            }
        }
        if constexpr (FULL) {
            if (a.update == UPD_MINSUM) {
                if (lp.ofdm_size == 32) return launch_sim<UPD_MINSUM, 32>(a, lp, s);
                if (lp.ofdm_size == 128) return launch_sim<UPD_MINSUM, 128>(a, lp, s);
                if (lp.ofdm_size == 256) return launch_sim<UPD_MINSUM, 256>(a, lp, s);
            }
        }
        return LDPC_EUNSUPPORTED;
    }

    // persistent, bulk-copy-fed form (decode_qc_pers.cuh): fixed iteration count >= 1, compiled for the headline code
    static int decode_tma(const DecodeArgs &a, cudaStream_t s) {
        if constexpr (FULL) {
            if (a.B <= 0) return LDPC_OK;
            if (a.early_exit || a.iters < 1) return LDPC_EUNSUPPORTED;
            switch (a.update) {
                case UPD_SP: return launch_qc_pers<Code, CW, UPD_SP, float, 6>(a, s);
                case UPD_MINSUM: return launch_qc_pers<Code, CW, UPD_MINSUM, float, 6>(a, s);
                default: return LDPC_EUNSUPPORTED;
            }
        }
        return LDPC_EUNSUPPORTED;
    }

    static QcCodeEntry entry(const char *name) { return QcCodeEntry{name, matches, decode, decode_h2, sim_fused, decode_tma, plan_info}; }
};

// codewords (fp32) / codeword pairs (f16x2) per CTA: as close to 256 threads as Z allows (two CTAs per SM)
template <int Z> struct QcTile { static constexpr int CW = (256 / Z) > 0 ? (256 / Z) : 1; };

// The minimal set for a code specialised at RUN TIME (ldpc_b200/jit.py compiles one translation unit with the system nvcc
// and registers it through ldpc_qc_register_plugin): the four update rules with a fixed iteration count.  Everything else
// (early termination, f16x2, the single-launch simulator) falls through to the kernels that take the prototype at run time.
template <class Code, int CW>
struct QcCodeImplLite {
    using F = QcCodeImpl<Code, CW, false>;
    static int decode(const DecodeArgs &a, cudaStream_t s) {
        if (a.early_exit) return LDPC_EUNSUPPORTED;
        return F::decode(a, s);
    }
    static int no_h2(const DecodeArgs &, cudaStream_t) {
        set_error("the f16x2 kernel is not part of a run-time specialised code: use LDPC_PREC_F32");
        return LDPC_EUNSUPPORTED;
    }
    static int no_tma(const DecodeArgs &, cudaStream_t) { return LDPC_EUNSUPPORTED; }
    static int no_sim(const DecodeArgs &, const LinkParams &, cudaStream_t) { return LDPC_EUNSUPPORTED; }
    static QcCodeEntry entry(const char *name) { return QcCodeEntry{name, F::matches, decode, no_h2, no_sim, no_tma, F::plan_info}; }
};

// what a plug-in and the library must agree on before any entry point is called through a QcCodeEntry
inline unsigned qc_plugin_abi_tag() {
    return (unsigned)(sizeof(DecodeArgs) * 1000003u + sizeof(LinkParams) * 10007u + sizeof(QcCodeEntry) * 101u + 2u /* revision */);
}

}  // namespace ldpc
