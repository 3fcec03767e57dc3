// qc_var_pipe.cuh - node-arithmetic adaptors (one spelling for fp32 and f16x2) and the software-pipelined variable
// phase shared by the code-specialised kernels (decode_qc_kernel.cuh, decode_qc_h2_kernel.cuh, decode_qc_pers.cuh).
#pragma once
#include <type_traits>

#include "node_math.cuh"
#include "node_math_h2.cuh"
#include "qc_plan.cuh"

namespace ldpc {

// ---- node arithmetic adaptors: one spelling for both number formats ----------------------------------------
struct NodeParams {
    float clampv, param;
    __half2 clamp_h, alpha_h;
};

template <int D, int UPD>
__device__ __forceinline__ void vnode(const float (&in)[D], float llr, float (&out)[D]) {
    var_node_m<D, (UPD == UPD_SP ? 1 : (UPD == UPD_SPF ? 2 : 0))>(in, D, llr, out);
}
template <int D, int UPD>
__device__ __forceinline__ void vnode(const __half2 (&in)[D], __half2 llr, __half2 (&out)[D]) {
    __half2 s[D];
    h2_sum_others<D>(in, s);
    const __half2 Lp = __hneg2(llr);
#pragma unroll
    for (int k = 0; k < D; ++k) out[k] = h2_add(Lp, s[k]);
}
template <int D, int UPD>
__device__ __forceinline__ void cnode(const float (&in)[D], const NodeParams &p, float (&out)[D]) {
    if constexpr (UPD == UPD_SP || UPD == UPD_SPF) check_node_sp<D, UPD == UPD_SPF>(in, D, p.clampv, out);
    else check_node_ms_ct<D, UPD>(in, p.clampv, p.param, out);
}
template <int D, int UPD>
__device__ __forceinline__ void cnode(const __half2 (&in)[D], const NodeParams &p, __half2 (&out)[D]) {
    if constexpr (UPD == UPD_NMS) {
        __half2 v[D];
#pragma unroll
        for (int j = 0; j < D; ++j) v[j] = __hmul2_rn(p.alpha_h, in[j]);
        h2_boxmin_others_clamped<D>(v, p.clamp_h, out);
    } else {
        h2_boxmin_others_clamped<D>(in, p.clamp_h, out);
    }
}
template <int D>
__device__ __forceinline__ float mnode(const float (&in)[D], int d, float llr) { return marginal_t<D>(in, d, llr); }
template <int D>
__device__ __forceinline__ __half2 mnode(const __half2 (&in)[D], int d, __half2 llr) {
    __half2 acc = __float2half2_rn(0.0f);
#pragma unroll
    for (int k = 0; k < D; ++k)
        if (k < d) acc = (k == 0) ? in[0] : h2_add(acc, in[k]);
    return __hmul2_rn(__float2half2_rn(0.5f), h2_add(__hneg2(llr), acc));
}
__device__ __forceinline__ float zero_of(float) { return 0.0f; }
__device__ __forceinline__ __half2 zero_of(__half2) { return __float2half2_rn(0.0f); }

// ---- software-pipelined variable phase (see VarBatches in qc_plan.cuh) ---------------------------------------
template <class Code, int CWT, int UPD, class T, int VBM>
struct VarPipe {
    static constexpr int Z = Code::Z, NB = Code::NB;
    static constexpr int VBW = kVarBatches<Code, VBM>.width, NBATCH = kVarBatches<Code, VBM>.n;
    static constexpr int NLOCA = kQc<Code>.n_local > 0 ? kQc<Code>.n_local : 1;

    template <int B>
    static __device__ __forceinline__ void load(int t, T *lo, T *hi, T (&in)[VBW], T *(&ptr)[VBW]) {
        constexpr int c0 = kVarBatches<Code, VBM>.first[B], c1 = kVarBatches<Code, VBM>.first[B + 1];
        static_for<c1 - c0>([&](auto ci) {
            constexpr int c = c0 + decltype(ci)::value;
            static_for<kQc<Code>.col_deg[c]>([&](auto kk) {
                constexpr int k = decltype(kk)::value;
                if constexpr (!kQc<Code>.col_loc[c][k]) {
                    constexpr int s = kQc<Code>.col_eff[c][k];
                    constexpr int off = (kQc<Code>.col_slot[c][k] * Z - s) * CWT;
                    constexpr int i = kVarBatches<Code, VBM>.idx[c][k];
                    ptr[i] = (t < s ? hi : lo) + off;
                    in[i] = *ptr[i];
                }
            });
        });
    }
    template <int B>
    static __device__ __forceinline__ void finish(const T (&llr)[NB], T (&loc)[NLOCA], T (&in)[VBW], T *(&ptr)[VBW]) {
        constexpr int c0 = kVarBatches<Code, VBM>.first[B], c1 = kVarBatches<Code, VBM>.first[B + 1];
        static_for<c1 - c0>([&](auto ci) {
            constexpr int c = c0 + decltype(ci)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            if constexpr (D > 0) {
                T x[D], y[D];
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr int slot = kQc<Code>.col_slot[c][k], i = kVarBatches<Code, VBM>.idx[c][k];
                    if constexpr (kQc<Code>.col_loc[c][k]) x[k] = loc[slot];
                    else x[k] = in[i];
                });
                vnode<D, UPD>(x, llr[c], y);
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr int slot = kQc<Code>.col_slot[c][k], i = kVarBatches<Code, VBM>.idx[c][k];
                    if constexpr (kQc<Code>.col_loc[c][k]) loc[slot] = y[k];
                    else *ptr[i] = y[k];
                });
            }
        });
    }
    // batch B is in (cur, pcur): prefetch B+1 into (nxt, pnxt), finish B, recurse with the buffers swapped
    template <int B>
    static __device__ __forceinline__ void run(int t, T *lo, T *hi, const T (&llr)[NB], T (&loc)[NLOCA], T (&cur)[VBW], T *(&pcur)[VBW],
                                               T (&nxt)[VBW], T *(&pnxt)[VBW]) {
        if constexpr (B < NBATCH) {
            if constexpr (B + 1 < NBATCH) load<B + 1>(t, lo, hi, nxt, pnxt);
            finish<B>(llr, loc, cur, pcur);
            run<B + 1>(t, lo, hi, llr, loc, nxt, pnxt, cur, pcur);
        }
    }
};

}  // namespace ldpc
