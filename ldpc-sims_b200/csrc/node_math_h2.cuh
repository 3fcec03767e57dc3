// node_math_h2.cuh - binary16x2 node arithmetic of the f16x2 decoder kernels (two codewords per 32-bit register).
// Arithmetic ("min-sum f16", defined bit-exactly by oracle/bp_oracle.py::bp_decode_f16 - the
// reference has neither min-sum nor half precision, bp/bp.py:27-31):
//   Lh   = fp16_rn(clip(llr, +-32768)),  L' = -Lh
//   V->C y_k = fp16(L' + S_k), S_k = P_k + Q_k two-sweep, every add rounded to fp16 (RN)
//   C->V x_j = c (+) (+)_{i!=j} g(y_i), a (+) b = sign(a)sign(b) min(|a|,|b|), c = fp16_rn(clamp),
//        g = identity (min-sum) | fp16(alpha_h * y) (normalized min-sum)
//   t    = fp16(0.5 * fp16(L' + ((x_0 + x_1) + ...)));  outputs from float(t) exactly as fp32 path.
#pragma once
#include <cuda_fp16.h>

namespace ldpc {

__device__ __forceinline__ __half2 h2_add(__half2 a, __half2 b) { return __hadd2_rn(a, b); }

__device__ __forceinline__ __half2 h2_boxmin(__half2 a, __half2 b) {
    unsigned d;
    asm("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(*reinterpret_cast<unsigned *>(&a)), "r"(*reinterpret_cast<unsigned *>(&b)));
    return *reinterpret_cast<__half2 *>(&d);
}

template <int D>
__device__ __forceinline__ void h2_sum_others(const __half2 (&in)[D], __half2 (&out)[D]) {
    if constexpr (D == 1) { out[0] = __float2half2_rn(0.0f); return; }
    __half2 pre[D];
    __half2 acc = in[0];
#pragma unroll
    for (int k = 1; k < D; ++k) { pre[k] = acc; acc = h2_add(acc, in[k]); }
    acc = in[D - 1];
    out[D - 1] = pre[D - 1];
#pragma unroll
    for (int k = D - 2; k >= 1; --k) { out[k] = h2_add(pre[k], acc); acc = h2_add(in[k], acc); }
    out[0] = acc;
}

// prefix / suffix form, 3D - 4 operations (see boxmin_others_clamped in node_math.cuh)
template <int D>
__device__ __forceinline__ void h2_boxmin_others_clamped(const __half2 (&v)[D], __half2 c, __half2 (&out)[D]) {
    if constexpr (D == 1) {
        out[0] = c;
    } else {
        __half2 pre[D];
        pre[0] = c;
#pragma unroll
        for (int j = 1; j < D; ++j) pre[j] = h2_boxmin(pre[j - 1], v[j - 1]);
        out[D - 1] = pre[D - 1];
        __half2 suf = v[D - 1];
#pragma unroll
        for (int j = D - 2; j >= 0; --j) {
            out[j] = h2_boxmin(pre[j], suf);
            if (j > 0) suf = h2_boxmin(v[j], suf);
        }
    }
}

__device__ __forceinline__ float sat_llr(float v) { return fminf(fmaxf(v, -32768.0f), 32768.0f); }

}  // namespace ldpc
