// decode_qc_h2_kernel.cuh - f16x2 variant of the code-specialised min-sum decoder: every thread
// carries TWO codewords packed in one 32-bit register (half2), so each shared-memory word,
// each HADD2 and each HMNMX2.XORSIGN serves two codewords.  Same plan (qc_plan.cuh), same
// schedule and phases as decode_qc_kernel.cuh; only the number format differs (node_math_h2.cuh).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "epilogue.cuh"
#include "node_math.cuh"
#include "node_math_h2.cuh"
#include "qc_plan.cuh"
#include "qc_var_pipe.cuh"

namespace ldpc {

template <class Code, int CW /* codeword PAIRS per CTA */, int UPD>
__global__ void __launch_bounds__((QcLayout<Code, CW>::THREADS)) decode_qc_h2_kernel(const DecodeArgs a) {
    using L = QcLayout<Code, CW>;
    constexpr int Z = Code::Z, NB = Code::NB, MB = Code::MB, N = L::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half2 *msg_s = reinterpret_cast<__half2 *>(smem_raw);
    uint8_t *hard_s = smem_raw + L::MSG_BYTES;                                           // [2*CW][HARD_STRIDE]
    int *scratch = reinterpret_cast<int *>(hard_s + 2 * CW * L::HARD_STRIDE);              // [4 + 2*CW]

    const int tid = threadIdx.x, T = blockDim.x;
    const long long cw0 = (long long)blockIdx.x * (2 * CW);
    const int ncw = (int)min((long long)(2 * CW), a.B - cw0);            // codewords in this tile
    const int t = tid / CW, pr = tid - t * CW;                            // lane, pair slot (pairs interleaved by lane, see decode_qc.cu)
    const bool active = 2 * pr < ncw && t < Z;                            // false for padding threads too
    const bool second = 2 * pr + 1 < ncw;                                 // .y lane holds a real codeword
    for (int i = tid; i < 4 + 2 * CW; i += T) scratch[i] = 0;

    __half2 *const msg = msg_s + (active ? tid : 0);
    __half2 *const lo = msg;
    __half2 *const hi = msg + Z * CW;

    __half2 llr[NB];
    __half2 loc[L::NLOC > 0 ? L::NLOC : 1];
    const long long gbase = (cw0 + (active ? 2 * pr : 0)) * N;
    if (active) {
        // issue all 2*NB independent global loads first (they overlap in flight), convert afterwards
        float raw0[NB], raw1[NB];
        const long long g1 = second ? gbase + N : gbase;                 // half-empty last pair: re-read row 0, discard
        // one uniform dtype branch, then 2*NB straight loads from two per-thread lane bases with compile-time offsets
        // (the wrapped position of a column is Z elements lower: a 32-bit select, see decode_qc_kernel.cuh)
        auto load_all = [&](auto *base, auto conv) {
            const auto *b0 = base + gbase + t, *b1 = base + g1 + t;
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int rho = kQc<Code>.rho[c];
                constexpr int o0 = c * Z + rho;
                const int off = (rho != 0 && t >= Z - rho) ? (o0 - Z) : o0;
                raw0[c] = conv(__ldg(b0 + off));
                raw1[c] = conv(__ldg(b1 + off));
            });
        };
        if (a.llr_dtype == LDPC_F32) load_all(reinterpret_cast<const float *>(a.llr), [](float v) { return v; });
        else if (a.llr_dtype == LDPC_F64) load_all(reinterpret_cast<const double *>(a.llr), [](double v) { return (float)v; });
        else if (a.llr_dtype == LDPC_I8) load_all(reinterpret_cast<const signed char *>(a.llr), [](signed char v) { return (float)v; });
        else load_all(reinterpret_cast<const __half *>(a.llr), [](__half v) { return __half2float(v); });
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            llr[c] = __floats2half2_rn(sat_llr(raw0[c]), second ? sat_llr(raw1[c]) : 0.0f);
        });
    }
    const __half2 clamp_h = __float2half2_rn(a.clampv);
    const __half2 alpha_h = __float2half2_rn(a.param);

    auto var_phase = [&](auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        if constexpr (FIRST) {
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int D = kQc<Code>.col_deg[c];
                if constexpr (D > 0) {
                    const __half2 y = h2_add(__hneg2(llr[c]), __float2half2_rn(0.0f));       // every C->V message is still zero
                    static_for<D>([&](auto kk) {
                        constexpr int k = decltype(kk)::value;
                        constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                        constexpr int slot = kQc<Code>.col_slot[c][k];
                        if constexpr (is_loc) loc[slot] = y;
                        else {
                            constexpr int sh = kQc<Code>.col_eff[c][k];
                            constexpr int off = (slot * Z - sh) * CW;
                            ((t < sh ? hi : lo) + off)[0] = y;
                        }
                    });
                }
            });
        } else {
            // in-order (the software-pipelined form of qc_var_pipe.cuh measured 2.8 % SLOWER here: two codewords per thread
            // already double the independent work between a load and its use, and the extra live registers cost more)
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int D = kQc<Code>.col_deg[c];
                if constexpr (D > 0) {
                    __half2 in[D], out[D];
                    __half2 *ptr[D];
                    static_for<D>([&](auto kk) {
                        constexpr int k = decltype(kk)::value;
                        constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                        constexpr int slot = kQc<Code>.col_slot[c][k];
                        if constexpr (is_loc) {
                            ptr[k] = nullptr;
                            in[k] = loc[slot];
                        } else {
                            constexpr int sh = kQc<Code>.col_eff[c][k];
                            constexpr int off = (slot * Z - sh) * CW;
                            ptr[k] = (t < sh ? hi : lo) + off;
                            in[k] = *ptr[k];
                        }
                    });
                    vnode<D, UPD>(in, llr[c], out);
                    static_for<D>([&](auto kk) {
                        constexpr int k = decltype(kk)::value;
                        constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                        constexpr int slot = kQc<Code>.col_slot[c][k];
                        if constexpr (is_loc) loc[slot] = out[k];
                        else *ptr[k] = out[k];
                    });
                }
            });
        }
    };
    auto check_phase = [&]() {
        static_for<MB>([&](auto rr) {
            constexpr int r = decltype(rr)::value;
            constexpr int D = kQc<Code>.row_deg[r];
            if constexpr (D > 0) {
                __half2 in[D], out[D];
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr bool is_loc = kQc<Code>.row_loc[r][j];
                    constexpr int slot = kQc<Code>.row_slot[r][j];
                    __half2 v;
                    if constexpr (is_loc) v = loc[slot];
                    else v = msg[slot * Z * CW];
                    in[j] = (UPD == UPD_NMS) ? __hmul2_rn(alpha_h, v) : v;
                });
                h2_boxmin_others_clamped<D>(in, clamp_h, out);
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr bool is_loc = kQc<Code>.row_loc[r][j];
                    constexpr int slot = kQc<Code>.row_slot[r][j];
                    if constexpr (is_loc) loc[slot] = out[j];
                    else msg[slot * Z * CW] = out[j];
                });
            }
        });
    };

    if (a.iters <= 0) {                                                   // no iteration: all messages are zero
        for (int i = tid; i < CW * L::MSG_STRIDE; i += T) msg_s[i] = __float2half2_rn(0.0f);
#pragma unroll
        for (int i = 0; i < (L::NLOC > 0 ? L::NLOC : 1); ++i) loc[i] = __float2half2_rn(0.0f);
        __syncthreads();
    }
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

    // ---- marginal, P(bit=1), hard decision (output pointers tested once, see decode_qc.cu) ---------------
    if (active) {
        const __half2 half_h = __float2half2_rn(0.5f);
        __half2 th[NB];
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            __half2 acc = __float2half2_rn(0.0f);
            static_for<D>([&](auto kk) {
                constexpr int k = decltype(kk)::value;
                constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                constexpr int slot = kQc<Code>.col_slot[c][k];
                __half2 v;
                if constexpr (is_loc) v = loc[slot];
                else {
                    constexpr int sh = kQc<Code>.col_eff[c][k];
                    constexpr int off = (slot * Z - sh) * CW;
                    v = ((t < sh ? hi : lo) + off)[0];
                }
                acc = (k == 0) ? v : h2_add(acc, v);
            });
            th[c] = __hmul2_rn(half_h, h2_add(__hneg2(llr[c]), acc));
        });
        unsigned hb0 = 0, hb1 = 0;
        float tmin = CUDART_INF_F;
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            const float2 tf = __half22float2(th[c]);
            hb0 |= (tf.x < 0.0f ? 1u : 0u) << c;
            hb1 |= (tf.y < 0.0f ? 1u : 0u) << c;
            tmin = fminf(tmin, fminf(fabsf(tf.x), fabsf(tf.y)));
        });
        if (!(tmin > 1e-5f)) {                                             // tie band (rare)
            hb0 = hb1 = 0;
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                const float2 tf = __half22float2(th[c]);
                hb0 |= (unsigned)hard_bit(tf.x) << c;
                hb1 |= (unsigned)hard_bit(tf.y) << c;
            });
        }
        uint8_t *const hrow = hard_s + (2 * pr) * L::HARD_STRIDE + t;     // lane base + compile-time offsets (wrapped: Z lower)
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int rho = kQc<Code>.rho[c];
            constexpr int o0 = c * Z + rho;
            const int off = (rho != 0 && t >= Z - rho) ? (o0 - Z) : o0;
            const float2 lf = __half22float2(llr[c]);
            hrow[off] = (uint8_t)(((hb0 >> c) & 1u) | ((lf.x > 0.0f) ? 2u : 0u));
            if (second) hrow[L::HARD_STRIDE + off] = (uint8_t)(((hb1 >> c) & 1u) | ((lf.y > 0.0f) ? 2u : 0u));
        });
        if (a.llr_post) {
            float *const post = a.llr_post + gbase + t;
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int rho = kQc<Code>.rho[c];
                constexpr int o0 = c * Z + rho;
                const float2 tf = __half22float2(th[c]);
                const float v0 = __fmul_rn(-2.0f, tf.x), v1 = __fmul_rn(-2.0f, tf.y);
                if (rho != 0 && t >= Z - rho) {
                    post[o0 - Z] = v0;
                    if (second) post[N + o0 - Z] = v1;
                } else {
                    post[o0] = v0;
                    if (second) post[N + o0] = v1;
                }
            });
        }
        if (a.prob || a.hard) {                                            // byte / probability outputs: cold path
#pragma unroll 1
            for (int c = 0; c < NB; ++c) {
                float2 tf = make_float2(0.0f, 0.0f);
                int rho = 0;
                static_for<NB>([&](auto cc) {
                    constexpr int c2 = decltype(cc)::value;
                    constexpr int rho2 = kQc<Code>.rho[c2];
                    if (c == c2) { tf = __half22float2(th[c2]); rho = rho2; }
                });
                int zv = t + rho;
                if (zv >= Z) zv -= Z;
                const long long o = gbase + c * Z + zv;
                if (a.prob) { a.prob[o] = prob_one(tf.x); if (second) a.prob[o + N] = prob_one(tf.y); }
                if (a.hard) { a.hard[o] = (uint8_t)((hb0 >> c) & 1u); if (second) a.hard[o + N] = (uint8_t)((hb1 >> c) & 1u); }
            }
        }
    }
    __syncthreads();

    // ---- syndrome weight --------------------------------------------------------------------------------
    if (a.syndrome) {
        if (active) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1 && !second) break;
                int w = 0;
                const uint8_t *hs = hard_s + (2 * pr + h) * L::HARD_STRIDE;
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
                if (w) atomicAdd(&scratch[4 + 2 * pr + h], w);
            }
        }
        __syncthreads();
        for (int i = tid; i < ncw; i += T) a.syndrome[cw0 + i] = scratch[4 + i];
        __syncthreads();
        for (int i = tid; i < 2 * CW; i += T) scratch[4 + i] = 0;
    }
    if (a.hard_packed) pack_hard(hard_s, L::HARD_STRIDE, ncw, N, a.hard_packed + cw0 * ((N + 7) >> 3));
    if (a.counters) {
        __syncthreads();
        count_errors(hard_s, L::HARD_STRIDE, ncw, N, a.k_info, a.ref_packed + cw0 * ((N + 7) >> 3), a.counters,
                     scratch + 1);
    }
}

}  // namespace ldpc
