// decode_qc_kernel.cuh - code-specialised belief-propagation decoder for quasi-cyclic codes whose
// prototype matrix is known at compile time (the IEEE 802.11n family, qc_protos.cuh).
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
#pragma once
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "epilogue.cuh"
#include "node_math.cuh"
#include "linksim_device.cuh"
#include "qc_plan.cuh"
#include "qc_var_pipe.cuh"

namespace ldpc {

// SIM = 0: channel LLRs come from global memory (ldpc_decode).
// SIM = N_ofdm (32/64/128/256): the link front end runs INSIDE this kernel - Philox information
// bits, linear-time dual-diagonal encoder, QPSK, per-codeword OFDM framing, warp IFFT, Philox
// AWGN, AGC + quantizer, warp FFT, exact LLR (linksim_device.cuh) - and writes the LLR tile into
// the (still unused) message region of shared memory; the error counters compare against the
// transmitted bits kept in shared memory.  One launch takes random bits to BER counts
// (replaces the per-SNR loop body of evaluate_quantized_snr.py:91-188).
template <class Code, int CW, int UPD, int SIM, bool EE>
__global__ void __launch_bounds__((QcLayout<Code, CW>::THREADS), (QcLayout<Code, CW>::MIN_CTAS)) decode_qc_kernel(const DecodeArgs a, const LinkParams lp) {
    using L = QcLayout<Code, CW>;
    constexpr bool IS_SP = (UPD == UPD_SP || UPD == UPD_SPF);
    constexpr int Z = Code::Z, NB = Code::NB, MB = Code::MB, N = L::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *msg_s = reinterpret_cast<float *>(smem_raw);
    uint8_t *hard_s = smem_raw + L::MSG_BYTES;
    int *scratch = reinterpret_cast<int *>(hard_s + CW * L::HARD_STRIDE);       // [4 + CW]

    const int tid = threadIdx.x, T = blockDim.x;
    const long long cw0 = (long long)blockIdx.x * CW;
    const int ncw = (int)min((long long)CW, a.B - cw0);
    // codewords interleaved by lane: thread = t * CW + cw, message slot (blk, z) of codeword cw at
    // (blk * Z + z) * CW + cw.  A block's CW rings form ONE ring of CW * Z words, so the linear access
    // of the check phase is conflict-free and the rotated window of the variable phase wraps once per
    // block and CTA (one two-wavefront warp) instead of once per block and codeword.
    const int t = tid / CW, cw = tid - t * CW;
    const bool active = cw < ncw && t < Z;   // false for the padding threads (t >= Z) too
    for (int i = tid; i < 4 + CW; i += T) scratch[i] = 0;

    float *const msg = msg_s + (active ? tid : 0);                         // slot (blk, t): msg[blk * Z * CW]
    // rotated window bases: slot (blk, (t - s') mod Z) = (t < s' ? hi : lo)[(blk*Z - s') * CW]
    float *const lo = msg;
    float *const hi = msg + Z * CW;

    // ---- channel LLRs of this thread's NB variables live in registers for the whole decode ----------
    float llr[NB];
    float loc[L::NLOC > 0 ? L::NLOC : 1];
    const long long gbase = (cw0 + (active ? cw : 0)) * N;
    if constexpr (SIM == 0) {
    if (active) {
        // One uniform dtype branch, then NB straight loads.  Variable (c, (t + rho_c) mod Z) sits at lane-base + c Z + rho_c,
        // or Z elements lower once t + rho_c wraps: two COMPILE-TIME offsets from one per-thread base pointer, selected by
        // a predicate.
        auto load_all = [&](auto *base, auto conv) {
            const auto *bp = base + gbase + t;
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int rho = kQc<Code>.rho[c];
                constexpr int o0 = c * Z + rho;
                const int off = (rho != 0 && t >= Z - rho) ? (o0 - Z) : o0;            // a 32-bit select, then one widening add
                llr[c] = conv(__ldg(bp + off));
            });
        };
        if (a.llr_dtype == LDPC_F32) load_all(reinterpret_cast<const float *>(a.llr), [](float v) { return v; });
        else if (a.llr_dtype == LDPC_F64) load_all(reinterpret_cast<const double *>(a.llr), [](double v) { return (float)v; });
        else if (a.llr_dtype == LDPC_I8) load_all(reinterpret_cast<const signed char *>(a.llr), [](signed char v) { return (float)v; });
        else load_all(reinterpret_cast<const __half *>(a.llr), [](__half v) { return __half2float(v); });
    }
    } else {
        static_assert(SIM == 0 || kQc<Code>.dual_diagonal, "the fused simulator needs the dual-diagonal encoder structure");
        constexpr int K = (NB - MB) * Z, KW = (K + 31) / 32, KB = NB - MB;
        constexpr int KWS = ((KW + 3) / 4) * 4;                            // packed information words per codeword (whole Philox blocks)
        static_assert(SIM == 0 || CW * L::MSG_STRIDE >= CW * N + CW * KWS + SIM, "LLR staging does not fit the message region");
        float *stage = msg_s;                                              // [CW][N] channel LLRs
        uint32_t *u_s = reinterpret_cast<uint32_t *>(msg_s + CW * N);       // [CW][KWS] packed information words
        cplx<float> *tw = reinterpret_cast<cplx<float> *>(u_s + CW * KWS);  // [SIM/2] twiddles
        const Philox rng(lp.seed);
        // -- information bits (same Philox blocks as sim.cu gen_codewords_kernel)
        for (int i = tid; i < ncw * ((KW + 3) / 4); i += T) {
            const int c = i / ((KW + 3) / 4), blk = i - c * ((KW + 3) / 4);
            info_words_block(rng, (unsigned long long)(lp.cw_first + cw0 + c), blk, K, u_s + c * KWS);
        }
        fill_twiddles_f<float>(tw, SIM);
        __syncthreads();
        // -- systematic encode, natural lane z = t: lambda_r = sum_c rot(u_c, s_rc); p0 = sum_r lambda_r;
        //    p_1 = lambda_0 + rot(p0, h_0); p_{r+1} = lambda_r + [h_r] rot(p0, h_r) + p_r
        uint8_t *bits = hard_s + (active ? cw : 0) * L::HARD_STRIDE;       // bit 2 of the byte = transmitted bit
        const uint32_t *u = u_s + (active ? cw : 0) * KWS;
        auto ubit = [&](int i) { return (u[i >> 5] >> (i & 31)) & 1u; };
        // the information bits as bytes first (bit 2 = transmitted bit): the encoder below then reads one byte per term instead
        // of extracting a bit from the packed Philox words (word index, shift, mask)
        if (active) {
            static_for<KB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                bits[c * Z + t] = (uint8_t)(ubit(c * Z + t) << 2);
            });
        }
        __syncthreads();
        unsigned lam = 0;                                                  // bit r = lambda_r[t]
        if (active) {
            static_for<MB>([&](auto rr) {
                constexpr int r = decltype(rr)::value;
                unsigned acc = 0;
                static_for<kQc<Code>.enc_deg[r]>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr int c = kQc<Code>.enc_col[r][j], sh = kQc<Code>.enc_shift[r][j];
                    int z = t + sh;
                    if (z >= Z) z -= Z;
                    acc ^= bits[c * Z + z];                                // bit 2 carries the XOR
                });
                lam |= ((acc >> 2) & 1u) << r;
            });
            bits[KB * Z + t] = (uint8_t)((__popc(lam) & 1) << 2);          // p0
        }
        __syncthreads();
        if (active) {
            unsigned prev = 0;
            static_for<MB - 1>([&](auto rr) {
                constexpr int r = decltype(rr)::value;
                unsigned v = (lam >> r) & 1u;
                if constexpr (kQc<Code>.hcol[r] >= 0) {
                    int z = t + kQc<Code>.hcol[r];
                    if (z >= Z) z -= Z;
                    v ^= (bits[KB * Z + z] >> 2) & 1u;
                }
                if constexpr (r >= 1) v ^= prev;
                prev = v;
                bits[(KB + r + 1) * Z + t] = (uint8_t)(v << 2);
            });
        }
        __syncthreads();
        // -- OFDM link, one warp per OFDM symbol
        {
            const LinkConsts kc(lp, SIM);
            const int lane = tid & 31, nsym = N / 2;
            constexpr int S = 1;                                         // OFDM symbols in flight per warp (2 measured no faster)
            constexpr int OFDM_PER_CW = (N / 2 + (SIM > 0 ? SIM : 32) - 1) / (SIM > 0 ? SIM : 32);   // = lp.n_ofdm_per_cw (sim.cu), as a constant divisor
            const int total = ncw * OFDM_PER_CW;
            for (int o0 = (tid >> 5) * S; o0 < total; o0 += (T >> 5) * S) {
                int osv[S];
                unsigned long long gcw[S];
                bool valid[S];
                const uint8_t *brow[S];
                float *orow[S];
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    const int o = o0 + q;
                    valid[q] = o < total;
                    const int c = valid[q] ? o / OFDM_PER_CW : 0;
                    osv[q] = valid[q] ? o - c * OFDM_PER_CW : 0;
                    gcw[q] = (unsigned long long)(lp.cw_first + cw0 + c);
                    brow[q] = hard_s + c * L::HARD_STRIDE;
                    orow[q] = stage + c * N;
                }
                ofdm_symbols_llr<(SIM > 0 ? SIM : 32), S>(lane, osv, gcw, valid, nsym, lp, kc, tw,
                                                          [&](int q, int i) { return (int)((brow[q][i] >> 2) & 1); },
                                                          [&](int q, int sidx, float l0, float l1) { *reinterpret_cast<float2 *>(orow[q] + 2 * sidx) = make_float2(l0, l1); });
            }
        }
        __syncthreads();
        if (active) {
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int rho = kQc<Code>.rho[c];
                int zv = t + rho;
                if (zv >= Z) zv -= Z;
                llr[c] = stage[cw * N + c * Z + zv];
            });
        }
        __syncthreads();                                                   // the message region is free again
    }

    // V -> C for all NB block columns of this thread.  FIRST: the C->V messages are still the zeros every reference
    // caller passes (ofdm_functions.py:157) - nothing is loaded.  Otherwise the shared-memory loads of the next batch of
    // block columns are issued before the current batch is computed and stored (VarPipe, qc_var_pipe.cuh): the rotated-
    // window pointers are run-time selections, so the compiler cannot move a load above a store of another block itself.
    auto var_phase = [&](auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        if constexpr (FIRST) {
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int D = kQc<Code>.col_deg[c];
                if constexpr (D > 0) {
                    float in[D], out[D];
                    static_for<D>([&](auto kk) { in[decltype(kk)::value] = 0.0f; });
                    vnode<D, UPD>(in, llr[c], out);
                    static_for<D>([&](auto kk) {
                        constexpr int k = decltype(kk)::value;
                        constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                        constexpr int slot = kQc<Code>.col_slot[c][k];
                        if constexpr (is_loc) loc[slot] = out[k];
                        else {
                            constexpr int s = kQc<Code>.col_eff[c][k];
                            constexpr int off = (slot * Z - s) * CW;
                            ((t < s ? hi : lo) + off)[0] = out[k];
                        }
                    });
                }
            });
        } else {
            using VP = VarPipe<Code, CW, UPD, float, 6>;
            float inA[VP::VBW], inB[VP::VBW];
            float *pA[VP::VBW], *pB[VP::VBW];
            VP::template load<0>(t, lo, hi, inA, pA);
            VP::template run<0>(t, lo, hi, llr, loc, inA, pA, inB, pB);
        }
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
                    else in[j] = msg[slot * Z * CW];
                });
                if constexpr (IS_SP) check_node_sp<D, UPD == UPD_SPF>(in, D, a.clampv, out);
                else check_node_ms_ct<D, UPD>(in, a.clampv, a.param, out);
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

    int *frozen_s = scratch + 4 + CW;                                   // [CW] iteration at which a codeword converged, 0 = running
    // forward declarations of the two tail phases (also used by the early-termination test)
    // Marginal, hard decision, outputs.  The output pointers are tested ONCE (uniform branches around
    // compact store loops); an iteration count of 0 is handled by zero-filling the messages up front.
    auto marginal_phase = [&](const bool final_pass) {
        float tm[NB];
        float tmin = CUDART_INF_F;
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            float in[D > 0 ? D : 1];
            static_for<D>([&](auto kk) {
                constexpr int k = decltype(kk)::value;
                constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                constexpr int slot = kQc<Code>.col_slot[c][k];
                if constexpr (is_loc) in[k] = loc[slot];
                else {
                    constexpr int s = kQc<Code>.col_eff[c][k];
                    constexpr int off = (slot * Z - s) * CW;
                    in[k] = ((t < s ? hi : lo) + off)[0];
                }
            });
            tm[c] = marginal_t<(D > 0 ? D : 1)>(in, D, llr[c]);
            tmin = fminf(tmin, fabsf(tm[c]));
        });
        // hard decision = (t < 0) outside the tie band; one rarely-taken branch per thread
        // re-evaluates the band cases the way the reference rounds them (node_math.cuh: hard_bit)
        unsigned hbits = 0;
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            hbits |= (tm[c] < 0.0f ? 1u : 0u) << c;
        });
        if (!(tmin > 1e-5f)) {
            hbits = 0;
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                hbits |= (unsigned)hard_bit(tm[c]) << c;
            });
        }
        uint8_t *const hrow = hard_s + cw * L::HARD_STRIDE + t;           // lane base: two compile-time offsets per column (see the LLR loads)
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int rho = kQc<Code>.rho[c];
            constexpr int o0 = c * Z + rho;
            const unsigned v = ((hbits >> c) & 1u) | ((llr[c] > 0.0f) ? 2u : 0u);
            uint8_t *hp = (rho != 0 && t >= Z - rho) ? hrow + (o0 - Z) : hrow + o0;
            *hp = (uint8_t)(SIM ? ((*hp & 4u) | v) : v);
        });
        if (final_pass) {
            if (a.llr_post) {
                float *const post = a.llr_post + gbase + t;
                static_for<NB>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    constexpr int rho = kQc<Code>.rho[c];
                    constexpr int o0 = c * Z + rho;
                    const float v = __fmul_rn(-2.0f, tm[c]);
                    if constexpr (rho == 0) post[o0] = v;
                    else if (t >= Z - rho) post[o0 - Z] = v;
                    else post[o0] = v;
                });
            }
            if (a.prob || a.hard) {                                          // byte / probability outputs: cold path
#pragma unroll 1
                for (int c = 0; c < NB; ++c) {
                    float tc = 0.0f;
                    int rho = 0;
                    static_for<NB>([&](auto cc) {
                        constexpr int c2 = decltype(cc)::value;
                        constexpr int rho2 = kQc<Code>.rho[c2];
                        if (c == c2) { tc = tm[c2]; rho = rho2; }
                    });
                    int zv = t + rho;
                    if (zv >= Z) zv -= Z;
                    if (a.prob) a.prob[gbase + c * Z + zv] = prob_one(tc);
                    if (a.hard) a.hard[gbase + c * Z + zv] = (uint8_t)((hbits >> c) & 1u);
                }
            }
        }
    };
    auto syndrome_phase = [&]() {                                        // adds this thread's unsatisfied checks
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
    };
    for (int i = tid; i < CW; i += T) frozen_s[i] = 0;                  // (visible after the first barrier below)
    if (a.iters <= 0) {                                                 // no iteration: the messages are the zeros every caller passes
        for (int i = tid; i < CW * L::MSG_STRIDE; i += T) msg_s[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < (L::NLOC > 0 ? L::NLOC : 1); ++i) loc[i] = 0.0f;
        __syncthreads();
    }
    if constexpr (!EE) {
        // fixed iteration count (the reference's schedule, bp/bp.py:46-47); first iteration peeled
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
        if (active) marginal_phase(true);
        __syncthreads();
        if (a.syndrome) {
            if (active) syndrome_phase();
            __syncthreads();
        }
    } else {
        // Early termination.  One loop in which every phase appears once (so each lambda is inlined
        // exactly once): iteration, then marginal + hard decision + syndrome; a codeword is frozen as
        // soon as its hard decision satisfies every check; the last pass writes the outputs.
        bool finished = false;
        int it = 0;
#pragma unroll 1
        for (;;) {
            const bool do_iter = (it < a.iters) && !finished;
            if (do_iter) {
                const bool run = active && (it == 0 || frozen_s[cw] == 0);
                if (it == 0) { if (run) var_phase(std::true_type{}); }
                else { if (run) var_phase(std::false_type{}); }
                __syncthreads();
                if (run) check_phase();
                __syncthreads();
                ++it;
            }
            const bool last = !do_iter || it >= a.iters;                // no further iteration will run
            const bool run2 = active && (last || frozen_s[cw] == 0);
            if (run2) marginal_phase(last);
            __syncthreads();
            if (!last || a.syndrome) {
                if (run2) syndrome_phase();
                __syncthreads();
            }
            if (last) break;
            if (tid < CW) {
                if (tid < ncw && frozen_s[tid] == 0 && scratch[4 + tid] == 0) frozen_s[tid] = it;
                scratch[4 + tid] = 0;
            }
            __syncthreads();
            bool all = true;
            for (int c = 0; c < ncw; ++c) all = all && (frozen_s[c] != 0);
            finished = all;
        }
    }
    if (a.iters_used)
        for (int i = tid; i < ncw; i += T) a.iters_used[cw0 + i] = (EE && frozen_s[i]) ? frozen_s[i] : a.iters;
    if (a.syndrome) {
        for (int i = tid; i < ncw; i += T) a.syndrome[cw0 + i] = scratch[4 + i];
        __syncthreads();
        for (int i = tid; i < CW; i += T) scratch[4 + i] = 0;
    }
    if (a.hard_packed) pack_hard(hard_s, L::HARD_STRIDE, ncw, N, a.hard_packed + cw0 * ((N + 7) >> 3));
    if (a.counters) {
        __syncthreads();
        count_errors<(SIM != 0)>(hard_s, L::HARD_STRIDE, ncw, N, a.k_info, SIM ? nullptr : a.ref_packed + cw0 * ((N + 7) >> 3),
                     a.counters, scratch + 1);
    }
}

}  // namespace ldpc
