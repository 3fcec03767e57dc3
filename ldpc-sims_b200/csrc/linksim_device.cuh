// linksim_device.cuh - the per-OFDM-symbol link chain of the simulator as ONE device function,
// shared by the standalone front-end kernel (sim.cu, K2b) and the single-launch fused
// simulator (decode_qc.cu): QPSK -> warp IFFT -> Philox AWGN -> AGC + quantizer -> warp FFT ->
// exact LLR.  Both translation units are compiled with -fmad=false so the two paths produce
// the same float bits and therefore the same integer counters.
// Reference chain: ofdm/ofdm_functions.py:17-35,63-78 and evaluate_quantized_snr.py:96-133.
#pragma once
#include "frontend.cuh"

namespace ldpc {

struct LinkParams {
    int n;                 // code length (bits)
    int n_ofdm_per_cw;     // ceil((n/2) / N)
    int ofdm_size;
    float snr;             // linear per-subcarrier Es/N0 (ofdm_functions.py:110)
    int qbits;             // 0 = no ADC model
    int agc_mode;          // 1 = script AGC (evaluate_quantized_snr.py:103-111), 2 = gen_qdata-style clip
    float agc_clip, clip_ratio;
    unsigned long long seed;
    long long cw_first;
    int channel = 0;       // 0 = AWGN (the reference, ofdm_functions.py:30-33); 1 = flat Rayleigh block fading, one CN(0,1) gain per OFDM
                           //     symbol, coherent receiver with perfect channel knowledge (north star: "AWGN/fading channel"; not in the reference)
    int compander = 0;     // 1 = tanh compander (soft clipping clip * tanh(x / clip)) in front of the uniform ADC (north star: "uniform/tanh quantizer")
    const float2 *noise = nullptr;   // ldpc_sim_frontend only: caller-supplied noise samples [codeword][ofdm symbol][time sample] instead of Philox
};

struct LinkConsts {
    float scale, a, nstd, two_np, factor, clip, levels;
    __device__ __forceinline__ LinkConsts(const LinkParams &p, int N) {
        scale = rsqrtf((float)N);
        a = 0.70710678118654752f;
        nstd = sqrtf(0.5f / p.snr);                      // per real dimension
        two_np = 1.0f / p.snr;                           // 2 * (0.5 / snr)
        factor = 1.0f; clip = 1.0f;
        if (p.qbits > 0) {
            if (p.agc_mode == 1) { clip = p.agc_clip; factor = p.agc_clip / (0.5f * (1.0f + 1.0f / p.snr)) * p.clip_ratio; }
            else { clip = sqrtf(1.0f + 1.0f / p.snr) * p.clip_ratio; }
        }
        levels = (float)(1 << (p.qbits > 0 ? p.qbits : 1));
    }
};

template <typename T>
__device__ __forceinline__ void fill_twiddles_f(cplx<T> *tw, int N) {
    for (int j = threadIdx.x; j < N / 2; j += blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)j / (double)N, &s, &c);
        tw[j] = {(T)c, (T)s};
    }
}

// One warp, S independent OFDM symbols at once (the S dependency chains of shuffles interleave and
// hide each other's latency).  Symbol slot s: OFDM symbol os[s] of global codeword gcw[s];
// valid[s] = false skips its output.  bit(s, i) -> 0/1 code bit i of that codeword;
// out(s, sidx, llr_b0, llr_b1) receives the LLR pair of QPSK symbol sidx < nsym.
// samp(s, t, re, im) receives the received (noisy, quantized and rescaled) TIME sample t of the symbol -
// the input of the MLP demappers (evaluate_quantized_snr.py:135-137); pass NoSamples to skip.
// Per-symbol arithmetic does not depend on S.
struct NoSamples { __device__ __forceinline__ void operator()(int, int, float, float) const {} };

// EXT_NOISE: the additive noise comes from p.noise (identical-input parity tests against the reference's own
// noise realisation, evaluate_quantized_snr.py:96-133); everything after the addition is the same code.
template <int N, int S, class BitFn, class OutFn, class SampFn = NoSamples, bool EXT_NOISE = false>
__device__ __forceinline__ void ofdm_symbols_llr(int lane, const int (&os)[S], const unsigned long long (&gcw)[S],
                                                 const bool (&valid)[S], int nsym, const LinkParams &p, const LinkConsts &k,
                                                 const cplx<float> *tw, BitFn bit, OutFn out, SampFn samp = SampFn()) {
    constexpr int P = N / 32, LOGN = ilog2(N);
    const Quantizer<float> quant(k.levels, k.clip);
    const Philox rng(p.seed);
    cplx<float> x[S][P];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int r = 0; r < P; ++r) {                     // QPSK, null subcarriers past the codeword
            const int sidx = os[s] * N + r * 32 + lane;
            if (valid[s] && sidx < nsym) x[s][r] = {bit(s, 2 * sidx) ? -k.a : k.a, bit(s, 2 * sidx + 1) ? -k.a : k.a};   // a (1 - 2 b), exactly
            else x[s][r] = {0.0f, 0.0f};
        }
#pragma unroll
    for (int s = 0; s < S; ++s) warp_fft<N, float, true>(x[s], lane, tw, k.scale);   // time sample t = bitrev(r*32+lane)
    // flat block fading: one complex gain per OFDM symbol (Philox stream RNG_FADE, counter = symbol index: every lane draws the same)
    cplx<float> hg[S];
    float inv_h2[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        hg[s] = {1.0f, 0.0f};
        inv_h2[s] = 1.0f;
        if (p.channel == 1) {
            uint32_t rnd[4];
            rng((uint32_t)gcw[s], (uint32_t)(gcw[s] >> 32), RNG_FADE, (uint32_t)os[s], rnd);
            float g0, g1;
            box_muller<float>(rnd[0], rnd[1], g0, g1);
            hg[s] = {k.a * g0, k.a * g1};                                              // CN(0, 1)
            inv_h2[s] = 1.0f / fmaxf(hg[s].re * hg[s].re + hg[s].im * hg[s].im, 1e-12f);
#pragma unroll
            for (int r = 0; r < P; ++r) x[s][r] = {x[s][r].re * hg[s].re - x[s][r].im * hg[s].im, x[s][r].re * hg[s].im + x[s][r].im * hg[s].re};
        }
    }
    auto adc = [&](float v) -> float {                                                 // AGC scale, optional compander, uniform ADC, rescale
        float w = k.factor * v;
        if (p.compander) w = k.clip * tanhf(w / k.clip);
        return div_rn_nochk(quant(w), k.factor);
    };
    // AWGN: one Philox block serves the two time samples 2j, 2j+1 (counter = j, words {0,1} / {2,3}).
    // With i = r*32 + lane the sample index t = bitrev(i) has bit 0 = bit LOGN-1 of i, so for N >= 64
    // registers r and r + P/2 of a lane hold exactly such a pair and share one Philox call.
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if constexpr (P >= 2) {
#pragma unroll
            for (int r = 0; r < P / 2; ++r) {
                const int t0 = bitrev(r * 32 + lane, LOGN);                   // even
                float z[4];
                if constexpr (EXT_NOISE) {
                    const float2 *nz = p.noise + ((long long)(gcw[s] - (unsigned long long)p.cw_first) * p.n_ofdm_per_cw + os[s]) * N + t0;
                    const float2 n0 = valid[s] ? nz[0] : make_float2(0.f, 0.f), n1 = valid[s] ? nz[1] : make_float2(0.f, 0.f);
                    z[0] = n0.x; z[1] = n0.y; z[2] = n1.x; z[3] = n1.y;
                } else {
                    uint32_t rnd[4];
                    rng((uint32_t)gcw[s], (uint32_t)(gcw[s] >> 32), RNG_NOISE, (uint32_t)((os[s] * N + t0) >> 1), rnd);
                    box_muller<float>(rnd[0], rnd[1], z[0], z[1]);
                    box_muller<float>(rnd[2], rnd[3], z[2], z[3]);
                }
                const float nscale = EXT_NOISE ? 1.0f : k.nstd;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int rr = r + h * (P / 2);
                    float re = x[s][rr].re + nscale * z[2 * h], im = x[s][rr].im + nscale * z[2 * h + 1];
                    if (p.qbits > 0) { re = adc(re); im = adc(im); }
                    x[s][rr] = {re, im};
                    if (valid[s]) samp(s, t0 + h, re, im);
                }
            }
        } else {
            const int t = bitrev(lane, LOGN);
            float z0, z1;
            if constexpr (EXT_NOISE) {
                const float2 nz = valid[s] ? p.noise[((long long)(gcw[s] - (unsigned long long)p.cw_first) * p.n_ofdm_per_cw + os[s]) * N + t] : make_float2(0.f, 0.f);
                z0 = nz.x; z1 = nz.y;
            } else {
                uint32_t rnd[4];
                rng((uint32_t)gcw[s], (uint32_t)(gcw[s] >> 32), RNG_NOISE, (uint32_t)((os[s] * N + t) >> 1), rnd);
                box_muller<float>(rnd[2 * (t & 1)], rnd[2 * (t & 1) + 1], z0, z1);
            }
            const float nscale = EXT_NOISE ? 1.0f : k.nstd;
            float re = x[s][0].re + nscale * z0, im = x[s][0].im + nscale * z1;
            if (p.qbits > 0) { re = adc(re); im = adc(im); }
            x[s][0] = {re, im};
            if (valid[s]) samp(s, t, re, im);
        }
    }
    // coherent receiver: divide by the (flat) channel gain before the FFT; the noise power seen by the demapper is sigma^2 / |h|^2
#pragma unroll
    for (int s = 0; s < S; ++s)
        if (p.channel == 1) {
#pragma unroll
            for (int r = 0; r < P; ++r)
                x[s][r] = {(x[s][r].re * hg[s].re + x[s][r].im * hg[s].im) * inv_h2[s], (x[s][r].im * hg[s].re - x[s][r].re * hg[s].im) * inv_h2[s]};
        }
#pragma unroll
    for (int s = 0; s < S; ++s) warp_fft_dit<N, float, false>(x[s], lane, tw, k.scale);  // back to natural subcarrier order
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const int sidx = os[s] * N + r * 32 + lane;
            const float tnp = k.two_np * inv_h2[s];
            if (valid[s] && sidx < nsym) out(s, sidx, qpsk_llr<float>(x[s][r].re, k.a, tnp), qpsk_llr<float>(x[s][r].im, k.a, tnp));
        }
}

// Philox information words of one codeword: word w (bit j = information bit 32 w + j) comes
// from block w/4, lane w%4 of stream RNG_BITS; bits at or beyond k are cleared.
__device__ __forceinline__ void info_words_block(const Philox &rng, unsigned long long gcw, int blk, int k, uint32_t *u /* [kw] */) {
    const int kw = (k + 31) / 32;
    uint32_t r[4];
    rng((uint32_t)gcw, (uint32_t)(gcw >> 32), RNG_BITS, (uint32_t)blk, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int w = blk * 4 + j;
        if (w < kw) {
            uint32_t v = r[j];
            if (32 * (w + 1) > k) v &= (k - 32 * w >= 32) ? 0xffffffffu : ((1u << (k - 32 * w)) - 1u);
            u[w] = v;
        }
    }
}

}  // namespace ldpc
