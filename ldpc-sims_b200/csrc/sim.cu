// sim.cu - the Monte-Carlo link simulator: random bits -> encode -> QPSK -> OFDM -> AWGN ->
// low-resolution ADC -> de-OFDM -> LLR -> BP decode -> exact error counters, all on the GPU.
//
// Replaces the per-SNR loop body of the evaluate scripts (evaluate_quantized_snr.py:91-188:
// create_bits / encode_bits / modulate_bits / gen_data / inline AGC quantizer / decode_bits /
// BER-BLER means).  Three launches per chunk of codewords, chained on one stream:
//   K2a gen_codewords   Philox info bits, systematic encode (bit-packed generator)
//   K2b linksim_llr     QPSK, per-codeword OFDM framing, warp IFFT, Philox AWGN, AGC+quantizer,
//                       warp FFT, exact LLR -> f32 [chunk, n] (stays L2-resident between K2b and K1)
//   K1+K3 decode        flooding BP + fused integer counters (epilogue.cuh)
// Every random draw is keyed by the GLOBAL codeword index, so counters do not depend on how the
// codewords are sharded over GPUs or chunks.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "linksim_device.cuh"

namespace ldpc {

// ---- K2a ----------------------------------------------------------------------------------------------
// One CTA per codeword (grid-stride).  Info bits: Philox(cw, RNG_BITS, block) -> 128 bits per call.
__global__ void __launch_bounds__(256) gen_codewords_kernel(const uint32_t *Pp /*[m][kw]*/, int n, int k, long long cw_first,
                                                            long long ncw, unsigned long long seed, uint8_t *cw_packed) {
    extern __shared__ uint32_t sm[];
    const int kw = (k + 31) / 32, m = n - k, nby = (n + 7) / 8;
    uint32_t *u_s = sm;                                   // [kw] info words, bit j of word w = bit 32w+j
    uint8_t *bits_s = reinterpret_cast<uint8_t *>(sm + kw);   // [n] one byte per code bit
    const Philox rng(seed);
    for (long long c = blockIdx.x; c < ncw; c += gridDim.x) {
        const unsigned long long gcw = (unsigned long long)(cw_first + c);
        for (int blk = threadIdx.x; blk * 4 < kw; blk += blockDim.x) info_words_block(rng, gcw, blk, k, u_s);
        __syncthreads();
        for (int i = threadIdx.x; i < k; i += blockDim.x) bits_s[i] = (u_s[i >> 5] >> (i & 31)) & 1u;
        for (int r = threadIdx.x; r < m; r += blockDim.x) {
            uint32_t acc = 0;
            for (int w = 0; w < kw; ++w) acc ^= __ldg(Pp + (long long)r * kw + w) & u_s[w];
            bits_s[k + r] = (uint8_t)(__popc(acc) & 1);
        }
        __syncthreads();
        for (int by = threadIdx.x; by < nby; by += blockDim.x) {
            unsigned v = 0;
            for (int b = 0; b < 8; ++b) {
                const int i = by * 8 + b;
                v |= (i < n ? (unsigned)bits_s[i] : 0u) << (7 - b);
            }
            cw_packed[c * nby + by] = (uint8_t)v;
        }
        __syncthreads();
    }
}

// Small codes (n <= 64, k <= 32: the reference's default (64,32) code): one THREAD per codeword - the information word
// is one Philox output, every parity bit one popc, the packed bytes one 64-bit store.  Same bits as the kernel above.
__global__ void __launch_bounds__(256) gen_codewords_small_kernel(const uint32_t *Pp /*[m][1]*/, int n, int k, long long cw_first,
                                                                  long long ncw, unsigned long long seed, uint8_t *cw_packed) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncw) return;
    const Philox rng(seed);
    uint32_t u = 0;
    info_words_block(rng, (unsigned long long)(cw_first + c), 0, k, &u);
    unsigned long long bits = u;                                        // bit i = code bit i (information bits first)
    const int m = n - k;
    for (int r = 0; r < m; ++r) bits |= (unsigned long long)(__popc(__ldg(Pp + r) & u) & 1) << (k + r);
    const int nby = (n + 7) / 8;
    unsigned long long w = 0;                                           // byte b = MSB-first packed byte b
    for (int b = 0; b < nby; ++b) w |= (unsigned long long)(__brev((unsigned)((bits >> (8 * b)) & 0xffu)) >> 24) << (8 * b);
    if (nby == 8) *reinterpret_cast<unsigned long long *>(cw_packed + c * 8) = w;
    else for (int b = 0; b < nby; ++b) cw_packed[c * nby + b] = (uint8_t)(w >> (8 * b));
}

// ---- K2b ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int cw_bit(const uint8_t *row, int i) { return (row[i >> 3] >> (7 - (i & 7))) & 1; }

// samples (nullable): [ncw * n_ofdm_per_cw][2N + 1] f32 rows (Re t = 0..N-1, Im t = 0..N-1, linear SNR): the
// input_samples of the MLP demappers (evaluate_quantized_snr.py:135-140)
template <int N, bool EXT_NOISE = false>
__global__ void __launch_bounds__(256) linksim_llr_kernel(const uint8_t *cw_packed, long long ncw, LinkParams p, float *llr,
                                                          float *samples) {
    __shared__ cplx<float> tw[N / 2];
    fill_twiddles_f<float>(tw, N);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = ncw * p.n_ofdm_per_cw;
    const int nby = (p.n + 7) / 8, nsym = p.n / 2;
    const LinkConsts k(p, N);
    constexpr int S = 1;                                     // OFDM symbols in flight per warp (2 measured no faster)
    for (long long o0 = warp * S; o0 < total; o0 += nwarps * S) {
        int os[S];
        unsigned long long gcw[S];
        bool valid[S];
        const uint8_t *row[S];
        float *orow[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const long long o = o0 + s;
            valid[s] = o < total;
            const long long c = valid[s] ? o / p.n_ofdm_per_cw : 0;
            os[s] = valid[s] ? (int)(o - c * p.n_ofdm_per_cw) : 0;
            gcw[s] = (unsigned long long)(p.cw_first + c);
            row[s] = cw_packed + c * nby;
            orow[s] = llr + c * p.n;
        }
        auto bitf = [&](int s, int i) { return cw_bit(row[s], i); };
        auto outf = [&](int s, int sidx, float l0, float l1) { *reinterpret_cast<float2 *>(orow[s] + 2 * sidx) = make_float2(l0, l1); };
        if (samples) {
            float *srow[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                srow[s] = samples + (o0 + s) * (2 * N + 1);
                if (valid[s] && lane == 0) srow[s][2 * N] = p.snr;
            }
            auto sampf = [&](int s, int t, float re, float im) { srow[s][t] = re; srow[s][N + t] = im; };
            ofdm_symbols_llr<N, S, decltype(bitf), decltype(outf), decltype(sampf), EXT_NOISE>(lane, os, gcw, valid, nsym, p, k, tw, bitf, outf, sampf);
        } else {
            ofdm_symbols_llr<N, S, decltype(bitf), decltype(outf), NoSamples, EXT_NOISE>(lane, os, gcw, valid, nsym, p, k, tw, bitf, outf);
        }
    }
}

}  // namespace ldpc

using namespace ldpc;

extern "C" {

void ldpc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    Philox rng(((uint64_t)key[1] << 32) | key[0]);
    uint32_t o[4];
    rng(ctr[0], ctr[1], ctr[2], ctr[3], o);
    for (int i = 0; i < 4; ++i) out[i] = o[i];
}

int ldpc_code_set_generator(ldpc_code_t *code, const uint32_t *parity_rows_packed_host, int k) {
    if (!code || !parity_rows_packed_host || k <= 0 || k >= code->n) { set_error("ldpc_code_set_generator: bad arguments"); return LDPC_EINVAL; }
    const int m = code->n - k, kw = (k + 31) / 32;
    if (code->d_gen) { cudaFree(code->d_gen); code->d_gen = nullptr; }
    LDPC_CUDA_TRY(cudaMalloc(&code->d_gen, (size_t)m * kw * sizeof(uint32_t)));
    LDPC_CUDA_TRY(cudaMemcpy(code->d_gen, parity_rows_packed_host, (size_t)m * kw * sizeof(uint32_t), cudaMemcpyHostToDevice));
    code->k_info = k;
    return LDPC_OK;
}

static int check_sim(const ldpc_code_t *code, const ldpc_sim_params_t *sp) {
    if (!code || !sp) { set_error("ldpc_sim: null argument"); return LDPC_EINVAL; }
    if (sp->struct_size != (int32_t)sizeof(ldpc_sim_params_t)) { set_error("ldpc_sim_params_t size mismatch (%d vs %d)", sp->struct_size, (int)sizeof(ldpc_sim_params_t)); return LDPC_EINVAL; }
    if (!code->d_gen) { set_error("code has no generator: call ldpc_code_set_generator first"); return LDPC_EINVAL; }
    if (code->n % 2) { set_error("QPSK needs an even code length"); return LDPC_EUNSUPPORTED; }
    if (sp->ofdm_size != 32 && sp->ofdm_size != 64 && sp->ofdm_size != 128 && sp->ofdm_size != 256) { set_error("ofdm_size must be 32, 64, 128 or 256"); return LDPC_EUNSUPPORTED; }
    if (sp->n_codewords < 0 || sp->qbits < 0 || sp->qbits > 16) { set_error("bad n_codewords / qbits"); return LDPC_EINVAL; }
    if (sp->qbits > 0 && sp->agc_mode != 1 && sp->agc_mode != 2) { set_error("agc_mode must be 1 (script AGC) or 2 (gen_qdata) when qbits > 0"); return LDPC_EINVAL; }
    return LDPC_OK;
}

static int launch_frontend(const ldpc_code_t *code, const ldpc_sim_params_t *sp, long long first, long long cnt,
                           uint8_t *cw_packed, float *llr, cudaStream_t s, float *samples = nullptr, bool given_codewords = false,
                           const float2 *noise = nullptr) {
    const int n = code->n, k = code->k_info, kw = (k + 31) / 32;
    const size_t sm = (size_t)kw * 4 + ((n + 3) & ~3);
    if (given_codewords) {
        // ldpc_sim_frontend: the caller supplies the transmitted codewords
    } else if (n <= 64 && k <= 32) {
        gen_codewords_small_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(code->d_gen, n, k, first, cnt, sp->seed, cw_packed);
    } else {
        const int g1 = (int)std::min<long long>(cnt, 148LL * 8);
        gen_codewords_kernel<<<g1, 256, sm, s>>>(code->d_gen, n, k, first, cnt, sp->seed, cw_packed);
    }
    LDPC_CUDA_TRY(cudaGetLastError());
    LinkParams lp;
    lp.n = n; lp.n_ofdm_per_cw = (n / 2 + sp->ofdm_size - 1) / sp->ofdm_size; lp.ofdm_size = sp->ofdm_size;
    lp.snr = powf(10.0f, sp->snr_db / 10.0f);
    lp.qbits = sp->qbits; lp.agc_mode = sp->agc_mode; lp.agc_clip = sp->agc_clip; lp.clip_ratio = sp->clip_ratio;
    lp.channel = (sp->reserved >> 1) & 1; lp.compander = (sp->reserved >> 2) & 1;
    lp.seed = sp->seed; lp.cw_first = first; lp.noise = noise;
    const long long warps = cnt * lp.n_ofdm_per_cw;
    const int g2 = (int)std::min<long long>((warps + 7) / 8, 148LL * 16);
    if (noise) {
        switch (sp->ofdm_size) {
            case 32: linksim_llr_kernel<32, true><<<g2, 256, 0, s>>>(cw_packed, cnt, lp, llr, samples); break;
            case 64: linksim_llr_kernel<64, true><<<g2, 256, 0, s>>>(cw_packed, cnt, lp, llr, samples); break;
            case 128: linksim_llr_kernel<128, true><<<g2, 256, 0, s>>>(cw_packed, cnt, lp, llr, samples); break;
            default: linksim_llr_kernel<256, true><<<g2, 256, 0, s>>>(cw_packed, cnt, lp, llr, samples); break;
        }
        LDPC_CUDA_TRY(cudaGetLastError());
        return LDPC_OK;
    }
    switch (sp->ofdm_size) {
        case 32: linksim_llr_kernel<32><<<g2, 256, 0, s>>>(cw_packed, cnt, lp, llr, samples); break;
        case 64: linksim_llr_kernel<64><<<g2, 256, 0, s>>>(cw_packed, cnt, lp, llr, samples); break;
        case 128: linksim_llr_kernel<128><<<g2, 256, 0, s>>>(cw_packed, cnt, lp, llr, samples); break;
        default: linksim_llr_kernel<256><<<g2, 256, 0, s>>>(cw_packed, cnt, lp, llr, samples); break;
    }
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int ldpc_sim_generate(const ldpc_code_t *code, const ldpc_sim_params_t *sp, uint8_t *cw_packed, float *llr,
                      ldpc_stream_t stream) {
    int rc = check_sim(code, sp);
    if (rc) return rc;
    if (!cw_packed || !llr) { set_error("ldpc_sim_generate: null output"); return LDPC_EINVAL; }
    if (sp->n_codewords == 0) return LDPC_OK;
    return launch_frontend(code, sp, sp->first_codeword, sp->n_codewords, cw_packed, llr, (cudaStream_t)stream);
}

int ldpc_sim_generate_ex(const ldpc_code_t *code, const ldpc_sim_params_t *sp, uint8_t *cw_packed, float *llr,
                         float *samples, ldpc_stream_t stream) {
    int rc = check_sim(code, sp);
    if (rc) return rc;
    if (!cw_packed || !llr) { set_error("ldpc_sim_generate_ex: null output"); return LDPC_EINVAL; }
    if (sp->n_codewords == 0) return LDPC_OK;
    return launch_frontend(code, sp, sp->first_codeword, sp->n_codewords, cw_packed, llr, (cudaStream_t)stream, samples);
}

int ldpc_sim_frontend(const ldpc_code_t *code, const ldpc_sim_params_t *sp, const uint8_t *cw_packed, const float *noise,
                      float *llr, float *samples, ldpc_stream_t stream) {
    if (!code || !sp) { set_error("ldpc_sim_frontend: null argument"); return LDPC_EINVAL; }
    if (sp->struct_size != (int32_t)sizeof(ldpc_sim_params_t)) { set_error("ldpc_sim_params_t size mismatch"); return LDPC_EINVAL; }
    if (code->n % 2) { set_error("QPSK needs an even code length"); return LDPC_EUNSUPPORTED; }
    if (sp->ofdm_size != 32 && sp->ofdm_size != 64 && sp->ofdm_size != 128 && sp->ofdm_size != 256) { set_error("ofdm_size must be 32, 64, 128 or 256"); return LDPC_EUNSUPPORTED; }
    if (sp->n_codewords < 0 || sp->qbits < 0 || sp->qbits > 16) { set_error("bad n_codewords / qbits"); return LDPC_EINVAL; }
    if (sp->qbits > 0 && sp->agc_mode != 1 && sp->agc_mode != 2) { set_error("agc_mode must be 1 (script AGC) or 2 (gen_qdata) when qbits > 0"); return LDPC_EINVAL; }
    if (!cw_packed || !llr) { set_error("ldpc_sim_frontend: null codewords / output"); return LDPC_EINVAL; }
    if (sp->n_codewords == 0) return LDPC_OK;
    return launch_frontend(code, sp, sp->first_codeword, sp->n_codewords, const_cast<uint8_t *>(cw_packed), llr, (cudaStream_t)stream, samples,
                           true, reinterpret_cast<const float2 *>(noise));
}

int ldpc_decode_count(const ldpc_code_t *code, const void *llr, int llr_dtype, int64_t B, int iters, int update,
                      float clamp_value, float param, const uint8_t *ref_packed, int k_info, int64_t *counters,
                      ldpc_stream_t stream) {
    if (!code || !llr || !ref_packed || !counters || B < 0) { set_error("ldpc_decode_count: null argument"); return LDPC_EINVAL; }
    if (iters < 0 || update < LDPC_UPDATE_SP || update > LDPC_UPDATE_OMS || !(clamp_value > 0.0f)) { set_error("ldpc_decode_count: bad decoder parameters"); return LDPC_EINVAL; }
    if (llr_dtype < LDPC_F32 || llr_dtype > LDPC_I8) { set_error("ldpc_decode_count: bad llr dtype"); return LDPC_EINVAL; }
    if (k_info <= 0 || k_info > code->n) { set_error("ldpc_decode_count: bad k"); return LDPC_EINVAL; }
    if (B == 0) return LDPC_OK;
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.llr = llr; a.llr_dtype = llr_dtype; a.B = B; a.iters = iters; a.update = update;
    a.clampv = clamp_value; a.param = param;
    a.ref_packed = ref_packed; a.counters = reinterpret_cast<unsigned long long *>(counters); a.k_info = k_info;
    a.precision = code->precision;
    return decode_dispatch(code, a, (cudaStream_t)stream);
}

int ldpc_sim_run(const ldpc_code_t *code, const ldpc_sim_params_t *sp, void *workspace, size_t workspace_bytes,
                 int64_t *counters, ldpc_stream_t stream) {
    int rc = check_sim(code, sp);
    if (rc) return rc;
    if (!counters) { set_error("ldpc_sim_run: null counters"); return LDPC_EINVAL; }
    if (sp->iters < 0 || sp->update < LDPC_UPDATE_SP || sp->update > LDPC_UPDATE_OMS || !(sp->clamp_value > 0.0f)) { set_error("ldpc_sim_run: bad decoder parameters"); return LDPC_EINVAL; }
    const int n = code->n, nby = (n + 7) / 8;
    cudaStream_t s = (cudaStream_t)stream;
    // ---- single launch: front end fused into the code-specialised decoder kernel -------------------
    if ((code->kernel == LDPC_KERNEL_QC || code->kernel == LDPC_KERNEL_QC_TMA) && code->precision == LDPC_PREC_F32 && !(sp->reserved & 1)) {
        DecodeArgs a;
        memset(&a, 0, sizeof(a));
        a.llr_dtype = LDPC_F32; a.B = sp->n_codewords; a.iters = sp->iters; a.update = sp->update;
        a.clampv = sp->clamp_value; a.param = sp->param;
        a.counters = reinterpret_cast<unsigned long long *>(counters); a.k_info = code->k_info;
        LinkParams lp;
        lp.n = n; lp.n_ofdm_per_cw = (n / 2 + sp->ofdm_size - 1) / sp->ofdm_size; lp.ofdm_size = sp->ofdm_size;
        lp.snr = powf(10.0f, sp->snr_db / 10.0f);
        lp.qbits = sp->qbits; lp.agc_mode = sp->agc_mode; lp.agc_clip = sp->agc_clip; lp.clip_ratio = sp->clip_ratio;
    lp.channel = (sp->reserved >> 1) & 1; lp.compander = (sp->reserved >> 2) & 1;
        lp.seed = sp->seed; lp.cw_first = sp->first_codeword;
        rc = launch_sim_fused_qc(code->qc_id, a, lp, s);
        if (rc != LDPC_EUNSUPPORTED) return rc;
    }
    // ---- three launches per chunk -----------------------------------------------------------------------
    const size_t per_cw = (size_t)n * sizeof(float) + ((nby + 15) & ~15);
    if (!workspace || workspace_bytes < per_cw * 1024) { set_error("ldpc_sim_run: workspace must hold at least 1024 codewords (%zu bytes)", per_cw * 1024); return LDPC_EINVAL; }
    long long chunk = (long long)(workspace_bytes / per_cw);
    chunk = std::min<long long>(chunk, 1 << 20);
    chunk &= ~1023LL;
    float *llr = reinterpret_cast<float *>(workspace);
    uint8_t *cwp = reinterpret_cast<uint8_t *>(workspace) + (size_t)chunk * n * sizeof(float);
    for (long long done = 0; done < sp->n_codewords; done += chunk) {
        const long long cnt = std::min<long long>(chunk, sp->n_codewords - done);
        rc = launch_frontend(code, sp, sp->first_codeword + done, cnt, cwp, llr, s);
        if (rc) return rc;
        DecodeArgs a;
        memset(&a, 0, sizeof(a));
        a.llr = llr; a.llr_dtype = LDPC_F32; a.B = cnt; a.iters = sp->iters; a.update = sp->update;
        a.clampv = sp->clamp_value; a.param = sp->param;
        a.ref_packed = cwp; a.counters = reinterpret_cast<unsigned long long *>(counters); a.k_info = code->k_info;
        rc = decode_dispatch(code, a, s);
        if (rc) return rc;
    }
    return LDPC_OK;
}

}  // extern "C"
