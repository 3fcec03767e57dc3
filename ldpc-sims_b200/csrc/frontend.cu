// frontend.cu - the link-simulator front end (K2) as standalone kernels + their C ABI.
//
// Reference functions replaced (ofdm/ofdm_functions.py): encode_bits :11-15, modulate_bits
// :17-22, transmit_symbols :25-35, quantizer :37-51, demodulate_signal :63-78, and the
// inline AGC-scaled quantizer front end of evaluate_quantized_snr.py:96-133.
//
// Layout convention = the reference's flat one: a (1, L) array is a sequence of OFDM symbols,
// symbol j occupying [j*N, (j+1)*N) (the reference reshapes to (-1, N).T and back).  Complex
// values are interleaved (re, im) pairs of the real type (numpy complex128 / complex64).
// One warp owns one OFDM symbol at a time; the N-point transform runs in registers with warp
// shuffles (frontend.cuh), never through cuFFT.
#include <math.h>

#include "common.cuh"
#include "frontend.cuh"

namespace ldpc {

template <typename T>
__device__ __forceinline__ void fill_twiddles(cplx<T> *tw, int N) {
    for (int j = threadIdx.x; j < N / 2; j += blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)j / (double)N, &s, &c);
        tw[j] = {(T)c, (T)s};
    }
}

// ---- transmit: time = W^H s (unitary), rx = time + noise ---------------------------------------------
template <int N, typename T>
__global__ void __launch_bounds__(256) ofdm_transmit_kernel(const cplx<T> *sym, long long n_ofdm, const cplx<T> *noise,
                                                            double snr, unsigned long long seed, cplx<T> *rx, cplx<T> *tx) {
    constexpr int P = N / 32, LOGN = ilog2(N);
    __shared__ cplx<T> tw[N / 2];
    __shared__ cplx<T> stage[8][N];
    fill_twiddles<T>(tw, N);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp = (long long)blockIdx.x * 8 + wib, nwarps = (long long)gridDim.x * 8;
    const T scale = (T)(1.0 / sqrt((double)N));
    const T nstd = (T)(1.0 / sqrt(2.0 * snr));             // per real dimension: N(0,1/sqrt(snr))/sqrt(2)
    const Philox rng(seed);
    for (long long o = warp; o < n_ofdm; o += nwarps) {
        cplx<T> x[P];
#pragma unroll
        for (int r = 0; r < P; ++r) x[r] = sym[o * N + r * 32 + lane];
        warp_fft<N, T, true>(x, lane, tw, scale);
#pragma unroll
        for (int r = 0; r < P; ++r) stage[wib][bitrev(r * 32 + lane, LOGN)] = x[r];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const int t = r * 32 + lane;
            const long long g = o * N + t;
            const cplx<T> v = stage[wib][t];
            cplx<T> w;
            if (noise) w = noise[g];
            else {
                uint32_t rnd[4];
                rng((uint32_t)g, (uint32_t)(g >> 32), RNG_NOISE, 0u, rnd);
                T z0, z1;
                box_muller<T>(rnd[0], rnd[1], z0, z1);
                w = {z0 * nstd, z1 * nstd};
            }
            if (tx) tx[g] = v;
            rx[g] = {v.re + w.re, v.im + w.im};
        }
        __syncwarp();
    }
}

// ---- demodulate: R = W r, exact QPSK LLRs, interleaved (b0, b1) per subcarrier --------------------------
template <int N, typename T>
__global__ void __launch_bounds__(256) ofdm_demod_kernel(const cplx<T> *sig, long long n_ofdm, double snr_est, T *llr,
                                                         cplx<T> *symbols) {
    constexpr int P = N / 32, LOGN = ilog2(N);
    __shared__ cplx<T> tw[N / 2];
    __shared__ cplx<T> stage[8][N];
    fill_twiddles<T>(tw, N);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp = (long long)blockIdx.x * 8 + wib, nwarps = (long long)gridDim.x * 8;
    const T scale = (T)(1.0 / sqrt((double)N));
    const T a = (T)(1.0 / sqrt(2.0));
    const T two_np = (T)(2.0 * (0.5 * (1.0 / snr_est)));
    for (long long o = warp; o < n_ofdm; o += nwarps) {
        cplx<T> x[P];
#pragma unroll
        for (int r = 0; r < P; ++r) x[r] = sig[o * N + r * 32 + lane];
        warp_fft<N, T, false>(x, lane, tw, scale);
#pragma unroll
        for (int r = 0; r < P; ++r) stage[wib][bitrev(r * 32 + lane, LOGN)] = x[r];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const int k = r * 32 + lane;
            const long long g = o * N + k;
            const cplx<T> R = stage[wib][k];
            if (symbols) symbols[g] = R;
            if (llr) {
                llr[2 * g] = qpsk_llr<T>(R.re, a, two_np);
                llr[2 * g + 1] = qpsk_llr<T>(R.im, a, two_np);
            }
        }
        __syncwarp();
    }
}

template <typename T>
__global__ void quantize_kernel(const T *in, long long n_real, T num_levels, T clip, T *out) {
    const Quantizer<T> q(num_levels, clip);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_real; i += (long long)gridDim.x * blockDim.x)
        out[i] = q(in[i]);
}

// bits (0 -> +1, 1 -> -1), pairs (b0, b1) -> ((1-2 b0) + j (1-2 b1)) / sqrt(2)
template <typename T>
__global__ void modulate_kernel(const uint8_t *bits, long long n_sym, cplx<T> *out) {
    const T a = (T)(1.0 / sqrt(2.0));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_sym; i += (long long)gridDim.x * blockDim.x) {
        const T b0 = (T)(-2 * (int)bits[2 * i] + 1), b1 = (T)(-2 * (int)bits[2 * i + 1] + 1);
        out[i] = {a * b0, a * b1};
    }
}

// c = G u mod 2 with bit-packed generator rows: bit j of Gp[r*kw + w] = G[r][32 w + j]
__global__ void encode_kernel(const uint8_t *bits, const uint32_t *Gp, int n, int k, long long ncw, uint8_t *out) {
    extern __shared__ uint32_t u_s[];                        // [kw] packed information bits
    const int kw = (k + 31) / 32;
    for (long long cw = blockIdx.x; cw < ncw; cw += gridDim.x) {
        for (int w = threadIdx.x; w < kw; w += blockDim.x) {
            uint32_t v = 0;
            for (int j = 0; j < 32 && 32 * w + j < k; ++j) v |= (uint32_t)(bits[cw * k + 32 * w + j] & 1) << j;
            u_s[w] = v;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < n; r += blockDim.x) {
            uint32_t acc = 0;
            for (int w = 0; w < kw; ++w) acc ^= __ldg(Gp + (long long)r * kw + w) & u_s[w];
            out[cw * n + r] = (uint8_t)(__popc(acc) & 1);
        }
        __syncthreads();
    }
}

static int grid_for(long long work, int per_block) {
    long long g = (work + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > 148LL * 16) g = 148LL * 16;
    return (int)g;
}

template <typename T>
static int transmit_t(const void *sym, long long n_ofdm, int N, const void *noise, double snr, unsigned long long seed,
                      void *rx, void *tx, cudaStream_t s) {
    const int grid = grid_for(n_ofdm, 8);
#define LAUNCH(NN)                                                                                                     \
    ofdm_transmit_kernel<NN, T><<<grid, 256, 0, s>>>((const cplx<T> *)sym, n_ofdm, (const cplx<T> *)noise, snr, seed, \
                                                     (cplx<T> *)rx, (cplx<T> *)tx)
    switch (N) {
        case 32: LAUNCH(32); break;
        case 64: LAUNCH(64); break;
        case 128: LAUNCH(128); break;
        case 256: LAUNCH(256); break;
        default: set_error("ofdm_size must be 32, 64, 128 or 256 (got %d)", N); return LDPC_EUNSUPPORTED;
    }
#undef LAUNCH
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

template <typename T>
static int demod_t(const void *sig, long long n_ofdm, int N, double snr_est, void *llr, void *symbols, cudaStream_t s) {
    const int grid = grid_for(n_ofdm, 8);
#define LAUNCH(NN) ofdm_demod_kernel<NN, T><<<grid, 256, 0, s>>>((const cplx<T> *)sig, n_ofdm, snr_est, (T *)llr, (cplx<T> *)symbols)
    switch (N) {
        case 32: LAUNCH(32); break;
        case 64: LAUNCH(64); break;
        case 128: LAUNCH(128); break;
        case 256: LAUNCH(256); break;
        default: set_error("ofdm_size must be 32, 64, 128 or 256 (got %d)", N); return LDPC_EUNSUPPORTED;
    }
#undef LAUNCH
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

}  // namespace ldpc

using namespace ldpc;

extern "C" {

int ldpc_encode_bits(const uint8_t *bits, const uint32_t *G_packed, int n, int k, int64_t ncw, uint8_t *out,
                     ldpc_stream_t stream) {
    if (!bits || !G_packed || !out || n <= 0 || k <= 0 || ncw < 0) { set_error("ldpc_encode_bits: bad arguments"); return LDPC_EINVAL; }
    if (ncw == 0) return LDPC_OK;
    const int kw = (k + 31) / 32;
    encode_kernel<<<grid_for(ncw, 1), 256, kw * sizeof(uint32_t), (cudaStream_t)stream>>>(bits, G_packed, n, k, ncw, out);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int ldpc_modulate_bits(const uint8_t *bits, int64_t n_symbols, int real_dtype, void *out, ldpc_stream_t stream) {
    if (!bits || !out || n_symbols < 0) { set_error("ldpc_modulate_bits: bad arguments"); return LDPC_EINVAL; }
    if (n_symbols == 0) return LDPC_OK;
    const int grid = grid_for(n_symbols, 256);
    if (real_dtype == LDPC_F64) modulate_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(bits, n_symbols, (cplx<double> *)out);
    else if (real_dtype == LDPC_F32) modulate_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(bits, n_symbols, (cplx<float> *)out);
    else { set_error("real_dtype must be LDPC_F32 or LDPC_F64"); return LDPC_EINVAL; }
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int ldpc_ofdm_transmit(const void *symbols, int64_t n_ofdm, int ofdm_size, int real_dtype, const void *noise, double snr,
                       uint64_t seed, void *rx, void *tx, ldpc_stream_t stream) {
    if (!symbols || !rx || n_ofdm < 0 || !(snr > 0.0)) { set_error("ldpc_ofdm_transmit: bad arguments"); return LDPC_EINVAL; }
    if (n_ofdm == 0) return LDPC_OK;
    if (real_dtype == LDPC_F64) return transmit_t<double>(symbols, n_ofdm, ofdm_size, noise, snr, seed, rx, tx, (cudaStream_t)stream);
    if (real_dtype == LDPC_F32) return transmit_t<float>(symbols, n_ofdm, ofdm_size, noise, snr, seed, rx, tx, (cudaStream_t)stream);
    set_error("real_dtype must be LDPC_F32 or LDPC_F64");
    return LDPC_EINVAL;
}

int ldpc_quantize(const void *in, int64_t n_real, int real_dtype, double num_levels, double clip, void *out,
                  ldpc_stream_t stream) {
    if (!in || !out || n_real < 0 || !(num_levels > 1.0)) { set_error("ldpc_quantize: bad arguments"); return LDPC_EINVAL; }
    if (n_real == 0) return LDPC_OK;
    const int grid = grid_for(n_real, 1024);
    if (real_dtype == LDPC_F64) quantize_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double *)in, n_real, num_levels, clip, (double *)out);
    else if (real_dtype == LDPC_F32) quantize_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)in, n_real, (float)num_levels, (float)clip, (float *)out);
    else { set_error("real_dtype must be LDPC_F32 or LDPC_F64"); return LDPC_EINVAL; }
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int ldpc_ofdm_demodulate(const void *signal, int64_t n_ofdm, int ofdm_size, int real_dtype, double snr_est, void *llrs,
                         void *symbols, ldpc_stream_t stream) {
    if (!signal || n_ofdm < 0 || !(snr_est > 0.0) || (!llrs && !symbols)) { set_error("ldpc_ofdm_demodulate: bad arguments"); return LDPC_EINVAL; }
    if (n_ofdm == 0) return LDPC_OK;
    if (real_dtype == LDPC_F64) return demod_t<double>(signal, n_ofdm, ofdm_size, snr_est, llrs, symbols, (cudaStream_t)stream);
    if (real_dtype == LDPC_F32) return demod_t<float>(signal, n_ofdm, ofdm_size, snr_est, llrs, symbols, (cudaStream_t)stream);
    set_error("real_dtype must be LDPC_F32 or LDPC_F64");
    return LDPC_EINVAL;
}

}  // extern "C"
