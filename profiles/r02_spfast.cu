// r02_spfast.cu - EXPERIMENT: sum-product with MUFU-based tanh / log (node_math.cuh: tanh_half_fast,
// log_ratio_fast) against the shipped sum-product (libm tanhf / logf + correctly rounded division) on the headline code:
// run time of both and the histogram of |dt| / max(|t|, 16.64) between them, plus the hard-decision differences.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I ldpc-sims_b200/csrc -I include \
//        -o spfast profiles/r02_spfast.cu && ./spfast [codewords=262144] [iters=10]
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "decode_qc_kernel.cuh"

namespace ldpc {
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
int cuda_fail(cudaError_t e, const char *what) { fprintf(stderr, "CUDA error %s at %s\n", cudaGetErrorString(e), what); return LDPC_ECUDA; }
}  // namespace ldpc
using namespace ldpc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void gen_llr(float *llr, long long n, float sigma) {      // all-zero codeword, BPSK/AWGN
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t a = hash32((uint32_t)i * 2u + 1u + (uint32_t)(i >> 31) * 0x9e3779b9u), b = hash32(a ^ 0x85ebca6bU);
        const float u1 = ((a >> 8) + 1) * (1.0f / 16777217.0f), u2 = (b >> 8) * (1.0f / 16777216.0f);
        llr[i] = -2.0f * (1.0f + sigma * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2)) / (sigma * sigma);
    }
}
// bins: e < 1e-7, < 1e-6, < 1e-5, < 1e-4, < 1e-3, < 1e-2, < 1e-1, >= 1e-1; [8] = hard-decision differences; [9] = relative > 1e-4
__global__ void compare(const float *x, const float *y, long long n, unsigned long long *h) {
    unsigned long long loc[10] = {};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = x[i], b = y[i];                              // posterior LLRs = -2 t
        const float d = fabsf(a - b) * 0.5f, sc = fmaxf(fabsf(a) * 0.5f, 16.64f), e = d / sc;
        int bin = 0;
        for (float th = 1e-7f; bin < 7 && e >= th; th *= 10.0f) ++bin;
        ++loc[bin];
        loc[8] += (a > 0.0f) != (b > 0.0f);
        loc[9] += (d > 1e-4f * fabsf(a) * 0.5f) && (d > 1e-6f);
    }
    for (int k = 0; k < 10; ++k) if (loc[k]) atomicAdd(&h[k], loc[k]);
}

template <int UPD>
static float run(const DecodeArgs &a, int reps) {
    using L = QcLayout<Wifi1944R12, 3>;
    auto k = decode_qc_kernel<Wifi1944R12, 3, UPD, 0, false>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = (int)((a.B + 2) / 3);
    k<<<grid, L::THREADS, L::SMEM>>>(a, LinkParams());
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) k<<<grid, L::THREADS, L::SMEM>>>(a, LinkParams());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char **argv) {
    const long long B = argc > 1 ? atoll(argv[1]) : 262144;
    const int iters = argc > 2 ? atoi(argv[2]) : 10;
    float *llr, *p0, *p1; unsigned long long *h;
    CK(cudaMalloc(&llr, B * 1944 * 4)); CK(cudaMalloc(&p0, B * 1944 * 4)); CK(cudaMalloc(&p1, B * 1944 * 4)); CK(cudaMalloc(&h, 80));
    for (float ebn0 : {2.0f, 3.0f}) {
        const float sigma = sqrtf(1.0f / (2.0f * 0.5f * powf(10.0f, ebn0 / 10.0f)));
        gen_llr<<<1184, 256>>>(llr, B * 1944, sigma);
        DecodeArgs a; memset(&a, 0, sizeof(a));
        a.llr = llr; a.llr_dtype = LDPC_F32; a.B = B; a.iters = iters; a.update = UPD_SP; a.clampv = 20.0f; a.param = 1.0f;
        a.llr_post = p0;
        const float ms0 = run<UPD_SP>(a, 3);
        a.llr_post = p1;
        const float ms1 = run<UPD_SPF>(a, 3);
        CK(cudaMemset(h, 0, 80));
        compare<<<1184, 256>>>(p0, p1, B * 1944, h);
        unsigned long long hh[10]; CK(cudaMemcpy(hh, h, 80, cudaMemcpyDeviceToHost));
        const double n = (double)B * 1944;
        printf("Eb/N0 %.0f dB, %lld codewords, %d iterations: libm sum-product %.3f ms (%.2f Gbit/s), fast %.3f ms (%.2f Gbit/s), x%.2f\n",
               ebn0, B, iters, ms0, B * 972.0 / (ms0 * 1e-3) / 1e9, ms1, B * 972.0 / (ms1 * 1e-3) / 1e9, ms0 / ms1);
        printf("  |dt|/max(|t|,16.64) histogram by decade [<1e-7 .. >=1e-1]: %llu %llu %llu %llu %llu %llu %llu %llu\n", hh[0], hh[1], hh[2], hh[3],
               hh[4], hh[5], hh[6], hh[7]);
        printf("  beyond 1e-4 of scale: %.3e of the values; beyond 1e-4 relative: %.3e; hard decisions that differ: %llu of %.0f (%.2e)\n",
               (double)(hh[4] + hh[5] + hh[6] + hh[7]) / n, hh[9] / n, hh[8], n, hh[8] / n);
    }
    return 0;
}
