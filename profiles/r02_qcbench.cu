// r02_qcbench.cu - A/B harness for the code-specialised decoder kernels (development tool, not product):
// times kernel variants on the headline workload (802.11n n=1944 r=1/2, min-sum x10, posterior f32 + packed
// bits out, LLRs resident in HBM) and checks every variant bit for bit against the round-1 kernel.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
//        -I ldpc-sims_b200/csrc -I include -o gpurun_out/qcbench profiles/r02_qcbench.cu
//   ./qcbench [codewords=1000000] [iters=10] [reps=5]
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "decode_qc_code.cuh"
#include "decode_qc_pers.cuh"

namespace ldpc {
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}
int cuda_fail(cudaError_t e, const char *what) {
    fprintf(stderr, "CUDA error %s at %s\n", cudaGetErrorString(e), what);
    return LDPC_ECUDA;
}
}  // namespace ldpc

using namespace ldpc;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// all-zero codeword over BPSK/AWGN at Eb/N0 (rate 1/2): llr = log P1/P0 = -2 y / sigma^2, y = 1 + sigma n
__global__ void gen_llr(float *llr, long long n, float sigma) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t a = hash32((uint32_t)i * 2u + 1u + (uint32_t)(i >> 31) * 0x9e3779b9u), b = hash32(a ^ 0x85ebca6bU);
        const float u1 = ((a >> 8) + 1) * (1.0f / 16777217.0f), u2 = (b >> 8) * (1.0f / 16777216.0f);
        const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        llr[i] = -2.0f * (1.0f + sigma * g) / (sigma * sigma);
    }
}
__global__ void count_diff(const uint32_t *x, const uint32_t *y, long long n, unsigned long long *out) {
    unsigned long long d = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) d += x[i] != y[i];
    if (d) atomicAdd(out, d);
}

struct Bufs {
    float *llr, *post_ref, *post;
    uint8_t *packed_ref, *packed;
    int32_t *synd_ref, *synd;
    unsigned long long *diff;
    long long B;
};

static DecodeArgs make_args(const Bufs &b, int iters, int upd, bool ref) {
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.llr = b.llr; a.llr_dtype = LDPC_F32; a.B = b.B; a.iters = iters; a.update = upd; a.clampv = 20.0f; a.param = 0.8125f;
    a.llr_post = ref ? b.post_ref : b.post;
    a.hard_packed = ref ? b.packed_ref : b.packed;
    return a;
}

template <class F>
static float time_ms(F &&launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

static unsigned long long diff_vs_ref(const Bufs &b) {
    CK(cudaMemset(b.diff, 0, 8));
    count_diff<<<1184, 256>>>((const uint32_t *)b.post_ref, (const uint32_t *)b.post, b.B * 1944, b.diff);
    count_diff<<<1184, 256>>>((const uint32_t *)b.packed_ref, (const uint32_t *)b.packed, b.B * 243 / 4, b.diff);
    unsigned long long h = 0;
    CK(cudaMemcpy(&h, b.diff, 8, cudaMemcpyDeviceToHost));
    return h;
}

static const char *g_only = nullptr;   // argv[4]: run only the variants whose name contains this string

template <class F>
static void run_variant(const char *name, const Bufs &b, int reps, F &&launch) {
    if (g_only && !strstr(name, g_only)) return;
    CK(cudaMemset(b.post, 0xff, (size_t)b.B * 1944 * 4));
    CK(cudaMemset(b.packed, 0xff, (size_t)b.B * 243));
    const float ms = time_ms(launch, reps);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
    const unsigned long long d = diff_vs_ref(b);
    printf("%-44s %8.3f ms  %7.2f Gbit/s  mismatching words vs round-1 kernel: %llu\n", name, ms, b.B * 972.0 / (ms * 1e-3) / 1e9, d);
    fflush(stdout);
}

int main(int argc, char **argv) {
    Bufs b;
    b.B = argc > 1 ? atoll(argv[1]) : 1000000;
    const int iters = argc > 2 ? atoi(argv[2]) : 10;
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    g_only = argc > 4 ? argv[4] : nullptr;
    const size_t nw = (size_t)b.B * 1944;
    CK(cudaMalloc(&b.llr, nw * 4));
    CK(cudaMalloc(&b.post_ref, nw * 4));
    CK(cudaMalloc(&b.post, nw * 4));
    CK(cudaMalloc(&b.packed_ref, (size_t)b.B * 243 + 16));
    CK(cudaMalloc(&b.packed, (size_t)b.B * 243 + 16));
    CK(cudaMalloc(&b.diff, 8));
    const float sigma = sqrtf(1.0f / (2.0f * 0.5f * powf(10.0f, 0.2f)));
    gen_llr<<<1184, 256>>>(b.llr, (long long)nw, sigma);
    CK(cudaDeviceSynchronize());
    printf("codewords %lld, iters %d, reps %d\n", b.B, iters, reps);

    for (int fmt = 0; fmt < 2; ++fmt) {
        // ---- reference: the round-1 kernels -------------------------------------------------------------------
        DecodeArgs ar = make_args(b, iters, UPD_MINSUM, true);
        if (fmt == 0) {
            const float ms = time_ms([&] { QcCodeImpl<Wifi1944R12, 3, true>::decode(ar, 0); }, reps);
            printf("%-44s %8.3f ms  %7.2f Gbit/s\n", "round-1 decode_qc_kernel<CW=3> fp32", ms, b.B * 972.0 / (ms * 1e-3) / 1e9);
        } else {
            const float ms = time_ms([&] { QcCodeImpl<Wifi1944R12, 3, true>::decode_h2(ar, 0); }, reps);
            printf("%-44s %8.3f ms  %7.2f Gbit/s\n", "round-1 decode_qc_h2_kernel<3 pairs> f16x2", ms, b.B * 972.0 / (ms * 1e-3) / 1e9);
        }
        CK(cudaDeviceSynchronize());
        DecodeArgs av = make_args(b, iters, UPD_MINSUM, false);
        if (fmt == 0) {   // occupancy experiment: the round-1 kernel with ONE CTA per SM (extra dynamic shared memory), i.e. half the warps
            using L1 = QcLayout<Wifi1944R12, 3>;
            auto k1 = decode_qc_kernel<Wifi1944R12, 3, UPD_MINSUM, 0, false>;
            CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
            run_variant("round-1 kernel, 1 CTA/SM (8 warps)", b, reps, [&] { k1<<<(int)((b.B + 2) / 3), L1::THREADS, 120 * 1024>>>(av, LinkParams()); });
            CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L1::SMEM));
        }
        if (fmt == 0) {
            run_variant("pers fp32 VB=0", b, reps, [&] { launch_qc_pers<Wifi1944R12, 3, UPD_MINSUM, float, 0>(av, 0); });
            run_variant("pers fp32 VB=6", b, reps, [&] { launch_qc_pers<Wifi1944R12, 3, UPD_MINSUM, float, 6>(av, 0); });
            run_variant("pers fp32 VB=11", b, reps, [&] { launch_qc_pers<Wifi1944R12, 3, UPD_MINSUM, float, 11>(av, 0); });
        } else {
            run_variant("pers f16x2 VB=0", b, reps, [&] { launch_qc_pers<Wifi1944R12, 3, UPD_MINSUM, __half2, 0>(av, 0); });
            run_variant("pers f16x2 VB=6", b, reps, [&] { launch_qc_pers<Wifi1944R12, 3, UPD_MINSUM, __half2, 6>(av, 0); });
        }
    }
    return 0;
}
