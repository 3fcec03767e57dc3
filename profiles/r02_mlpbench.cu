// r02_mlpbench.cu - stand-alone timing / accuracy harness of the MLP demapper kernels (csrc/mlp.cu is #included, so every
// -DMLP_EXP_* experiment switch of that file can be timed without touching the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ldpc-sims_b200/csrc -I include [-DMLP_EXP_...] \
//        -o mlpbench profiles/r02_mlpbench.cu -lcuda && ./mlpbench [rows=1048576] [reps=5] [chunk_rows=0 (library default)] [mode=0 auto|1 per layer|2 chain]
// 65 -> 512 -> 512 -> 512 -> 64 tanh chain (nn/llr.py:62-73) with seeded random weights of torch's default scale; prints the
// time per forward, the 16-bit MMA rate, and the error of 512 sampled rows against a float64 evaluation on the host.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "mlp.cu"

namespace ldpc {
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
int cuda_fail(cudaError_t e, const char *what) { fprintf(stderr, "CUDA error %s at %s\n", cudaGetErrorString(e), what); return LDPC_ECUDA; }
}  // namespace ldpc

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

int main(int argc, char **argv) {
    const long long B = argc > 1 ? atoll(argv[1]) : (1LL << 20);
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const long long chunk = argc > 3 ? atoll(argv[3]) : 0;
    const int mode = argc > 4 ? atoi(argv[4]) : 0;
    const int32_t dims[5] = {65, 512, 512, 512, 64};
    std::mt19937 rng(1234);
    std::vector<std::vector<float>> W(4), Bv(4);
    const float *wp[4], *bp[4];
    for (int l = 0; l < 4; ++l) {
        const float a = 1.0f / std::sqrt((float)dims[l]);
        std::uniform_real_distribution<float> u(-a, a);
        W[l].resize((size_t)dims[l + 1] * dims[l]);
        Bv[l].resize(dims[l + 1]);
        for (auto &v : W[l]) v = u(rng);
        for (auto &v : Bv[l]) v = u(rng);
        wp[l] = W[l].data(); bp[l] = Bv[l].data();
    }
    std::vector<float> x((size_t)B * 65);
    std::normal_distribution<float> nd(0.0f, 0.7f);
    for (auto &v : x) v = nd(rng);
    float *dx, *dy;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dy, (size_t)B * 64 * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
#ifdef MLP_EXP_PERSIST
    {   // experiment: set the persisting L2 carve-out to its maximum
        cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
        printf("L2 %d MB, persisting max %d MB, window max %d MB\n", pr.l2CacheSize >> 20, pr.persistingL2CacheMaxSize >> 20, pr.accessPolicyMaxWindowSize >> 20);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, pr.persistingL2CacheMaxSize));
    }
#endif
    ldpc_mlp_t *h = nullptr;
    if (ldpc_mlp_create(4, dims, wp, bp, nullptr, 2, chunk, &h)) return 1;
    if (ldpc_mlp_set_mode(h, mode)) return 1;
    if (ldpc_mlp_forward(h, dx, B, dy, nullptr)) return 1;
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) if (ldpc_mlp_forward(h, dx, B, dy, nullptr)) return 1;
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    const double flops = 2.0 * B * (128.0 * 512 + 2.0 * 512 * 512 + 512.0 * 64);      // as issued: K of the input layer padded to 128
    printf("%lld rows, chunk %lld, mode %d: %.3f ms per forward, %.0f TFLOP/s of 16-bit MMA issued (x3 plane pairs)\n", B, chunk, mode, ms, 3 * flops / (ms * 1e-3) / 1e12);
    // accuracy: 512 rows spread over the batch against float64
    std::vector<float> y((size_t)B * 64);
    CK(cudaMemcpy(y.data(), dy, y.size() * 4, cudaMemcpyDeviceToHost));
    unsigned long long hash = 1469598103934665603ULL;
    for (size_t i = 0; i < y.size(); ++i) { unsigned int u; memcpy(&u, &y[i], 4); hash = (hash ^ u) * 1099511628211ULL; }
    printf("output hash %016llx\n", hash);
    double emax = 0, scale = 0, esum = 0;
    const int NSAMP = 512;
    for (int sidx = 0; sidx < NSAMP; ++sidx) {
        const long long r = (B - 1) * sidx / (NSAMP - 1);
        std::vector<double> a(x.begin() + r * 65, x.begin() + r * 65 + 65), o;
        for (int l = 0; l < 4; ++l) {
            o.assign(dims[l + 1], 0.0);
            for (int n = 0; n < dims[l + 1]; ++n) {
                double acc = Bv[l][n];
                for (int k = 0; k < dims[l]; ++k) acc += a[k] * (double)W[l][(size_t)n * dims[l] + k];
                o[n] = l < 3 ? std::tanh(acc) : acc;
            }
            a = o;
        }
        for (int n = 0; n < 64; ++n) {
            const double e = std::fabs(a[n] - (double)y[r * 64 + n]);
            emax = std::max(emax, e); esum += e; scale = std::max(scale, std::fabs(a[n]));
        }
    }
    printf("error against float64 on %d rows: max %.3e, mean %.3e, output scale %.3f -> max / scale %.3e\n", NSAMP, emax, esum / (NSAMP * 64), scale, emax / scale);
#ifdef MLP_TRACE
    {   // pipeline events of CTA 0 during one more forward (clock64 relative to the first event; see TRACE() in mlp.cu)
        static long long tr[3][256]; int cnt[3];
        ldpc_mlp_debug_trace(&tr[0][0], cnt, 1);
        if (ldpc_mlp_forward(h, dx, B, dy, nullptr)) return 1;
        ldpc_mlp_debug_trace(&tr[0][0], cnt, 1);
        long long t0 = tr[1][0];
        for (int r = 0; r < 3; ++r) if (cnt[r] && tr[r][0] < t0) t0 = tr[r][0];
        const char *names[3] = {"producer (tile may load)", "mma (tile start, last commit)", "epilogue (start, end)"};
        for (int r = 0; r < 3; ++r) {
            printf("%s:", names[r]);
            for (int i = 0; i < cnt[r] && i < 120; ++i) printf(" %lld", tr[r][i] - t0);
            printf("\n");
        }
    }
#endif
    ldpc_mlp_destroy(h);
    return 0;
}
