// micro-benchmark of issue rates: FADD, FADD2, FMNMX.XORSIGN, FMNMX3, HMNMX2.XORSIGN, mixes
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, float seed, long long* clk) {
    float r[8]; unsigned long long q[8];
    for (int i = 0; i < 8; ++i) { r[i] = seed + threadIdx.x * 0.001f + i; q[i] = ((unsigned long long)__float_as_uint(r[i]) << 32) | __float_as_uint(r[i] + 1.f); }
    float c = seed * 0.5f;
    unsigned long long cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c));
                if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(cc));
                if (OP == 2) asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c));
                if (OP == 3) asm volatile("min.abs.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(c), "f"(seed));
                if (OP == 4) { unsigned x = __float_as_uint(r[i]); asm volatile("min.xorsign.abs.f16x2 %0, %0, %1;" : "+r"(x) : "r"(__float_as_uint(c))); r[i] = __uint_as_float(x); }
                if (OP == 5) { if (i & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c)); else asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c)); }
                if (OP == 6) { if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(cc)); else asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c)); }
                if (OP == 7) { unsigned x = __float_as_uint(r[i]); asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x) : "r"(__float_as_uint(c))); r[i] = __uint_as_float(x); }
                if (OP == 8) { if ((i & 3) == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c)); else asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c)); }
                if (OP == 9) { int x = __float_as_int(r[i]); asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(__float_as_int(c)), "r"(it)); r[i] = __int_as_float(x); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += r[i] + __uint_as_float((unsigned)(q[i] >> 32)) + __uint_as_float((unsigned)q[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int OP> void run(const char* name, float* out, long long* clk) {
    for (int warps = 1; warps <= 32; warps *= 2) {
        if (warps != 4 && warps != 8 && warps != 16 && warps != 32) continue;
        k<OP><<<148, warps * 32>>>(out, 1.5f, clk);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        double instr = (double)N_IT * 32 * warps;   // warp-instructions per SM
        printf("%-28s warps/SM %2d: %.3f warp-instr/clk/SM (%.3f per SMSP)\n", name, warps, instr / h, instr / h / 4);
    }
}
int main() {
    float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
    run<0>("FADD", out, clk); run<1>("FADD2", out, clk); run<2>("FMNMX.XORSIGN", out, clk); run<3>("FMNMX3.abs", out, clk);
    run<4>("HMNMX2.XORSIGN", out, clk); run<7>("HADD2", out, clk); run<5>("FADD+FMNMX 1:1", out, clk); run<6>("FADD2+FMNMX 1:1", out, clk); run<8>("FADD+FMNMX 1:3", out, clk); run<9>("IMAD.HI", out, clk);
    return 0;
}
