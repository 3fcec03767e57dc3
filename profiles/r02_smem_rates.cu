// r02_smem_rates.cu - shared-memory INSTRUCTION throughput on B200 (sm_100a): is a kernel that moves one 32-bit word per
// LDS/STS limited by bytes (128 B/clk/SM = one wavefront per clock) or by the number of LSU instructions?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_rates profiles/r02_smem_rates.cu && ./smem_rates
// Every access is conflict-free (consecutive lanes -> consecutive words / 8-byte / 16-byte units).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float *out, int iters, long long *clk) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 16384; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const unsigned base0 = (unsigned)__cvta_generic_to_shared(sm);
    float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, bcc0 = 0, bcc1 = 0, bcc2 = 0, bcc3 = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // loop-variant address (another 16-byte-aligned window every iteration) and 16 DIFFERENT offsets in the unrolled body:
        // a first version revisited 8 / 4 / 2 offsets per iteration and nvcc merged the repeated `asm volatile` loads (8 LDS
        // feeding 16 FADD in the SASS), which read as 2.0 - 3.7 "loads"/clk - a compiler artefact, not bandwidth
        const unsigned base = base0 + (((unsigned)it * 2064u) & 0x3ff0u);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) {        // LDS.32
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                acc0 += v;
            } else if (MODE == 1) { // LDS.64
                float v0, v1;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v0), "=f"(v1) : "r"(base + (unsigned)(tid * 8 + u * 2048)) : "memory");
                acc0 += v0; acc1 += v1;
            } else if (MODE == 2) { // LDS.128
                float v0, v1, v2, v3;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(base + (unsigned)(tid * 16 + u * 1024)) : "memory");
                acc0 += v0; acc1 += v1; acc2 += v2; acc3 += v3;
            } else if (MODE == 3) { // STS.32
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + u * 2048)), "f"(acc0 + (float)u) : "memory");
            } else if (MODE == 4) { // STS.64
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(base + (unsigned)(tid * 8 + u * 2048)), "f"(acc0 + (float)u), "f"(acc1) : "memory");
            } else if (MODE == 5) { // LDS.32 + STS.32 alternating (the decoder's mix)
                if (u & 1) {
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + u * 2048)), "f"(acc0) : "memory");
                } else {
                    float v;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                    acc1 += v;
                }
            } else if (MODE == 6) { // LDS.64 + STS.64 alternating
                if (u & 1) {
                    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(base + (unsigned)(tid * 8 + u * 2048)), "f"(acc0), "f"(acc1) : "memory");
                } else {
                    float v0, v1;
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v0), "=f"(v1) : "r"(base + (unsigned)(tid * 8 + u * 2048)) : "memory");
                    acc2 += v0; acc3 += v1;
                }
            } else if (MODE == 7) { // LDS.32 + 2 FMNMX (check-phase mix)
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc0) : "f"(v));
                asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc1) : "f"(v));
            } else if (MODE == 9 || MODE == 10) { // the same mix, stores depend on the loaded values through the min chain (as in a check node), + barriers
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                float m0, m1, m2;
                asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m0) : "f"(acc0), "f"(v));
                asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m1) : "f"(m0), "f"(acc1));
                if (u & 1) asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m2) : "f"(m1), "f"(acc2)); else m2 = m1;
                acc3 = __fadd_rn(acc3, v);
                acc0 = __fadd_rn(m2, acc3);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + (u ^ 5) * 2048)), "f"(m2) : "memory");
            } else if (MODE == 15 || MODE == 16) {
                // the decoder's two phases with their real per-exchanged-edge ratios: check-like unit = LDS + 5 dependent FMNMX + STS
                // (248 / 51), variable-like unit = LDS + 4 FADD + ISETP/SEL + STS.  15: PHASED - a whole loop iteration of one kind,
                // 8-warp barrier, then the other (16 warps = two groups, as the kernel).  16: both units INTERLEAVED in one thread.
                const bool do_check = (MODE == 16) || !(it & 1), do_var = (MODE == 16) || (it & 1);
                if (do_check) {
                    float v;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                    float m0, m1, m2, m3, m4;
                    asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m0) : "f"(acc0), "f"(v));
                    asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m1) : "f"(acc1), "f"(v));
                    asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m2) : "f"(m0), "f"(acc2));
                    asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m3) : "f"(m1), "f"(m2));
                    asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m4) : "f"(m3), "f"(acc3));
                    acc0 = m1; acc1 = m2; acc2 = m4; acc3 = v;
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + (u ^ 5) * 2048)), "f"(m4) : "memory");
                }
                if (do_var) {
                    float w;
                    const unsigned sel = (tid < 7 * u + 3) ? 12u : 0u;                  // the rotated-window select
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(base + sel + (unsigned)(tid * 4 + (u ^ 3) * 2048)) : "memory");
                    bcc0 = __fadd_rn(bcc0, w);
                    bcc1 = __fadd_rn(bcc1, bcc0);
                    bcc2 = __fadd_rn(bcc2, w);
                    bcc3 = __fadd_rn(bcc1, bcc2);
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + sel + (unsigned)(tid * 4 + (u ^ 6) * 4096 % 32768)), "f"(bcc3) : "memory");
                }
            } else if (MODE == 14) { // TWO independent dependent-chain edge mixes per thread, interleaved (intra-thread anti-phase probe)
                float v, w;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(base + (unsigned)(tid * 4 + (u ^ 3) * 2048)) : "memory");
                float m0, m1, m2;
                asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m0) : "f"(acc0), "f"(v));
                asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m1) : "f"(m0), "f"(acc1));
                if (u & 1) asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(m2) : "f"(m1), "f"(acc2)); else m2 = m1;
                // second stream: the variable-node side (adds only), independent of the first
                bcc0 = __fadd_rn(bcc0, w);
                bcc1 = __fadd_rn(bcc1, bcc0);
                bcc2 = __fadd_rn(bcc2, w);
                bcc3 = __fadd_rn(bcc1, bcc2);
                acc3 = __fadd_rn(acc3, v);
                acc0 = __fadd_rn(m2, acc3);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + (u ^ 5) * 2048)), "f"(m2) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + (u ^ 6) * 4096 % 32768)), "f"(bcc3) : "memory");
            } else if (MODE == 11 || MODE == 12) { // bursts: a whole loop iteration of loads, then one of stores (12: CTA-wide barrier between them)
                if (it & 1) {
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + u * 2048)), "f"(acc0) : "memory");
                } else {
                    float v;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                    acc1 += v;
                }
            } else if (MODE == 13) { // 3 loads : 1 store
                if ((u & 3) == 3) {
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + u * 2048)), "f"(acc0) : "memory");
                } else {
                    float v;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                    acc1 += v;
                }
            } else if (MODE == 8) { // the decoder's per-exchanged-edge mix: LDS + 2.5 FMNMX + 2 FADD + STS (ALU-pipe share 0.38 of the instructions)
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (unsigned)(tid * 4 + u * 2048)) : "memory");
                asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc0) : "f"(v));
                asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc1) : "f"(v));
                if (u & 1) asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(acc2) : "f"(v));
                acc3 = __fadd_rn(acc3, v);
                acc2 = __fadd_rn(acc2, acc3);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + (unsigned)(tid * 4 + (u ^ 5) * 2048)), "f"(acc2) : "memory");
            }
        }
        if (MODE == 12) __syncthreads();
        if (MODE == 15 || MODE == 16) asm volatile("bar.sync %0, 256;" ::"r"(1 + tid / 256) : "memory");
        if (MODE == 10 && (it & 1)) asm volatile("bar.sync %0, 256;" ::"r"(1 + tid / 256) : "memory");   // 8-warp groups, every ~230 instructions
    }
    const long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + tid] = acc0 + acc1 + acc2 + acc3 + bcc3;
}

template <int MODE>
static void run(const char *name, int wf_per_instr) {
    float *out;
    long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&clk, 148 * 8);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    printf("%-36s", name);
    for (int warps : {8, 16, 32}) {
        const int iters = 2000;
        k<MODE><<<148, warps * 32, 65536>>>(out, iters, clk);
        k<MODE><<<148, warps * 32, 65536>>>(out, iters, clk);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < 148; ++i) mean += h[i];
        mean /= 148;
        const double instr = (double)iters * 16 * warps;
        printf("  %2dw: %.3f instr/clk %.3f wf/clk", warps, instr / mean, instr * wf_per_instr / mean);
    }
    printf("\n");
}

int main() {
    printf("# shared-memory instruction rate per SM (148 CTAs, one per SM), conflict-free; wf = 128-byte wavefronts\n");
    run<0>("LDS.32", 1);
    run<1>("LDS.64", 2);
    run<2>("LDS.128", 4);
    run<3>("STS.32", 1);
    run<4>("STS.64", 2);
    run<5>("LDS.32 + STS.32 1:1", 1);
    run<6>("LDS.64 + STS.64 1:1", 2);
    run<7>("LDS.32 + 2 FMNMX.XORSIGN (instr = LDS)", 1);
    run<11>("16 LDS.32 then 16 STS.32 per warp", 1);
    run<12>("16 LDS.32 | barrier | 16 STS.32 | barrier", 1);
    run<13>("LDS.32 : STS.32 3:1", 1);
    run<15>("PHASED check-like / variable-like iterations, 8-warp barriers (instr = LDS+STS)", 2);
    run<16>("INTERLEAVED check-like + variable-like in one thread (instr = 2 LDS + 2 STS)", 4);
    run<14>("two interleaved streams/thread: check-like + var-like (instr = 2 LDS + 2 STS)", 4);
    run<9>("edge mix, dependent chain (instr = LDS+STS)", 2);
    run<10>("edge mix, dependent chain + 8-warp barriers", 2);
    run<8>("edge mix LDS+2.5FMNMX+2FADD+STS (instr = LDS+STS)", 2);
    return 0;
}
