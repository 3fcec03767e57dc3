// epilogue.cuh - shared tail of the decoder kernels: bit packing of the hard decision and
// the fused exact integer link metrics (replaces evaluate_quantized_snr.py:169-188).
#pragma once
#include "common.cuh"

namespace ldpc {

// hard_s: [ncw][hs_stride] u8 in shared memory, bit 0 = decoded bit, bit 1 = uncoded channel
// decision (llr > 0); packed output MSB-first (numpy.packbits).
__device__ __forceinline__ void pack_hard(const uint8_t *hard_s, int hs_stride, int ncw, int n,
                                          uint8_t *packed_g /* [ncw][ceil(n/8)] */) {
    const int nbytes = (n + 7) >> 3;
    // fast path (8-byte aligned rows): one 64-bit load per output byte, bits 0 of the eight bytes gathered
    // MSB-first by one multiplication (bit 8b of y lands on bit 63 - b, no two terms share a position)
    const bool wide = ((reinterpret_cast<uintptr_t>(hard_s) | (uintptr_t)hs_stride) & 7u) == 0;
    const int nfull = wide ? (n >> 3) : 0;
    for (int i = threadIdx.x; i < ncw * nbytes; i += blockDim.x) {
        const int cw = i / nbytes, by = i - cw * nbytes;
        const uint8_t *h = hard_s + cw * hs_stride + by * 8;
        unsigned v = 0;
        if (by < nfull) {
            const unsigned long long y = *reinterpret_cast<const unsigned long long *>(h) & 0x0101010101010101ull;
            v = (unsigned)((y * 0x8040201008040201ull) >> 56);
        } else {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int idx = by * 8 + b;
                v |= (idx < n ? (unsigned)(h[b] & 1) : 0u) << (7 - b);
            }
        }
        packed_g[(long long)cw * nbytes + by] = (uint8_t)v;
    }
}

// counters: {uncoded bit errors, info-bit errors, frame errors, bits, frames} (u64, +=).
// Uncoded decision = bit 1 of hard_s (llr > 0 -> 1, llr == 0 -> 0, i.e. (sign+1)//2);
// ref_packed_g: transmitted codewords, MSB-first.
// scratch: 3 ints + ncw ints of shared memory, zeroed by the caller before a barrier.
// SELF_REF (single-launch simulator): the transmitted bit is bit 2 of the shared byte itself and ref_packed_g is null.
template <bool SELF_REF = false>
__device__ __forceinline__ void count_errors(const uint8_t *hard_s, int hs_stride, int ncw, int n, int k_info,
                                             const uint8_t *ref_packed_g, unsigned long long *counters,
                                             int *scratch /* [3 + ncw] */) {
    const int nbytes = (n + 7) >> 3;
    int unc = 0, inf = 0;
    if (SELF_REF && (n & 3) == 0 && (hs_stride & 3) == 0) {
        // single-launch simulator: decision (bit 0), uncoded decision (bit 1) and transmitted bit (bit 2) share a byte -
        // four code bits per 32-bit word, compared with three shifts and counted with popc
        const int words = n >> 2;
        for (int cw = 0; cw < ncw; ++cw) {
            const uint32_t *row = reinterpret_cast<const uint32_t *>(hard_s + cw * hs_stride);
            unsigned any = 0;
            for (int w = threadIdx.x; w < words; w += blockDim.x) {
                const uint32_t hv = row[w], ref = (hv >> 2) & 0x01010101u;
                const uint32_t e = (hv ^ ref) & 0x01010101u;                   // decoded bit != transmitted bit
                unc += __popc(((hv >> 1) ^ ref) & 0x01010101u);
                const int left = k_info - 4 * w;                               // information bits in this word: all four, none, or the first `left`
                inf += __popc(left >= 4 ? e : (left <= 0 ? 0u : (e & (0xffffffffu >> (32 - 8 * left)))));
                any |= e;
            }
            if (any) scratch[3 + cw] = 1;                                      // benign race: everyone writes 1
        }
    } else
    for (int i = threadIdx.x; i < ncw * n; i += blockDim.x) {
        const int cw = i / n, v = i - cw * n;
        const int hv = hard_s[cw * hs_stride + v];
        // transmitted bit: packed global array, or bit 2 of the shared byte (single-launch simulator)
        const int ref = ref_packed_g ? (ref_packed_g[(long long)cw * nbytes + (v >> 3)] >> (7 - (v & 7))) & 1 : (hv >> 2) & 1;
        const int hb = hv & 1, ub = (hv >> 1) & 1;
        unc += (ub != ref);
        const int e = (hb != ref);
        inf += e & (v < k_info);
        if (e) scratch[3 + cw] = 1;            // benign race: everyone writes 1
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unc += __shfl_xor_sync(0xffffffffu, unc, o);
        inf += __shfl_xor_sync(0xffffffffu, inf, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (unc) atomicAdd(&scratch[0], unc);
        if (inf) atomicAdd(&scratch[1], inf);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int fe = 0;
        for (int c = 0; c < ncw; ++c) fe += scratch[3 + c];
        if (scratch[0]) atomicAdd(&counters[0], (unsigned long long)scratch[0]);
        if (scratch[1]) atomicAdd(&counters[1], (unsigned long long)scratch[1]);
        if (fe) atomicAdd(&counters[2], (unsigned long long)fe);
        atomicAdd(&counters[3], (unsigned long long)ncw * (unsigned long long)n);
        atomicAdd(&counters[4], (unsigned long long)ncw);
    }
}

}  // namespace ldpc
