// decode_qc_rt.cu - belief-propagation decoder for ANY quasi-cyclic code (circulant weight <= 1), prototype
// matrix given at RUN time (ldpc_code_create with qc_Z / qc_proto and no compiled specialisation): the other
// 802.11n lengths and rates, 5G-style base graphs, anything loaded from an alist / .mat file.
//
// Same mapping as the compiled kernel (decode_qc.cu): thread = (lane z of Z, codeword of the CTA tile), codewords
// interleaved by lane, one fp32 slot per edge in shared memory at (block * Z + check lane) * CW + cw, so the check
// phase is a linear access and the variable phase a rotated window that wraps once per block and CTA.  The
// difference: the block lists (which blocks a block row / column holds, their shifts) are small tables in shared
// memory instead of immediates, every message and the channel LLRs live in shared memory (no register-resident
// blocks), and the node degrees are run-time values below a compile-time cap.
// Arithmetic = node_math.cuh with the generic kernel's edge order (a variable's edges ascending in the check index,
// a check's edges ascending in the variable index), so the results are bit-identical to decode_generic.cu.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "epilogue.cuh"
#include "node_math.cuh"

namespace ldpc {

// volatile: never moved across a barrier or each other; "memory" on the store orders it against the plain accesses
__device__ __forceinline__ float lds_f32(int addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(int addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

struct QcRtParams {
    DecodeArgs a;
    int Z, MB, NB, nblk, CW;
    int hard_stride;
    const int32_t *tab;        // device: row_ptr[MB+1] | row_blk[nblk] | col_ptr[NB+1] | col_blk[nblk] | blk_shift[nblk] | blk_col[nblk]
};

template <int MAXDV, int MAXDC, int UPD>
__global__ void __launch_bounds__(512) decode_qc_rt_kernel(const QcRtParams p) {
    constexpr bool IS_SP = (UPD == UPD_SP);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DecodeArgs &a = p.a;
    const int Z = p.Z, MB = p.MB, NB = p.NB, nblk = p.nblk, CW = p.CW, n = NB * Z;
    const int ntab = (MB + 1) + nblk + (NB + 1) + 3 * nblk;
    int32_t *tab_s = reinterpret_cast<int32_t *>(smem_raw);
    const int32_t *row_ptr = tab_s, *row_blk = row_ptr + MB + 1, *col_ptr = row_blk + nblk, *col_blk = col_ptr + NB + 1,
                  *blk_shift = col_blk + nblk, *blk_col = blk_shift + nblk;
    // per-CTA derived tables (word offsets for this CW): a block's base, and for the rotated access base - shift and the
    // wrap threshold, so an edge address is one add and one select on the thread index
    // (shared-window BYTE addresses, so that an edge access is one table load, one add and - in the variable phase - one
    // select on the thread index; the accesses below are ld/st.shared on those addresses)
    int2 *col_tab = reinterpret_cast<int2 *>(tab_s + ((ntab + 3) & ~3));          // {address of block - shift, wrap threshold}
    int32_t *row_base = reinterpret_cast<int32_t *>(col_tab + nblk);
    float *msg = reinterpret_cast<float *>(row_base + ((nblk + 3) & ~3));         // [nblk][Z][CW]
    float *llr_s = msg + (size_t)nblk * Z * CW;                                    // [NB][Z][CW]
    uint8_t *hard_s = reinterpret_cast<uint8_t *>(llr_s + (size_t)n * CW);        // [CW][hard_stride]
    int *scratch = reinterpret_cast<int *>(hard_s + (size_t)CW * p.hard_stride);  // [4 + CW]

    const int tid = threadIdx.x, T = blockDim.x;
    const long long cw0 = (long long)blockIdx.x * CW;
    const int ncw = (int)min((long long)CW, a.B - cw0);
    const int z = tid / CW, cw = tid - z * CW;
    const bool active = z < Z && cw < ncw;
    const int tid4 = tid * 4, ZCW4 = Z * CW * 4;
    const uint32_t msg_a = (uint32_t)__cvta_generic_to_shared(msg);

    for (int i = tid; i < ntab; i += T) tab_s[i] = __ldg(p.tab + i);
    {
        const int32_t *g_row_blk = p.tab + MB + 1, *g_col_blk = g_row_blk + nblk + NB + 1, *g_shift = g_col_blk + nblk;
        for (int i = tid; i < nblk; i += T) {
            row_base[i] = msg_a + __ldg(g_row_blk + i) * ZCW4;
            const int b = __ldg(g_col_blk + i), sh = __ldg(g_shift + b);
            col_tab[i] = make_int2((int)(msg_a + b * ZCW4 - sh * CW * 4), sh * CW);
        }
    }
    for (int i = tid; i < nblk * Z * CW; i += T) msg[i] = 0.0f;                   // the zeros every reference caller passes
    for (int i = tid; i < 4 + CW; i += T) scratch[i] = 0;
    if (active)
        for (int c = 0; c < NB; ++c) llr_s[(c * Z + z) * CW + cw] = load_llr(a.llr, a.llr_dtype, (cw0 + cw) * n + c * Z + z);
    __syncthreads();

    for (int it = 0; it < a.iters; ++it) {
        // ---- V -> C: variable (c, z); its edge in block b = (r, c, s) is check lane (z - s) mod Z ----------------
        if (active) {
#pragma unroll 1
            for (int c = 0; c < NB; ++c) {
                const int b0 = col_ptr[c], d = col_ptr[c + 1] - b0;
                if (d == 0) continue;
                const float l = llr_s[(c * Z + z) * CW + cw];
                degree_switch<1, MAXDV>(d, [&](auto dd) {
                    constexpr int D = decltype(dd)::value;
                    int slot[D];
                    float in[D], out[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const int2 ct = col_tab[b0 + k];
                        slot[k] = ct.x + tid4 + (tid < ct.y ? ZCW4 : 0);
                        in[k] = lds_f32(slot[k]);
                    }
                    var_node<D, IS_SP>(in, D, l, out);
#pragma unroll
                    for (int k = 0; k < D; ++k) sts_f32(slot[k], out[k]);
                });
            }
        }
        __syncthreads();
        // ---- C -> V: check (r, z): linear access ---------------------------------------------------------------
        if (active) {
#pragma unroll 1
            for (int r = 0; r < MB; ++r) {
                const int b0 = row_ptr[r], d = row_ptr[r + 1] - b0;
                if (d == 0) continue;
                degree_switch<1, MAXDC>(d, [&](auto dd) {
                    constexpr int D = decltype(dd)::value;
                    int slot[D];
                    float in[D], out[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        slot[j] = row_base[b0 + j] + tid4;
                        in[j] = lds_f32(slot[j]);
                    }
                    if constexpr (IS_SP) check_node_sp<D>(in, D, a.clampv, out);
                    else check_node_ms_ct<D, UPD>(in, a.clampv, a.param, out);      // box-min tree: same bits as check_node_ms
#pragma unroll
                    for (int j = 0; j < D; ++j) sts_f32(slot[j], out[j]);
                });
            }
        }
        __syncthreads();
    }

    // ---- marginal, P(bit=1), hard decision --------------------------------------------------------------------------
    if (active) {
#pragma unroll 1
        for (int c = 0; c < NB; ++c) {
            const int b0 = col_ptr[c], d = col_ptr[c + 1] - b0;
            const float l = llr_s[(c * Z + z) * CW + cw];
            float t = __fmul_rn(0.5f, __fadd_rn(-l, 0.0f));
            degree_switch<1, MAXDV>(d, [&](auto dd) {
                constexpr int D = decltype(dd)::value;
                float in[D];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int2 ct = col_tab[b0 + k];
                    in[k] = lds_f32(ct.x + tid4 + (tid < ct.y ? ZCW4 : 0));
                }
                t = marginal_t<D>(in, D, l);
            });
            const uint8_t hb = hard_bit(t);
            hard_s[cw * p.hard_stride + c * Z + z] = hb | ((l > 0.0f) ? 2 : 0);
            const long long o = (cw0 + cw) * n + c * Z + z;
            if (a.prob) a.prob[o] = prob_one(t);
            if (a.llr_post) a.llr_post[o] = __fmul_rn(-2.0f, t);
            if (a.hard) a.hard[o] = hb;
        }
    }
    __syncthreads();
    if (a.syndrome) {
        if (active) {
            int w = 0;
            const uint8_t *h = hard_s + cw * p.hard_stride;
#pragma unroll 1
            for (int r = 0; r < MB; ++r) {
                unsigned par = 0;
                for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
                    const int b = row_blk[e];
                    int zv = z + blk_shift[b];                      // check lane z touches variable lane (z + s) mod Z
                    if (zv >= Z) zv -= Z;
                    par ^= h[blk_col[b] * Z + zv] & 1u;
                }
                w += (int)par;
            }
            if (w) atomicAdd(&scratch[4 + cw], w);
        }
        __syncthreads();
        for (int i = tid; i < ncw; i += T) a.syndrome[cw0 + i] = scratch[4 + i];
        __syncthreads();
        for (int i = tid; i < CW; i += T) scratch[4 + i] = 0;
    }
    if (a.iters_used)
        for (int i = tid; i < ncw; i += T) a.iters_used[cw0 + i] = a.iters;
    if (a.hard_packed) pack_hard(hard_s, p.hard_stride, ncw, n, a.hard_packed + cw0 * ((n + 7) >> 3));
    if (a.counters) {
        __syncthreads();
        count_errors(hard_s, p.hard_stride, ncw, n, a.k_info, a.ref_packed + cw0 * ((n + 7) >> 3), a.counters, scratch + 1);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------
// Builds the device tables of a prototype matrix; returns the number of int32 words (0 on failure).
int qc_rt_build_tables(int Z, int mb, int nb, const int16_t *proto, std::vector<int32_t> &out, int *max_dv, int *max_dc) {
    std::vector<int> blk_row, blk_col, blk_shift;
    std::vector<int32_t> row_ptr(mb + 1, 0), col_ptr(nb + 1, 0);
    for (int r = 0; r < mb; ++r) {
        for (int c = 0; c < nb; ++c)
            if (proto[r * nb + c] >= 0) { blk_row.push_back(r); blk_col.push_back(c); blk_shift.push_back(proto[r * nb + c] % Z); }
        row_ptr[r + 1] = (int32_t)blk_row.size();
    }
    const int nblk = (int)blk_row.size();
    std::vector<int32_t> row_blk(nblk), col_blk;
    for (int b = 0; b < nblk; ++b) row_blk[b] = b;                    // check-major enumeration: ascending column inside a row
    *max_dc = 0; *max_dv = 0;
    for (int r = 0; r < mb; ++r) *max_dc = std::max(*max_dc, row_ptr[r + 1] - row_ptr[r]);
    for (int c = 0; c < nb; ++c) {
        for (int b = 0; b < nblk; ++b)
            if (blk_col[b] == c) col_blk.push_back(b);               // ascending block row = ascending check index
        col_ptr[c + 1] = (int32_t)col_blk.size();
        *max_dv = std::max(*max_dv, col_ptr[c + 1] - col_ptr[c]);
    }
    out.clear();
    out.insert(out.end(), row_ptr.begin(), row_ptr.end());
    out.insert(out.end(), row_blk.begin(), row_blk.end());
    out.insert(out.end(), col_ptr.begin(), col_ptr.end());
    out.insert(out.end(), col_blk.begin(), col_blk.end());
    out.insert(out.end(), blk_shift.begin(), blk_shift.end());
    out.insert(out.end(), blk_col.begin(), blk_col.end());
    return (int)out.size();
}

static size_t qc_rt_smem(int Z, int MB, int NB, int nblk, int CW) {
    const int ntab = (MB + 1) + nblk + (NB + 1) + 3 * nblk;
    const int hard_stride = (NB * Z + 15) & ~15;
    return (size_t)(((ntab + 3) & ~3) + 2 * nblk + ((nblk + 3) & ~3)) * 4 + sizeof(float) * ((size_t)nblk * Z * CW + (size_t)NB * Z * CW) + (size_t)CW * hard_stride +
           sizeof(int) * (8 + 2 * CW);
}

bool qc_rt_supported(int Z, int mb, int nb, int nblk, int max_dv, int max_dc) {
    return Z <= 512 && max_dv <= 16 && max_dc <= 32 && qc_rt_smem(Z, mb, nb, nblk, 1) <= 220 * 1024;
}

template <int MAXDV, int MAXDC>
static int launch_rt_t(const QcRtParams &p, size_t smem, int threads, int grid, cudaStream_t s) {
    void (*k)(const QcRtParams) = nullptr;
    switch (p.a.update) {
        case UPD_SP: k = decode_qc_rt_kernel<MAXDV, MAXDC, UPD_SP>; break;
        case UPD_MINSUM: k = decode_qc_rt_kernel<MAXDV, MAXDC, UPD_MINSUM>; break;
        case UPD_NMS: k = decode_qc_rt_kernel<MAXDV, MAXDC, UPD_NMS>; break;
        default: k = decode_qc_rt_kernel<MAXDV, MAXDC, UPD_OMS>; break;
    }
    LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, threads, smem, s>>>(p);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int launch_decode_qc_rt(const int32_t *d_tab, int Z, int mb, int nb, int nblk, int max_dv, int max_dc, const DecodeArgs &a, cudaStream_t s) {
    if (a.B <= 0) return LDPC_OK;
    QcRtParams p;
    p.a = a; p.Z = Z; p.MB = mb; p.NB = nb; p.nblk = nblk; p.tab = d_tab;
    p.hard_stride = (nb * Z + 15) & ~15;
    // codewords per CTA: the tile that keeps the most codewords resident per SM (228 KB of shared memory, 1 KB reserved
    // per CTA, 2048 threads); occupancy is what this kernel lives on (three CTAs instead of two: +20 %)
    int CW = 1, best = 0;
    for (int c = 1; c <= 8 && c * Z <= 512; ++c) {
        const size_t sm = qc_rt_smem(Z, mb, nb, nblk, c);
        if (sm > 220 * 1024) break;
        const int thr = ((c * Z + 31) / 32) * 32;
        const int ctas = std::min<int>({(int)((228 * 1024) / (sm + 1024)), 2048 / thr, 32});
        if (ctas * c >= best && ctas > 0) { best = ctas * c; CW = c; }
    }
    if (qc_rt_smem(Z, mb, nb, nblk, CW) > 220 * 1024) { set_error("code too large for the run-time QC kernel"); return LDPC_EUNSUPPORTED; }
    p.CW = CW;
    const int threads = ((CW * Z + 31) / 32) * 32;
    const long long grid = (a.B + CW - 1) / CW;
    if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
    const size_t smem = qc_rt_smem(Z, mb, nb, nblk, CW);
    if (max_dv <= 4 && max_dc <= 8) return launch_rt_t<4, 8>(p, smem, threads, (int)grid, s);
    if (max_dv <= 12 && max_dc <= 8) return launch_rt_t<12, 8>(p, smem, threads, (int)grid, s);
    if (max_dv <= 12 && max_dc <= 24) return launch_rt_t<12, 24>(p, smem, threads, (int)grid, s);
    return launch_rt_t<16, 32>(p, smem, threads, (int)grid, s);
}

}  // namespace ldpc
