// decode_generic.cu - belief-propagation decoder for an ARBITRARY parity-check matrix.
//
// One CTA owns a tile of CW codewords.  Their edge messages (one fp32 slot per edge,
// check-major slot order) and channel LLRs stay resident in shared memory for all
// iterations; HBM traffic is the LLR load and the final posterior / hard-bit store only.
// The Tanner graph comes from CSR/CSC edge tables in global memory (read-only, L1/L2
// resident) - this replaces the reference's dense E x E masks (bp/masking.py:12-147) and
// its [B,E,E] masked product (bp/bp_cv.py:24-42).
//
// Per iteration: variable phase (thread per (codeword, variable): gather the dv slots of
// the variable, write the V->C values back IN PLACE), barrier, check phase (thread per
// (codeword, check): its dc slots are contiguous), barrier.  Flooding schedule, fixed
// iteration count (bp/bp.py:46-47).  Arithmetic: node_math.cuh.
#include "common.cuh"
#include "epilogue.cuh"
#include "node_math.cuh"

namespace ldpc {

struct GenericParams {
    GraphTables g;
    DecodeArgs a;
    int CW;            // codewords per CTA
    int llr_stride;    // floats per codeword in the llr tile
    int msg_stride;    // floats per codeword in the message tile
    int hard_stride;   // bytes per codeword in the hard-bit tile
};

template <int MAXDV, int MAXDC, bool IS_SP, bool SW>
__global__ void __launch_bounds__(256) decode_generic_kernel(const GenericParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphTables &g = p.g;
    const DecodeArgs &a = p.a;
    const int n = g.n, m = g.m, E = g.E;
    const long long cw0 = (long long)blockIdx.x * p.CW;
    const int ncw = (int)min((long long)p.CW, a.B - cw0);
    float *llr_s = reinterpret_cast<float *>(smem_raw);
    float *msg = llr_s + p.CW * p.llr_stride;
    uint8_t *hard_s = reinterpret_cast<uint8_t *>(msg + p.CW * p.msg_stride);
    int *scratch = reinterpret_cast<int *>(hard_s + p.CW * p.hard_stride);   // [4 + CW]
    const int tid = threadIdx.x, T = blockDim.x;

    // ---- load tile: LLRs (any dtype -> f32), initial messages ------------------------------
    for (int i = tid; i < ncw * n; i += T) {
        const int cw = i / n, v = i - cw * n;
        llr_s[cw * p.llr_stride + v] = load_llr(a.llr, a.llr_dtype, cw0 * n + i);
    }
    for (int i = tid; i < ncw * E; i += T) {
        const int cw = i / E, e = i - cw * E;
        msg[cw * p.msg_stride + e] = a.x0 ? __ldg(a.x0 + cw0 * E + i) : 0.0f;
    }
    for (int i = tid; i < 4 + p.CW; i += T) scratch[i] = 0;
    __syncthreads();

    int *frozen_s = scratch + 4 + p.CW;                                  // [CW] iteration of convergence, 0 = running
    for (int i = tid; i < p.CW; i += T) frozen_s[i] = 0;
    __syncthreads();
    bool finished = false;
    int it = 0;
    for (;;) {
        const bool do_iter = (it < a.iters) && !finished;
        if (do_iter) {
            // ---- V -> C -----------------------------------------------------------------------
            for (int i = tid; i < ncw * n; i += T) {
                const int cw = i / n, v = i - cw * n;
                if (a.early_exit && frozen_s[cw]) continue;
                const int b = __ldg(g.var_ptr + v), d = __ldg(g.var_ptr + v + 1) - b;
                if (d == 0) continue;
                float *mrow = msg + cw * p.msg_stride;
                with_degree<MAXDV, SW>(d, [&](auto cap, int d) {          // compile-time degree: register arrays
                    constexpr int D = decltype(cap)::value;
                    int slot[D];
                    float in[D], out[D];
#pragma unroll
                    for (int k = 0; k < D; ++k)
                        if (k < d) { slot[k] = __ldg(g.cm_of_vm + b + k); in[k] = mrow[slot[k]]; }
                    if (a.w_edge)
                        var_node_weighted<D, IS_SP>(in, d, llr_s[cw * p.llr_stride + v], __ldg(a.w_llr + (long long)it * n + v),
                                                    a.w_edge + ((long long)it * E + b) * a.w_stride, a.w_stride, out);
                    else
                        var_node<D, IS_SP>(in, d, llr_s[cw * p.llr_stride + v], out);
#pragma unroll
                    for (int k = 0; k < D; ++k)
                        if (k < d) mrow[slot[k]] = out[k];
                });
            }
            __syncthreads();
            // ---- C -> V -----------------------------------------------------------------------
            for (int i = tid; i < ncw * m; i += T) {
                const int cw = i / m, c = i - cw * m;
                if (a.early_exit && frozen_s[cw]) continue;
                const int b = __ldg(g.chk_ptr + c), d = __ldg(g.chk_ptr + c + 1) - b;
                if (d == 0) continue;
                float *mrow = msg + cw * p.msg_stride + b;
                with_degree<MAXDC, SW>(d, [&](auto cap, int d) {
                    constexpr int D = decltype(cap)::value;
                    float in[D], out[D];
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j < d) in[j] = mrow[j];
                    if (IS_SP) check_node_sp<D>(in, d, a.clampv, out);
                    else check_node_ms<D>(in, d, a.update, a.clampv, a.param, out);
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j < d) mrow[j] = out[j];
                });
            }
            __syncthreads();
            ++it;
        }
        const bool last = !do_iter || it >= a.iters;
        if (!last && !a.early_exit) continue;

        // ---- marginal, P(bit=1), hard decision (outputs only on the last pass) ---------------------
        for (int i = tid; i < ncw * n; i += T) {
            const int cw = i / n, v = i - cw * n;
            if (!last && frozen_s[cw]) continue;
            const int b = __ldg(g.var_ptr + v), d = __ldg(g.var_ptr + v + 1) - b;
            const float *mrow = msg + cw * p.msg_stride;
            float t = __fmul_rn(0.5f, __fadd_rn(-llr_s[cw * p.llr_stride + v], 0.0f));      // a variable without edges
            if (a.wf_edge && d == 0) t = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(__ldg(a.wf_llr + v), -llr_s[cw * p.llr_stride + v]), 0.0f));
            with_degree<MAXDV, SW>(d, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                float in[D];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) in[k] = mrow[__ldg(g.cm_of_vm + b + k)];
                t = a.wf_edge ? marginal_t_weighted<D>(in, d, llr_s[cw * p.llr_stride + v], __ldg(a.wf_llr + v), a.wf_edge + b)
                              : marginal_t<D>(in, d, llr_s[cw * p.llr_stride + v]);
            });
            const uint8_t hb = hard_bit(t);                  // np.round(prob): tie 0.5 -> 0
            hard_s[cw * p.hard_stride + v] = hb | ((llr_s[cw * p.llr_stride + v] > 0.0f) ? 2 : 0);
            if (last) {
                const long long o = cw0 * n + i;
                if (a.prob) a.prob[o] = prob_one(t);
                if (a.llr_post) a.llr_post[o] = __fmul_rn(-2.0f, t);
                if (a.hard) a.hard[o] = hb;
            }
        }
        __syncthreads();
        // ---- syndrome weight of the hard decision ---------------------------------------------------
        if (!last || a.syndrome) {
            for (int i = tid; i < ncw * m; i += T) {
                const int cw = i / m, c = i - cw * m;
                if (!last && frozen_s[cw]) continue;
                unsigned par = 0;
                for (int e = __ldg(g.chk_ptr + c); e < __ldg(g.chk_ptr + c + 1); ++e)
                    par ^= hard_s[cw * p.hard_stride + __ldg(g.chk_var + e)] & 1u;
                if (par) atomicAdd(&scratch[4 + cw], 1);
            }
            __syncthreads();
        }
        if (last) break;
        for (int i = tid; i < p.CW; i += T) {
            if (i < ncw && frozen_s[i] == 0 && scratch[4 + i] == 0) frozen_s[i] = it;
            scratch[4 + i] = 0;
        }
        __syncthreads();
        bool all = true;
        for (int c = 0; c < ncw; ++c) all = all && (frozen_s[c] != 0);
        finished = all;
    }
    if (a.iters_used)
        for (int i = tid; i < ncw; i += T) a.iters_used[cw0 + i] = (a.early_exit && frozen_s[i]) ? frozen_s[i] : a.iters;
    if (a.x_out)
        for (int i = tid; i < ncw * E; i += T) {
            const int cw = i / E, e = i - cw * E;
            a.x_out[cw0 * E + i] = msg[cw * p.msg_stride + e];
        }
    if (a.syndrome) {
        for (int i = tid; i < ncw; i += T) a.syndrome[cw0 + i] = scratch[4 + i];
        __syncthreads();
        for (int i = tid; i < p.CW; i += T) scratch[4 + i] = 0;
    }
    if (a.hard_packed) pack_hard(hard_s, p.hard_stride, ncw, n, a.hard_packed + cw0 * ((n + 7) >> 3));
    if (a.counters) {
        __syncthreads();
        count_errors(hard_s, p.hard_stride, ncw, n, a.k_info,
                     a.ref_packed + cw0 * ((n + 7) >> 3), a.counters, scratch + 1);
    }
}

template <int MAXDV, int MAXDC, bool SW>
static int launch_t(const GenericParams &p, size_t smem, int grid, cudaStream_t s) {
    if (p.a.update == UPD_SP) {
        auto k = decode_generic_kernel<MAXDV, MAXDC, true, SW>;
        LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, s>>>(p);
    } else {
        auto k = decode_generic_kernel<MAXDV, MAXDC, false, SW>;
        LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, s>>>(p);
    }
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int launch_decode_generic(const GraphTables &g, int max_dv, int max_dc, const DecodeArgs &a, cudaStream_t s) {
    if (a.B <= 0) return LDPC_OK;
    GenericParams p;
    p.g = g; p.a = a;
    p.llr_stride = g.n | 1;                         // odd strides spread codewords over banks
    p.msg_stride = g.E | 1;
    p.hard_stride = (g.n + 15) & ~15;
    const size_t per_cw = sizeof(float) * (p.llr_stride + p.msg_stride) + p.hard_stride;
    const size_t fixed = sizeof(int) * 8 + 64;
    const size_t budget_small = 100 * 1024, budget_max = 227 * 1024;
    if (per_cw + fixed + 64 * sizeof(int) > budget_max) {
        set_error("code too large for shared-memory residency: %zu bytes per codeword", per_cw);
        return LDPC_EUNSUPPORTED;
    }
    int CW = (int)(budget_small / per_cw);
    if (CW < 1) CW = (int)((budget_max - fixed - 256) / per_cw);
    if (CW > 32) CW = 32;
    if (CW < 1) CW = 1;
    // do not starve the grid for small batches
    while (CW > 1 && (a.B + CW - 1) / CW < 2 * 148) CW >>= 1;
    p.CW = CW;
    const size_t smem = CW * per_cw + sizeof(int) * (8 + 2 * CW) + 16;
    const long long grid = (a.B + CW - 1) / CW;
    if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
    if (max_dv <= 4 && max_dc <= 4) return launch_t<4, 4, true>(p, smem, (int)grid, s);
    if (max_dv <= 12 && max_dc <= 8) return launch_t<12, 8, true>(p, smem, (int)grid, s);
    if (max_dv <= 32 && max_dc <= 32) return launch_t<32, 32, false>(p, smem, (int)grid, s);
    set_error("node degree above 32 is not supported (max_dv=%d, max_dc=%d)", max_dv, max_dc);
    return LDPC_EUNSUPPORTED;
}

}  // namespace ldpc
