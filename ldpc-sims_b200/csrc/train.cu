// train.cu - weighted ("neural") belief propagation with a tape and its exact sparse backward pass.
//
// Replaces, for training, what the reference does with dense [E,E] masks: BeliefPropagationVC_Function
// (bp/bp_vc.py:16-58) and BeliefPropagationCV_Function (bp/bp_cv.py:22-96; its backward materialises a
// [B,E,E,E] tensor) unrolled by BeliefPropagation.forward (bp/bp.py:43-51).  Sum-product only (the reference's rule).
//
// Two kernel pairs, chosen by batch size (wpc_warps below).
// Large batches: ONE THREAD PER CODEWORD.  Every per-edge array (the tape of C->V messages entering each iteration, the
// recomputed tanh values, the message gradients) lives in global memory as [edge][B], so a warp's 32 codewords read
// and write 128 contiguous bytes per edge and walk the Tanner graph in lock step (node tables are warp-uniform
// loads).  That makes the weight gradients - sums over the batch - a warp shuffle reduction followed by one atomic per
// warp and weight.  Batches up to 16 384: ONE WARP PER CODEWORD (further down).  In both, the forward arithmetic is node_math.cuh in the generic kernel's order, so prob equals
// ldpc_decode_weighted bit for bit; the backward is the derivative of exactly that forward (clamps pass the gradient
// where the clamped value is inside or on the bound, as torch.clamp does; a saturated product therefore has zero
// gradient, where the reference's hand-written backward keeps 2/(1-q^2) ~ 1e7 - a documented deviation).
#include "common.cuh"
#include "node_math.cuh"

namespace ldpc {

struct TrainArgs {
    GraphTables g;
    const float *llr;          // [B][n]
    long long B;
    int iters, w_stride;
    float clampv;
    const float *w_edge, *w_llr, *wf_edge, *wf_llr;
    const float *x0;           // [B][E] or null
    float *prob;               // [B][n]
    float *tape;               // C->V messages entering iteration it (tape[iters] feeds the marginal): [iters+1][E][B] for the
                               // thread-per-codeword pair, [iters+1][B][E] for the warp-per-codeword pair (opaque to callers)
    // backward only
    const float *grad_prob;    // [B][n]
    float *grad_llr;           // [B][n]
    float *g_w_edge, *g_w_llr, *g_wf_edge, *g_wf_llr;
    float *ws_u, *ws_g;        // [E][B] each
};

// (Keeping the working arrays in per-thread columns of shared memory was measured too: no faster at a batch of 512 -
// the chains are instruction-, not load-bound once the degrees are compile-time - and 2x slower at 262 144.)
template <int MAXDV, int MAXDC, bool SW>
__global__ void __launch_bounds__(128) bp_train_forward_kernel(const TrainArgs a) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const GraphTables &g = a.g;
    const int n = g.n, m = g.m, E = g.E;
    const long long B = a.B;
    const long long S = B;                                         // stride of the [edge][B] arrays
    const float *L = a.llr + b * n;
    for (int e = 0; e < E; ++e) a.tape[(long long)e * B + b] = a.x0 ? __ldg(a.x0 + b * E + e) : 0.0f;
    for (int it = 0; it < a.iters; ++it) {
        float *xout = a.tape + (long long)(it + 1) * E * B + b;
        const float *xin = a.tape + (long long)it * E * B + b;
        float *xw = xout;                                         // the check phase works in place: a node's edges are its own
        for (int v = 0; v < n; ++v) {
            const int b0 = __ldg(g.var_ptr + v);
            with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                int slot[D];
                float in[D], out[D];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) { slot[k] = __ldg(g.cm_of_vm + b0 + k); in[k] = xin[slot[k] * S]; }
                var_node_weighted<D, true>(in, d, __ldg(L + v), __ldg(a.w_llr + (long long)it * n + v),
                                           a.w_edge + ((long long)it * E + b0) * a.w_stride, a.w_stride, out);
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) xw[slot[k] * S] = out[k];
            });
        }
        for (int c = 0; c < m; ++c) {
            const int b0 = __ldg(g.chk_ptr + c);
            with_degree<MAXDC, SW>(__ldg(g.chk_ptr + c + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                float in[D], out[D];
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j < d) in[j] = xw[(b0 + j) * S];
                check_node_sp<D>(in, d, a.clampv, out);
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j < d) xw[(b0 + j) * S] = out[j];
            });
        }
    }
    const float *xl = a.tape + (long long)a.iters * E * B + b;
    for (int v = 0; v < n; ++v) {
        const int b0 = __ldg(g.var_ptr + v);
        with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
            constexpr int D = decltype(cap)::value;
            float in[D];
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k < d) in[k] = xl[__ldg(g.cm_of_vm + b0 + k) * S];
            a.prob[b * n + v] = prob_one(marginal_t_weighted<D>(in, d, __ldg(L + v), __ldg(a.wf_llr + v), a.wf_edge + b0));
        });
        if (SW && __ldg(g.var_ptr + v + 1) == b0)             // a variable without edges has no switch case
            a.prob[b * n + v] = prob_one(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(__ldg(a.wf_llr + v), -__ldg(L + v)), 0.0f)));
    }
}

// sum over the warp's codewords, then one atomic (all 32 lanes call this together)
__device__ __forceinline__ void batch_add(float *dst, float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0f) atomicAdd(dst, v);
}

template <int MAXDV, int MAXDC, bool SW>
__global__ void __launch_bounds__(128) bp_train_backward_kernel(const TrainArgs a) {
    const long long bb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = bb < a.B;
    const long long b = valid ? bb : a.B - 1;             // idle lanes shadow the last codeword and contribute zeros
    const float live = valid ? 1.0f : 0.0f;
    const GraphTables &g = a.g;
    const int n = g.n, m = g.m, E = g.E;
    const long long B = a.B;
    const long long S = B;
    const float *L = a.llr + b * n;
    float *U = a.ws_u + bb, *GX = a.ws_g + bb;            // tanh(V->C), message gradients; idle lanes never store (guarded below)

    // ---- marginal + sigmoid (bp/bp.py:36-39,51): prob = 1 - sigmoid(t), dprob/dt = -prob (1 - prob) -------------
    {
        const float *xl = a.tape + (long long)a.iters * E * B + b;
        for (int v = 0; v < n; ++v) {
            const int b0 = __ldg(g.var_ptr + v);
            if (SW && __ldg(g.var_ptr + v + 1) == b0) {          // a variable without edges has no switch case
                const float l = __ldg(L + v), wl = __ldg(a.wf_llr + v);
                const float P = prob_one(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(wl, -l), 0.0f)));
                const float h = 0.5f * live * __ldg(a.grad_prob + b * n + v) * (-(P * (1.0f - P)));
                batch_add(a.g_wf_llr + v, h * -l);
                if (valid) a.grad_llr[b * n + v] = h * -wl;
            }
            with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                int slot[D];
                float in[D];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) { slot[k] = __ldg(g.cm_of_vm + b0 + k); in[k] = xl[(long long)slot[k] * B]; }
                const float l = __ldg(L + v), wl = __ldg(a.wf_llr + v);
                const float P = prob_one(marginal_t_weighted<D>(in, d, l, wl, a.wf_edge + b0));
                const float h = 0.5f * live * __ldg(a.grad_prob + b * n + v) * (-(P * (1.0f - P)));
                batch_add(a.g_wf_llr + v, h * -l);
                if (valid) a.grad_llr[b * n + v] = h * -wl;
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) {
                        batch_add(a.g_wf_edge + b0 + k, h * in[k]);
                        if (valid) GX[slot[k] * S] = h * __ldg(a.wf_edge + b0 + k);
                    }
            });
        }
    }

    for (int it = a.iters - 1; it >= 0; --it) {
        const float *xin = a.tape + (long long)it * E * B + b;
        const float *wE = a.w_edge + (long long)it * E * a.w_stride;
        // ---- recompute tanh(V->C) of this iteration --------------------------------------------------------------
        for (int v = 0; v < n; ++v) {
            const int b0 = __ldg(g.var_ptr + v);
            with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                int slot[D];
                float in[D], out[D];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) { slot[k] = __ldg(g.cm_of_vm + b0 + k); in[k] = xin[slot[k] * S]; }
                var_node_weighted<D, true>(in, d, __ldg(L + v), __ldg(a.w_llr + (long long)it * n + v),
                                           wE + (long long)b0 * a.w_stride, a.w_stride, out);
                if (valid)
#pragma unroll
                    for (int k = 0; k < D; ++k)
                        if (k < d) U[slot[k] * S] = out[k];
            });
        }
        // ---- check node backward: GX (d/d message out) -> GX (d/d pre-tanh V->C value) ------------------------------
        if (valid)
            for (int c = 0; c < m; ++c) {
                const int b0 = __ldg(g.chk_ptr + c);
                with_degree<MAXDC, SW>(__ldg(g.chk_ptr + c + 1) - b0, [&](auto cap, int d) {
                    constexpr int D = decltype(cap)::value;
                    float u[D], p[D], gp[D], gu[D];
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j < d) u[j] = U[(b0 + j) * S];
                    prod_others<D>(u, d, p);
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        gu[j] = 0.0f;
                        if (j < d) {
                            const float q = clampf(p[j], LDPC_P_CLAMP);
                            const float o = logf(div_rn_one_plus_minus(q));
                            const bool pass = fabsf(p[j]) <= LDPC_P_CLAMP && fabsf(o) <= a.clampv;
                            gp[j] = pass ? GX[(b0 + j) * S] * (2.0f / ((1.0f - q) * (1.0f + q))) : 0.0f;
                        }
                    }
                    // d/du_i = sum_{j != i} gp_j prod_{k != i,j} u_k: for every j, the leave-one-out products of u with u_j := 1
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j < d) {
                            float uj[D], pj[D];
#pragma unroll
                            for (int k = 0; k < D; ++k) uj[k] = (k == j) ? 1.0f : u[k];
                            prod_others<D>(uj, d, pj);
#pragma unroll
                            for (int i = 0; i < D; ++i)
                                if (i < d && i != j) gu[i] = fmaf(gp[j], pj[i], gu[i]);
                        }
#pragma unroll
                    for (int i = 0; i < D; ++i)
                        if (i < d) GX[(b0 + i) * S] = gu[i] * (1.0f - u[i] * u[i]);
                });
            }
        // ---- variable node backward -------------------------------------------------------------------------------------
        for (int v = 0; v < n; ++v) {
            const int b0 = __ldg(g.var_ptr + v);
            with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                int slot[D];
                float x[D], hs[D];
                float sum = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) {
                        slot[k] = __ldg(g.cm_of_vm + b0 + k);
                        x[k] = xin[slot[k] * S];
                        hs[k] = valid ? 0.5f * GX[slot[k] * S] : 0.0f;
                        sum += hs[k];
                    }
                const float l = __ldg(L + v), wl = __ldg(a.w_llr + (long long)it * n + v);
                batch_add(a.g_w_llr + (long long)it * n + v, sum * -l);
                if (valid) a.grad_llr[b * n + v] += sum * -wl;
                const float *w = wE + (long long)b0 * a.w_stride;
                float *gw = a.g_w_edge + ((long long)it * E + b0) * a.w_stride;
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j < d) {
                        float gx = 0.0f;
#pragma unroll
                        for (int k = 0; k < D; ++k)
                            if (k < d && k != j) {
                                batch_add(gw + k * a.w_stride + j, hs[k] * x[j]);
                                gx = fmaf(hs[k], __ldg(w + k * a.w_stride + j), gx);
                            }
                        if (valid) GX[slot[j] * S] = gx;
                    }
            });
        }
    }
}

// ---- small batches (up to 16 384 codewords): ONE WARP PER CODEWORD ---------------------------------------------------------------------------
// With a few hundred codewords (the reference trains on minibatches of 512, ofdm/ofdm_nn.py:262) one thread per
// codeword leaves the GPU with 16 warps, each walking the whole graph serially.  Here the 32 lanes of a warp take the
// nodes of ONE codeword (variable v = lane, lane + 32, ...), the working arrays are a per-warp slice of shared memory,
// phases are separated by __syncwarp, and every weight-gradient term goes straight to a float atomic (B atomics per
// weight - cheap while B is small, which is exactly when this variant is selected).  Tape layout here: [it][B][E].
template <int MAXDV, int MAXDC, bool SW>
__global__ void __launch_bounds__(128) bp_train_forward_wpc_kernel(const TrainArgs a) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const long long cw = (long long)blockIdx.x * W + w;
    if (cw >= a.B) return;                                   // whole warps leave; only __syncwarp below
    const GraphTables &g = a.g;
    const int n = g.n, m = g.m, E = g.E;
    float *X = sm + (size_t)w * E;
    const float *L = a.llr + cw * n;
    float *t0 = a.tape + cw * E;
    for (int e = lane; e < E; e += 32) { const float x = a.x0 ? __ldg(a.x0 + cw * E + e) : 0.0f; X[e] = x; t0[e] = x; }
    __syncwarp();
    for (int it = 0; it < a.iters; ++it) {
        float *tout = a.tape + ((long long)(it + 1) * a.B + cw) * E;
        for (int v = lane; v < n; v += 32) {                 // in place: a variable's slots are its own
            const int b0 = __ldg(g.var_ptr + v);
            with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                int slot[D];
                float in[D], out[D];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) { slot[k] = __ldg(g.cm_of_vm + b0 + k); in[k] = X[slot[k]]; }
                var_node_weighted<D, true>(in, d, __ldg(L + v), __ldg(a.w_llr + (long long)it * n + v),
                                           a.w_edge + ((long long)it * E + b0) * a.w_stride, a.w_stride, out);
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) X[slot[k]] = out[k];
            });
        }
        __syncwarp();
        for (int c = lane; c < m; c += 32) {
            const int b0 = __ldg(g.chk_ptr + c);
            with_degree<MAXDC, SW>(__ldg(g.chk_ptr + c + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                float in[D], out[D];
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j < d) in[j] = X[b0 + j];
                check_node_sp<D>(in, d, a.clampv, out);
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j < d) { X[b0 + j] = out[j]; tout[b0 + j] = out[j]; }
            });
        }
        __syncwarp();
    }
    for (int v = lane; v < n; v += 32) {
        const int b0 = __ldg(g.var_ptr + v);
        with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
            constexpr int D = decltype(cap)::value;
            float in[D];
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k < d) in[k] = X[__ldg(g.cm_of_vm + b0 + k)];
            a.prob[cw * n + v] = prob_one(marginal_t_weighted<D>(in, d, __ldg(L + v), __ldg(a.wf_llr + v), a.wf_edge + b0));
        });
        if (__ldg(g.var_ptr + v + 1) == b0) a.prob[cw * n + v] = prob_one(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(__ldg(a.wf_llr + v), -__ldg(L + v)), 0.0f)));
    }
}

__device__ __forceinline__ void red_add(float *dst, float v) {
    if (v != 0.0f) atomicAdd(dst, v);
}

template <int MAXDV, int MAXDC, bool SW>
__global__ void __launch_bounds__(128) bp_train_backward_wpc_kernel(const TrainArgs a) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const long long cw = (long long)blockIdx.x * W + w;
    if (cw >= a.B) return;
    const GraphTables &g = a.g;
    const int n = g.n, m = g.m, E = g.E;
    float *X = sm + (size_t)w * 3 * E, *U = X + E, *GX = U + E;
    const float *L = a.llr + cw * n;
    {
        const float *xl = a.tape + ((long long)a.iters * a.B + cw) * E;
        for (int v = lane; v < n; v += 32) {
            const int b0 = __ldg(g.var_ptr + v), dd = __ldg(g.var_ptr + v + 1) - b0;
            const float l = __ldg(L + v), wl = __ldg(a.wf_llr + v);
            if (SW && dd == 0) {                             // a variable without edges: only the channel term (no switch case)
                const float P = prob_one(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(wl, -l), 0.0f)));
                const float h = 0.5f * __ldg(a.grad_prob + cw * n + v) * (-(P * (1.0f - P)));
                red_add(a.g_wf_llr + v, h * -l);
                a.grad_llr[cw * n + v] = h * -wl;
            }
            with_degree<MAXDV, SW>(dd, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                int slot[D];
                float in[D];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) { slot[k] = __ldg(g.cm_of_vm + b0 + k); in[k] = xl[slot[k]]; }
                const float P = prob_one(marginal_t_weighted<D>(in, d, l, wl, a.wf_edge + b0));
                const float h = 0.5f * __ldg(a.grad_prob + cw * n + v) * (-(P * (1.0f - P)));
                red_add(a.g_wf_llr + v, h * -l);
                a.grad_llr[cw * n + v] = h * -wl;
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) {
                        red_add(a.g_wf_edge + b0 + k, h * in[k]);
                        GX[slot[k]] = h * __ldg(a.wf_edge + b0 + k);
                    }
            });
        }
    }
    __syncwarp();
    for (int it = a.iters - 1; it >= 0; --it) {
        const float *xg = a.tape + ((long long)it * a.B + cw) * E;
        const float *wE = a.w_edge + (long long)it * E * a.w_stride;
        for (int e = lane; e < E; e += 32) X[e] = xg[e];
        __syncwarp();
        for (int v = lane; v < n; v += 32) {                 // recompute tanh(V->C)
            const int b0 = __ldg(g.var_ptr + v);
            with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                int slot[D];
                float in[D], out[D];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) { slot[k] = __ldg(g.cm_of_vm + b0 + k); in[k] = X[slot[k]]; }
                var_node_weighted<D, true>(in, d, __ldg(L + v), __ldg(a.w_llr + (long long)it * n + v),
                                           wE + (long long)b0 * a.w_stride, a.w_stride, out);
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) U[slot[k]] = out[k];
            });
        }
        __syncwarp();
        for (int c = lane; c < m; c += 32) {                 // check node backward, in place on GX
            const int b0 = __ldg(g.chk_ptr + c);
            with_degree<MAXDC, SW>(__ldg(g.chk_ptr + c + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                float u[D], p[D], gp[D], gu[D];
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j < d) u[j] = U[b0 + j];
                prod_others<D>(u, d, p);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    gu[j] = 0.0f;
                    if (j < d) {
                        const float q = clampf(p[j], LDPC_P_CLAMP);
                        const float o = logf(div_rn_one_plus_minus(q));
                        const bool pass = fabsf(p[j]) <= LDPC_P_CLAMP && fabsf(o) <= a.clampv;
                        gp[j] = pass ? GX[b0 + j] * (2.0f / ((1.0f - q) * (1.0f + q))) : 0.0f;
                    }
                }
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j < d) {
                        float uj[D], pj[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) uj[k] = (k == j) ? 1.0f : u[k];
                        prod_others<D>(uj, d, pj);
#pragma unroll
                        for (int i = 0; i < D; ++i)
                            if (i < d && i != j) gu[i] = fmaf(gp[j], pj[i], gu[i]);
                    }
#pragma unroll
                for (int i = 0; i < D; ++i)
                    if (i < d) GX[b0 + i] = gu[i] * (1.0f - u[i] * u[i]);
            });
        }
        __syncwarp();
        for (int v = lane; v < n; v += 32) {                 // variable node backward
            const int b0 = __ldg(g.var_ptr + v);
            with_degree<MAXDV, SW>(__ldg(g.var_ptr + v + 1) - b0, [&](auto cap, int d) {
                constexpr int D = decltype(cap)::value;
                int slot[D];
                float x[D], hs[D];
                float sum = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d) {
                        slot[k] = __ldg(g.cm_of_vm + b0 + k);
                        x[k] = X[slot[k]];
                        hs[k] = 0.5f * GX[slot[k]];
                        sum += hs[k];
                    }
                const float l = __ldg(L + v), wl = __ldg(a.w_llr + (long long)it * n + v);
                red_add(a.g_w_llr + (long long)it * n + v, sum * -l);
                a.grad_llr[cw * n + v] += sum * -wl;
                const float *wr = wE + (long long)b0 * a.w_stride;
                float *gw = a.g_w_edge + ((long long)it * E + b0) * a.w_stride;
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j < d) {
                        float gx = 0.0f;
#pragma unroll
                        for (int k = 0; k < D; ++k)
                            if (k < d && k != j) {
                                red_add(gw + k * a.w_stride + j, hs[k] * x[j]);
                                gx = fmaf(hs[k], __ldg(wr + k * a.w_stride + j), gx);
                            }
                        GX[slot[j]] = gx;
                    }
            });
        }
        __syncwarp();
    }
}

// Which variant serves a batch: forward and backward MUST agree (the tape layout differs).
static int wpc_warps(const TrainArgs &a) {
    if (a.B > 16384) return 0;             // measured crossover on B200 (default code): beyond it one thread per codeword wins
    const size_t per_warp = (size_t)a.g.E * 3 * sizeof(float);
    if (per_warp * 4 <= 96 * 1024) return 4;
    if (per_warp * 2 <= 200 * 1024) return 2;
    if (per_warp <= 200 * 1024) return 1;
    return 0;
}

template <int MAXDV, int MAXDC, bool SW>
static int launch_train_t(const TrainArgs &a, bool backward, cudaStream_t s) {
    if (const int W = wpc_warps(a)) {
        const long long grid = (a.B + W - 1) / W;
        const size_t bytes = (size_t)W * a.g.E * sizeof(float) * (backward ? 3 : 1);
        void (*k)(const TrainArgs) = backward ? bp_train_backward_wpc_kernel<MAXDV, MAXDC, SW> : bp_train_forward_wpc_kernel<MAXDV, MAXDC, SW>;
        if (bytes > 48 * 1024) LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        k<<<(int)grid, 32 * W, bytes, s>>>(a);
        LDPC_CUDA_TRY(cudaGetLastError());
        return LDPC_OK;
    }
    const int threads = 128;
    const long long grid = (a.B + threads - 1) / threads;
    if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
    if (backward) bp_train_backward_kernel<MAXDV, MAXDC, SW><<<(int)grid, threads, 0, s>>>(a);
    else bp_train_forward_kernel<MAXDV, MAXDC, SW><<<(int)grid, threads, 0, s>>>(a);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

static int launch_train(const ldpc_code *code, const TrainArgs &a, bool backward, cudaStream_t s) {
    if (code->max_dv <= 4 && code->max_dc <= 8) return launch_train_t<4, 8, true>(a, backward, s);     // the default (64,32) code
    if (code->max_dv <= 12 && code->max_dc <= 8) return launch_train_t<12, 8, true>(a, backward, s);   // 802.11n
    if (code->max_dv <= 32 && code->max_dc <= 32) return launch_train_t<32, 32, false>(a, backward, s);
    set_error("node degree above 32");
    return LDPC_EUNSUPPORTED;
}

}  // namespace ldpc

using namespace ldpc;

static int check_train(const ldpc_code_t *code, const float *llr, int64_t B, int iters, float clamp_value, const float *w_edge,
                       const float *w_llr, const float *wf_edge, const float *wf_llr, int w_stride, const float *tape) {
    if (!code || ((!llr || !tape) && B > 0)) { set_error("ldpc_bp_train: null argument"); return LDPC_EINVAL; }
    if (B < 0 || iters < 0 || !(clamp_value > 0.0f)) { set_error("ldpc_bp_train: bad B / iters / clamp"); return LDPC_EINVAL; }
    if (((!w_edge || !w_llr) && iters > 0) || !wf_edge || !wf_llr || w_stride < code->max_dv) {   // zero iterations: empty per-iteration tables
        set_error("ldpc_bp_train: weight tables missing or w_stride < max_dv (%d)", code->max_dv);
        return LDPC_EINVAL;
    }
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != code->device) { set_error("ldpc_bp_train: current device %d, code tables on %d", dev, code->device); return LDPC_EINVAL; }
    return LDPC_OK;
}

extern "C" {

int ldpc_bp_train_forward(const ldpc_code_t *code, const float *llr, int64_t B, int iters, float clamp_value,
                          const float *w_edge, const float *w_llr, const float *wf_edge, const float *wf_llr, int w_stride,
                          const float *x0, float *prob, float *tape, ldpc_stream_t stream) {
    int rc = check_train(code, llr, B, iters, clamp_value, w_edge, w_llr, wf_edge, wf_llr, w_stride, tape);
    if (rc) return rc;
    if (!prob && B > 0) { set_error("ldpc_bp_train_forward: prob is null"); return LDPC_EINVAL; }
    if (B == 0) return LDPC_OK;
    TrainArgs a = {};
    a.g = code->g; a.llr = llr; a.B = B; a.iters = iters; a.w_stride = w_stride; a.clampv = clamp_value;
    a.w_edge = w_edge; a.w_llr = w_llr; a.wf_edge = wf_edge; a.wf_llr = wf_llr; a.x0 = x0; a.prob = prob; a.tape = tape;
    return launch_train(code, a, false, (cudaStream_t)stream);
}

int ldpc_bp_train_backward(const ldpc_code_t *code, const float *llr, int64_t B, int iters, float clamp_value,
                           const float *w_edge, const float *w_llr, const float *wf_edge, const float *wf_llr, int w_stride,
                           const float *tape, const float *grad_prob, float *grad_llr, float *g_w_edge, float *g_w_llr,
                           float *g_wf_edge, float *g_wf_llr, float *workspace, ldpc_stream_t stream) {
    int rc = check_train(code, llr, B, iters, clamp_value, w_edge, w_llr, wf_edge, wf_llr, w_stride, tape);
    if (rc) return rc;
    if (((!g_w_edge || !g_w_llr) && iters > 0) || ((!grad_prob || !grad_llr) && B > 0) || !g_wf_edge || !g_wf_llr || !workspace) {
        set_error("ldpc_bp_train_backward: null output or workspace");
        return LDPC_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t E = code->E, n = code->n;
    if (iters > 0) {
        LDPC_CUDA_TRY(cudaMemsetAsync(g_w_edge, 0, sizeof(float) * (size_t)iters * E * w_stride, s));
        LDPC_CUDA_TRY(cudaMemsetAsync(g_w_llr, 0, sizeof(float) * (size_t)iters * n, s));
    }
    LDPC_CUDA_TRY(cudaMemsetAsync(g_wf_edge, 0, sizeof(float) * E, s));
    LDPC_CUDA_TRY(cudaMemsetAsync(g_wf_llr, 0, sizeof(float) * n, s));
    if (B == 0) return LDPC_OK;
    TrainArgs a = {};
    a.g = code->g; a.llr = llr; a.B = B; a.iters = iters; a.w_stride = w_stride; a.clampv = clamp_value;
    a.w_edge = w_edge; a.w_llr = w_llr; a.wf_edge = wf_edge; a.wf_llr = wf_llr; a.tape = const_cast<float *>(tape);
    a.grad_prob = grad_prob; a.grad_llr = grad_llr;
    a.g_w_edge = g_w_edge; a.g_w_llr = g_w_llr; a.g_wf_edge = g_wf_edge; a.g_wf_llr = g_wf_llr;
    a.ws_u = workspace; a.ws_g = workspace + E * (size_t)B;
    return launch_train(code, a, true, s);
}

}  // extern "C"
