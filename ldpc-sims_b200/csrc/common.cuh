// common.cuh - parameter blocks and error plumbing shared by the kernels and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <vector>
#include "../../include/ldpc_b200.h"

namespace ldpc {

struct DecodeArgs {
    // inputs
    const void *llr;        // [B,n] of llr_dtype
    int llr_dtype;
    long long B;
    int iters, update;
    float clampv, param;
    const float *x0;        // [B,E] or null
    // outputs (nullable)
    float *prob, *llr_post;
    uint8_t *hard, *hard_packed;
    int32_t *syndrome;
    float *x_out;
    // fused error counting (nullable): ref_bits [B,n] u8 or packed, counters[5] i64
    const uint8_t *ref_packed;  // [B,ceil(n/8)] transmitted codeword, MSB-first
    unsigned long long *counters;
    int k_info;
    // syndrome-based early termination (off in every reference-parity run: the reference has a fixed
    // iteration count, bp/bp.py:46-47): a codeword whose hard decision satisfies all checks after an
    // iteration is frozen; iters_used [B] i32 (nullable) receives the iterations actually run
    int early_exit;
    int32_t *iters_used;
    void *host_pipe;    // lazily created staging state of ldpc_decode_host
    int precision;      // LDPC_PREC_*
    // the reference's trainable weights (bp_vc.py:16-32), nullable (generic kernel only):
    // w_edge [iters][E][w_stride]: row = variable-major OUT edge, column j = weight of the variable's j-th edge as input;
    // w_llr [iters][n]; wf_edge [E] (variable-major) and wf_llr [n] for the final marginal
    const float *w_edge, *w_llr, *wf_edge, *wf_llr;
    int w_stride;
};

struct GraphTables {        // device pointers
    int m, n, E;
    const int32_t *chk_ptr;   // [m+1]
    const int32_t *chk_var;   // [E]  variable of a check-major edge
    const int32_t *var_ptr;   // [n+1]
    const int32_t *cm_of_vm;  // [E]  check-major slot of a variable-major edge
};

__device__ __forceinline__ float load_llr(const void *p, int dtype, long long i) {
    if (dtype == LDPC_F32) return __ldg(reinterpret_cast<const float *>(p) + i);
    if (dtype == LDPC_F64) return static_cast<float>(__ldg(reinterpret_cast<const double *>(p) + i));
    if (dtype == LDPC_I8) return static_cast<float>(__ldg(reinterpret_cast<const signed char *>(p) + i));
    return __half2float(__ldg(reinterpret_cast<const __half *>(p) + i));
}

}  // namespace ldpc

// the opaque handle of include/ldpc_b200.h
struct ldpc_code {
    int m, n, E, max_dc, max_dv;
    int kernel;         // LDPC_KERNEL_*
    int qc_id;          // index of the compiled specialisation or -1
    int tiny_id;        // index of the compiled register-resident specialisation (decode_tiny.cu) or -1
    int32_t *d_qc_rt;   // device tables of the run-time QC kernel (decode_qc_rt.cu) or null
    int qc_mb, qc_nb, qc_nblk;
    int qc_Z;
    int device;
    int32_t *d_tables;  // one allocation: chk_ptr | chk_var | var_ptr | cm_of_vm
    ldpc::GraphTables g;
    uint32_t *d_gen;    // bit-packed parity rows of the systematic generator [m][ceil(k/32)] or null
    int k_info;
    void *host_pipe;    // lazily created staging state of ldpc_decode_host
    int precision;      // LDPC_PREC_*
};

namespace ldpc {

void set_error(const char *fmt, ...);
int decode_dispatch(const ldpc_code *code, const DecodeArgs &a, cudaStream_t s);
int cuda_fail(cudaError_t e, const char *what);

#define LDPC_CUDA_TRY(expr)                                         \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return ::ldpc::cuda_fail(_e, #expr); \
    } while (0)

// kernels' host launchers
int launch_decode_generic(const GraphTables &g, int max_dv, int max_dc, const DecodeArgs &a, cudaStream_t s);
bool qc_kernel_available(int Z, int mb, int nb, const int16_t *proto);
int launch_decode_qc(int qc_id, const DecodeArgs &a, cudaStream_t s);
int launch_decode_qc_h2(int qc_id, const DecodeArgs &a, cudaStream_t s);
int launch_decode_qc_tma(int qc_id, const DecodeArgs &a, cudaStream_t s);   // LDPC_EUNSUPPORTED -> launch_decode_qc
struct LinkParams;
int launch_sim_fused_qc(int qc_id, const DecodeArgs &a, const LinkParams &lp, cudaStream_t s);   // LDPC_EUNSUPPORTED -> use the 3-launch chain
int qc_lookup(int Z, int mb, int nb, const int16_t *proto);   // -1 if no compiled specialisation
int qc_register_plugin(const char *so_path);                   // -> registry id (>= 0) or a negative LDPC_E* code
int tiny_lookup(int m, int n, const int32_t *row_ptr, const int32_t *col_idx);   // -1 if no compiled register-resident specialisation
int launch_decode_tiny(int tiny_id, const DecodeArgs &a, cudaStream_t s);
int launch_decode_qc_rt(const int32_t *d_tab, int Z, int mb, int nb, int nblk, int max_dv, int max_dc, const DecodeArgs &a, cudaStream_t s);
bool qc_rt_supported(int Z, int mb, int nb, int nblk, int max_dv, int max_dc);
int qc_rt_build_tables(int Z, int mb, int nb, const int16_t *proto, std::vector<int32_t> &out, int *max_dv, int *max_dc);
void qc_plan_info(int qc_id, int out[4]);                     // {register-resident blocks, shared-memory blocks, threads/CTA, codewords/CTA}

}  // namespace ldpc
