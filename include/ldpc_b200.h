/* ldpc_b200.h - C ABI of the B200-native LDPC belief-propagation / OFDM link-simulation
 * hot path (libldpc_b200.so, hand-written sm_100a CUDA).
 *
 * This is the drop-in boundary for the reference's decode path.  Every entry point cites
 * the reference interface (file:line under pytorch/ of realjwin/ldpc-sims) it replaces.
 * Plain pointers and sizes only; no torch types.  Unless a name ends in _host every
 * data pointer is a DEVICE pointer owned by the caller, the call is asynchronous on the
 * given stream, returns 0 or a negative LDPC_E* code and never throws;
 * ldpc_last_error() gives the message for the calling thread.  A code handle is
 * immutable after creation, so concurrent decodes on different streams are safe.
 * There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------ */
enum {
    LDPC_OK = 0,
    LDPC_EINVAL = -1,       /* bad argument (shape, null pointer, unsupported degree ...) */
    LDPC_ECUDA = -2,        /* CUDA runtime error (message has cudaGetErrorString)        */
    LDPC_ENOMEM = -3,
    LDPC_EUNSUPPORTED = -4  /* valid request this build cannot serve                      */
};

/* ---- check-node update rule (reference: tanh-domain sum-product only,
 *      bp/bp.py:27-31 + bp/bp_cv.py:42-50; the min-sum family is new) ------------------- */
enum {
    LDPC_UPDATE_SP = 0,      /* tanh / 2*atanh sum-product, the reference's arithmetic   */
    LDPC_UPDATE_MINSUM = 1,  /* min-sum                                                   */
    LDPC_UPDATE_NMS = 2,     /* normalized min-sum, param = alpha                         */
    LDPC_UPDATE_OMS = 3      /* offset min-sum,     param = beta                          */
};

/* ---- element type of an LLR buffer --------------------------------------------------- */
enum { LDPC_F32 = 0, LDPC_F64 = 1, LDPC_F16 = 2 };

/* ---- which kernel a code handle dispatches to ---------------------------------------- */
enum { LDPC_KERNEL_GENERIC = 0, LDPC_KERNEL_QC = 1 };

typedef struct ldpc_code ldpc_code_t;
typedef void *ldpc_stream_t;             /* a cudaStream_t (CUstream); NULL = default stream */

/* Library / device probes (no reference counterpart). */
int ldpc_abi_version(void);
const char *ldpc_last_error(void);
int ldpc_device_count(void);             /* 0 when no usable CUDA device */

/* ldpc_code_create - compile a parity-check matrix into device-resident sparse edge tables.
 * Replaces generate_masks(H) (bp/masking.py:12-147: four dense E x E / n x E masks) and the
 * mask upload in BeliefPropagation.__init__ (bp/bp.py:19-39).
 *   row_ptr[m+1], col_idx[E]: CSR of H, columns ascending inside a row (this IS the
 *   reference's check-major edge order, masking.py:85-88).
 *   qc_Z > 0 with qc_proto[(m/Z)*(n/Z)] (shift or -1 per block, row-major) declares the
 *   quasi-cyclic structure; when it matches a prototype compiled into the library the
 *   handle dispatches to the code-specialised kernel, otherwise to the generic one.
 * Host-side, synchronous.  Free with ldpc_code_destroy. */
int ldpc_code_create(const int32_t *row_ptr, const int32_t *col_idx, int m, int n,
                     int qc_Z, const int16_t *qc_proto, ldpc_code_t **out);
void ldpc_code_destroy(ldpc_code_t *code);

/* Geometry of a compiled code.  E replaces BeliefPropagation.layer_size() (bp/bp.py:61-62). */
typedef struct {
    int32_t m, n, E, max_dc, max_dv, kernel, qc_Z, reserved;
} ldpc_code_info_t;
int ldpc_code_info(const ldpc_code_t *code, ldpc_code_info_t *info);

/* Force the generic kernel for a handle (testing / A-B comparison). */
int ldpc_code_set_kernel(ldpc_code_t *code, int kernel);

/* ldpc_decode - one batch through `iters` flooding iterations.
 * Replaces BeliefPropagation.forward(x, llr, clamp_value) (bp/bp.py:43-51) together with
 * the per-layer ops it unrolls: BeliefPropagationVC (bp/bp_vc.py:16-32), nn.Tanh
 * (bp/bp.py:29), BeliefPropagationCV (bp/bp_cv.py:22-55), the outer clamp (bp/bp.py:47),
 * the final marginal + sigmoid (bp/bp.py:36-39,51) and np.round (ofdm/ofdm_functions.py:161).
 *   llr        [B,n] row-major, log(P1/P0) as the callers pass it (ofdm_functions.py:72);
 *              element type llr_dtype (the reference casts f64 -> f32 at ofdm_functions.py:156)
 *   x0         [B,E] f32 initial C->V messages, check-major, or NULL for the zeros every
 *              reference caller passes (ofdm_functions.py:157)
 *   outputs, each nullable:
 *   prob       [B,n] f32  P(bit=1) = 1 - sigmoid(t)            (what forward() returns)
 *   llr_post   [B,n] f32  posterior log(P1/P0) = -2 t          (never exposed by the reference)
 *   hard       [B,n] u8   round-half-even(prob) in {0,1}       (ofdm_functions.py:161)
 *   hard_packed[B,ceil(n/8)] u8, MSB-first per byte (numpy.packbits layout)
 *   syndrome   [B] i32    number of unsatisfied checks of the hard decision (new)
 *   x_out      [B,E] f32  final C->V messages, check-major
 */
int ldpc_decode(const ldpc_code_t *code, const void *llr, int llr_dtype, int64_t B,
                int iters, int update, float clamp_value, float param, const float *x0,
                float *prob, float *llr_post, uint8_t *hard, uint8_t *hard_packed,
                int32_t *syndrome, float *x_out, ldpc_stream_t stream);

/* ldpc_decode_host - the decode_bits batching loop (ofdm/ofdm_functions.py:131-163) with
 * HOST buffers: chunked, double-buffered H2D copy -> ldpc_decode -> D2H copy on internal
 * streams; synchronous.  llr_host [N,n] of llr_dtype; outputs (each nullable):
 * hard_host [N,n] u8, hard_packed_host [N,ceil(n/8)] u8, llr_post_host [N,n] f32,
 * syndrome_host [N] i32.  Every one of the N rows is decoded (the reference's silent
 * drop of a ragged tail, ofdm_functions.py:135, is reproduced by the Python wrapper). */
int ldpc_decode_host(const ldpc_code_t *code, const void *llr_host, int llr_dtype, int64_t N,
                     int iters, int update, float clamp_value, float param,
                     uint8_t *hard_host, uint8_t *hard_packed_host, float *llr_post_host,
                     int32_t *syndrome_host, int64_t chunk_codewords);

/* ldpc_count_errors - exact integer link metrics, accumulated (+=) into counters[5] (i64):
 *   {uncoded bit errors over n, decoded info-bit errors over the first k, frame errors
 *    (any of n decoded bits wrong), bits = B*n, frames = B}
 * Replaces evaluate_quantized_snr.py:169-188 / compute_ber (ofdm/ofdm_functions.py:83-84).
 *   llr [B,n] channel LLRs (hard decision (sign+1)//2, llr==0 -> 0) or NULL to skip the
 *   uncoded count; hard [B,n] u8 decoded bits; ref_bits [B,n] u8 transmitted codeword. */
int ldpc_count_errors(const void *llr, int llr_dtype, const uint8_t *hard, const uint8_t *ref_bits,
                      int64_t B, int n, int k, int64_t *counters, ldpc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
