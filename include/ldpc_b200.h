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
enum { LDPC_F32 = 0, LDPC_F64 = 1, LDPC_F16 = 2, LDPC_I8 = 3 };   /* I8: receiver-quantised LLRs, value = the integer (decoder
                                                                      input only; a quarter of the f32 bytes over PCIe / HBM) */

/* ---- which kernel a code handle dispatches to ---------------------------------------- */
enum { LDPC_KERNEL_GENERIC = 0, LDPC_KERNEL_QC = 1, LDPC_KERNEL_TINY = 2 /* register-resident, one thread per codeword: the reference's default (64,32) code */,
       LDPC_KERNEL_QC_RT = 3 /* any quasi-cyclic code, prototype matrix at run time (qc_Z / qc_proto without a compiled specialisation) */,
       LDPC_KERNEL_QC_TMA = 4 /* opt-in (ldpc_code_set_kernel): the persistent form of the compiled kernel - CTAs loop over codeword tiles,
                                 the next tile's LLRs arrive by cp.async.bulk while the current one decodes; same bits as LDPC_KERNEL_QC,
                                 measured 2 % slower on B200 (HBM is at 9 % of its bandwidth: there is no load latency left to hide) */ };

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

/* Register a quasi-cyclic code that was specialised at RUN TIME: `so_path` is a shared object built from one generated
 * translation unit (ldpc_b200/jit.py writes it from the caller's prototype matrix and compiles it with the system nvcc
 * against the library's own kernel templates, csrc/decode_qc_code.cuh).  After registration ldpc_code_create selects the
 * code-compiled kernel (LDPC_KERNEL_QC) for that prototype exactly as for the built-in IEEE 802.11n family; without it
 * the prototype runs on LDPC_KERNEL_QC_RT.  Returns the registry id (>= 0) or a negative LDPC_E* code (file not loadable,
 * not a plug-in, built against other headers).  Replaces nothing in the reference: it has one hard-wired code
 * (bp/parity.py:7-47) and mentions a parity.mat it does not ship (bp/masking.py:151-153). */
int ldpc_qc_register_plugin(const char *so_path);

/* Execution plan of the code-specialised kernel (zeros for the generic kernel):
 * out = {blocks whose messages stay in registers, blocks exchanged through shared memory,
 *        threads per CTA, codewords per CTA}.  Used by bench.py for the roofline arithmetic. */
int ldpc_code_plan_info(const ldpc_code_t *code, int32_t out[4]);

/* Message precision of the code-specialised min-sum kernels.  LDPC_PREC_F32 (default) is the
 * reference's arithmetic type (bp/bp.py runs in torch.float, ofdm_functions.py:156).
 * LDPC_PREC_F16X2 packs two codewords per thread in half2 registers / shared-memory words
 * (min-sum and normalized min-sum only; sum-product, offset min-sum and warm starts keep fp32);
 * its results are defined bit-exactly by oracle/bp_oracle.py::bp_decode_f16. */
enum { LDPC_PREC_F32 = 0, LDPC_PREC_F16X2 = 1 };
int ldpc_code_set_precision(ldpc_code_t *code, int precision);

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

/* ldpc_decode_weighted - ldpc_decode with the reference's TRAINABLE weights (bp/bp_vc.py:16-32, 60-122: input_weight
 * per (output edge, input edge) pair of a variable and llr_weight per variable, one set per iteration layer plus the
 * final layer, bp/bp.py:26-39): V->C = 0.5 (w_llr[v] llr' + sum_{j != k} w_edge[e_k][j] x_j), marginal likewise.
 * DEVICE tables: w_edge [iters][E][w_stride] (row = variable-major output edge, column j = weight of the variable's
 * j-th edge as input, w_stride >= max_dv), w_llr [iters][n], wf_edge [E] (variable-major), wf_llr [n].
 * Runs on the register-resident kernel for the default (64,32) code (sum-product / min-sum), else on the generic kernel. */
int ldpc_decode_weighted(const ldpc_code_t *code, const void *llr, int llr_dtype, int64_t B, int iters, int update,
                         float clamp_value, float param, const float *w_edge, const float *w_llr,
                         const float *wf_edge, const float *wf_llr, int w_stride, float *prob, float *llr_post,
                         uint8_t *hard, uint8_t *hard_packed, int32_t *syndrome, float *x_out,
                         ldpc_stream_t stream);

/* ldpc_bp_train_forward / ldpc_bp_train_backward - the TRAINING path of the weighted decoder (sum-product): what
 * autograd does in the reference through BeliefPropagationVC_Function / BeliefPropagationCV_Function (bp/bp_vc.py:16-58,
 * bp/bp_cv.py:22-96, unrolled by bp/bp.py:43-51; joint training loop ofdm/ofdm_nn.py:257-396) with dense [E,E] masks
 * and a [B,E,E,E] intermediate.  forward: prob [B,n] = P(bit=1), identical to ldpc_decode_weighted, and the tape
 * ((iters+1)*E*B floats, layout private to the pair) of C->V messages entering every iteration (x0 [B,E] check-major or null = zeros).
 * backward: given grad_prob [B,n] writes grad_llr [B,n] and the batch-summed weight gradients g_w_edge
 * [iters][E][w_stride], g_w_llr [iters][n], g_wf_edge [E], g_wf_llr [n] (overwritten; unused (k,k) / padding entries
 * stay 0).  workspace: 2*E*B floats.  All DEVICE pointers, f32, asynchronous on `stream`.  The gradient is the exact
 * derivative of the forward (clamps pass it inside or on the bound, like torch.clamp); the weight sums use float
 * atomics, so their last bits depend on the order of arrival. */
int ldpc_bp_train_forward(const ldpc_code_t *code, const float *llr, int64_t B, int iters, float clamp_value,
                          const float *w_edge, const float *w_llr, const float *wf_edge, const float *wf_llr,
                          int w_stride, const float *x0, float *prob, float *tape, ldpc_stream_t stream);
int ldpc_bp_train_backward(const ldpc_code_t *code, const float *llr, int64_t B, int iters, float clamp_value,
                           const float *w_edge, const float *w_llr, const float *wf_edge, const float *wf_llr,
                           int w_stride, const float *tape, const float *grad_prob, float *grad_llr,
                           float *g_w_edge, float *g_w_llr, float *g_wf_edge, float *g_wf_llr, float *workspace,
                           ldpc_stream_t stream);

/* ldpc_decode_ex - ldpc_decode with the extensible parameter block; adds syndrome-based early
 * termination (north star; NOT in the reference, whose iteration count is fixed, bp/bp.py:46-47,
 * so it is off in every parity run).  With early_exit != 0 a codeword is frozen after the first
 * iteration whose hard decision satisfies every check; its outputs are those of that iteration.
 * iters_used [B] i32 (nullable) receives the iterations run per codeword.  The f16x2 kernel does
 * not implement early termination (the fp32 kernel is used). */
typedef struct {
    int32_t struct_size;      /* = sizeof(ldpc_decode_params_t) */
    int32_t llr_dtype;
    const void *llr;
    int64_t B;
    int32_t iters, update;
    float clamp_value, param;
    const float *x0;
    float *prob, *llr_post;
    uint8_t *hard, *hard_packed;
    int32_t *syndrome;
    float *x_out;
    int32_t early_exit, reserved;
    int32_t *iters_used;
} ldpc_decode_params_t;
int ldpc_decode_ex(const ldpc_code_t *code, const ldpc_decode_params_t *params, ldpc_stream_t stream);

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

/* ldpc_decode_bits_host - decode_bits (ofdm/ofdm_functions.py:131-163) for ORDINARY (pageable) host arrays: llr_host
 * [N,n] of llr_dtype (the reference passes float64 and casts it to f32, :156), bits_out [N,n] of {0,1} as out_dtype
 * (LDPC_F64 = the reference's float64 result, :161; LDPC_F32; LDPC_I8 = one byte per bit).  Host threads cast / copy each
 * chunk into pinned staging while the previous chunk is on the GPU, only packed bits return over PCIe, and host threads
 * expand them into bits_out while the next chunk decodes.  threads <= 0: min(16, hardware threads).  Synchronous. */
int ldpc_decode_bits_host(const ldpc_code_t *code, const void *llr_host, int llr_dtype, int64_t N, int iters,
                          int update, float clamp_value, float param, void *bits_out, int out_dtype,
                          int64_t chunk_codewords, int threads);

/* ldpc_count_errors - exact integer link metrics, accumulated (+=) into counters[5] (i64):
 *   {uncoded bit errors over n, decoded info-bit errors over the first k, frame errors
 *    (any of n decoded bits wrong), bits = B*n, frames = B}
 * Replaces evaluate_quantized_snr.py:169-188 / compute_ber (ofdm/ofdm_functions.py:83-84).
 *   llr [B,n] channel LLRs (hard decision (sign+1)//2, llr==0 -> 0) or NULL to skip the
 *   uncoded count; hard [B,n] u8 decoded bits; ref_bits [B,n] u8 transmitted codeword. */
int ldpc_count_errors(const void *llr, int llr_dtype, const uint8_t *hard, const uint8_t *ref_bits,
                      int64_t B, int n, int k, int64_t *counters, ldpc_stream_t stream);

/* ==== link-simulator front end (K2).  Complex arrays are interleaved (re, im) pairs of
 * real_dtype (LDPC_F64 = numpy complex128, LDPC_F32 = complex64); an array of L complex values
 * is a sequence of OFDM symbols, symbol j occupying [j*N, (j+1)*N) - the reference's flat
 * (1, L) layout (it reshapes to (-1, N).T and back).  ofdm_size N in {32, 64, 128, 256};
 * transforms are unitary (1/sqrt(N)) like DFT(N), ofdm/ofdm_functions.py:86-93. ==== */

/* encode_bits (ofdm/ofdm_functions.py:11-15): out[c, :] = G bits[c, :] mod 2.
 *   bits [ncw,k] u8 0/1; G_packed [n, ceil(k/32)] u32, bit j of word w = G[r][32 w + j];
 *   out [ncw,n] u8. */
int ldpc_encode_bits(const uint8_t *bits, const uint32_t *G_packed, int n, int k, int64_t ncw,
                     uint8_t *out, ldpc_stream_t stream);

/* modulate_bits (ofdm/ofdm_functions.py:17-22): bit pairs (b0,b1) -> ((1-2 b0) + j(1-2 b1))/sqrt(2).
 *   bits [2*n_symbols] u8; out [n_symbols] complex. */
int ldpc_modulate_bits(const uint8_t *bits, int64_t n_symbols, int real_dtype, void *out,
                       ldpc_stream_t stream);

/* transmit_symbols (ofdm/ofdm_functions.py:25-35): tx = W^H s per OFDM symbol, rx = tx + noise.
 *   noise: [n_ofdm*N] complex, ALREADY scaled (the reference's (N(0,1/sqrt(snr)) + j N(..))/sqrt(2)),
 *   or NULL to draw it on the device from Philox4x32-10(seed) with that variance.
 *   tx may be NULL. */
int ldpc_ofdm_transmit(const void *symbols, int64_t n_ofdm, int ofdm_size, int real_dtype,
                       const void *noise, double snr, uint64_t seed, void *rx, void *tx,
                       ldpc_stream_t stream);

/* quantizer (ofdm/ofdm_functions.py:37-51) applied to n_real real values (a complex array is
 * 2 reals per element): step = 2 clip/(L-1); q = step*floor(x/step + .5);
 * clip(q, -(L/2) step + 1, (L/2) step - 1) - the +-1 in SIGNAL units is the reference's
 * behaviour and is reproduced.  num_levels = 2^num_bits. */
int ldpc_quantize(const void *in, int64_t n_real, int real_dtype, double num_levels, double clip,
                  void *out, ldpc_stream_t stream);

/* demodulate_signal (ofdm/ofdm_functions.py:63-78): R = W r per OFDM symbol;
 *   llrs [2*n_ofdm*N] real, interleaved (b0,b1) per subcarrier,
 *   llr = ((R - a)^2 - (R + a)^2) / (2 * 0.5/snr_est), a = 1/sqrt(2)  (log P1/P0);
 *   symbols [n_ofdm*N] complex.  Either output may be NULL. */
int ldpc_ofdm_demodulate(const void *signal, int64_t n_ofdm, int ofdm_size, int real_dtype,
                         double snr_est, void *llrs, void *symbols, ldpc_stream_t stream);

/* ==== fused Monte-Carlo link simulation (replaces the per-SNR loop body of
 * evaluate_quantized_snr.py:91-188: create_bits, encode_bits, modulate_bits, gen_data, the
 * inline AGC quantizer, decode_bits and the BER/BLER means) ==== */

/* Philox4x32-10 block function exactly as the simulator's kernels use it (host-callable, for
 * known-answer tests): counter = {low32(codeword), high32(codeword), stream, block},
 * streams: 0 = information bits, 1 = channel noise. */
void ldpc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Attach the systematic encoder: parity rows of G (rows k..n-1 of the [n,k] generator,
 * ofdm/ofdm_functions.py:11-15 / bp/parity.py:44), bit-packed [n-k, ceil(k/32)] u32 on the HOST. */
int ldpc_code_set_generator(ldpc_code_t *code, const uint32_t *parity_rows_packed_host, int k);

typedef struct {
    int32_t struct_size;    /* = sizeof(ldpc_sim_params_t) */
    int32_t ofdm_size;      /* 32 (every reference script), 64, 128, 256 */
    int32_t qbits;          /* 0 = unquantized (gen_data path); > 0 = low-resolution ADC */
    int32_t agc_mode;       /* 1 = factor = agc_clip/(0.5(1+1/snr))*clip_ratio, clip = agc_clip
                               (evaluate_quantized_snr.py:103-111);
                               2 = clip = sqrt(1+1/snr)*clip_ratio, the analytic value of gen_qdata's
                               std(rx_signal)*clip_ratio (ofdm/ofdm_functions.py:121-123) */
    float agc_clip;         /* 10 in the reference scripts */
    float clip_ratio;
    float snr_db;           /* per-subcarrier Es/N0 in dB (ofdm/ofdm_functions.py:110) */
    int32_t iters;          /* BP iterations */
    int32_t update;         /* LDPC_UPDATE_* */
    float clamp_value;
    float param;
    int32_t reserved;       /* option bits.  bit 0: force the three-launch chain (A/B test of the single-launch kernel);
                               bit 1: flat Rayleigh block fading - one CN(0,1) gain per OFDM symbol (Philox stream 2), coherent
                                      receiver with perfect channel knowledge: r = h x + n, z = r / h, LLR noise power sigma^2 / |h|^2
                                      (north star "AWGN/fading channel"; the reference has AWGN only, ofdm_functions.py:30-33);
                               bit 2: tanh compander clip * tanh(x / clip) in front of the uniform ADC (north star "uniform/tanh
                                      quantizer"; the reference's quantizer is ofdm_functions.py:37-51) */
    uint64_t seed;          /* Philox key */
    int64_t first_codeword; /* global index of the first codeword (Philox subsequence) */
    int64_t n_codewords;
} ldpc_sim_params_t;

/* Run the link for codewords [first_codeword, first_codeword + n_codewords) and ADD the exact
 * integer metrics into counters[5] (i64, device): {uncoded bit errors, info-bit errors, frame
 * errors, bits, frames} (evaluate_quantized_snr.py:169-188).  Each codeword is framed into
 * ceil((n/2)/N) OFDM symbols (null subcarriers after its last QPSK symbol; identical to the
 * reference when n = 2N).  For a code with a compiled specialisation the whole chain runs in ONE
 * kernel launch (front end fused in front of the decoder, nothing touches HBM but the counters);
 * otherwise three launches per chunk (generate, link, decode+count) with `workspace` as device
 * scratch for the LLR tile (>= 1024 codewords of 4 n + ceil(n/8) bytes; the run is chunked to
 * fit).  Both paths give identical counters; they depend only on (seed, global codeword index),
 * never on the sharding. */
int ldpc_sim_run(const ldpc_code_t *code, const ldpc_sim_params_t *params, void *workspace,
                 size_t workspace_bytes, int64_t *counters, ldpc_stream_t stream);

/* The front end alone (K2): transmitted codewords, MSB-first packed [n_codewords, ceil(n/8)],
 * and channel LLRs f32 [n_codewords, n], for tests and for feeding ldpc_decode directly. */
int ldpc_sim_generate(const ldpc_code_t *code, const ldpc_sim_params_t *params, uint8_t *cw_packed,
                      float *llr, ldpc_stream_t stream);

/* ldpc_sim_generate plus the received time-domain samples (nullable) that the MLP demappers take as
 * input (evaluate_quantized_snr.py:135-140): f32 [n_codewords * ceil((n/2)/ofdm_size), 2 ofdm_size + 1],
 * one row per OFDM symbol = Re[0..N), Im[0..N) of the (noisy, quantized, rescaled) time signal and the
 * linear SNR. */
int ldpc_sim_generate_ex(const ldpc_code_t *code, const ldpc_sim_params_t *params, uint8_t *cw_packed,
                         float *llr, float *samples, ldpc_stream_t stream);

/* The same link chain on CALLER-SUPPLIED transmitted codewords (MSB-first packed, device) and, when `noise` is
 * non-NULL, caller-supplied additive noise: f32 (re, im) pairs [n_codewords][ceil((n/2)/ofdm_size)][ofdm_size]
 * (device), one pair per received time sample, already scaled (the reference's own draw,
 * ofdm/ofdm_functions.py:30-33).  With noise == NULL the Philox stream of ldpc_sim_generate is used.  This is the
 * identical-input form of the AGC-scaled quantizer front end (evaluate_quantized_snr.py:96-133): everything after
 * the addition - AGC, ADC, rescale, FFT, LLR - is the simulator's own device code. */
int ldpc_sim_frontend(const ldpc_code_t *code, const ldpc_sim_params_t *params, const uint8_t *cw_packed,
                      const float *noise, float *llr, float *samples, ldpc_stream_t stream);

/* ldpc_decode_count - ldpc_decode with the exact link metrics fused at its tail
 * (evaluate_quantized_snr.py:169-188): decodes llr [B,n] and ADDS {uncoded bit errors, info-bit errors,
 * frame errors, bits, frames} into counters[5] (i64, device) against ref_packed (transmitted
 * codewords, MSB-first [B, ceil(n/8)], device); k = number of information bits. */
int ldpc_decode_count(const ldpc_code_t *code, const void *llr, int llr_dtype, int64_t B, int iters,
                      int update, float clamp_value, float param, const uint8_t *ref_packed, int k,
                      int64_t *counters, ldpc_stream_t stream);

/* ---- MLP demapper (next row of the scope table: pytorch/nn/llr.py:7-73) ---------------------------
 * The reference's LLR estimators are chains of nn.Linear (+ tanh) evaluated in fp32
 * (LLRestimator_withSNR: [2N+1] -> 16N -> 16N -> 16N -> 2N, nn/llr.py:54-73; called from
 * evaluate_quantized_snr.py:150-160).  ldpc_mlp_create uploads the weights (HOST pointers,
 * weights[l] = nn.Linear.weight [dims[l+1], dims[l]] row-major, biases[l] = [dims[l+1]] or NULL,
 * activations[l] != 0 -> tanh after layer l; NULL = tanh after every layer but the last) and
 * splits them exactly into `splits` binary16 planes for the tensor cores (2 = fp32-equivalent:
 * 22+ significant bits, 3 MMAs per product; 3 = 33 bits, 6 MMAs; 1 = plain fp16).  Operands must stay
 * below 65504 in magnitude.  Every dims[l+1] must be a multiple of 64.
 * ldpc_mlp_forward: x [B, dims[0]] f32 row-major DEVICE -> y [B, dims[n_layers]] f32 DEVICE,
 * asynchronous on `stream`, chunked internally (chunk_rows, 0 = default).  A handle owns scratch
 * buffers: use it from one stream at a time. */
typedef struct ldpc_mlp ldpc_mlp_t;
int ldpc_mlp_create(int n_layers, const int32_t *dims, const float *const *weights,
                    const float *const *biases, const int32_t *activations, int splits,
                    int64_t chunk_rows, ldpc_mlp_t **out);
int ldpc_mlp_forward(ldpc_mlp_t *mlp, const float *x, int64_t B, float *y, ldpc_stream_t stream);
void ldpc_mlp_destroy(ldpc_mlp_t *mlp);
/* How ldpc_mlp_forward schedules the chain (results are bit-identical):
 * LDPC_MLP_PER_LAYER - one launch per layer, activation planes of a chunk ping-pong through device memory;
 * LDPC_MLP_CHAIN     - one cooperative launch per chunk: groups of 4 SMs carry blocks of 128 rows through all layers while
 *                      their activation planes are still in L2 (splits = 2, at most 6 layers of at most 512 outputs;
 *                      LDPC_EUNSUPPORTED otherwise);
 * LDPC_MLP_CHAIN_PAIRS - the chain on pairs of SMs (clusters of 2, cta_group::2 MMAs of M = 256, each SM staging half of every
 *                      weight operand): 4 % faster than LDPC_MLP_CHAIN on B200 with 1.4x its DRAM traffic;
 * LDPC_MLP_AUTO      - LDPC_MLP_CHAIN_PAIRS where it applies and the batch has at least 32 768 rows, else LDPC_MLP_CHAIN, else
 *                      LDPC_MLP_PER_LAYER (default). */
enum { LDPC_MLP_AUTO = 0, LDPC_MLP_PER_LAYER = 1, LDPC_MLP_CHAIN = 2, LDPC_MLP_CHAIN_PAIRS = 3 };
int ldpc_mlp_set_mode(ldpc_mlp_t *mlp, int mode);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
