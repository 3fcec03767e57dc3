// decode_qc.cu - code-specialised belief-propagation decoder for quasi-cyclic codes whose
// prototype matrix is known at compile time (IEEE 802.11n n=1944 R=1/2 Z=81 first).
//
// The parity-check matrix is compiled INTO the instruction stream: every block's shift and
// shared-memory offset is an immediate, every node degree a compile-time loop bound, so the
// inner loops carry no index loads at all (the reference multiplies by dense E x E masks,
// bp/masking.py:12-147, bp/bp_vc.py:19, bp/bp_cv.py:24-42).
//
// Mapping: thread = (codeword cw of the CTA's tile, lane z in [0,Z)).  Messages live in
// shared memory for all iterations, one fp32 slot per edge at  blk*Z + zc  (zc = the
// CHECK's lane), so the check phase is a pure linear access and the variable phase reads
// a rotated window (z - shift mod Z).  Channel LLRs sit beside them in natural order.
// Per iteration: variable phase (NB unrolled block columns per thread), barrier, check
// phase (MB unrolled block rows per thread), barrier.  Arithmetic = node_math.cuh, so the
// results are bit-identical to the generic kernel and to the CPU oracle's definition.
#include <utility>

#include "common.cuh"
#include "epilogue.cuh"
#include "node_math.cuh"

namespace ldpc {

// ---- compile-time prototype matrices ---------------------------------------------------------
struct Wifi1944R12 {
    static constexpr int Z = 81, MB = 12, NB = 24;
    static constexpr int16_t proto[MB][NB] = {
        {57, -1, -1, -1, 50, -1, 11, -1, 50, -1, 79, -1, 1, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1},
        {3, -1, 28, -1, 0, -1, -1, -1, 55, 7, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1},
        {30, -1, -1, -1, 24, 37, -1, -1, 56, 14, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1, -1, -1, -1},
        {62, 53, -1, -1, 53, -1, -1, 3, 35, -1, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1, -1, -1},
        {40, -1, -1, 20, 66, -1, -1, 22, 28, -1, -1, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1, -1},
        {0, -1, -1, -1, 8, -1, 42, -1, 50, -1, -1, 8, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1, -1},
        {69, 79, 79, -1, -1, -1, 56, -1, 52, -1, -1, -1, 0, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1, -1},
        {65, -1, -1, -1, 38, 57, -1, -1, 72, -1, 27, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, -1, -1, -1},
        {64, -1, -1, -1, 14, 52, -1, -1, 30, -1, -1, 32, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, -1, -1},
        {-1, 45, -1, 70, 0, -1, -1, -1, 77, 9, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, -1},
        {2, 56, -1, 57, 35, -1, -1, -1, -1, -1, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0},
        {24, -1, 61, -1, 60, -1, -1, 27, 51, -1, -1, 16, 1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0}};
};

// Derived tables, all evaluated at compile time.
template <class Code>
struct QcTables {
    static constexpr int Z = Code::Z, MB = Code::MB, NB = Code::NB;
    int nblk = 0;
    int row_deg[MB] = {}, row_col[MB][NB] = {}, row_shift[MB][NB] = {}, row_blk[MB][NB] = {};
    int col_deg[NB] = {}, col_row[NB][MB] = {}, col_shift[NB][MB] = {}, col_blk[NB][MB] = {};
    int max_dv = 0, max_dc = 0;
    constexpr QcTables() {
        int id = 0;
        for (int r = 0; r < MB; ++r)
            for (int c = 0; c < NB; ++c)
                if (Code::proto[r][c] >= 0) {
                    const int j = row_deg[r]++;
                    row_col[r][j] = c; row_shift[r][j] = Code::proto[r][c]; row_blk[r][j] = id;
                    const int k = col_deg[c]++;
                    col_row[c][k] = r; col_shift[c][k] = Code::proto[r][c]; col_blk[c][k] = id;
                    ++id;
                }
        nblk = id;
        for (int r = 0; r < MB; ++r) if (row_deg[r] > max_dc) max_dc = row_deg[r];
        for (int c = 0; c < NB; ++c) if (col_deg[c] > max_dv) max_dv = col_deg[c];
    }
};

template <class Code>
inline constexpr QcTables<Code> kQc{};

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    static_for_impl(static_cast<F &&>(f), std::make_integer_sequence<int, N>{});
}

template <class Code, int CW>
struct QcLayout {
    static constexpr int Z = Code::Z;
    static constexpr int N = Code::NB * Z;
    static constexpr int M = Code::MB * Z;
    static constexpr int E = kQc<Code>.nblk * Z;
    // codeword strides == Z (mod 32): lanes of two codewords sharing a warp stay on distinct banks
    static constexpr int pad_to(int v) { return v + ((Z % 32) - (v % 32) + 32) % 32; }
    static constexpr int LLR_STRIDE = pad_to(N);
    static constexpr int MSG_STRIDE = pad_to(E);
    static constexpr int HARD_STRIDE = (N + 15) & ~15;
    static constexpr int THREADS = ((CW * Z + 31) / 32) * 32;
    static constexpr size_t SMEM = sizeof(float) * CW * (LLR_STRIDE + MSG_STRIDE) + (size_t)CW * HARD_STRIDE +
                                   sizeof(int) * (8 + CW);
};

template <class Code, int CW, int UPD>
__global__ void __launch_bounds__((QcLayout<Code, CW>::THREADS)) decode_qc_kernel(const DecodeArgs a) {
    using L = QcLayout<Code, CW>;
    constexpr bool IS_SP = (UPD == UPD_SP);
    constexpr int Z = Code::Z, NB = Code::NB, MB = Code::MB, N = L::N, M = L::M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *llr_s = reinterpret_cast<float *>(smem_raw);
    float *msg_s = llr_s + CW * L::LLR_STRIDE;
    uint8_t *hard_s = reinterpret_cast<uint8_t *>(msg_s + CW * L::MSG_STRIDE);
    int *scratch = reinterpret_cast<int *>(hard_s + CW * L::HARD_STRIDE);       // [4 + CW]

    const int tid = threadIdx.x, T = blockDim.x;
    const long long cw0 = (long long)blockIdx.x * CW;
    const int ncw = (int)min((long long)CW, a.B - cw0);
    const int cw = tid / Z, z = tid - cw * Z;
    const bool active = cw < ncw;            // also false for the padding threads (cw >= CW)

    // ---- load LLR tile, clear messages --------------------------------------------------------
    for (int i = tid; i < ncw * N; i += T) {
        const int c = i / N, v = i - c * N;
        llr_s[c * L::LLR_STRIDE + v] = load_llr(a.llr, a.llr_dtype, cw0 * N + i);
    }
    for (int i = tid; i < CW * L::MSG_STRIDE; i += T) msg_s[i] = 0.0f;
    for (int i = tid; i < 4 + CW; i += T) scratch[i] = 0;
    __syncthreads();

    float *const msg = msg_s + (active ? cw : 0) * L::MSG_STRIDE;
    const float *const llr = llr_s + (active ? cw : 0) * L::LLR_STRIDE;
    // rotated window bases: slot (blk, (z - s) mod Z) = (z < s ? hi : lo)[blk*Z - s]
    float *const lo = msg + z;
    float *const hi = msg + z + Z;

    for (int it = 0; it < a.iters; ++it) {
        if (active) {
            static_for<NB>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                constexpr int D = kQc<Code>.col_deg[c];
                if constexpr (D > 0) {
                    float in[D], out[D];
                    float *ptr[D];
                    static_for<D>([&](auto kk) {
                        constexpr int k = decltype(kk)::value;
                        constexpr int s = kQc<Code>.col_shift[c][k];
                        constexpr int off = kQc<Code>.col_blk[c][k] * Z - s;
                        ptr[k] = (z < s ? hi : lo) + off;
                        in[k] = *ptr[k];
                    });
                    var_node<D, IS_SP>(in, D, llr[c * Z + z], out);
                    static_for<D>([&](auto kk) { *ptr[decltype(kk)::value] = out[decltype(kk)::value]; });
                }
            });
        }
        __syncthreads();
        if (active) {
            static_for<MB>([&](auto rr) {
                constexpr int r = decltype(rr)::value;
                constexpr int D = kQc<Code>.row_deg[r];
                if constexpr (D > 0) {
                    float in[D], out[D];
                    static_for<D>([&](auto jj) {
                        constexpr int j = decltype(jj)::value;
                        constexpr int off = kQc<Code>.row_blk[r][j] * Z;
                        in[j] = msg[off + z];
                    });
                    if constexpr (IS_SP) check_node_sp<D>(in, D, a.clampv, out);
                    else check_node_ms_ct<D, UPD>(in, a.clampv, a.param, out);
                    static_for<D>([&](auto jj) {
                        constexpr int j = decltype(jj)::value;
                        constexpr int off = kQc<Code>.row_blk[r][j] * Z;
                        msg[off + z] = out[j];
                    });
                }
            });
        }
        __syncthreads();
    }

    // ---- marginal, P(bit=1), hard decision --------------------------------------------------------
    if (active) {
        const long long obase = (cw0 + cw) * N;
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            float in[D > 0 ? D : 1];
            static_for<D>([&](auto kk) {
                constexpr int k = decltype(kk)::value;
                constexpr int s = kQc<Code>.col_shift[c][k];
                constexpr int off = kQc<Code>.col_blk[c][k] * Z - s;
                in[k] = ((z < s ? hi : lo) + off)[0];
            });
            const float t = marginal_t<(D > 0 ? D : 1)>(in, D, llr[c * Z + z]);
            const float pr = prob_one(t);
            const uint8_t hb = hard_bit(t);
            hard_s[cw * L::HARD_STRIDE + c * Z + z] = hb;
            const long long o = obase + c * Z + z;
            if (a.prob) a.prob[o] = pr;
            if (a.llr_post) a.llr_post[o] = __fmul_rn(-2.0f, t);
            if (a.hard) a.hard[o] = hb;
        });
    }
    __syncthreads();

    // ---- syndrome weight ------------------------------------------------------------------------
    if (a.syndrome) {
        if (active) {
            int w = 0;
            const uint8_t *h = hard_s + cw * L::HARD_STRIDE;
            static_for<MB>([&](auto rr) {
                constexpr int r = decltype(rr)::value;
                constexpr int D = kQc<Code>.row_deg[r];
                unsigned par = 0;
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr int s = kQc<Code>.row_shift[r][j];
                    int zv = z + s;
                    if (zv >= Z) zv -= Z;
                    constexpr int cbase = kQc<Code>.row_col[r][j] * Z;
                    par ^= h[cbase + zv];
                });
                w += (int)par;
            });
            if (w) atomicAdd(&scratch[4 + cw], w);
        }
        __syncthreads();
        for (int i = tid; i < ncw; i += T) a.syndrome[cw0 + i] = scratch[4 + i];
        __syncthreads();
        for (int i = tid; i < CW; i += T) scratch[4 + i] = 0;
    }
    if (a.hard_packed) pack_hard(hard_s, L::HARD_STRIDE, ncw, N, a.hard_packed + cw0 * ((N + 7) >> 3));
    if (a.counters) {
        __syncthreads();
        count_errors(llr_s, L::LLR_STRIDE, hard_s, L::HARD_STRIDE, ncw, N, a.k_info,
                     a.ref_packed + cw0 * ((N + 7) >> 3), a.counters, scratch + 1);
    }
    (void)M;
}

// ---- registry of compiled specialisations ----------------------------------------------------------
template <class Code>
static bool proto_matches(int Z, int mb, int nb, const int16_t *proto) {
    if (Z != Code::Z || mb != Code::MB || nb != Code::NB) return false;
    for (int r = 0; r < mb; ++r)
        for (int c = 0; c < nb; ++c)
            if (proto[r * nb + c] != Code::proto[r][c]) return false;
    return true;
}

int qc_lookup(int Z, int mb, int nb, const int16_t *proto) {
    if (proto_matches<Wifi1944R12>(Z, mb, nb, proto)) return 0;
    return -1;
}

template <class Code, int CW>
static int launch_qc_t(const DecodeArgs &a, cudaStream_t s) {
    using L = QcLayout<Code, CW>;
    const long long grid = (a.B + CW - 1) / CW;
    if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
    void (*k)(const DecodeArgs) = nullptr;
    switch (a.update) {
        case UPD_SP: k = decode_qc_kernel<Code, CW, UPD_SP>; break;
        case UPD_MINSUM: k = decode_qc_kernel<Code, CW, UPD_MINSUM>; break;
        case UPD_NMS: k = decode_qc_kernel<Code, CW, UPD_NMS>; break;
        default: k = decode_qc_kernel<Code, CW, UPD_OMS>; break;
    }
    LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
    k<<<(int)grid, L::THREADS, L::SMEM, s>>>(a);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

int launch_decode_qc(int qc_id, const DecodeArgs &a, cudaStream_t s) {
    if (a.B <= 0) return LDPC_OK;
    switch (qc_id) {
        case 0: return launch_qc_t<Wifi1944R12, 3>(a, s);
        default: set_error("unknown QC specialisation %d", qc_id); return LDPC_EINVAL;
    }
}

}  // namespace ldpc
