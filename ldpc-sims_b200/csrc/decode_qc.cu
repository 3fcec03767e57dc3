// decode_qc.cu - code-specialised belief-propagation decoder for quasi-cyclic codes whose
// prototype matrix is known at compile time (IEEE 802.11n n=1944 R=1/2 Z=81 first).
//
// The parity-check matrix is compiled INTO the instruction stream: every block's shift and
// shared-memory offset is an immediate, every node degree a compile-time loop bound, so the
// inner loops carry no index loads at all (the reference multiplies by dense E x E masks,
// bp/masking.py:12-147, bp/bp_vc.py:19, bp/bp_cv.py:24-42).
//
// Mapping: thread = (codeword cw of the CTA's tile, lane t in [0,Z)); see QcPlan for which
// check / variable of each block row / column a lane computes.  Channel LLRs (NB per thread)
// and the messages of the thread-local blocks stay in REGISTERS for the whole decode; the
// other blocks keep one fp32 slot per edge in shared memory at  blk*Z + tc  (tc = the check's
// thread), so the check phase is a pure linear access and the variable phase reads a rotated
// window ((t - s') mod Z).  HBM traffic is the LLR load and the result store only.
// Per iteration: variable phase (NB unrolled block columns per thread), barrier, check
// phase (MB unrolled block rows per thread), barrier.  Arithmetic = node_math.cuh, so the
// results are bit-identical to the generic kernel and to the CPU oracle's definition.
#include <cstdlib>
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

// ---- compile-time plan ------------------------------------------------------------------------------
// Lane relabelling.  Thread t of a codeword processes check (r, (t + sigma_r) mod Z) of every
// block row r and variable (c, (t + rho_c) mod Z) of every block column c.  An edge of block
// (r, c, shift s) then joins check-thread tc with variable-thread tc + s', where
//     s' = (s + sigma_r - rho_c) mod Z.
// Blocks with s' == 0 connect a thread to ITSELF: their messages never leave the register
// file.  A spanning tree of the (block row, block column) graph fixes sigma/rho so that
// MB + NB - 1 blocks (35 of the 86 for 802.11n n=1944) become thread-local; only the remaining
// blocks are exchanged through shared memory.  Which node a thread computes does not change
// the node's arithmetic, so results stay bit-identical to the generic kernel.
template <class Code>
struct QcPlan {
    static constexpr int Z = Code::Z, MB = Code::MB, NB = Code::NB;
    int sigma[MB] = {}, rho[NB] = {};
    int nblk = 0, n_local = 0, n_smem = 0;
    int row_deg[MB] = {}, row_col[MB][NB] = {}, row_eff[MB][NB] = {}, row_slot[MB][NB] = {}, row_shift[MB][NB] = {};
    int col_deg[NB] = {}, col_row[NB][MB] = {}, col_eff[NB][MB] = {}, col_slot[NB][MB] = {};
    bool row_loc[MB][NB] = {}, col_loc[NB][MB] = {};
    constexpr QcPlan() {
        // breadth-first spanning tree over block rows / block columns
        bool row_seen[MB] = {}, col_seen[NB] = {};
        int queue[MB + NB] = {}, head = 0, tail = 0;     // entries: r (>=0) or -(c+1)
        for (int root = 0; root < MB; ++root) {
            if (row_seen[root]) continue;
            row_seen[root] = true; sigma[root] = 0; queue[tail++] = root;
            while (head < tail) {
                const int q = queue[head++];
                if (q >= 0) {
                    const int r = q;
                    for (int c = 0; c < NB; ++c)
                        if (Code::proto[r][c] >= 0 && !col_seen[c]) {
                            col_seen[c] = true;
                            rho[c] = (Code::proto[r][c] + sigma[r]) % Z;
                            queue[tail++] = -(c + 1);
                        }
                } else {
                    const int c = -q - 1;
                    for (int r = 0; r < MB; ++r)
                        if (Code::proto[r][c] >= 0 && !row_seen[r]) {
                            row_seen[r] = true;
                            sigma[r] = ((rho[c] - Code::proto[r][c]) % Z + Z) % Z;
                            queue[tail++] = r;
                        }
                }
            }
        }
        for (int r = 0; r < MB; ++r)
            for (int c = 0; c < NB; ++c)
                if (Code::proto[r][c] >= 0) {
                    const int eff = ((Code::proto[r][c] + sigma[r] - rho[c]) % Z + Z) % Z;
                    const bool loc = (eff == 0);
                    const int slot = loc ? n_local++ : n_smem++;
                    const int j = row_deg[r]++;
                    row_col[r][j] = c; row_eff[r][j] = eff; row_slot[r][j] = slot; row_loc[r][j] = loc;
                    row_shift[r][j] = Code::proto[r][c];
                    const int k = col_deg[c]++;
                    col_row[c][k] = r; col_eff[c][k] = eff; col_slot[c][k] = slot; col_loc[c][k] = loc;
                    ++nblk;
                }
    }
};

template <class Code>
inline constexpr QcPlan<Code> kQc{};

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
    static constexpr int NLOC = kQc<Code>.n_local, NSM = kQc<Code>.n_smem;
    // codeword strides == Z (mod 32): lanes of two codewords sharing a warp stay on distinct banks
    static constexpr int pad_to(int v) { return v + ((Z % 32) - (v % 32) + 32) % 32; }
    static constexpr int MSG_STRIDE = pad_to(NSM * Z);
    static constexpr int HARD_STRIDE = (N + 15) & ~15;
    static constexpr int THREADS = ((CW * Z + 31) / 32) * 32;
    static constexpr size_t SMEM = sizeof(float) * CW * MSG_STRIDE + (size_t)CW * HARD_STRIDE + sizeof(int) * (8 + CW);
};

template <class Code, int CW, int UPD>
__global__ void __launch_bounds__((QcLayout<Code, CW>::THREADS)) decode_qc_kernel(const DecodeArgs a) {
    using L = QcLayout<Code, CW>;
    constexpr bool IS_SP = (UPD == UPD_SP);
    constexpr int Z = Code::Z, NB = Code::NB, MB = Code::MB, N = L::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *msg_s = reinterpret_cast<float *>(smem_raw);
    uint8_t *hard_s = reinterpret_cast<uint8_t *>(msg_s + CW * L::MSG_STRIDE);
    int *scratch = reinterpret_cast<int *>(hard_s + CW * L::HARD_STRIDE);       // [4 + CW]

    const int tid = threadIdx.x, T = blockDim.x;
    const long long cw0 = (long long)blockIdx.x * CW;
    const int ncw = (int)min((long long)CW, a.B - cw0);
    const int cw = tid / Z, t = tid - cw * Z;
    const bool active = cw < ncw;            // also false for the padding threads (cw >= CW)
    for (int i = tid; i < 4 + CW; i += T) scratch[i] = 0;

    float *const msg = msg_s + (active ? cw : 0) * L::MSG_STRIDE;
    // rotated window bases: slot (blk, (t - s') mod Z) = (t < s' ? hi : lo)[blk*Z - s']
    float *const lo = msg + t;
    float *const hi = msg + t + Z;

    // ---- channel LLRs of this thread's NB variables live in registers for the whole decode ----------
    float llr[NB];
    float loc[L::NLOC > 0 ? L::NLOC : 1];
    const long long gbase = (cw0 + (active ? cw : 0)) * N;
    if (active) {
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int rho = kQc<Code>.rho[c];
            int zv = t + rho;
            if (zv >= Z) zv -= Z;
            llr[c] = load_llr(a.llr, a.llr_dtype, gbase + c * Z + zv);
        });
    }

    // V -> C for all NB block columns of this thread.  FIRST: the C->V messages are still the
    // zeros every reference caller passes (ofdm_functions.py:157) - nothing is loaded.
    auto var_phase = [&](auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            if constexpr (D > 0) {
                float in[D], out[D];
                float *ptr[D];
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                    constexpr int slot = kQc<Code>.col_slot[c][k];
                    if constexpr (is_loc) {
                        ptr[k] = nullptr;
                        in[k] = FIRST ? 0.0f : loc[slot];
                    } else {
                        constexpr int s = kQc<Code>.col_eff[c][k];
                        constexpr int off = slot * Z - s;
                        ptr[k] = (t < s ? hi : lo) + off;
                        in[k] = FIRST ? 0.0f : *ptr[k];
                    }
                });
                var_node<D, IS_SP>(in, D, llr[c], out);
                static_for<D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                    constexpr int slot = kQc<Code>.col_slot[c][k];
                    if constexpr (is_loc) loc[slot] = out[k];
                    else *ptr[k] = out[k];
                });
            }
        });
    };
    auto check_phase = [&]() {
        static_for<MB>([&](auto rr) {
            constexpr int r = decltype(rr)::value;
            constexpr int D = kQc<Code>.row_deg[r];
            if constexpr (D > 0) {
                float in[D], out[D];
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr bool is_loc = kQc<Code>.row_loc[r][j];
                    constexpr int slot = kQc<Code>.row_slot[r][j];
                    if constexpr (is_loc) in[j] = loc[slot];
                    else in[j] = msg[slot * Z + t];
                });
                if constexpr (IS_SP) check_node_sp<D>(in, D, a.clampv, out);
                else check_node_ms_ct<D, UPD>(in, a.clampv, a.param, out);
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr bool is_loc = kQc<Code>.row_loc[r][j];
                    constexpr int slot = kQc<Code>.row_slot[r][j];
                    if constexpr (is_loc) loc[slot] = out[j];
                    else msg[slot * Z + t] = out[j];
                });
            }
        });
    };

    if (a.iters > 0) {
        if (active) var_phase(std::true_type{});
        __syncthreads();
        if (active) check_phase();
        __syncthreads();
    }
#pragma unroll 1
    for (int it = 1; it < a.iters; ++it) {
        if (active) var_phase(std::false_type{});
        __syncthreads();
        if (active) check_phase();
        __syncthreads();
    }

    // ---- marginal, P(bit=1), hard decision --------------------------------------------------------
    if (active) {
        static_for<NB>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr int D = kQc<Code>.col_deg[c];
            constexpr int rho = kQc<Code>.rho[c];
            float in[D > 0 ? D : 1];
            static_for<D>([&](auto kk) {
                constexpr int k = decltype(kk)::value;
                constexpr bool is_loc = kQc<Code>.col_loc[c][k];
                constexpr int slot = kQc<Code>.col_slot[c][k];
                if constexpr (is_loc) in[k] = (a.iters == 0) ? 0.0f : loc[slot];
                else {
                    constexpr int s = kQc<Code>.col_eff[c][k];
                    constexpr int off = slot * Z - s;
                    in[k] = (a.iters == 0) ? 0.0f : ((t < s ? hi : lo) + off)[0];
                }
            });
            const float tm = marginal_t<(D > 0 ? D : 1)>(in, D, llr[c]);
            const uint8_t hb = hard_bit(tm);
            int zv = t + rho;
            if (zv >= Z) zv -= Z;
            hard_s[cw * L::HARD_STRIDE + c * Z + zv] = hb | ((llr[c] > 0.0f) ? 2 : 0);
            const long long o = gbase + c * Z + zv;
            if (a.prob) a.prob[o] = prob_one(tm);
            if (a.llr_post) a.llr_post[o] = __fmul_rn(-2.0f, tm);
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
                constexpr int sg = kQc<Code>.sigma[r];
                int zc = t + sg;
                if (zc >= Z) zc -= Z;
                unsigned par = 0;
                static_for<D>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    constexpr int s = kQc<Code>.row_shift[r][j];
                    constexpr int cbase = kQc<Code>.row_col[r][j] * Z;
                    int zv = zc + s;
                    if (zv >= Z) zv -= Z;
                    par ^= h[cbase + zv] & 1u;
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
        count_errors(hard_s, L::HARD_STRIDE, ncw, N, a.k_info, a.ref_packed + cw0 * ((N + 7) >> 3), a.counters,
                     scratch + 1);
    }
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

void qc_plan_info(int qc_id, int out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    if (qc_id == 0) {
        using L = QcLayout<Wifi1944R12, 3>;
        out[0] = L::NLOC; out[1] = L::NSM; out[2] = L::THREADS; out[3] = 3;
    }
}

int qc_lookup(int Z, int mb, int nb, const int16_t *proto) {
    if (proto_matches<Wifi1944R12>(Z, mb, nb, proto)) return 0;
    return -1;
}

template <class Code, int CW, int UPD>
static int launch_qc_one(const DecodeArgs &a, cudaStream_t s) {
    using L = QcLayout<Code, CW>;
    const long long grid = (a.B + CW - 1) / CW;
    if (grid > 0x7fffffffLL) { set_error("batch too large"); return LDPC_EINVAL; }
    auto k = decode_qc_kernel<Code, CW, UPD>;
    LDPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
    k<<<(int)grid, L::THREADS, L::SMEM, s>>>(a);
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
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
        case 0: {
            static const int cw = [] { const char *e = getenv("LDPC_QC_CW"); return e ? atoi(e) : 3; }();
            if (a.update == UPD_MINSUM) {
                if (cw == 6) return launch_qc_one<Wifi1944R12, 6, UPD_MINSUM>(a, s);
                if (cw == 1) return launch_qc_one<Wifi1944R12, 1, UPD_MINSUM>(a, s);
            }
            return launch_qc_t<Wifi1944R12, 3>(a, s);
        }
        default: set_error("unknown QC specialisation %d", qc_id); return LDPC_EINVAL;
    }
}

}  // namespace ldpc
