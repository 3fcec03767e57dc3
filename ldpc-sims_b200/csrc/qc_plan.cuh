// qc_plan.cuh - compile-time prototype matrices and the execution plan shared by the
// code-specialised decoder kernels (fp32 and f16x2 variants).
#pragma once
#include <stdint.h>
#include <type_traits>
#include <utility>

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
    // Spanning tree (block row, block column) that fixes the lane relabelling (QcPlan below).  ANY spanning tree makes
    // MB + NB - 1 = 35 blocks thread-local; this one (hill-climbing over edge swaps, scratch search recorded in
    // profiles/r02_decoder_schedule_experiments.md) additionally leaves only 24 DISTINCT effective shifts among the 51
    // exchanged blocks instead of the 33 of the breadth-first tree: the variable phase needs one ISETP + one SEL per
    // distinct shift for its rotated-window pointers.
    static constexpr int forest[35][2] = {{6, 18}, {9, 8}, {9, 4}, {2, 14}, {1, 8}, {6, 1}, {11, 0}, {5, 11}, {10, 23}, {1, 9}, {6, 8}, {3, 0}, {9, 3}, {10, 22}, {3, 7}, {0, 4}, {8, 20}, {11, 8}, {10, 10}, {2, 15}, {2, 5}, {0, 13}, {7, 20}, {8, 11}, {11, 23}, {9, 21}, {4, 4}, {11, 12}, {6, 6}, {8, 8}, {4, 17}, {6, 2}, {8, 5}, {7, 19}, {3, 16}};
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
template <class Code, class = void>
struct has_forest : std::false_type {};
template <class Code>
struct has_forest<Code, std::void_t<decltype(Code::forest)>> : std::true_type {};

template <class Code>
struct QcPlan {
    static constexpr int Z = Code::Z, MB = Code::MB, NB = Code::NB;
    int sigma[MB] = {}, rho[NB] = {};
    int nblk = 0, n_local = 0, n_smem = 0;
    int row_deg[MB] = {}, row_col[MB][NB] = {}, row_eff[MB][NB] = {}, row_slot[MB][NB] = {}, row_shift[MB][NB] = {};
    int col_deg[NB] = {}, col_row[NB][MB] = {}, col_eff[NB][MB] = {}, col_slot[NB][MB] = {};
    bool row_loc[MB][NB] = {}, col_loc[NB][MB] = {};
    // systematic encoder (802.11n style): H = [A | h | T], T dual-diagonal (shift 0), column h with the
    // same shift at its top and bottom entries and shift 0 at the middle one
    int kb = NB - MB, enc_deg[MB] = {}, enc_col[MB][NB] = {}, enc_shift[MB][NB] = {}, hcol[MB] = {};
    bool dual_diagonal = true;
    constexpr QcPlan() {
        for (int r = 0; r < MB; ++r) {
            for (int c = 0; c < NB - MB; ++c)
                if (Code::proto[r][c] >= 0) { enc_col[r][enc_deg[r]] = c; enc_shift[r][enc_deg[r]] = Code::proto[r][c]; ++enc_deg[r]; }
            hcol[r] = Code::proto[r][NB - MB];
            for (int j = 1; j < MB; ++j) {
                const bool want = (r == j - 1) || (r == j);
                const int v = Code::proto[r][NB - MB + j];
                if (want ? (v != 0) : (v >= 0)) dual_diagonal = false;
            }
        }
        {
            int cnt = 0, mid = -1;
            for (int r = 0; r < MB; ++r) if (hcol[r] >= 0) { ++cnt; if (r != 0 && r != MB - 1) mid = r; }
            if (cnt != 3 || hcol[0] < 0 || hcol[MB - 1] < 0 || hcol[0] != hcol[MB - 1] || mid < 0 || hcol[mid] != 0) dual_diagonal = false;
        }
        bool row_seen[MB] = {}, col_seen[NB] = {};
        if constexpr (has_forest<Code>::value) {
            // the code brings its own spanning forest: propagate sigma / rho along its edges
            constexpr int NE = (int)(sizeof(Code::forest) / sizeof(Code::forest[0]));
            for (int pass = 0; pass < MB + NB; ++pass) {
                bool any = false;
                for (int e = 0; e < NE; ++e) {
                    const int r = Code::forest[e][0], c = Code::forest[e][1];
                    if (row_seen[r] && !col_seen[c]) { rho[c] = (Code::proto[r][c] + sigma[r]) % Z; col_seen[c] = true; any = true; }
                    else if (col_seen[c] && !row_seen[r]) { sigma[r] = ((rho[c] - Code::proto[r][c]) % Z + Z) % Z; row_seen[r] = true; any = true; }
                }
                if (!any) {                              // start (the next) component at the first forest edge not reached yet
                    for (int e = 0; e < NE && !any; ++e)
                        if (!row_seen[Code::forest[e][0]] && !col_seen[Code::forest[e][1]]) { row_seen[Code::forest[e][0]] = true; sigma[Code::forest[e][0]] = 0; any = true; }
                    if (!any) break;
                }
            }
        }
        // breadth-first spanning tree over whatever block rows / block columns are still unassigned
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

// Variable-phase batches of the persistent kernel (decode_qc_pers.cuh): consecutive block columns whose
// shared-memory edges are loaded TOGETHER, one batch ahead of the batch being computed.  The rotated-window
// pointers of the variable phase are run-time selections, so the compiler cannot prove that a store of one
// node and a load of the next do not alias and keeps them in program order; issuing the loads of batch b+1
// before the stores of batch b in the SOURCE is what lets the shared-memory latency overlap the additions.
template <class Code, int VB_MAX>
struct VarBatches {
    static constexpr int MB = Code::MB, NB = Code::NB;
    int n = 0, first[NB + 1] = {}, width = 1;
    int idx[NB][MB] = {};                                // (column, edge) -> index inside its batch (shared-memory edges only)
    constexpr VarBatches() {
        int cnt = 0;
        for (int c = 0; c < NB; ++c) {
            int nsm = 0;
            for (int k = 0; k < kQc<Code>.col_deg[c]; ++k) nsm += kQc<Code>.col_loc[c][k] ? 0 : 1;
            if (c > 0 && (cnt + nsm > VB_MAX || VB_MAX <= 0)) { first[++n] = c; cnt = 0; }
            for (int k = 0; k < kQc<Code>.col_deg[c]; ++k)
                if (!kQc<Code>.col_loc[c][k]) idx[c][k] = cnt++;
            if (cnt > width) width = cnt;
        }
        first[++n] = NB;
    }
};
template <class Code, int VB_MAX>
inline constexpr VarBatches<Code, VB_MAX> kVarBatches{};

template <class Code, int CW>
struct QcLayout {
    static constexpr int Z = Code::Z;
    static constexpr int N = Code::NB * Z;
    static constexpr int M = Code::MB * Z;
    static constexpr int NLOC = kQc<Code>.n_local, NSM = kQc<Code>.n_smem;
    static constexpr int MSG_STRIDE = NSM * Z;       // message words per codeword (codewords interleaved by lane)
    static constexpr int HARD_STRIDE = (N + 15) & ~15;
    static constexpr int THREADS = ((CW * Z + 31) / 32) * 32;
    static constexpr int MIN_CTAS = THREADS <= 96 ? 5 : (THREADS <= 256 ? 2 : 1);   // register budget: 64K / (THREADS * MIN_CTAS)
    static constexpr size_t MSG_BYTES = (sizeof(float) * CW * MSG_STRIDE + 15) & ~size_t(15);   // 4-byte message words; keeps hard_s 16-byte aligned
    static constexpr size_t SMEM = MSG_BYTES + (size_t)CW * HARD_STRIDE + sizeof(int) * (8 + 2 * CW);
};

}  // namespace ldpc
