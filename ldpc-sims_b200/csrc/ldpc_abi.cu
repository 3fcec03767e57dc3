// ldpc_abi.cu - the C ABI declared in include/ldpc_b200.h: code handles, dispatch, the
// host-buffer decode pipeline and error plumbing.  No torch types, no CPU fallback.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "common.cuh"

namespace ldpc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return LDPC_ECUDA;
}

}  // namespace ldpc

using namespace ldpc;

extern "C" void ldpc_host_pipe_free(void *p);

extern "C" {

int ldpc_abi_version(void) { return LDPC_B200_ABI_VERSION; }

const char *ldpc_last_error(void) { return g_err; }

int ldpc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ldpc_code_create(const int32_t *row_ptr, const int32_t *col_idx, int m, int n, int qc_Z,
                     const int16_t *qc_proto, ldpc_code_t **out) {
    if (!row_ptr || !col_idx || !out || m <= 0 || n <= 0) { set_error("ldpc_code_create: bad arguments"); return LDPC_EINVAL; }
    *out = nullptr;
    if (row_ptr[0] != 0) { set_error("row_ptr[0] must be 0"); return LDPC_EINVAL; }
    const int E = row_ptr[m];
    if (E <= 0) { set_error("H has no edges"); return LDPC_EINVAL; }
    int max_dc = 0;
    std::vector<int32_t> dv(n, 0);
    for (int c = 0; c < m; ++c) {
        const int b = row_ptr[c], e = row_ptr[c + 1];
        if (e < b) { set_error("row_ptr not monotone at row %d", c); return LDPC_EINVAL; }
        max_dc = std::max(max_dc, e - b);
        for (int i = b; i < e; ++i) {
            const int v = col_idx[i];
            if (v < 0 || v >= n) { set_error("col_idx out of range at row %d", c); return LDPC_EINVAL; }
            if (i > b && col_idx[i - 1] >= v) { set_error("columns must be strictly ascending inside row %d", c); return LDPC_EINVAL; }
            dv[v]++;
        }
    }
    // variable-major numbering: column-major non-zeros (masking.py:92-95)
    std::vector<int32_t> var_ptr(n + 1, 0), cm_of_vm(E), fill(n, 0);
    int max_dv = 0;
    for (int v = 0; v < n; ++v) { var_ptr[v + 1] = var_ptr[v] + dv[v]; max_dv = std::max(max_dv, dv[v]); }
    for (int c = 0; c < m; ++c)
        for (int i = row_ptr[c]; i < row_ptr[c + 1]; ++i) {
            const int v = col_idx[i];
            cm_of_vm[var_ptr[v] + fill[v]++] = i;       // rows visited ascending => checks ascending
        }
    int dev = 0;
    LDPC_CUDA_TRY(cudaGetDevice(&dev));
    ldpc_code *h = new (std::nothrow) ldpc_code();
    if (!h) { set_error("out of host memory"); return LDPC_ENOMEM; }
    h->m = m; h->n = n; h->E = E; h->max_dc = max_dc; h->max_dv = max_dv; h->device = dev;
    h->qc_Z = 0; h->qc_id = -1; h->kernel = LDPC_KERNEL_GENERIC; h->d_tables = nullptr;
    h->d_qc_rt = nullptr; h->qc_mb = h->qc_nb = h->qc_nblk = 0;
    h->tiny_id = tiny_lookup(m, n, row_ptr, col_idx);
    if (h->tiny_id >= 0) h->kernel = LDPC_KERNEL_TINY;
    h->d_gen = nullptr; h->k_info = 0; h->host_pipe = nullptr; h->precision = LDPC_PREC_F32;
    const size_t words = (size_t)(m + 1) + E + (n + 1) + E;
    cudaError_t e = cudaMalloc(&h->d_tables, words * sizeof(int32_t));
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaMalloc(tables)"); }
    std::vector<int32_t> host(words);
    int32_t *p = host.data();
    memcpy(p, row_ptr, sizeof(int32_t) * (m + 1)); p += m + 1;
    memcpy(p, col_idx, sizeof(int32_t) * E); p += E;
    memcpy(p, var_ptr.data(), sizeof(int32_t) * (n + 1)); p += n + 1;
    memcpy(p, cm_of_vm.data(), sizeof(int32_t) * E);
    e = cudaMemcpy(h->d_tables, host.data(), words * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(h->d_tables); delete h; return cuda_fail(e, "cudaMemcpy(tables)"); }
    h->g.m = m; h->g.n = n; h->g.E = E;
    h->g.chk_ptr = h->d_tables;
    h->g.chk_var = h->d_tables + (m + 1);
    h->g.var_ptr = h->g.chk_var + E;
    h->g.cm_of_vm = h->g.var_ptr + (n + 1);
    if (qc_Z > 0 && qc_proto) {
        if (m % qc_Z || n % qc_Z) { cudaFree(h->d_tables); delete h; set_error("qc_Z does not divide H"); return LDPC_EINVAL; }
        // verify that the declared prototype really describes H (edge count + every edge)
        const int mb = m / qc_Z, nb = n / qc_Z;
        long long cnt = 0;
        for (int i = 0; i < mb * nb; ++i) cnt += qc_proto[i] >= 0 ? qc_Z : 0;
        bool ok = (cnt == E);
        for (int c = 0; ok && c < m; ++c)
            for (int i = row_ptr[c]; ok && i < row_ptr[c + 1]; ++i) {
                const int v = col_idx[i], s = qc_proto[(c / qc_Z) * nb + v / qc_Z];
                ok = s >= 0 && ((c % qc_Z + s) % qc_Z) == v % qc_Z;
            }
        if (!ok) { cudaFree(h->d_tables); delete h; set_error("qc_proto does not match H"); return LDPC_EINVAL; }
        h->qc_Z = qc_Z;
        h->qc_id = qc_lookup(qc_Z, mb, nb, qc_proto);
        if (h->qc_id >= 0) h->kernel = LDPC_KERNEL_QC;
        {   // run-time QC tables: the fast path of every QC code without a compiled specialisation
            std::vector<int32_t> tab;
            int mdv = 0, mdc = 0;
            const int words = qc_rt_build_tables(qc_Z, mb, nb, qc_proto, tab, &mdv, &mdc);
            const int nblk = (int)(cnt / qc_Z);
            if (words > 0 && qc_rt_supported(qc_Z, mb, nb, nblk, mdv, mdc) &&
                cudaMalloc(&h->d_qc_rt, (size_t)words * sizeof(int32_t)) == cudaSuccess) {
                cudaMemcpy(h->d_qc_rt, tab.data(), (size_t)words * sizeof(int32_t), cudaMemcpyHostToDevice);
                h->qc_mb = mb; h->qc_nb = nb; h->qc_nblk = nblk;
                if (h->qc_id < 0 && h->tiny_id < 0) h->kernel = LDPC_KERNEL_QC_RT;
            } else {
                h->d_qc_rt = nullptr;
                cudaGetLastError();
            }
        }
    }
    *out = h;
    return LDPC_OK;
}

void ldpc_code_destroy(ldpc_code_t *code) {
    if (!code) return;
    if (code->d_tables) cudaFree(code->d_tables);
    if (code->d_gen) cudaFree(code->d_gen);
    if (code->d_qc_rt) cudaFree(code->d_qc_rt);
    if (code->host_pipe) ldpc_host_pipe_free(code->host_pipe);
    delete code;
}

int ldpc_code_info(const ldpc_code_t *code, ldpc_code_info_t *info) {
    if (!code || !info) { set_error("ldpc_code_info: null argument"); return LDPC_EINVAL; }
    info->m = code->m; info->n = code->n; info->E = code->E;
    info->max_dc = code->max_dc; info->max_dv = code->max_dv;
    info->kernel = code->kernel; info->qc_Z = code->qc_Z; info->reserved = 0;
    return LDPC_OK;
}

int ldpc_qc_register_plugin(const char *so_path) { return qc_register_plugin(so_path); }

int ldpc_code_plan_info(const ldpc_code_t *code, int32_t out[4]) {
    if (!code || !out) { set_error("ldpc_code_plan_info: null argument"); return LDPC_EINVAL; }
    int v[4] = {0, 0, 0, 0};
    if (code->qc_id >= 0) qc_plan_info(code->qc_id, v);
    for (int i = 0; i < 4; ++i) out[i] = v[i];
    return LDPC_OK;
}

int ldpc_code_set_precision(ldpc_code_t *code, int precision) {
    if (!code) { set_error("null code"); return LDPC_EINVAL; }
    if (precision == LDPC_PREC_F32) { code->precision = precision; return LDPC_OK; }
    if (precision == LDPC_PREC_F16X2 && code->qc_id >= 0) { code->precision = precision; return LDPC_OK; }
    set_error("precision %d not available for this code (f16x2 needs a code-specialised kernel)", precision);
    return LDPC_EUNSUPPORTED;
}

int ldpc_code_set_kernel(ldpc_code_t *code, int kernel) {
    if (!code) { set_error("null code"); return LDPC_EINVAL; }
    if (kernel == LDPC_KERNEL_GENERIC) { code->kernel = kernel; return LDPC_OK; }
    if ((kernel == LDPC_KERNEL_QC || kernel == LDPC_KERNEL_QC_TMA) && code->qc_id >= 0) { code->kernel = kernel; return LDPC_OK; }
    if (kernel == LDPC_KERNEL_TINY && code->tiny_id >= 0) { code->kernel = kernel; return LDPC_OK; }
    if (kernel == LDPC_KERNEL_QC_RT && code->d_qc_rt) { code->kernel = kernel; return LDPC_OK; }
    set_error("kernel %d not available for this code", kernel);
    return LDPC_EUNSUPPORTED;
}

static int check_decode_args(const ldpc_code_t *code, const void *llr, int llr_dtype, int64_t B, int iters,
                             int update, float clamp_value) {
    if (!code) { set_error("null code handle"); return LDPC_EINVAL; }
    if (B < 0 || iters < 0) { set_error("negative batch or iteration count"); return LDPC_EINVAL; }
    if (B > 0 && !llr) { set_error("null llr pointer"); return LDPC_EINVAL; }
    if (llr_dtype < LDPC_F32 || llr_dtype > LDPC_I8) { set_error("bad llr_dtype %d", llr_dtype); return LDPC_EINVAL; }
    if (update < LDPC_UPDATE_SP || update > LDPC_UPDATE_OMS) { set_error("bad update rule %d", update); return LDPC_EINVAL; }
    if (!(clamp_value > 0.0f)) { set_error("clamp_value must be positive"); return LDPC_EINVAL; }
    return LDPC_OK;
}

// device-pointer entry points: the handle's tables live on code->device; a launch from another current device would
// read them through an invalid context
static int check_current_device(const ldpc_code_t *code) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); set_error("no CUDA device: libldpc_b200 has no CPU fallback"); return LDPC_ECUDA; }
    if (dev != code->device) { set_error("code handle belongs to device %d but device %d is current", code->device, dev); return LDPC_EINVAL; }
    return LDPC_OK;
}

}  // extern "C"

namespace ldpc {
// the register-resident kernel moves whole rows with 16-byte accesses: raw C-ABI callers may pass any pointer
static bool tiny_alignment_ok(const DecodeArgs &a) {
    auto al = [](const void *p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & (m - 1)) == 0; };
    return (a.llr_dtype != LDPC_F32 || al(a.llr, 16)) && al(a.llr_post, 16) && al(a.prob, 16) && al(a.hard, 4) && al(a.hard_packed, 8);
}

int decode_dispatch(const ldpc_code *code, const DecodeArgs &a, cudaStream_t s) {
    if (a.w_edge || a.wf_edge) {                       // trainable weights: register-resident kernel where compiled, else generic
        if (code->kernel == LDPC_KERNEL_TINY && a.x0 == nullptr && a.x_out == nullptr && !a.early_exit && tiny_alignment_ok(a) &&
            (a.update == LDPC_UPDATE_SP || a.update == LDPC_UPDATE_MINSUM))
            return launch_decode_tiny(code->tiny_id, a, s);
        return launch_decode_generic(code->g, code->max_dv, code->max_dc, a, s);
    }
    if ((code->kernel == LDPC_KERNEL_QC || code->kernel == LDPC_KERNEL_QC_TMA) && a.x0 == nullptr && a.x_out == nullptr) {
        if (code->kernel == LDPC_KERNEL_QC_TMA && code->precision == LDPC_PREC_F32) {
            const int rc = launch_decode_qc_tma(code->qc_id, a, s);
            if (rc != LDPC_EUNSUPPORTED) return rc;        // early termination, other update rules, other codes: the one-tile-per-CTA kernel
        }
        if (code->precision == LDPC_PREC_F16X2 && (a.update == LDPC_UPDATE_MINSUM || a.update == LDPC_UPDATE_NMS) &&
            !a.early_exit && !a.iters_used)
            return launch_decode_qc_h2(code->qc_id, a, s);
        const int rc = launch_decode_qc(code->qc_id, a, s);
        if (rc != LDPC_EUNSUPPORTED) return rc;            // a combination this code was not compiled for: generic kernel below
    }
    if (code->kernel == LDPC_KERNEL_QC_RT && a.x0 == nullptr && a.x_out == nullptr && !a.early_exit)
        return launch_decode_qc_rt(code->d_qc_rt, code->qc_Z, code->qc_mb, code->qc_nb, code->qc_nblk, code->max_dv, code->max_dc, a, s);
    if (code->kernel == LDPC_KERNEL_TINY && a.x0 == nullptr && a.x_out == nullptr && tiny_alignment_ok(a))
        return launch_decode_tiny(code->tiny_id, a, s);
    return launch_decode_generic(code->g, code->max_dv, code->max_dc, a, s);
}
}  // namespace ldpc

extern "C" {

int ldpc_decode(const ldpc_code_t *code, const void *llr, int llr_dtype, int64_t B, int iters, int update,
                float clamp_value, float param, const float *x0, float *prob, float *llr_post,
                uint8_t *hard, uint8_t *hard_packed, int32_t *syndrome, float *x_out, ldpc_stream_t stream) {
    int rc = check_decode_args(code, llr, llr_dtype, B, iters, update, clamp_value);
    if (rc) return rc;
    if ((rc = check_current_device(code))) return rc;
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.llr = llr; a.llr_dtype = llr_dtype; a.B = B; a.iters = iters; a.update = update;
    a.clampv = clamp_value; a.param = param; a.x0 = x0;
    a.prob = prob; a.llr_post = llr_post; a.hard = hard; a.hard_packed = hard_packed;
    a.syndrome = syndrome; a.x_out = x_out;
    return decode_dispatch(code, a, (cudaStream_t)stream);
}

int ldpc_decode_weighted(const ldpc_code_t *code, const void *llr, int llr_dtype, int64_t B, int iters, int update,
                         float clamp_value, float param, const float *w_edge, const float *w_llr, const float *wf_edge,
                         const float *wf_llr, int w_stride, float *prob, float *llr_post, uint8_t *hard, uint8_t *hard_packed,
                         int32_t *syndrome, float *x_out, ldpc_stream_t stream) {
    int rc = check_decode_args(code, llr, llr_dtype, B, iters, update, clamp_value);
    if (rc) return rc;
    if ((rc = check_current_device(code))) return rc;
    if (!w_edge || !w_llr || !wf_edge || !wf_llr || w_stride < code->max_dv) { set_error("ldpc_decode_weighted: weight tables missing or w_stride < max_dv (%d)", code->max_dv); return LDPC_EINVAL; }
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.llr = llr; a.llr_dtype = llr_dtype; a.B = B; a.iters = iters; a.update = update;
    a.clampv = clamp_value; a.param = param;
    a.prob = prob; a.llr_post = llr_post; a.hard = hard; a.hard_packed = hard_packed;
    a.syndrome = syndrome; a.x_out = x_out;
    a.w_edge = w_edge; a.w_llr = w_llr; a.wf_edge = wf_edge; a.wf_llr = wf_llr; a.w_stride = w_stride;
    return decode_dispatch(code, a, (cudaStream_t)stream);
}

int ldpc_decode_ex(const ldpc_code_t *code, const ldpc_decode_params_t *p, ldpc_stream_t stream) {
    if (!p || p->struct_size != (int32_t)sizeof(ldpc_decode_params_t)) { set_error("ldpc_decode_ex: bad params struct"); return LDPC_EINVAL; }
    int rc = check_decode_args(code, p->llr, p->llr_dtype, p->B, p->iters, p->update, p->clamp_value);
    if (rc) return rc;
    if ((rc = check_current_device(code))) return rc;
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.llr = p->llr; a.llr_dtype = p->llr_dtype; a.B = p->B; a.iters = p->iters; a.update = p->update;
    a.clampv = p->clamp_value; a.param = p->param; a.x0 = p->x0;
    a.prob = p->prob; a.llr_post = p->llr_post; a.hard = p->hard; a.hard_packed = p->hard_packed;
    a.syndrome = p->syndrome; a.x_out = p->x_out;
    a.early_exit = p->early_exit ? 1 : 0; a.iters_used = p->iters_used;
    return decode_dispatch(code, a, (cudaStream_t)stream);
}

// ---- host-buffer pipeline (decode_bits, ofdm_functions.py:131-163) ---------------------------
// Staging buffers, streams and events are created once per code handle and reused by later
// calls (a mutex serialises host-buffer decodes on one handle).
namespace {
struct HostPipe {
    static const int NBUF = 3;
    struct Buf { void *llr = nullptr; uint8_t *hard = nullptr, *packed = nullptr; float *post = nullptr; int32_t *synd = nullptr; cudaStream_t s = nullptr; } buf[NBUF];
    size_t llr_bytes = 0, hard_bytes = 0, packed_bytes = 0, post_bytes = 0, synd_bytes = 0;
    // staged pipeline of ldpc_decode_bits_host: pinned staging on both sides of the device buffers
    struct SBuf { void *h_llr = nullptr, *d_llr = nullptr; uint8_t *h_packed = nullptr, *d_packed = nullptr; cudaStream_t s = nullptr;
                  cudaEvent_t done = nullptr; int64_t first = -1, cnt = 0; } sb[NBUF];
    size_t s_llr_bytes = 0, s_packed_bytes = 0;
    std::mutex mu;
    void release_staged() {
        for (auto &b : sb) {
            if (b.h_llr) cudaFreeHost(b.h_llr);
            if (b.h_packed) cudaFreeHost(b.h_packed);
            if (b.d_llr) cudaFree(b.d_llr);
            if (b.d_packed) cudaFree(b.d_packed);
            if (b.done) cudaEventDestroy(b.done);
            if (b.s) cudaStreamDestroy(b.s);
            b = SBuf();
        }
        s_llr_bytes = s_packed_bytes = 0;
    }
    void release() {
        release_staged();
        for (auto &b : buf) {
            if (b.llr) cudaFree(b.llr);
            if (b.hard) cudaFree(b.hard);
            if (b.packed) cudaFree(b.packed);
            if (b.post) cudaFree(b.post);
            if (b.synd) cudaFree(b.synd);
            if (b.s) cudaStreamDestroy(b.s);
            b = Buf();
        }
        llr_bytes = hard_bytes = packed_bytes = post_bytes = synd_bytes = 0;
    }
};
}  // namespace

// One mutex for the lazy creation of a handle's HostPipe, shared by every entry point that creates it (two
// function-local mutexes let concurrent first calls of ldpc_decode_host and ldpc_decode_bits_host both allocate).
static std::mutex g_host_pipe_create_mu;

// The host pipelines own device buffers and streams: they run on the handle's device whatever device is current in the
// calling thread, and restore the caller's device on return.
namespace {
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = (err == cudaSuccess); }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
}  // namespace

extern "C" void ldpc_host_pipe_free(void *p) {
    HostPipe *hp = static_cast<HostPipe *>(p);
    if (!hp) return;
    hp->release();
    delete hp;
}

int ldpc_decode_host(const ldpc_code_t *code, const void *llr_host, int llr_dtype, int64_t N, int iters,
                     int update, float clamp_value, float param, uint8_t *hard_host,
                     uint8_t *hard_packed_host, float *llr_post_host, int32_t *syndrome_host,
                     int64_t chunk) {
    int rc = check_decode_args(code, llr_host, llr_dtype, N, iters, update, clamp_value);
    if (rc) return rc;
    if (N == 0) return LDPC_OK;
    const int n = code->n, nby = (n + 7) / 8;
    const size_t esz = llr_dtype == LDPC_F64 ? 8 : (llr_dtype == LDPC_F16 ? 2 : (llr_dtype == LDPC_I8 ? 1 : 4));
    if (chunk <= 0) chunk = 16384;
    chunk = std::min<int64_t>(chunk, N);
    ldpc_code *mc = const_cast<ldpc_code *>(code);
    DeviceGuard dev_guard(code->device);
    if (dev_guard.err != cudaSuccess) return cuda_fail(dev_guard.err, "cudaSetDevice(code->device)");
    {
        std::lock_guard<std::mutex> g(g_host_pipe_create_mu);
        if (!mc->host_pipe) mc->host_pipe = new (std::nothrow) HostPipe();
        if (!mc->host_pipe) { set_error("out of host memory"); return LDPC_ENOMEM; }
    }
    HostPipe *hp = static_cast<HostPipe *>(mc->host_pipe);
    std::lock_guard<std::mutex> lock(hp->mu);
#define HTRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { hp->release(); return cuda_fail(_e, #x); } } while (0)
    const size_t need_llr = (size_t)chunk * n * esz, need_hard = hard_host ? (size_t)chunk * n : 0,
                 need_packed = hard_packed_host ? (size_t)chunk * nby : 0,
                 need_post = llr_post_host ? (size_t)chunk * n * sizeof(float) : 0,
                 need_synd = syndrome_host ? (size_t)chunk * sizeof(int32_t) : 0;
    if (need_llr > hp->llr_bytes || need_hard > hp->hard_bytes || need_packed > hp->packed_bytes ||
        need_post > hp->post_bytes || need_synd > hp->synd_bytes || !hp->buf[0].s) {
        hp->release();
        for (auto &b : hp->buf) {
            HTRY(cudaStreamCreateWithFlags(&b.s, cudaStreamNonBlocking));
            HTRY(cudaMalloc(&b.llr, need_llr));
            if (need_hard) HTRY(cudaMalloc(&b.hard, need_hard));
            if (need_packed) HTRY(cudaMalloc(&b.packed, need_packed));
            if (need_post) HTRY(cudaMalloc(&b.post, need_post));
            if (need_synd) HTRY(cudaMalloc(&b.synd, need_synd));
        }
        hp->llr_bytes = need_llr; hp->hard_bytes = need_hard; hp->packed_bytes = need_packed;
        hp->post_bytes = need_post; hp->synd_bytes = need_synd;
    }
    int64_t done = 0;
    int i = 0;
    while (done < N) {
        HostPipe::Buf &b = hp->buf[i % HostPipe::NBUF];
        const int64_t cnt = std::min<int64_t>(chunk, N - done);
        HTRY(cudaMemcpyAsync(b.llr, (const char *)llr_host + (size_t)done * n * esz, (size_t)cnt * n * esz,
                             cudaMemcpyHostToDevice, b.s));
        DecodeArgs a;
        memset(&a, 0, sizeof(a));
        a.llr = b.llr; a.llr_dtype = llr_dtype; a.B = cnt; a.iters = iters; a.update = update;
        a.clampv = clamp_value; a.param = param;
        a.hard = hard_host ? b.hard : nullptr; a.hard_packed = hard_packed_host ? b.packed : nullptr;
        a.llr_post = llr_post_host ? b.post : nullptr; a.syndrome = syndrome_host ? b.synd : nullptr;
        rc = decode_dispatch(code, a, b.s);
        if (rc) { hp->release(); return rc; }
        if (hard_host) HTRY(cudaMemcpyAsync(hard_host + (size_t)done * n, b.hard, (size_t)cnt * n, cudaMemcpyDeviceToHost, b.s));
        if (hard_packed_host) HTRY(cudaMemcpyAsync(hard_packed_host + (size_t)done * nby, b.packed, (size_t)cnt * nby, cudaMemcpyDeviceToHost, b.s));
        if (llr_post_host) HTRY(cudaMemcpyAsync(llr_post_host + (size_t)done * n, b.post, (size_t)cnt * n * sizeof(float), cudaMemcpyDeviceToHost, b.s));
        if (syndrome_host) HTRY(cudaMemcpyAsync(syndrome_host + done, b.synd, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, b.s));
        done += cnt;
        ++i;
    }
    for (auto &b : hp->buf) HTRY(cudaStreamSynchronize(b.s));
#undef HTRY
    return LDPC_OK;
}

}  // extern "C"

// ---- decode_bits proper: pageable numpy arrays in, {0,1} array out -------------------------------------------------
// The reference's decode_bits (ofdm_functions.py:131-163) takes a float64 ndarray and returns a float64 ndarray of
// {0.,1.}: 15.5 KB in and 15.5 KB out per n=1944 codeword of ordinary (pageable) host memory, against 1.9 us of decoding.
// Here host threads convert f64 -> f32 (the reference's own cast, ofdm_functions.py:156) into pinned staging while the
// previous chunk is on the GPU, only PACKED bits come back over PCIe, and host threads expand them into the caller's
// array while the next chunk decodes.
template <class F>
static void parallel_for(int64_t n, int threads, F f) {          // f(begin, end) on `threads` host threads
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n));
    std::vector<std::thread> pool;
    const int64_t per = (n + threads - 1) / threads;
    for (int t = 1; t < threads; ++t) {
        const int64_t b = std::min(n, t * per), e = std::min(n, b + per);
        if (b < e) pool.emplace_back([=] { f(b, e); });
    }
    f(0, std::min(n, per));
    for (auto &th : pool) th.join();
}

// Both host loops below are memory-bound (31 KB of caller memory per n=1944 codeword): the float64 results and the
// staged float32 LLRs are written with non-temporal stores where SSE2 is available, which saves the read-for-ownership
// of every destination line.
template <class T>
static void expand_bits(const uint8_t *packed, int64_t rows, int n, int nby, T *out, int threads) {
    parallel_for(rows, threads, [=](int64_t b, int64_t e) {
        for (int64_t r = b; r < e; ++r) {
            const uint8_t *p = packed + r * nby;
            T *o = out + r * n;
            int v = 0;
            for (int j = 0; j < nby; ++j) {
                const unsigned byte = p[j];
                const int lim = std::min(8, n - v);
                for (int k = 0; k < lim; ++k) o[v + k] = (T)((byte >> (7 - k)) & 1u);   // np.packbits order: first bit = MSB
                v += lim;
            }
        }
    });
}

#if defined(__SSE2__)
template <>
void expand_bits<double>(const uint8_t *packed, int64_t rows, int n, int nby, double *out, int threads) {
    static double lut[256][8];
    static std::once_flag once;
    std::call_once(once, [] {
        for (int b = 0; b < 256; ++b)
            for (int k = 0; k < 8; ++k) lut[b][k] = (double)((b >> (7 - k)) & 1);
    });
    parallel_for(rows, threads, [=](int64_t b, int64_t e) {
        for (int64_t r = b; r < e; ++r) {
            const uint8_t *p = packed + r * nby;
            double *o = out + r * n;
            const bool aligned = (reinterpret_cast<uintptr_t>(o) & 15) == 0;
            int v = 0;
            for (int j = 0; j < nby; ++j, v += 8) {
                const double *l = lut[p[j]];
                if (aligned && v + 8 <= n) {
                    _mm_stream_pd(o + v, _mm_loadu_pd(l));
                    _mm_stream_pd(o + v + 2, _mm_loadu_pd(l + 2));
                    _mm_stream_pd(o + v + 4, _mm_loadu_pd(l + 4));
                    _mm_stream_pd(o + v + 6, _mm_loadu_pd(l + 6));
                } else {
                    for (int k = 0; k < std::min(8, n - v); ++k) o[v + k] = l[k];
                }
            }
        }
        _mm_sfence();
    });
}
#endif

// dst (16-byte aligned pinned staging) = (float)src
static void cast_f64_to_f32(const double *src, float *dst, int64_t elems, int threads) {
    parallel_for((elems + 3) / 4, threads, [=](int64_t lo4, int64_t hi4) {
        const int64_t lo = lo4 * 4, hi = std::min(elems, hi4 * 4);
        int64_t k = lo;
#if defined(__SSE2__)
        for (; k + 4 <= hi; k += 4) {
            const __m128 a = _mm_cvtpd_ps(_mm_loadu_pd(src + k)), b = _mm_cvtpd_ps(_mm_loadu_pd(src + k + 2));   // round to nearest even
            _mm_stream_ps(dst + k, _mm_movelh_ps(a, b));
        }
        _mm_sfence();
#endif
        for (; k < hi; ++k) dst[k] = (float)src[k];
    });
}

extern "C" {

int ldpc_decode_bits_host(const ldpc_code_t *code, const void *llr_host, int llr_dtype, int64_t N, int iters, int update,
                          float clamp_value, float param, void *bits_out, int out_dtype, int64_t chunk, int threads) {
    int rc = check_decode_args(code, llr_host, llr_dtype, N, iters, update, clamp_value);
    if (rc) return rc;
    if (!bits_out && N > 0) { set_error("ldpc_decode_bits_host: bits_out is null"); return LDPC_EINVAL; }
    if (out_dtype != LDPC_F64 && out_dtype != LDPC_F32 && out_dtype != LDPC_I8) { set_error("ldpc_decode_bits_host: out_dtype must be F64, F32 or I8"); return LDPC_EINVAL; }
    if (N == 0) return LDPC_OK;
    const int n = code->n, nby = (n + 7) / 8;
    const int dev_dtype = llr_dtype == LDPC_F64 ? LDPC_F32 : llr_dtype;            // f64 is cast on the host
    const size_t src_esz = llr_dtype == LDPC_F64 ? 8 : (llr_dtype == LDPC_F16 ? 2 : (llr_dtype == LDPC_I8 ? 1 : 4));
    const size_t esz = llr_dtype == LDPC_F64 ? 4 : src_esz;
    if (chunk <= 0) chunk = 16384;
    chunk = std::min<int64_t>(chunk, N);
    if (threads <= 0) threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));   // host-memory-bound: scales to 16 on the test box
    ldpc_code *mc = const_cast<ldpc_code *>(code);
    DeviceGuard dev_guard(code->device);
    if (dev_guard.err != cudaSuccess) return cuda_fail(dev_guard.err, "cudaSetDevice(code->device)");
    {
        std::lock_guard<std::mutex> g(g_host_pipe_create_mu);
        if (!mc->host_pipe) mc->host_pipe = new (std::nothrow) HostPipe();
        if (!mc->host_pipe) { set_error("out of host memory"); return LDPC_ENOMEM; }
    }
    HostPipe *hp = static_cast<HostPipe *>(mc->host_pipe);
    std::lock_guard<std::mutex> lock(hp->mu);
#define HTRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { hp->release_staged(); return cuda_fail(_e, #x); } } while (0)
    const size_t need_llr = (size_t)chunk * n * esz, need_packed = (size_t)chunk * nby;
    if (need_llr > hp->s_llr_bytes || need_packed > hp->s_packed_bytes || !hp->sb[0].s) {
        hp->release_staged();
        for (auto &b : hp->sb) {
            HTRY(cudaStreamCreateWithFlags(&b.s, cudaStreamNonBlocking));
            HTRY(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
            HTRY(cudaHostAlloc(&b.h_llr, need_llr, cudaHostAllocDefault));
            HTRY(cudaHostAlloc((void **)&b.h_packed, need_packed, cudaHostAllocDefault));
            HTRY(cudaMalloc(&b.d_llr, need_llr));
            HTRY(cudaMalloc((void **)&b.d_packed, need_packed));
        }
        hp->s_llr_bytes = need_llr; hp->s_packed_bytes = need_packed;
    }
    for (auto &b : hp->sb) { b.first = -1; b.cnt = 0; }
    auto drain = [&](HostPipe::SBuf &b) -> cudaError_t {       // expand a finished chunk into the caller's array
        if (b.first < 0) return cudaSuccess;
        cudaError_t e = cudaEventSynchronize(b.done);
        if (e != cudaSuccess) return e;
        if (out_dtype == LDPC_F64) expand_bits(b.h_packed, b.cnt, n, nby, (double *)bits_out + (size_t)b.first * n, threads);
        else if (out_dtype == LDPC_F32) expand_bits(b.h_packed, b.cnt, n, nby, (float *)bits_out + (size_t)b.first * n, threads);
        else expand_bits(b.h_packed, b.cnt, n, nby, (uint8_t *)bits_out + (size_t)b.first * n, threads);
        b.first = -1;
        return cudaSuccess;
    };
    int64_t done = 0;
    int i = 0;
    while (done < N) {
        HostPipe::SBuf &b = hp->sb[i % HostPipe::NBUF];
        HTRY(drain(b));                                         // its previous chunk (also frees the staging for reuse)
        const int64_t cnt = std::min<int64_t>(chunk, N - done);
        const int64_t elems = cnt * n;
        if (llr_dtype == LDPC_F64) {
            cast_f64_to_f32((const double *)llr_host + (size_t)done * n, (float *)b.h_llr, elems, threads);
        } else {
            const char *src = (const char *)llr_host + (size_t)done * n * src_esz;
            char *dst = (char *)b.h_llr;
            parallel_for(elems, threads, [=](int64_t lo, int64_t hi) { memcpy(dst + lo * src_esz, src + lo * src_esz, (size_t)(hi - lo) * src_esz); });
        }
        HTRY(cudaMemcpyAsync(b.d_llr, b.h_llr, (size_t)elems * esz, cudaMemcpyHostToDevice, b.s));
        DecodeArgs a;
        memset(&a, 0, sizeof(a));
        a.llr = b.d_llr; a.llr_dtype = dev_dtype; a.B = cnt; a.iters = iters; a.update = update;
        a.clampv = clamp_value; a.param = param; a.hard_packed = b.d_packed;
        rc = decode_dispatch(code, a, b.s);
        if (rc) { hp->release_staged(); return rc; }
        HTRY(cudaMemcpyAsync(b.h_packed, b.d_packed, (size_t)cnt * nby, cudaMemcpyDeviceToHost, b.s));
        HTRY(cudaEventRecord(b.done, b.s));
        b.first = done; b.cnt = cnt;
        done += cnt;
        ++i;
    }
    for (int k = 0; k < HostPipe::NBUF; ++k) HTRY(drain(hp->sb[(i + k) % HostPipe::NBUF]));   // oldest first
#undef HTRY
    return LDPC_OK;
}

}  // extern "C"

// ---- standalone error counting ---------------------------------------------------------------
namespace ldpc {
__global__ void count_errors_kernel(const void *llr, int llr_dtype, const uint8_t *hard, const uint8_t *ref,
                                    long long B, int n, int k, unsigned long long *counters) {
    // one warp per codeword, grid-stride
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long unc = 0, inf = 0, fr = 0;
    for (long long cw = warp; cw < B; cw += nwarps) {
        int e_any = 0;
        for (int v = lane; v < n; v += 32) {
            const long long o = cw * n + v;
            const int r = ref[o];
            if (llr) unc += ((load_llr(llr, llr_dtype, o) > 0.0f) != r);
            const int e = hard[o] != r;
            inf += e & (v < k);
            e_any |= e;
        }
        e_any = __any_sync(0xffffffffu, e_any);
        if (lane == 0) fr += e_any;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unc += __shfl_xor_sync(0xffffffffu, unc, o);
        inf += __shfl_xor_sync(0xffffffffu, inf, o);
    }
    if (lane == 0) {
        if (unc) atomicAdd(&counters[0], unc);
        if (inf) atomicAdd(&counters[1], inf);
        if (fr) atomicAdd(&counters[2], fr);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters[3], (unsigned long long)B * n);
        atomicAdd(&counters[4], (unsigned long long)B);
    }
}
}  // namespace ldpc

extern "C" int ldpc_count_errors(const void *llr, int llr_dtype, const uint8_t *hard, const uint8_t *ref_bits,
                                 int64_t B, int n, int k, int64_t *counters, ldpc_stream_t stream) {
    if (!hard || !ref_bits || !counters || B < 0 || n <= 0 || k < 0 || k > n) { set_error("ldpc_count_errors: bad arguments"); return LDPC_EINVAL; }
    if (B == 0) return LDPC_OK;
    const int threads = 256;
    long long blocks = std::min<long long>((B * 32 + threads - 1) / threads, 148 * 8);
    count_errors_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(
        llr, llr_dtype, hard, ref_bits, B, n, k, reinterpret_cast<unsigned long long *>(counters));
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}
