/* CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
 * product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Plain-C restatement of the reference's belief-propagation decoder with the SAME
 * association order as oracle/bp_oracle.py (see that file's header):
 *   pytorch/bp/masking.py:75-138   edge numbering (check-major x, variable-major y)
 *   pytorch/bp/bp_vc.py:16-32      V->C  0.5*(llr' + sum of others)
 *   pytorch/bp/bp.py:29            tanh
 *   pytorch/bp/bp_cv.py:22-55      C->V  log((1+p)/(1-p)), p clamped to +-0.99999988f
 *   pytorch/bp/bp.py:43-51         loop, outer clamp, marginal, 1 - sigmoid
 * Min-sum family (not in the reference) is DEFINED here and in bp_oracle.py: only
 * add/min/abs/sign, so it is bit-exact against the CUDA kernels.  Sum-product uses
 * glibc tanhf/logf/expf, which differ from ATen's by ulps: compare with tolerance
 * (bp_oracle.py is the bit-exact SP pin against the reference).
 *
 * Build: make -C oracle   (gcc -O2 -pthread -ffp-contract=off; no fast-math)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

enum { UPD_SP = 0, UPD_MINSUM = 1, UPD_NMS = 2, UPD_OMS = 3 };

#define P_CLAMP 0.99999988f /* fp32(1 - 1e-7), bp_cv.py:44-47 */

typedef struct {
    int m, n, E, max_d;
    const int32_t *chk_ptr, *vm_of_cm, *var_ptr, *cm_of_vm, *chk_var;
} graph_t;

static void sum_others(const float *in, int d, float *out) {
    /* out[k] = P_k + Q_k (see header of bp_oracle.py); d==1 -> +0.0f */
    float pre[64], acc = 0.f;
    if (d == 1) { out[0] = 0.0f; return; }
    for (int k = 0; k < d; ++k) { pre[k] = acc; acc = (k == 0) ? in[0] : acc + in[k]; }
    acc = 0.f;
    for (int k = d - 1; k >= 0; --k) {
        if (k == d - 1) out[k] = pre[k];
        else if (k == 0) out[k] = acc;
        else out[k] = pre[k] + acc;
        acc = (k == d - 1) ? in[k] : in[k] + acc;
    }
}

static void prod_others(const float *in, int d, float *out) {
    float pre[64], acc = 1.f;
    if (d == 1) { out[0] = 1.0f; return; }
    for (int k = 0; k < d; ++k) { pre[k] = acc; acc = (k == 0) ? in[0] : acc * in[k]; }
    acc = 1.f;
    for (int k = d - 1; k >= 0; --k) {
        if (k == d - 1) out[k] = pre[k];
        else if (k == 0) out[k] = acc;
        else out[k] = pre[k] * acc;
        acc = (k == d - 1) ? in[k] : in[k] * acc;
    }
}

static void decode_one(const graph_t *g, const float *llr, int iters, int update, float clampv,
                       float param, const float *x0, float *x, float *y, float *t_out,
                       float *prob_out, uint8_t *hard_out, int32_t *synd_out) {
    float in[64], ot[64];
    const int E = g->E;
    if (x0) memcpy(x, x0, sizeof(float) * E); else memset(x, 0, sizeof(float) * E);
    for (int it = 0; it < iters; ++it) {
        for (int v = 0; v < g->n; ++v) {                       /* V -> C */
            const int b = g->var_ptr[v], d = g->var_ptr[v + 1] - b;
            if (d == 0) continue;
            const float Lp = -llr[v];
            for (int k = 0; k < d; ++k) in[k] = x[g->cm_of_vm[b + k]];
            sum_others(in, d, ot);
            for (int k = 0; k < d; ++k) {
                const float s = Lp + ot[k];
                y[b + k] = (update == UPD_SP) ? tanhf(0.5f * s) : s;
            }
        }
        for (int c = 0; c < g->m; ++c) {                       /* C -> V */
            const int b = g->chk_ptr[c], d = g->chk_ptr[c + 1] - b;
            if (d == 0) continue;
            for (int j = 0; j < d; ++j) in[j] = y[g->vm_of_cm[b + j]];
            if (update == UPD_SP) {
                prod_others(in, d, ot);
                for (int j = 0; j < d; ++j) {
                    float p = ot[j];
                    p = p < -P_CLAMP ? -P_CLAMP : (p > P_CLAMP ? P_CLAMP : p);
                    float o = logf((1.0f + p) / (1.0f - p));
                    o = o < -clampv ? -clampv : (o > clampv ? clampv : o);
                    x[b + j] = o;
                }
            } else {
                float m1 = INFINITY, m2 = INFINITY; int i1 = -1; unsigned par = 0;
                for (int j = 0; j < d; ++j) {
                    const float a = fabsf(in[j]);
                    par ^= (unsigned)(signbit(in[j]) != 0);
                    if (a < m1) { m2 = m1; m1 = a; i1 = j; } else if (a < m2) m2 = a;
                }
                for (int j = 0; j < d; ++j) {
                    float mg = (j == i1) ? m2 : m1;
                    if (update == UPD_NMS) mg = param * mg;
                    else if (update == UPD_OMS) { mg = mg - param; mg = mg > 0.f ? mg : 0.f; }
                    mg = mg < clampv ? mg : clampv;
                    const unsigned s = par ^ (unsigned)(signbit(in[j]) != 0);
                    x[b + j] = s ? -mg : mg;
                }
            }
        }
    }
    for (int v = 0; v < g->n; ++v) {                           /* marginal */
        const int b = g->var_ptr[v], d = g->var_ptr[v + 1] - b;
        float acc = 0.f;
        for (int k = 0; k < d; ++k) acc = (k == 0) ? x[g->cm_of_vm[b]] : acc + x[g->cm_of_vm[b + k]];
        const float t = 0.5f * (-llr[v] + acc);
        const float prob = 1.0f - 1.0f / (1.0f + expf(-t));
        if (t_out) t_out[v] = t;
        if (prob_out) prob_out[v] = prob;
        hard_out[v] = prob > 0.5f;
    }
    if (synd_out) {
        int w = 0;
        for (int c = 0; c < g->m; ++c) {
            unsigned p = 0;
            for (int e = g->chk_ptr[c]; e < g->chk_ptr[c + 1]; ++e) p ^= hard_out[g->chk_var[e]];
            w += (int)p;
        }
        *synd_out = w;
    }
}

typedef struct {
    const graph_t *g; const float *llr; int64_t b0, b1; int iters, update; float clampv, param;
    const float *x0; float *t_out, *prob_out; uint8_t *hard_out; int32_t *synd_out; float *x_out;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    const graph_t *g = j->g;
    const int E = g->E, n = g->n;
    float *x = (float *)malloc(sizeof(float) * (E + 1));
    float *y = (float *)malloc(sizeof(float) * (E + 1));
    for (int64_t b = j->b0; b < j->b1; ++b) {
        decode_one(g, j->llr + b * n, j->iters, j->update, j->clampv, j->param,
                   j->x0 ? j->x0 + b * E : NULL, x, y, j->t_out ? j->t_out + b * n : NULL,
                   j->prob_out ? j->prob_out + b * n : NULL, j->hard_out + b * n,
                   j->synd_out ? j->synd_out + b : NULL);
        if (j->x_out) memcpy(j->x_out + b * E, x, sizeof(float) * E);
    }
    free(x); free(y);
    return NULL;
}

int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* Returns 0, or -1 on bad arguments (degree > 64). hard_out is required ([B,n] u8);
 * t_out, prob_out ([B,n] f32), synd_out ([B] i32), x_out ([B,E] f32) may be NULL.
 * nthreads <= 0 uses every online core (pthreads, contiguous codeword ranges). */
int oracle_decode(int m, int n, int E, const int32_t *chk_ptr, const int32_t *chk_var,
                  const int32_t *vm_of_cm, const int32_t *var_ptr, const int32_t *cm_of_vm,
                  const float *llr, int64_t B, int iters, int update, float clampv, float param,
                  const float *x0, float *t_out, float *prob_out, uint8_t *hard_out,
                  int32_t *synd_out, float *x_out, int nthreads) {
    graph_t g = {m, n, E, 0, chk_ptr, vm_of_cm, var_ptr, cm_of_vm, chk_var};
    for (int c = 0; c < m; ++c) if (chk_ptr[c + 1] - chk_ptr[c] > 64) return -1;
    for (int v = 0; v < n; ++v) if (var_ptr[v + 1] - var_ptr[v] > 64) return -1;
    if (nthreads <= 0) nthreads = oracle_max_threads();
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > B) nthreads = B > 0 ? (int)B : 1;
    pthread_t th[256]; job_t jobs[256];
    for (int i = 0; i < nthreads; ++i) {
        job_t j = {&g, llr, B * i / nthreads, B * (i + 1) / nthreads, iters, update, clampv, param,
                   x0, t_out, prob_out, hard_out, synd_out, x_out};
        jobs[i] = j;
        if (nthreads == 1) worker(&jobs[0]);
        else pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    if (nthreads > 1) for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    return 0;
}
