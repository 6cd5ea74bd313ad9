/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference quantiser arithmetic.
 *
 * Restates, scalar and in the reference's evaluation order, what
 *   vq_ae/layers/vq.py:121-130  torch.argmin(torch.cdist(flat_input, embed, p=ndim=4), dim=1)
 *                               + F.embedding(idx, embed)
 *   vq_ae/layers/vq.py:143      F.mse_loss(inputs, quantized)
 * compute on the CPU.  The arithmetic itself lives in a third-party dependency that is not
 * under /root/reference: PyTorch (pinned torch==1.11.0+cu115, pyproject.toml:9) --
 * ATen/native/cpu/DistanceOpsKernel.cpp, the generic-p `pdist_calc` path used by
 * `cdist` : per pair, agg = sum over d (in d order) of std::pow(|a-b|, p), then
 * std::pow(agg, 1/p).  `torch.argmin` returns the first minimal index.
 *
 * Pinned in tests/test_oracle.py against torch.cdist (bit-exact distance matrix) and the
 * committed golden vectors.  Only tests/, smoke() and bench.py's cpu_baseline leg may load
 * this library; the product path never does.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include <pthread.h>
#include <unistd.h>

/* rooted L4 distance of one (row, code) pair, exactly as ATen's scalar loop does it */
static inline float l4_rooted(const float* a, const float* b, int d, float* unrooted) {
    float agg = 0.0f;
    for (int x = 0; x < d; ++x) {
        float diff = fabsf(a[x] - b[x]);
        agg = agg + powf(diff, 4.0f);
    }
    *unrooted = agg;
    return powf(agg, 1.0f / 4.0f);
}

/* Full N x K distance matrix (small N only; used to pin the restatement bit-exactly). */
void oracle_l4_cdist(const float* flat, int64_t n, const float* embed, int k, int d, float* out) {
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) {
            float un;
            out[i * k + j] = l4_rooted(flat + i * d, embed + (int64_t)j * d, d, &un);
        }
}

/* idx[i] = first argmin_j rooted distance; q[i,:] = embed[idx[i],:];
 * gap[i] = (d2 - d1) / d2 on the un-rooted sums (near-tie report);
 * returns sum over all elements of (flat - q)^2 accumulated in double. */
static double quantize_range(const float* flat, int64_t lo, int64_t hi, const float* embed, int k,
                             int d, int64_t* idx, float* q, float* gap) {
    double sq_total = 0.0;
    for (int64_t i = lo; i < hi; ++i) {
        const float* a = flat + i * d;
        float best = INFINITY, best_un = INFINITY, second_un = INFINITY;
        int best_j = 0;
        for (int j = 0; j < k; ++j) {
            float un;
            float r = l4_rooted(a, embed + (int64_t)j * d, d, &un);
            if (r < best) {            /* strict: first index wins ties */
                second_un = best_un;
                best = r; best_un = un; best_j = j;
            } else if (un < second_un) {
                second_un = un;
            }
        }
        idx[i] = best_j;
        if (gap) gap[i] = second_un > 0.0f ? fabsf(second_un - best_un) / second_un : 0.0f;
        const float* e = embed + (int64_t)best_j * d;
        for (int x = 0; x < d; ++x) {
            if (q) q[i * d + x] = e[x];
            double df = (double)a[x] - (double)e[x];
            sq_total += df * df;
        }
    }
    return sq_total;
}

typedef struct {
    const float* flat; int64_t lo, hi; const float* embed; int k, d;
    int64_t* idx; float* q; float* gap; double sq;
} job_t;

static void* job_main(void* arg) {
    job_t* j = (job_t*)arg;
    j->sq = quantize_range(j->flat, j->lo, j->hi, j->embed, j->k, j->d, j->idx, j->q, j->gap);
    return NULL;
}

int oracle_l4_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* nthreads <= 0: all online cores */
double oracle_l4_quantize_mt(const float* flat, int64_t n, const float* embed, int k, int d,
                             int64_t* idx, float* q, float* gap, int nthreads) {
    if (nthreads <= 0) nthreads = oracle_l4_num_threads();
    if (nthreads > 256) nthreads = 256;
    if (nthreads <= 1 || n < 1024)
        return quantize_range(flat, 0, n, embed, k, d, idx, q, gap);
    pthread_t th[256];
    job_t jobs[256];
    int64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        int64_t lo = t * per, hi = lo + per;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        jobs[t] = (job_t){flat, lo, hi, embed, k, d, idx, q, gap, 0.0};
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    double total = 0.0;
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].sq;
    }
    return total;
}

double oracle_l4_quantize(const float* flat, int64_t n, const float* embed, int k, int d,
                          int64_t* idx, float* q, float* gap) {
    return quantize_range(flat, 0, n, embed, k, d, idx, q, gap);
}
