/*
 * Plain-C CPU restatement for the matrix-correction path.  TEST INFRASTRUCTURE ONLY: loaded by
 * tests/ and by bench.py's cpu_baseline / --impl reference legs through oracle/hic_cpu.py, never by
 * the product.
 *
 *   ice_loop_coo()   the hot loop of the reference, hicexplorer/iterativeCorrection.py:40-74, in the
 *                    reference's own formulation (COO triplets, row sums, two in-place rescaling
 *                    passes, overflow scan), one OpenMP thread team.  threads=1 reproduces the
 *                    reference's single-threaded numpy arithmetic order exactly for the row sums
 *                    (sequential accumulation in stored order, like scipy's coo_matvec).
 *   kr_matvec()      y = x o ((A + eps I) u) + v o p : the mat-vec inside krbalancing's CG loop
 *                    (restated in oracle/kr_oracle.py), CSR, OpenMP over rows.
 *   synth_*()        the benchmark's synthetic matrix, same function as
 *                    hicexplorer_b200/csrc/hb_synth.cu and oracle/synth_ref.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- the same loop on all host cores with plain pthreads (this image has no libgomp) ---------------------------------
 * The reference itself cannot use more than one core (numpy ufuncs + scipy's coo_matvec are single-threaded); this is the
 * "what if the reference's algorithm were parallelised" figure bench.py reports next to the single-threaded port.
 * COO triplets must be sorted by row (a CSR expanded to COO is): thread t owns the entries of a contiguous row range,
 * so the row sums need no atomics and come out in the same order as the serial loop. */
#include <pthread.h>

/* numpy's float64 add.reduce over a contiguous array (np.sum / np.mean): pairwise summation with an 8-way unrolled
 * base case of at most 128 elements (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum).  The reference
 * takes np.mean(s[s != 0]) (iterativeCorrection.py:44), so the restatement sums in the same order. */
static double np_pairwise_sum(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

/* np.mean(s[keep & (s != 0)]) ; tmp has room for n doubles */
static double mean_nonzero(const double* s, const uint8_t* keep, int64_t n, double* tmp) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i)
        if ((keep == NULL || keep[i]) && s[i] != 0.0) tmp[cnt++] = s[i];
    return np_pairwise_sum(tmp, cnt) / (double)cnt;
}

double np_sum_f64(const double* a, int64_t n) { return np_pairwise_sum(a, n); }

typedef struct {
    int64_t k0, k1;
    const int32_t* row;
    const int32_t* col;
    double* data;
    double* s;
    int phase; /* 0: row sums, 1: rescale */
    int overflow;
} ice_mt_job;

static void* ice_mt_worker(void* arg) {
    ice_mt_job* j = (ice_mt_job*)arg;
    if (j->phase == 0) {
        for (int64_t k = j->k0; k < j->k1; ++k) j->s[j->row[k]] += j->data[k];
    } else {
        int of = 0;
        for (int64_t k = j->k0; k < j->k1; ++k) {
            double v = j->data[k];
            v *= j->s[j->row[k]];
            v *= j->s[j->col[k]];
            j->data[k] = v;
            if (v > 1e100) of = 1;
        }
        j->overflow = of;
    }
    return NULL;
}

int ice_loop_coo_mt(int64_t n, int64_t nnz, const int32_t* row, const int32_t* col, double* data, double* bias, int max_iter,
                    double tol, int threads, double* deviation_out) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    double* s = (double*)malloc(sizeof(double) * (size_t)n);
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    ice_mt_job* jobs = (ice_mt_job*)calloc((size_t)threads, sizeof(ice_mt_job));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    /* entry ranges cut at row boundaries */
    for (int t = 0; t < threads; ++t) {
        int64_t k = nnz * t / threads;
        while (t > 0 && k > 0 && k < nnz && row[k] == row[k - 1]) ++k;
        jobs[t].k0 = k;
        jobs[t].row = row;
        jobs[t].col = col;
        jobs[t].data = data;
        jobs[t].s = s;
    }
    for (int t = 0; t < threads; ++t) jobs[t].k1 = t + 1 < threads ? jobs[t + 1].k0 : nnz;
    int done = 0, overflow = 0;
    double dev = NAN;
    for (int it = 0; it < max_iter && !overflow; ++it) {
        done++;
        memset(s, 0, sizeof(double) * (size_t)n);
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) {
                const double mean = mean_nonzero(s, NULL, n, tmp);
                dev = 0.0;
                for (int64_t i = 0; i < n; ++i) {
                    s[i] = s[i] / mean;
                    bias[i] *= s[i];
                    const double d = fabs(s[i] - 1.0);
                    if (d > dev || d != d) dev = d;
                    s[i] = 1.0 / s[i];
                }
            }
            for (int t = 0; t < threads; ++t) {
                jobs[t].phase = phase;
                jobs[t].overflow = 0;
                if (t > 0) pthread_create(&th[t], NULL, ice_mt_worker, &jobs[t]);
            }
            ice_mt_worker(&jobs[0]);
            for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
            if (phase == 1)
                for (int t = 0; t < threads; ++t) overflow |= jobs[t].overflow;
        }
        if (dev < tol) break;
    }
    free(s);
    free(tmp);
    free(jobs);
    free(th);
    if (deviation_out) *deviation_out = dev;
    return overflow ? -1 : done;
}

/* ---- the reference's loop on a CSR store with a bin mask, all host cores ------------------------------------------------
 * What iterativeCorrection() computes on matrix[keep][:, keep] (the matrix hicCorrectMatrix hands it after maskBins),
 * without building that sub-matrix: entries whose row or column is masked are skipped.  Arithmetic and order are the
 * reference's: row sums accumulate sequentially in stored order (scipy's coo_matvec over row-major triplets,
 * iterativeCorrection.py:42), every kept entry is rescaled in place by s[row] then s[col] (:56-57), the 1e100 scan (:58),
 * the strict '<' test after the rescale (:72).  The O(n) part runs over the kept bins in order, like numpy on the
 * compacted vectors.  Threads own contiguous row ranges (balanced by entries), so nothing depends on the thread count.
 * data[] is overwritten with the scaled values of the last iteration (masked entries untouched).
 * bias[] is indexed by ORIGINAL bin (masked bins keep 1).  Returns iterations done, -1 on overflow. */
typedef struct {
    int64_t r0, r1;
    const int64_t* indptr;
    const int32_t* idx;
    double* data;
    const uint8_t* keep;
    double* s;
    int phase;
    int overflow;
} ice_csr_job;

static void* ice_csr_worker(void* arg) {
    ice_csr_job* j = (ice_csr_job*)arg;
    if (j->phase == 0) {
        for (int64_t i = j->r0; i < j->r1; ++i) {
            if (!j->keep[i]) continue;
            double acc = 0.0;
            for (int64_t k = j->indptr[i]; k < j->indptr[i + 1]; ++k)
                if (j->keep[j->idx[k]]) acc += j->data[k];
            j->s[i] = acc;
        }
    } else {
        int of = 0;
        for (int64_t i = j->r0; i < j->r1; ++i) {
            if (!j->keep[i]) continue;
            const double si = j->s[i];
            for (int64_t k = j->indptr[i]; k < j->indptr[i + 1]; ++k) {
                const int32_t c = j->idx[k];
                if (!j->keep[c]) continue;
                double v = j->data[k];
                v *= si;
                v *= j->s[c];
                j->data[k] = v;
                if (v > 1e100) of = 1;
            }
        }
        j->overflow = of;
    }
    return NULL;
}

int ice_loop_csr_masked_mt(int64_t n, const int64_t* indptr, const int32_t* idx, double* data, const uint8_t* keep,
                           double* bias, int max_iter, double tol, int threads, double* deviation_out) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    double* s = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    ice_csr_job* jobs = (ice_csr_job*)calloc((size_t)threads, sizeof(ice_csr_job));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    const int64_t nnz = indptr[n] - indptr[0];
    {
        int64_t r = 0;
        for (int t = 0; t < threads; ++t) {
            jobs[t].r0 = r;
            const int64_t target = indptr[0] + nnz * (t + 1) / threads;
            while (r < n && indptr[r] < target) ++r;
            if (t == threads - 1) r = n;
            jobs[t].r1 = r;
            jobs[t].indptr = indptr;
            jobs[t].idx = idx;
            jobs[t].data = data;
            jobs[t].keep = keep;
            jobs[t].s = s;
        }
    }
    int done = 0, overflow = 0;
    double dev = NAN;
    for (int it = 0; it < max_iter && !overflow; ++it) {
        done++;
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == 1) {
                const double mean = mean_nonzero(s, keep, n, tmp);
                dev = 0.0;
                for (int64_t i = 0; i < n; ++i) {
                    if (!keep[i]) continue;
                    s[i] = s[i] / mean;
                    bias[i] *= s[i];
                    const double d = fabs(s[i] - 1.0);
                    if (d > dev || d != d) dev = d;
                    s[i] = 1.0 / s[i];
                }
            }
            for (int t = 0; t < threads; ++t) {
                jobs[t].phase = phase;
                jobs[t].overflow = 0;
                if (t > 0) pthread_create(&th[t], NULL, ice_csr_worker, &jobs[t]);
            }
            ice_csr_worker(&jobs[0]);
            for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
            if (phase == 1)
                for (int t = 0; t < threads; ++t) overflow |= jobs[t].overflow;
        }
        if (dev < tol) break;
    }
    free(s);
    free(tmp);
    free(jobs);
    free(th);
    if (deviation_out) *deviation_out = dev;
    return overflow ? -1 : done;
}

/* row sums and diagonal over the kept bins (hicCorrectMatrix.py:588, 614): sequential per row, threads over rows */
typedef struct {
    int64_t r0, r1;
    const int64_t* indptr;
    const int32_t* idx;
    const double* data;
    const uint8_t* keep;
    double* rs;
    double* dg;
} rowstat_job;

static void* rowstat_worker(void* arg) {
    rowstat_job* j = (rowstat_job*)arg;
    for (int64_t i = j->r0; i < j->r1; ++i) {
        double acc = 0.0, d = 0.0;
        if (j->keep == NULL || j->keep[i]) {
            for (int64_t k = j->indptr[i]; k < j->indptr[i + 1]; ++k) {
                const int32_t c = j->idx[k];
                if (j->keep != NULL && !j->keep[c]) continue;
                acc += j->data[k];
                if (c == i) d += j->data[k];
            }
        }
        j->rs[i] = acc;
        j->dg[i] = d;
    }
    return NULL;
}

void row_stats_csr_mt(int64_t n, const int64_t* indptr, const int32_t* idx, const double* data, const uint8_t* keep,
                      double* rs, double* dg, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    rowstat_job* jobs = (rowstat_job*)calloc((size_t)threads, sizeof(rowstat_job));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; ++t) {
        jobs[t].r0 = n * t / threads;
        jobs[t].r1 = n * (t + 1) / threads;
        jobs[t].indptr = indptr;
        jobs[t].idx = idx;
        jobs[t].data = data;
        jobs[t].keep = keep;
        jobs[t].rs = rs;
        jobs[t].dg = dg;
        if (t > 0) pthread_create(&th[t], NULL, rowstat_worker, &jobs[t]);
    }
    rowstat_worker(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
}

/* y = x o ((A + eps I) u) [+ v o p] with threads over rows (same per-row arithmetic as kr_matvec) */
typedef struct {
    int64_t r0, r1;
    const int64_t* indptr;
    const int32_t* idx;
    const double* val;
    const double *u, *x, *v, *p;
    double eps;
    double* out;
} krmv_job;

static void* krmv_worker(void* arg) {
    krmv_job* j = (krmv_job*)arg;
    for (int64_t i = j->r0; i < j->r1; ++i) {
        double acc = 0.0;
        for (int64_t k = j->indptr[i]; k < j->indptr[i + 1]; ++k) acc += j->val[k] * j->u[j->idx[k]];
        double r = j->x[i] * (acc + j->eps * j->u[i]);
        if (j->v && j->p) r += j->v[i] * j->p[i];
        j->out[i] = r;
    }
    return NULL;
}

void kr_matvec_mt(int64_t n, const int64_t* indptr, const int32_t* idx, const double* val, const double* u, const double* x,
                  const double* v, const double* p, double eps, double* out, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    krmv_job* jobs = (krmv_job*)calloc((size_t)threads, sizeof(krmv_job));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    const int64_t nnz = indptr[n] - indptr[0];
    int64_t r = 0;
    for (int t = 0; t < threads; ++t) {
        jobs[t].r0 = r;
        const int64_t target = indptr[0] + nnz * (t + 1) / threads;
        while (r < n && indptr[r] < target) ++r;
        if (t == threads - 1) r = n;
        jobs[t].r1 = r;
        jobs[t].indptr = indptr;
        jobs[t].idx = idx;
        jobs[t].val = val;
        jobs[t].u = u;
        jobs[t].x = x;
        jobs[t].v = v;
        jobs[t].p = p;
        jobs[t].eps = eps;
        jobs[t].out = out;
        if (t > 0) pthread_create(&th[t], NULL, krmv_worker, &jobs[t]);
    }
    krmv_worker(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
}

/* returns iterations done; *deviation_out = last max|s-1|; -1 on overflow (> 1e100) */
int ice_loop_coo(int64_t n, int64_t nnz, const int32_t* row, const int32_t* col, double* data, double* bias,
                 int max_iter, double tol, int threads, double* deviation_out) {
    double* s = (double*)malloc(sizeof(double) * (size_t)n);
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int done = 0;
    double dev = NAN;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    for (int it = 0; it < max_iter; ++it) {
        done++;
        memset(s, 0, sizeof(double) * (size_t)n);
        if (threads == 1) {
            for (int64_t k = 0; k < nnz; ++k) s[row[k]] += data[k]; /* W.sum(axis=1), line 42 */
        } else {
#pragma omp parallel for schedule(static)
            for (int64_t k = 0; k < nnz; ++k) {
#pragma omp atomic
                s[row[k]] += data[k];
            }
        }
        const double mean = mean_nonzero(s, NULL, n, tmp); /* line 44 */
        dev = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            s[i] = s[i] / mean;
            bias[i] *= s[i];
            const double d = fabs(s[i] - 1.0);
            if (d > dev || d != d) dev = d;
            s[i] = 1.0 / s[i];
        }
        int overflow = 0;
#pragma omp parallel for schedule(static) reduction(| : overflow)
        for (int64_t k = 0; k < nnz; ++k) {
            double v = data[k];
            v *= s[row[k]]; /* line 56 */
            v *= s[col[k]]; /* line 57 */
            data[k] = v;
            if (v > 1e100) overflow |= 1; /* line 58 */
        }
        if (overflow) {
            free(s);
            free(tmp);
            return -1;
        }
        if (dev < tol) break; /* line 72 */
    }
    free(s);
    free(tmp);
    if (deviation_out) *deviation_out = dev;
    return done;
}

void kr_matvec(int64_t n, const int64_t* indptr, const int32_t* idx, const double* val, const double* u,
               const double* x, const double* v, const double* p, double eps, double* out, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) acc += val[k] * u[idx[k]];
        double r = x[i] * (acc + eps * u[i]);
        if (v && p) r += v[i] * p[i];
        out[i] = r;
    }
}

/* ---------------- synthetic matrix (same function as hb_synth.cu) ---------------- */
static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xbf58476d1ce4e5b9ull;
    z ^= z >> 27;
    z *= 0x94d049bb133111ebull;
    z ^= z >> 31;
    return z;
}
static inline uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t z = mix64(seed + 0x9E3779B97F4A7C15ull * (a + 1));
    return mix64(z ^ (0xD1B54A32D192ED03ull * (b + 1)));
}
static inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
#define TAG_BASE (1ull << 40)

typedef struct {
    uint64_t seed;
    int32_t nchrom, band, n_trans, reserved;
    double d0, gamma, lambda0, low_frac, empty_frac, lambda_trans;
} synth_params;

static double count_of(uint64_t hp, double lam, double ga, double gb) {
    const double u2 = u01(mix64(hp + 0x9E3779B97F4A7C15ull));
    volatile double t = lam * ga; /* volatile: forbid FMA contraction, one rounding per operation */
    t = t * gb;
    t = t * u2;
    return 1.0 + floor(t);
}

static int cmp_i64(const void* a, const void* b) {
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

/* pass 1 (idx == NULL): fills indptr[n+1]; pass 2: fills idx/val.  Returns nnz. */
int64_t synth_generate(const synth_params* p, const int64_t* chrom_bins, int64_t* indptr, int32_t* idx, double* val) {
    int64_t n = 0;
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->nchrom + 1));
    off[0] = 0;
    for (int c = 0; c < p->nchrom; ++c) off[c + 1] = off[c] + chrom_bins[c];
    n = off[p->nchrom];
    double* g = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (u01(hash3(p->seed, (uint64_t)i, TAG_BASE + 0)) < p->empty_frac) g[i] = 0.0;
        else if (u01(hash3(p->seed, (uint64_t)i, TAG_BASE + 1)) < p->low_frac) g[i] = 0.02;
        else {
            const double u = u01(hash3(p->seed, (uint64_t)i, TAG_BASE + 2));
            volatile double t = u * u;
            t = t * u;
            t = 4.8 * t;
            g[i] = 0.2 + t;
        }
    }
    int64_t cs[64];
    int ncs = 0;
    for (int k = 0; k < p->n_trans && k < 32; ++k) cs[ncs++] = (int64_t)(hash3(p->seed, (uint64_t)k, TAG_BASE + 3) % (uint64_t)n);
    qsort(cs, (size_t)ncs, sizeof(int64_t), cmp_i64);
    {
        int m = 0;
        for (int k = 0; k < ncs; ++k)
            if (m == 0 || cs[k] != cs[m - 1]) cs[m++] = cs[k];
        ncs = m;
    }
    const int fill = idx != NULL;
    int64_t* chrom_of = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    for (int c = 0; c < p->nchrom; ++c)
        for (int64_t i = off[c]; i < off[c + 1]; ++i) chrom_of[i] = c;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        int64_t w = fill ? indptr[i] : 0;
        int64_t c_n = 0;
        if (g[i] != 0.0) {
            const int64_t clo = off[chrom_of[i]], chi = off[chrom_of[i] + 1];
            const int64_t blo = i - p->band > clo ? i - p->band : clo;
            const int64_t bhi = i + p->band < chi - 1 ? i + p->band : chi - 1;
            int64_t tj[64];
            int nt = 0;
            for (int k = 0; k < ncs; ++k) {
                int64_t j = (cs[k] - i) % n;
                if (j < 0) j += n;
                if ((j < blo || j > bhi) && g[j] != 0.0) tj[nt++] = j;
            }
            qsort(tj, (size_t)nt, sizeof(int64_t), cmp_i64);
            int t = 0;
            for (; t < nt && tj[t] < blo; ++t) {
                if (fill) {
                    const int64_t a = i < tj[t] ? i : tj[t], b = i < tj[t] ? tj[t] : i;
                    idx[w] = (int32_t)tj[t];
                    val[w] = count_of(hash3(p->seed, (uint64_t)a, (uint64_t)b), p->lambda_trans, g[a], g[b]);
                    w++;
                }
                c_n++;
            }
            for (int64_t j = blo; j <= bhi; ++j) {
                if (g[j] == 0.0) continue;
                const int64_t a = i < j ? i : j, b = i < j ? j : i;
                const uint64_t hp = hash3(p->seed, (uint64_t)a, (uint64_t)b);
                const double d1 = (double)(b - a + 1);
                const double u = u01(hp);
                int ex;
                if (p->gamma == 1.0) {
                    volatile double t2 = u * d1;
                    ex = t2 < p->d0;
                } else ex = u < pow(p->d0 / d1, p->gamma);
                if (!ex) continue;
                if (fill) {
                    volatile double q = (double)(b - a) / p->d0;
                    q = 1.0 + q;
                    q = p->lambda0 / q;
                    idx[w] = (int32_t)j;
                    val[w] = count_of(hp, q, g[a], g[b]);
                    w++;
                }
                c_n++;
            }
            for (; t < nt; ++t) {
                if (fill) {
                    const int64_t a = i < tj[t] ? i : tj[t], b = i < tj[t] ? tj[t] : i;
                    idx[w] = (int32_t)tj[t];
                    val[w] = count_of(hash3(p->seed, (uint64_t)a, (uint64_t)b), p->lambda_trans, g[a], g[b]);
                    w++;
                }
                c_n++;
            }
        }
        if (!fill) indptr[i + 1] = c_n;
    }
    int64_t nnz;
    if (!fill) {
        indptr[0] = 0;
        for (int64_t i = 0; i < n; ++i) indptr[i + 1] += indptr[i];
    }
    nnz = indptr[n];
    free(off);
    free(g);
    free(chrom_of);
    return nnz;
}
