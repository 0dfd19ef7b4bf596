/* CPU oracle (plain C) for the pyremap weight-application path.
 * TEST INFRASTRUCTURE ONLY -- never linked into or called by the product.
 *
 * Restates the published algorithm of scipy's `_sparsetools.csr_matvecs`
 * (third-party dependency of /root/reference, called from
 * pyremap/remapper/remap_numpy.py:264,265,268) and the closed form of
 * pyremap/remapper/remap_numpy.py:258-278 around it.
 *
 * Build with -ffp-contract=off: the reference kernel rounds the product and the
 * sum separately (no FMA), and bit-for-bit parity depends on that.
 * Rows are independent, so the optional pthread row split (threads > 1) does not
 * change any result bit; it only provides a many-core "port" baseline.
 */
#include <stdint.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

typedef struct {
    int64_t row_begin, row_end, n_vecs;
    const int32_t *Ap, *Aj;
    const double *Ax, *frac_b, *X;
    const uint8_t *valid;
    int mode;
    double thr;
    double *Y;
    uint8_t *keep_out;
} job_t;

/* Y[i,:] = sum_jj Ax[jj] * X[Aj[jj],:]  -- stored order, mul then add, from +0.0 */
static void *matvecs_rows(void *arg) {
    const job_t *j = (const job_t *)arg;
    const int64_t n_vecs = j->n_vecs;
    for (int64_t i = j->row_begin; i < j->row_end; ++i) {
        double *y = j->Y + i * n_vecs;
        for (int64_t k = 0; k < n_vecs; ++k) y[k] = 0.0;
        for (int32_t jj = j->Ap[i]; jj < j->Ap[i + 1]; ++jj) {
            const double a = j->Ax[jj];
            const double *x = j->X + (int64_t)j->Aj[jj] * n_vecs;
            for (int64_t k = 0; k < n_vecs; ++k) {
                double p = a * x[k];
                y[k] = y[k] + p;
            }
        }
    }
    return 0;
}

/* mode 0: raw product.
 * mode 1: unmasked branch (remap_numpy.py:268-278): keep = frac_b[i] > 0,
 *         value = (S@x)/frac_b where kept, NaN elsewhere.
 * mode 2: masked branch (remap_numpy.py:263-266,277-278) with the validity
 *         either taken from `valid` (bytes, [n_col,K]) or, when valid == NULL,
 *         from x == x: num = S@(valid ? x : 0.0), den = S@(valid ? 1.0 : 0.0),
 *         keep = den > thr, value = num/den where kept, NaN elsewhere.
 * keep_out (nullable) receives the keep flags. */
static void *fused_rows(void *arg) {
    const job_t *j = (const job_t *)arg;
    const int64_t n_vecs = j->n_vecs;
    const int mode = j->mode;
    for (int64_t i = j->row_begin; i < j->row_end; ++i) {
        double *y = j->Y + i * n_vecs;
        for (int64_t k = 0; k < n_vecs; ++k) {
            double num = 0.0, den = 0.0;
            for (int32_t jj = j->Ap[i]; jj < j->Ap[i + 1]; ++jj) {
                const double a = j->Ax[jj];
                const int64_t at = (int64_t)j->Aj[jj] * n_vecs + k;
                double x = j->X[at];
                if (mode == 2) {
                    int ok = j->valid ? (j->valid[at] != 0) : (x == x);
                    double m = ok ? 1.0 : 0.0;
                    double x0 = ok ? x : 0.0;
                    double p = a * x0;
                    num = num + p;
                    double q = a * m;
                    den = den + q;
                } else {
                    double p = a * x;
                    num = num + p;
                }
            }
            int keep = 1;
            if (mode == 1) { den = j->frac_b[i]; keep = den > 0.0; }
            if (mode == 2) { keep = den > j->thr; }
            if (mode != 0) num = keep ? num / den : NAN;
            y[k] = num;
            if (j->keep_out) j->keep_out[i * n_vecs + k] = (uint8_t)keep;
        }
    }
    return 0;
}

/* split rows so every thread gets about the same number of stored entries */
static void run_split(void *(*fn)(void *), job_t base, int64_t n_row, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads == 1 || n_row < 2 * threads) {
        base.row_begin = 0; base.row_end = n_row; fn(&base); return;
    }
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)threads);
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    const int64_t nnz = base.Ap[n_row];
    int64_t row = 0;
    for (int t = 0; t < threads; ++t) {
        jobs[t] = base;
        jobs[t].row_begin = row;
        if (t == threads - 1) {
            row = n_row;
        } else {
            const int64_t target = (nnz + n_row) * (t + 1) / threads;
            while (row < n_row && (int64_t)base.Ap[row] + row < target) ++row;
        }
        jobs[t].row_end = row;
        pthread_create(&tid[t], 0, fn, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], 0);
    free(jobs); free(tid);
}

void oracle_csr_matvecs(int64_t n_row, int64_t n_vecs, const int32_t *Ap,
                        const int32_t *Aj, const double *Ax, const double *X,
                        double *Y, int threads) {
    job_t b = {0, 0, n_vecs, Ap, Aj, Ax, 0, X, 0, 0, 0.0, Y, 0};
    run_split(matvecs_rows, b, n_row, threads);
}

void oracle_remap_fused(int64_t n_row, int64_t n_vecs, const int32_t *Ap,
                        const int32_t *Aj, const double *Ax,
                        const double *frac_b, const double *X,
                        const uint8_t *valid, int mode, double thr, double *Y,
                        uint8_t *keep_out, int threads) {
    job_t b = {0, 0, n_vecs, Ap, Aj, Ax, frac_b, X, valid, mode, thr, Y, keep_out};
    run_split(fused_rows, b, n_row, threads);
}
