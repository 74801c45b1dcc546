/* oracle/replay.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C restatement of the reference's CPU executor for the H.C hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this; the CUDA library never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks b2o_seq_matvec against
 * sigma vectors produced by the unmodified reference (oracle/_ref/b2ref_*,
 * TensorFunctions::operator() -> BatchGEMMSeq::operator()) stored in the
 * committed fixtures under tests/golden/.
 *
 * Reference being restated (paths under /root/reference/src/core):
 *   batch_gemm.hpp:219-235   single_xgemm : one row-major GEMM,
 *                            C = (alpha*scale) * op(A) * op(B) + beta * C
 *   batch_gemm.hpp:1613-1688 BatchGEMMSeq::operator() (Tasked branch):
 *                            for every pair i:  work = op(A0[i] + cshift) * op(B0[i])
 *                                               C1[i] + vshift += alpha1*scale * op(A1[i]) * work
 *                            per-thread sigma replicas, then a sum over threads
 *   batch_gemm.hpp:81-111    cblas_xgemm_batch : grouped GEMM list
 *   matrix_functions.hpp:968-992  row-major GEMM through column-major dgemm
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* C[m x n] (ldc) = alpha * op(A) * op(B) + beta * C, row-major, ta/tb in {0,1}.
 * beta == 0 overwrites (BLAS semantics: C is not read). */
static void gemm_rm(int ta, int tb, int m, int n, int k, double alpha,
                    const double *a, int lda, const double *b, int ldb,
                    double beta, double *c, int ldc) {
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int l = 0; l < k; l++) {
                double av = ta ? a[(size_t)l * lda + i] : a[(size_t)i * lda + l];
                double bv = tb ? b[(size_t)j * ldb + l] : b[(size_t)l * ldb + j];
                s += av * bv;
            }
            double *cp = &c[(size_t)i * ldc + j];
            *cp = beta == 0.0 ? alpha * s : alpha * s + beta * *cp;
        }
}

/* Grouped GEMM list with the cblas_dgemm_batch signature the reference's
 * BatchGEMM::perform uses (batch_gemm.hpp:81-111, 339-357). */
void b2o_dgemm_batch(int64_t ngroups, const int32_t *ta, const int32_t *tb,
                     const int32_t *m, const int32_t *n, const int32_t *k,
                     const double *alpha, const double *const *a,
                     const int32_t *lda, const double *const *b,
                     const int32_t *ldb, const double *beta, double *const *c,
                     const int32_t *ldc, const int32_t *group_size) {
    int64_t z = 0;
    for (int64_t g = 0; g < ngroups; g++)
        for (int32_t j = 0; j < group_size[g]; j++, z++)
            gemm_rm(ta[g], tb[g], m[g], n[g], k[g], alpha[g], a[z], lda[g], b[z],
                    ldb[g], beta[g], c[z], ldc[g]);
}

/* sigma += H.c by replaying the pair list.
 *   a0_off : offset (doubles) of the wavefunction window inside c
 *   b0     : host pointer of the operator block of GEMM 0
 *   a1     : host pointer of the operator block of GEMM 1
 *   c1_off : offset (doubles) of the sigma window inside v
 * nthreads > 1 reproduces the reference's static partition over pairs with
 * one private sigma per thread, summed afterwards in thread order. */
void b2o_seq_matvec(int64_t npairs, const int32_t *ta0, const int32_t *tb0,
                    const int32_t *m0, const int32_t *n0, const int32_t *k0,
                    const int32_t *lda0, const int32_t *ldb0,
                    const int32_t *ldc0, const double *alpha0,
                    const double *beta0, const int64_t *a0_off,
                    const double *const *b0, const int32_t *ta1,
                    const int32_t *tb1, const int32_t *m1, const int32_t *n1,
                    const int32_t *k1, const int32_t *lda1,
                    const int32_t *ldb1, const int32_t *ldc1,
                    const double *alpha1, const double *beta1,
                    const double *const *a1, const int64_t *c1_off,
                    int64_t max_work, const double *c, double *v, int64_t vsize,
                    double scale, int nthreads) {
    if (npairs == 0)
        return;
    if (nthreads < 1)
        nthreads = 1;
    double **vts = (double **)calloc((size_t)nthreads, sizeof(double *));
    vts[0] = v;
    for (int t = 1; t < nthreads; t++)
        vts[t] = (double *)calloc((size_t)vsize, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        int tid = 0, nt = 1;
#ifdef _OPENMP
        tid = omp_get_thread_num(), nt = omp_get_num_threads();
#endif
        double *work = (double *)malloc(sizeof(double) * (size_t)(max_work > 0 ? max_work : 1));
        /* schedule(static): contiguous chunks */
        int64_t chunk = (npairs + nt - 1) / nt;
        int64_t lo = tid * chunk, hi = lo + chunk < npairs ? lo + chunk : npairs;
        for (int64_t i = lo; i < hi; i++) {
            gemm_rm(ta0[i], tb0[i], m0[i], n0[i], k0[i], alpha0[i], c + a0_off[i],
                    lda0[i], b0[i], ldb0[i], beta0[i], work, ldc0[i]);
            gemm_rm(ta1[i], tb1[i], m1[i], n1[i], k1[i], alpha1[i] * scale, a1[i],
                    lda1[i], work, ldb1[i], beta1[i], vts[tid] + c1_off[i],
                    ldc1[i]);
        }
        free(work);
    }
    for (int t = 1; t < nthreads; t++) {
        for (int64_t j = 0; j < vsize; j++)
            v[j] += vts[t][j];
        free(vts[t]);
    }
    free(vts);
}

/* Same replay with every GEMM handed to a Fortran-ABI dgemm (the reference's
 * xgemm<double>, matrix_functions.hpp:335-351; row-major through the
 * (B, A) operand swap of single_xgemm, batch_gemm.hpp:219-235).  The caller
 * passes the dgemm_ entry of the BLAS the reference would link (OpenBLAS here);
 * used for the CPU baseline timing, on all the host threads it is given. */
typedef void (*b2o_dgemm_fn)(const char *, const char *, const int *, const int *,
                             const int *, const double *, const double *,
                             const int *, const double *, const int *,
                             const double *, double *, const int *);

void b2o_seq_matvec_blas(b2o_dgemm_fn dgemm, int64_t npairs, const int32_t *ta0,
                         const int32_t *tb0, const int32_t *m0, const int32_t *n0,
                         const int32_t *k0, const int32_t *lda0, const int32_t *ldb0,
                         const int32_t *ldc0, const double *alpha0,
                         const double *beta0, const int64_t *a0_off,
                         const double *const *b0, const int32_t *ta1,
                         const int32_t *tb1, const int32_t *m1, const int32_t *n1,
                         const int32_t *k1, const int32_t *lda1, const int32_t *ldb1,
                         const int32_t *ldc1, const double *alpha1,
                         const double *beta1, const double *const *a1,
                         const int64_t *c1_off, int64_t max_work, const double *c,
                         double *v, int64_t vsize, double scale, int nthreads) {
    if (npairs == 0)
        return;
    if (nthreads < 1)
        nthreads = 1;
    double **vts = (double **)calloc((size_t)nthreads, sizeof(double *));
    vts[0] = v;
    for (int t = 1; t < nthreads; t++)
        vts[t] = (double *)calloc((size_t)vsize, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *work = (double *)malloc(sizeof(double) * (size_t)(max_work > 0 ? max_work : 1));
        double t0 = omp_get_wtime();
#ifdef _OPENMP
#pragma omp for schedule(static) nowait
#endif
        for (int64_t i = 0; i < npairs; i++) {
            int m = m0[i], n = n0[i], k = k0[i], la = lda0[i], lb = ldb0[i], lc = ldc0[i];
            dgemm(tb0[i] ? "t" : "n", ta0[i] ? "t" : "n", &n, &m, &k, &alpha0[i], b0[i], &lb,
                  c + a0_off[i], &la, &beta0[i], work, &lc);
            double al = alpha1[i] * scale;
            m = m1[i], n = n1[i], k = k1[i], la = lda1[i], lb = ldb1[i], lc = ldc1[i];
            dgemm(tb1[i] ? "t" : "n", ta1[i] ? "t" : "n", &n, &m, &k, &al, work, &lb, a1[i], &la,
                  &beta1[i], vts[tid] + c1_off[i], &lc);
        }
        if (getenv("B2O_DEBUG")) printf("tid %d loop %.4f s\n", tid, omp_get_wtime() - t0);
        free(work);
    }
    for (int t = 1; t < nthreads; t++) {
        for (int64_t j = 0; j < vsize; j++)
            v[j] += vts[t][j];
        free(vts[t]);
    }
    free(vts);
}
