/* CPU restatement (TEST / BASELINE INFRASTRUCTURE ONLY -- never on the product path) of the reference's
 * compute_MM(::SPMF_NEP, S, V) for S = lambda*I  (src/NEPTypes.jl:276-319):
 *
 *     Z = zeros(n, k);  for i = 1:p   VFi = V * f_i(lambda);  Z .+= A_i * VFi   end
 *
 * i.e. p separate sparse-times-dense products, each with its own index arrays, accumulated into Z -- the loop structure of the
 * reference (:296-316), with Julia's SparseMatrixCSC * Matrix product (SparseArrays `mul!`: for every column of the dense
 * block, for every column j of A, for every stored entry (i,j): C[i,col] += A[i,j]*B[j,col]) as the inner kernel.  The
 * reference runs this on one thread; `threads > 1` splits the k dense columns (and, for k < threads, row blocks of a CSR copy
 * built by the caller) over OpenMP threads so that the baseline gets every host core.
 *
 * Arrays: CSC (colptr[n+1], rowval[nnz], nzval[nnz] real), 0-based int64 indices like the fixtures; V, Z column-major n x k
 * complex (interleaved re/im).  Parity pinned in tests/test_oracle_golden.py against oracle/nep.py:compute_MM.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Z(:, c) += A * (f * V(:, c)) for the columns [c0, c1), CSC, the SparseArrays loop order */
static void csc_times_dense_cols(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval, double fr, double fi,
                                 const double* V, double* Z, int c0, int c1) {
    for (int c = c0; c < c1; ++c) {
        const double* v = V + 2 * (size_t)n * c;
        double* z = Z + 2 * (size_t)n * c;
        for (int64_t j = 0; j < n; ++j) {
            const double xr = fr * v[2 * j] - fi * v[2 * j + 1];
            const double xi = fr * v[2 * j + 1] + fi * v[2 * j];
            for (int64_t e = colptr[j]; e < colptr[j + 1]; ++e) {
                const int64_t i = rowval[e];
                z[2 * i] += nzval[e] * xr;
                z[2 * i + 1] += nzval[e] * xi;
            }
        }
    }
}

/* the reference's loop, dense columns split over `threads` threads (threads = 1: exactly the reference's serial order) */
int oracle_spmf_mm_csc(int64_t n, int p, const int64_t* const* colptr, const int64_t* const* rowval, const double* const* nzval,
                       const double* f /* p complex */, int k, const double* V, double* Z, int threads) {
    memset(Z, 0, sizeof(double) * 2 * (size_t)n * k);
    if (threads < 1) threads = 1;
    if (threads > k) threads = k;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int t = 0; t < threads; ++t) {
        const int c0 = (int)((int64_t)k * t / threads), c1 = (int)((int64_t)k * (t + 1) / threads);
        for (int i = 0; i < p; ++i) csc_times_dense_cols(n, colptr[i], rowval[i], nzval[i], f[2 * i], f[2 * i + 1], V, Z, c0, c1);
    }
    return 0;
}

/* the same sum with CSR copies of the A_i, rows split over the threads: the all-core variant for narrow blocks (k = 1), where
 * splitting dense columns leaves cores idle.  rowptr / colind / nzval per term, V and Z as above. */
int oracle_spmf_mm_csr(int64_t n, int p, const int64_t* const* rowptr, const int64_t* const* colind, const double* const* nzval,
                       const double* f, int k, const double* V, double* Z, int threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static, 4096) num_threads(threads)
    for (int64_t r = 0; r < n; ++r) {
        for (int c = 0; c < k; ++c) {
            double zr = 0.0, zi = 0.0;
            for (int i = 0; i < p; ++i) {
                double sr = 0.0, si = 0.0;
                const double* v = V + 2 * (size_t)n * c;
                for (int64_t e = rowptr[i][r]; e < rowptr[i][r + 1]; ++e) {
                    const int64_t j = colind[i][e];
                    sr += nzval[i][e] * v[2 * j];
                    si += nzval[i][e] * v[2 * j + 1];
                }
                zr += f[2 * i] * sr - f[2 * i + 1] * si;
                zi += f[2 * i] * si + f[2 * i + 1] * sr;
            }
            Z[2 * ((size_t)n * c + r)] = zr;
            Z[2 * ((size_t)n * c + r) + 1] = zi;
        }
    }
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
