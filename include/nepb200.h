/*
 * nepb200.h -- C ABI of libnepb200.so: the B200-native (sm_100a) hot path for NEP-PACK
 * (NonlinearEigenproblems.jl v1.1.1).  Reference citations are relative to the reference repo.
 *
 * The reference has no FFI on this path today (it is 100 % Julia, multiple dispatch); these entry
 * points are what a Julia shim `ccall`s from new methods of the reference's own plugin interfaces
 * (see INTEGRATION.md and nonlineareigenproblems.jl_b200/julia/NEPB200.jl):
 *
 *   compute_Mder / compute_Mlincomb[!] / compute_MM   src/NEPCore.jl:89,113-160,192
 *   AbstractSPMF, get_Av / get_fv, SPMF_NEP           src/NEPTypes.jl:96-113,162-237
 *   LinSolver / lin_solve                             src/LinSolvers.jl:100,135-137,157-159
 *   LinSolverCreator / create_linsolver               src/LinSolverCreators.jl:11,35,107
 *   integrate_interval(MatrixTrapezoidal, ...)        src/method_contour_common.jl:61-94
 *   orthogonalize_and_normalize!(V,w,h,DGKS())        call sites src/method_iar.jl:107, method_tiar.jl:128
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (NEPB_E_*); nepb_last_error() gives the
 *     message of the last failure on the calling thread.  There is no CPU fallback: without a
 *     CUDA device every compute entry point fails with NEPB_E_CUDA.
 *   - host pointers unless the name ends in _dev; host dense arrays are column-major (Julia layout),
 *     complex = interleaved (re,im) doubles (layout of Julia ComplexF64 / C99 double _Complex).
 *   - sparse input is Julia's SparseMatrixCSC: int64 colptr[n+1], int64 rowval[nnz], index_base 1
 *     (0 accepted for C callers), row indices ascending within a column.
 *   - functions f_i never cross the ABI: the caller evaluates them (scalars or small matrix
 *     functions, exactly as src/NEPTypes.jl:993-1004 does) and passes coefficient blocks.
 */
#ifndef NEPB200_H
#define NEPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEPB_OK 0
#define NEPB_E_INVALID (-1)   /* bad argument / inconsistent sizes (reference: error("...")) */
#define NEPB_E_CUDA (-2)      /* CUDA runtime failure, or no device */
#define NEPB_E_SINGULAR (-3)  /* zero / non-finite pivot (reference: LinearAlgebra.SingularException) */
#define NEPB_E_NOMEM (-4)
#define NEPB_E_UNSUPPORTED (-5)

/* coefficient modes of nepb_spmf_apply*: Z(n x q) = sum_i A_i * (V(n x k) * C_i(k x q)) */
#define NEPB_COEF_SCALAR 0  /* C_i = c_i*I, q == k;          C = c[p]                              */
#define NEPB_COEF_DIAG 1    /* C_i = diag(c_i[0..k-1]), q==k; C = p x k column-major (C[i + p*s])  */
#define NEPB_COEF_GENERAL 2 /* C_i dense k x q;              C = p blocks, block i column-major    */

typedef struct nepb_spmf nepb_spmf;   /* SPMF operator resident in HBM (union-pattern CSR)        */
typedef struct nepb_block nepb_block; /* dense n x k complex block resident in HBM (row-major)    */

/* ---- runtime ------------------------------------------------------------------------------- */
const char* nepb_version(void);
const char* nepb_last_error(void);
int nepb_device_count(int* count);
int nepb_set_device(int device);
/* All work of the calling process is enqueued on one stream; by default a stream the library owns.
 * Pass a cudaStream_t (e.g. torch's current stream) to time with external events. */
int nepb_set_stream(void* cuda_stream);
int nepb_synchronize(void);
/* CUDA-event timer on the library stream: start / stop return elapsed milliseconds. */
int nepb_timer_start(void);
int nepb_timer_stop(float* ms);
/* number of kernels this library launched since process start (bench.py "gpu_launches") */
int64_t nepb_launch_count(void);

/* ---- a1: SPMF operator (src/NEPTypes.jl:162-237; alignment :244-274) -------------------------- */
/* Builds the union sparsity pattern of the p matrices (what form_aligned_sparsity_patterns does),
 * converts CSC -> CSR, int64 -> int32 (overflow checked) and uploads one int32 column index per
 * union nonzero plus p interleaved values.  nzval[i] is double[nnz_i] (val_is_complex=0) or
 * interleaved complex (val_is_complex=1). */
int nepb_spmf_create(int64_t n, int p, const int64_t* const* colptr, const int64_t* const* rowval,
                     const void* const* nzval, int val_is_complex, int index_base, nepb_spmf** out);
int nepb_spmf_destroy(nepb_spmf* h);
int nepb_spmf_info(const nepb_spmf* h, int64_t* n, int* p, int64_t* nnz_union, int* val_is_complex);
/* union pattern, CSC, in the index base given at creation: colptr[n+1], rowval[nnz_union] */
int nepb_spmf_pattern(const nepb_spmf* h, int64_t* colptr, int64_t* rowval);
/* the same pattern as the device sees it: CSR, 0-based int32, plus csr_of_csc[nnz] (position in CSR
 * order of the j-th CSC nonzero) -- exposed so the integer work can be checked bit-exactly */
int nepb_spmf_pattern_csr(const nepb_spmf* h, int32_t* rowptr, int32_t* colind, int32_t* csr_of_csc);

/* ---- a5: compute_Mder for SPMF (src/NEPTypes.jl:336-367) ---------------------------------------- */
/* nzval_out[nnz_union] (complex, CSC order of nepb_spmf_pattern) = sum_i coef[i] * A_i.nzval */
int nepb_spmf_mder(const nepb_spmf* h, const double* coef /* p complex */, double* nzval_out);

/* ---- a2,a3,a4,a12,a13: the fused multi-term SpMM ------------------------------------------------ */
/* Z = sum_i A_i (V C_i); V is n x k, Z is n x q, host column-major with leading dimensions ldv, ldz
 * (in complex elements).  Covers compute_MM with diagonal S (NEPTypes.jl:299-311), M(lambda)*V,
 * batched residuals (errmeasure.jl:128-130) and, through NEPB_COEF_GENERAL, compute_Mlincomb
 * (NEPTypes.jl:972-1011,1130-1160) and the nleigs stacked product (method_nleigs.jl:456-472). */
int nepb_spmf_apply(const nepb_spmf* h, int mode, int k, int q, const double* V, int64_t ldv,
                    const double* C, double* Z, int64_t ldz);

/* ---- device-resident dense blocks (Krylov bases, probe matrices stay in HBM) ------------------- */
int nepb_block_create(int64_t n, int k, nepb_block** out);
int nepb_block_destroy(nepb_block* b);
/* copy columns [k0, k0+kc) from/to a host column-major array (ld in complex elements) */
int nepb_block_upload(nepb_block* b, int k0, int kc, const double* host, int64_t ld);
int nepb_block_download(const nepb_block* b, int k0, int kc, double* host, int64_t ld);
/* Page-lock a host array that is passed to the host-buffer entry points repeatedly (bases, probes, moments): its transfers
 * then run as direct DMA at PCIe rate.  Unregistered (pageable) arrays go through a pinned three-slot ring with the host copy,
 * the DMA and the layout change overlapped (csrc/hostcopy.cu). */
int nepb_host_register(void* ptr, int64_t bytes);
int nepb_host_unregister(void* ptr);
/* raw device pointer of the row-major n x k storage (complex interleaved) */
void* nepb_block_dev_ptr(nepb_block* b);
/* same product with operands already in HBM: no host traffic, asynchronous on the library stream */
int nepb_spmf_apply_block(const nepb_spmf* h, int mode, const nepb_block* V, int q, const double* C,
                          nepb_block* Z);
/* the same on column windows: Z[:, zcol0 : zcol0+q) = sum_i A_i (V[:, vcol0 : vcol0+k) C_i) */
int nepb_spmf_apply_block_ex(const nepb_spmf* h, int mode, const nepb_block* V, int vcol0, int k, int q, const double* C,
                             nepb_block* Z, int zcol0);
/* row tiles of the multi-column kernel (built lazily, host integer work): tiles of <= 32 consecutive rows whose distinct
 * columns (<= 192) are staged in shared memory; *distinct_total = sum over tiles = rows of V gathered per product
 * (vs nnz_union without tiling); all zero when the operator has a row with more than 192 nonzeros (untiled kernels) */
int nepb_spmf_tiles_info(const nepb_spmf* h, int64_t* ntiles, int64_t* distinct_total, int* max_distinct);
/* the two-dimensional tiles of the multi-column product (5..32 columns): line = dominant column offset found in the pattern
 * (0: none, the one-dimensional tiles are used), segments x seg_rows rows per tile, staged rows of V summed over the tiles */
int nepb_spmf_tiles2d_info(const nepb_spmf* h, int* line, int* segments, int* seg_rows, int64_t* ntiles, int64_t* distinct_total);
/* algorithmic HBM bytes of one nepb_spmf_apply_block call (SURVEY.md 8(d) formula) */
int64_t nepb_spmf_apply_bytes(const nepb_spmf* h, int mode, int k, int q);

/* ---- a6,a7: shifted solves -- device multifrontal LU of M(sigma) = sum_i coef_i A_i ------------------------
 * Replaces FactorizeLinSolver / BackslashLinSolver (src/LinSolvers.jl:109-159) and what their creators build
 * (src/LinSolverCreators.jl:21-37,62-122).  The symbolic analysis (fill-reducing ordering of A+A^T, elimination
 * tree, supernodes, index maps) is computed once per operator, lazily, and shared by every factorisation. */
typedef struct nepb_lu nepb_lu; /* numeric factors of one or several shifts, resident in HBM */
/* optional, before the first factorisation: ordering 0 = approximate minimum degree, 1 = natural; relax_leaf /
 * max_np = supernode relaxation and pivot-block width (<=64); user_perm[n] (index base of the operator, new -> old)
 * overrides the ordering.  Negative / zero / NULL keep the defaults. */
int nepb_lu_set_options(nepb_spmf* h, int ordering, int relax_leaf, int max_np, const int64_t* user_perm);
int nepb_lu_symbolic_info(const nepb_spmf* h, int64_t* nnz_factor, int64_t* front_entries, int* nfronts, int* nlevels,
                          int* max_front, double* flops);
/* integer results of the analysis (0-based): perm[n] new->old, etree parent[n] (-1 root), sn_ptr[nfronts+1],
 * sn_parent[nfronts], sn_rows[nfronts] (front order nf), sn_level[nfronts]; any pointer may be NULL */
int nepb_lu_symbolic_get(const nepb_spmf* h, int32_t* perm, int32_t* parent, int32_t* sn_ptr, int32_t* sn_parent,
                         int32_t* sn_rows, int32_t* sn_level);
/* the same analysis for an arbitrary pattern on the host only (no device needed); stats[8] = nnz(L+U), front
 * entries, fronts, levels, max front, max pivot block, factor multiply-adds, solve work rows; sn_ptr[n+1],
 * sn_rows[n], sn_level[n] (optional, caller-allocated for the worst case of n fronts) describe the fronts */
int nepb_lu_analyse_pattern(int64_t n, const int64_t* colptr, const int64_t* rowval, int index_base, int ordering,
                            int relax_leaf, int max_np, int32_t* perm, int32_t* parent, int32_t* colcount, double* stats,
                            int32_t* sn_ptr, int32_t* sn_rows, int32_t* sn_level);
/* Static pivoting, host only (lu_matching.cpp): the row matching that maximises the product of the diagonal magnitudes of
 * |A| (CSC, absval[nnz] >= 0) and the scalings that make matched entries 1 and all others <= 1 in modulus (Duff & Koster
 * 2001).  row_of_col[n] (0-based) is the row matched to each column, dr[n] / dc[n] the row / column scalings.  This is what
 * stands in for UMFPACK's numerical pivoting (src/LinSolvers.jl:116) when a factorisation on the plain pattern meets zero or
 * tiny pivots: nepb_lu_create then permutes and scales with the matching of the offending shift, repeats the symbolic
 * analysis on the row-permuted pattern and factorises again (NEPB_LU_MATCHING = 0 never / 1 on demand / 2 always).
 * Returns NEPB_E_SINGULAR for a structurally singular pattern. */
int nepb_lu_matching(int64_t n, const int64_t* colptr, const int64_t* rowval, int index_base, const double* absval,
                     int32_t* row_of_col, double* dr, double* dc);
/* factorise nshift matrices at once: coef is nshift x p complex, row s = (f_1(sigma_s) .. f_p(sigma_s)) */
int nepb_lu_create(const nepb_spmf* h, int nshift, const double* coef, nepb_lu** out);
int nepb_lu_destroy(nepb_lu* lu);
/* flags: bit0 = an exactly zero pivot was replaced (matrix singular to working precision), bit1 = non-finite pivot,
 * bit3 (8) = this factorisation uses the static-pivoting row matching / scaling;
 * nperturbed = pivots below eps*max|M_ij| that were lifted to that threshold; min_pivot_ratio = min|pivot|/max|M_ij| */
int nepb_lu_status(const nepb_lu* lu, int shift, int* flags, int* nperturbed, double* min_pivot_ratio);
/* lin_solve (src/LinSolvers.jl:135-137): X = M(sigma_shift)^-1 B, B and X host column-major n x nrhs (may alias).
 * refine_steps = maximum iterative-refinement steps (the reference's umfpack_refinements); berr_out (optional)
 * receives the final normwise backward error. */
int nepb_lu_solve(nepb_lu* lu, int shift, int nrhs, const double* B, int64_t ldb, double* X, int64_t ldx, int refine_steps,
                  double* berr_out);

/* device-resident lin_solve: X[:, xcol0 : xcol0+nrhs) = alpha * M(sigma_shift)^-1 B[:, bcol0 : bcol0+nrhs), alpha complex
 * (NULL = 1); used by iar / tiar / resinv loops that keep their vectors in HBM (y[:,1] = -lin_solve(..), method_iar.jl:103) */
int nepb_lu_solve_block(nepb_lu* lu, int shift, const nepb_block* B, int bcol0, int nrhs, nepb_block* X, int xcol0,
                        const double* alpha);
/* the same with iterative refinement against the fused SpMM residual (umfpack_refinements of LinSolverCreators.jl:66 on the
 * device-resident path) and the final backward error (berr_out optional).  Like nepb_lu_solve it returns NEPB_E_SINGULAR when
 * the backward error stays above 1e-9 after refinement, or after any solve with a factorisation whose pivots were replaced. */
int nepb_lu_solve_block_ex(nepb_lu* lu, int shift, const nepb_block* B, int bcol0, int nrhs, nepb_block* X, int xcol0,
                           const double* alpha, int refine_steps, double* berr_out);

/* ---- a11: dense tall-skinny blocks of iar / tiar (src/method_iar.jl:100-116, src/method_tiar.jl:119,128,187-189) ---
 * orthogonalize_and_normalize!(V[:, 0:k), w, h, DGKS()) with w = W[:, wcol] (V and W may be the same block), over the
 * first `rows` rows (0 = all): classical Gram-Schmidt, re-orthogonalised while ||w|| < ||h||/sqrt(2), w normalised in
 * place; h[k] (host complex) receives the accumulated coefficients, *nrm_out the norm before normalisation. */
int nepb_orth_dgks(const nepb_block* V, int k, nepb_block* W, int wcol, int64_t rows, double* h, double* nrm_out, int* sweeps);
/* Y[:, ycol0 : ycol0+q) = A[:, acol0 : acol0+ka) * C, C host column-major ka x q complex (leading dimension ldc),
 * FP64 tensor cores (DMMA); Z*a', VV*W, Q = V*Z of the Arnoldi callers */
int nepb_block_gemm(const nepb_block* A, int acol0, int ka, const double* C, int64_t ldc, int q, nepb_block* Y, int ycol0,
                    int64_t rows);
/* dst[:, d0 : d0+nc) = alpha * src[:, s0 : s0+nc) (alpha complex, NULL = 1) */
int nepb_block_copy_cols(const nepb_block* src, int s0, int nc, nepb_block* dst, int d0, const double* alpha, int64_t rows);

/* iar's block shift and re-packing (src/method_iar.jl:100-101,105) on a basis block of n*(m+1) rows:
 * expand: Y[i, ycol0+b] = V[b*n+i, vcol] / (b+1 if scale_by_index)   pack: V[b*n+i, vcol] = Y[i, ycol0+b],  b < nb */
int nepb_iar_expand(const nepb_block* V, int vcol, int64_t n, int nb, nepb_block* Y, int ycol0, int scale_by_index);
int nepb_iar_pack(const nepb_block* Y, int ycol0, int nb, int64_t n, nepb_block* V, int vcol);
/* out[c] = ||A[0:rows, c0+c]||_2 (residual norms of all Ritz pairs, src/method_iar.jl:134-135) */
int nepb_block_colnorms(const nepb_block* A, int c0, int nc, int64_t rows, double* out);

/* ---- a9,a10: contour quadrature, sharded over ranks (src/method_contour_common.jl:61-94) ---------------------
 * S[:,:,j] = sum_i w[i,j] * M(lambda_i)^-1 Vh over the quadrature nodes this rank owns.  coef is nnodes x p complex
 * (row i = f_1(lambda_i)..f_p(lambda_i)); weights is nnodes x mg complex (row i = h*gp(t_i)*g_j(t_i)[/(2 pi i)], i.e.
 * `temp*G[i,j]` of :86-90 including the step h).  Nodes are factorised and solved `batch` at a time. */
typedef struct nepb_contour nepb_contour;
int nepb_contour_create(const nepb_spmf* h, int k, int mg, int batch, nepb_contour** out);
int nepb_contour_destroy(nepb_contour* c);
/* reduce != 0: sum the moments over all ranks of the communicator (one ncclAllReduce, in place in HBM); reduce == 2: only
 * rank 0 copies the moments back to its host array S (the extraction of method_beyncontour.jl:114-184 runs once), the other
 * ranks leave S untouched.
 * node_flags[nnodes] (optional): bit0 zero pivot, bit1 non-finite pivot, bit2 perturbed pivots, bit3 (8) the static-pivoting
 * row matching is in use (the integration is repeated once on a matched analysis when a node on the plain pattern shows
 * any of the other bits), bit4 (16) a pivot below 1e-8 max|M_ij| (the node is close to an eigenvalue). */
int nepb_contour_integrate(nepb_contour* c, int nnodes, const double* coef, const double* weights, const double* Vh,
                           int64_t ldv, int reduce, double* S /* n x k x mg, column-major */, int* node_flags);
/* the same in three steps, so that a benchmark can time the device part alone */
int nepb_contour_set_probe(nepb_contour* c, const double* Vh, int64_t ldv);
int nepb_contour_integrate_dev(nepb_contour* c, int nnodes, const double* coef, const double* weights, int reduce);
int nepb_contour_get_moments(nepb_contour* c, double* S);

/* ---- multi-GPU plumbing: one process per GPU, NCCL loaded at run time ------------------------------------------
 * Rank 0 calls nepb_comm_unique_id and ships the 128 bytes to the other ranks with whatever the host has
 * (torch.distributed, MPI, Julia Distributed); then every rank calls nepb_comm_init. */
int nepb_comm_unique_id(char id[128]);
int nepb_comm_init(int nranks, int rank, const char id[128]);
int nepb_comm_destroy(void);
int nepb_comm_info(int* nranks, int* rank, int* nccl_version);
int nepb_comm_allreduce_sum_dev(void* dev_ptr, int64_t count /* doubles */);

/* ---- (f)3: WEP-native path -- the waveguide eigenvalue problem in its own format ------------------------------
 * Replaces, for `WEP_FD` (src/gallery_extra/waveguide/Waveguide.jl:203-240), the matrix-free methods
 *   compute_Mlincomb(nep::WEP_FD, lambda, V, a)      Waveguide.jl:324-379
 *   *(M::SchurMatVec, v)                             Waveguide.jl:393-402
 *   Pinv(nep, lambda, x)                             Waveguide.jl:160-163
 * n = nx*nz + 2*nz; the interior unknowns are vec(X), X nz x nx (index z + nz*x), followed by the 2 nz boundary unknowns.
 * K_scaled = K - mean(K) (nz x nx complex, column-major), bb = the reference's bb (nz complex).  Scalar functions stay
 * on the host: the caller passes the Gegenbauer derivative table of sqrt_derivative (Waveguide.jl:574-616) as `coef`. */
typedef struct nepb_wep nepb_wep;
int nepb_wep_create(int nx, int nz, double hx, double hz, const double* K_scaled, const double* k_bar /* complex */,
                    const double* bb, nepb_wep** out);
int nepb_wep_destroy(nepb_wep* h);
int nepb_wep_info(const nepb_wep* h, int* nx, int* nz, int64_t* n);
/* Derivative table of the boundary functions at one lambda, kept in HBM: D (2 nz x ncols complex, ROW-major),
 * D[m, j] = 1im * d^j/dlambda^j sqrt(beta_m(lambda)) (+ d0 for j = 0), the `D` of Waveguide.jl:351-361. */
int nepb_wep_set_table(nepb_wep* h, int ncols, const double* D);
/* Z[:, zcol] = sum_j a_j M^{(j)}(lambda) V[:, vcol0 + j], j = 0..na-1.  coef: 2 nz x na complex, ROW-major,
 * coef[m, j] = a_j * D[m, j]; or NULL: the table of nepb_wep_set_table (for this lambda, >= na columns) is used and a_j is
 * applied on the device (a solver loop at a fixed shift then uploads 16 na bytes per call). */
int nepb_wep_mlincomb_block(const nepb_wep* h, const double* lambda, const nepb_block* V, int vcol0, int na, const double* a,
                            const double* coef, nepb_block* Z, int zcol);
/* y = [R(coef[0:nz] .* Rinv(x[0:nz])); R(coef[nz:2nz] .* Rinv(x[nz:2nz]))]; coef = 1 ./ [sM; sP] gives Pinv */
int nepb_wep_pinv(const nepb_wep* h, const double* coef, const double* x, double* y);
/* Y[:, ycol] = (A(lambda) X + X B + K .* X) - C1 Pinv(lambda, C2T x) for blocks with nx*nz rows; sinv = 1 ./ [sM; sP] */
int nepb_wep_schur_matvec_block(const nepb_wep* h, const double* lambda, const double* sinv, const nepb_block* X, int xcol,
                                nepb_block* Y, int ycol);
/* algorithmic HBM bytes of one nepb_wep_mlincomb_block call */
int64_t nepb_wep_mlincomb_bytes(const nepb_wep* h, int na);

/* ---- deterministic synthetic data (bench / tests): Middle-Square-Weyl stream ------------------- */
/* state = {x_lo,x_hi,w_lo,w_hi,s_lo,s_hi}; fills out[count] with uniform doubles in [0,1) exactly as
 * gen_rng_float of src/gallery_extra/basic_random_examples.jl:86-95 and advances the state. */
int nepb_msws_init(uint64_t seed_lo, uint64_t seed_hi, uint64_t state[6]);
int nepb_msws_fill(uint64_t state[6], int64_t count, double* out);

#ifdef __cplusplus
}
#endif
#endif /* NEPB200_H */
