/*
 * ngs_oracle.h -- CPU restatement of NGSolve's assembled-system solve hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ngsolve_b200/, include/)
 * may include, link or call this.  Allowed users: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * NGSolve source tree).  Parity status: PINNED -- checked against fixtures under
 * tests/golden/ that were produced by the reference itself (built from the
 * mounted tree in this container; generator script tests/golden/make_golden.py).
 * The distributed (ParallelDofs / Cumulate) part has no runnable reference here
 * (MPI is absent), it is a restatement only: "parity unpinned" for that part.
 *
 * Conventions: complex numbers are interleaved (re, im) doubles == C99
 * `double _Complex`; 3x3 block entries are 9 doubles row-major (ngbla Mat<3,3>
 * stores data[i*W+j], basiclinalg/matrix.hpp); row pointers are uint64 and column
 * indices int32 exactly as SparseMatrix::CSR() hands them out
 * (linalg/python_linalg.cpp:121-138).
 */
#ifndef NGS_ORACLE_H
#define NGS_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double _Complex orc_cplx;

/* number of worker threads the OpenMP loops use (0 = leave default) */
void orc_set_num_threads(int nthreads);
int orc_get_max_threads(void);

/* ---- SparseMatrix<TM>::MultAdd / BaseMatrix::Mult ------------------------------- */
void orc_csr_multadd_d(size_t n, const uint64_t *firsti, const int32_t *colnr, const double *data,
                       double s, const double *x, double *y);
void orc_csr_multadd_z(size_t n, const uint64_t *firsti, const int32_t *colnr, const orc_cplx *data,
                       double s, const orc_cplx *x, orc_cplx *y);
/* complex scale passed as (re, im) */
void orc_csr_multadd_zs(size_t n, const uint64_t *firsti, const int32_t *colnr, const orc_cplx *data,
                        double s_re, double s_im, const orc_cplx *x, orc_cplx *y);
void orc_csr_multadd_b3(size_t n, const uint64_t *firsti, const int32_t *colnr, const double *data,
                        double s, const double *x, double *y);
/* Mult = SetZero + MultAdd(1.0)   kind: 0 real, 1 complex, 3 block3 */
void orc_csr_mult(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr, const void *data,
                  const void *x, void *y);

/* ---- BaseVector updates ---------------------------------------------------------- */
void orc_vec_set_scalar_d(size_t N, double *x, double s);
void orc_vec_scale_d(size_t N, double *x, double s);
void orc_vec_set_d(size_t N, double *y, double s, const double *x);
void orc_vec_add_d(size_t N, double *y, double s, const double *x);
void orc_vec_scale_z(size_t N, orc_cplx *x, orc_cplx s);
void orc_vec_set_z(size_t N, orc_cplx *y, orc_cplx s, const orc_cplx *x);
void orc_vec_add_z(size_t N, orc_cplx *y, orc_cplx s, const orc_cplx *x);

/* ---- reductions -------------------------------------------------------------------- */
double orc_vec_inner_d(size_t N, const double *x, const double *y);
/* writes (re, im) to out[2]; conj applies to the ARGUMENT y */
void orc_vec_inner_z(size_t N, const orc_cplx *x, const orc_cplx *y, int conjugate, double *out);
/* is_complex selects FV<Complex> chunks (N complex entries) vs FV<double> (N doubles) */
double orc_vec_l2norm(size_t N, const void *x, int is_complex);

/* ---- JacobiPrecond ------------------------------------------------------------------ */
/* kind 0/1/3.  freebits may be NULL (no mask).  invdiag gets n entries of the kind.
 * returns 0, or -1 if a 3x3 block is singular, -2 if a diagonal entry is missing
 * where required. */
int orc_jacobi_setup(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr,
                     const void *data, const uint8_t *freebits, void *invdiag);
void orc_jacobi_multadd(int kind, size_t n, const void *invdiag, const uint8_t *freebits,
                        double s, const void *x, void *y);

/* ---- Krylov solvers ------------------------------------------------------------------ */
/* ip_mode: 0 = double, 1 = Complex (bilinear), 2 = ComplexConjugate.
 * kind must be 0 for ip_mode 0 (or 3 for block vectors), 1 otherwise.
 * invdiag == NULL  =>  no preconditioner (w = d).
 * history (may be NULL) receives Abs(wdn) after the initial residual and after every
 * iteration (at most maxsteps+1 values); *nhist receives how many were written.
 * Returns GetSteps() (= iterations + 1 when the loop ends by its condition). */
int orc_cg_solve(int kind, int ip_mode, size_t n, const uint64_t *firsti, const int32_t *colnr,
                 const void *data, const void *invdiag, const uint8_t *freebits,
                 const void *f, void *u, double prec, int maxsteps, int initialize,
                 double *history, int *nhist);

/* GMRESSolver<double|Complex>::Mult ; returns steps (= j after the loop). */
int orc_gmres_solve(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr,
                    const void *data, const void *invdiag, const uint8_t *freebits,
                    const void *f, void *x, double prec, int maxsteps, int initialize,
                    double *history, int *nhist);

/* ---- SparseMatrix::Reorder ---------------------------------------------------------- */
/* new(i, inv[col]) = old(reorder[i], col); new rows sorted ascending.  kind 0/1/3. */
void orc_csr_reorder(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr, const void *data,
                     const uint64_t *reorder, uint64_t *nfirsti, int32_t *ncolnr, void *ndata);
/* Cuthill-McKee ordering of the pattern (own specification, see ngs_oracle.c); perm is Reorder's argument */
void orc_rcm(size_t n, const uint64_t *firsti, const int32_t *colnr, int max_components, uint64_t *perm);

/* ---- ParallelDofs tables (restated; no MPI here) -------------------------------------- */
/* dist_procs as CSR table (dp_first[ndof+1], dp_data).  Output: exchangedofs as CSR table
 * over ranks (ex_first[ntasks+1], ex_data sized dp_first[ndof]) and ismaster bytes (0/1). */
void orc_pardofs_build(int ntasks, int id, size_t ndof, const uint64_t *dp_first, const int32_t *dp_data,
                       uint64_t *ex_first, int32_t *ex_data, uint8_t *ismaster);

/* ---- BlockJacobiPrecond<double> (linalg/blockjacobi.cpp:380-500, 594-681) ---------------- */
int orc_blockjacobi_setup(size_t n, const uint64_t *firsti, const int32_t *colnr, const double *data, size_t nblocks,
                          const uint64_t *bfirst, const int32_t *bdofs, double *inverses);
void orc_blockjacobi_multadd(size_t nblocks, const uint64_t *bfirst, const int32_t *bdofs, const double *inverses, double s,
                             const double *x, double *y, int transpose);

/* ---- MultTransAdd (linalg/sparsematrix_impl.hpp:344-352) and symmetric storage (:967-983) ---- */
void orc_csr_multtransadd(int kind, size_t h, const uint64_t *firsti, const int32_t *colnr, const void *data, double s,
                          const void *x, void *y);
void orc_csrsym_multadd_d(size_t n, const uint64_t *firsti, const int32_t *colnr, const double *data, double s,
                          const double *x, double *y);

#ifdef __cplusplus
}
#endif
#endif
