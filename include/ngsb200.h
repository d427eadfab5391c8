/*
 * ngsb200.h -- C ABI of libngsb200: NGSolve's assembled-system solve hot path on B200.
 *
 * This is the drop-in boundary.  Every entry point replaces one piece of the
 * reference's device layer (ngscuda/) or of the ngla interface that layer plugs
 * into; the reference location is cited next to each declaration (paths relative
 * to the NGSolve source tree).  The NGSolve-side adapter that binds these calls
 * (BaseVector/BaseMatrix subclasses registered through
 * BaseMatrix::RegisterDeviceMatrixCreator / BaseVector::RegisterDeviceVectorCreator,
 * linalg/basematrix.hpp:209-215, linalg/basevector.hpp:321-327) is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *  - plain C: opaque handles, pointers and sizes only; no C++/torch types.
 *  - every call returns an int status (NGSB_OK = 0); ngsb_last_error() gives the
 *    message of the last failing call on the calling thread.  The reference
 *    throws ngstd::Exception instead (e.g. size mismatch, linalg/basevector.cpp:151);
 *    the adapter turns a non-zero status into that exception.
 *  - scalars are passed as double[2] = (re, im); im is ignored for real objects.
 *  - "kind": NGSB_REAL (double), NGSB_COMPLEX (std::complex<double>, interleaved),
 *    NGSB_BLOCK3 (Mat<3,3,double> entries / Vec<3,double> vector entries, row-major).
 *  - host pointers are borrowed for the duration of the call only; handles own
 *    their device memory (same as DevSparseMatrix, ngscuda/cuda_linalg.cpp:187-231).
 *  - all work of a context is enqueued on that context's stream; calls that
 *    return a value to the host synchronise that stream, all others are async.
 *  - there is no CPU fallback: without a CUDA device ngsb_ctx_create fails.
 */
#ifndef NGSB200_H
#define NGSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGSB_OK 0
#define NGSB_ERR_INVALID 1   /* bad argument / size mismatch */
#define NGSB_ERR_CUDA 2      /* CUDA runtime error          */
#define NGSB_ERR_NOMEM 3
#define NGSB_ERR_UNSUPPORTED 4
#define NGSB_ERR_COMM 5      /* NCCL error */

#define NGSB_REAL 0
#define NGSB_COMPLEX 1
#define NGSB_BLOCK3 3

/* inner-product flavours of CGSolver<IPTYPE>, linalg/basevector.hpp:1099-1124 */
#define NGSB_IP_REAL 0
#define NGSB_IP_COMPLEX 1          /* bilinear  sum x_i y_i        */
#define NGSB_IP_COMPLEX_CONJ 2     /* sum x_i conj(y_i)            */

typedef struct ngsb_ctx ngsb_ctx;
typedef struct ngsb_vec ngsb_vec;
typedef struct ngsb_scalar ngsb_scalar;
typedef struct ngsb_csr ngsb_csr;
typedef struct ngsb_jacobi ngsb_jacobi;
typedef struct ngsb_blockjacobi ngsb_blockjacobi;
typedef struct ngsb_comm ngsb_comm;
typedef struct ngsb_parmat ngsb_parmat;

const char *ngsb_last_error(void);
const char *ngsb_version(void);

/* ---- context: replaces InitCUDA + InitCuLinalg + the global ngs_cuda_stream ------------
 * ngscuda/cuda_ngstd.cpp:84-108 (device choice by NGS_CUDA_DEVICE_INDEX, stream),
 * ngscuda/cuda_linalg.cpp:70-171 (handles + creator registration).  device < 0 reads
 * NGS_CUDA_DEVICE_INDEX like the reference (default 0). */
int ngsb_ctx_create(int device, ngsb_ctx **out);
int ngsb_ctx_destroy(ngsb_ctx *ctx);
int ngsb_ctx_sync(ngsb_ctx *ctx);                       /* SyncNGSStream, ngscuda/cuda_core.hpp */
int ngsb_ctx_device(const ngsb_ctx *ctx, int *device, int *sm_count);
void *ngsb_ctx_stream(ngsb_ctx *ctx);                   /* cudaStream_t */
/* kernels of this library launched on ctx so far (bench.py's gpu_launches) */
int ngsb_ctx_launch_count(const ngsb_ctx *ctx, uint64_t *count);
/* options (name, long value); unknown names fail with NGSB_ERR_INVALID.
 *   read at every launch / solve:
 *     "spmv_algo"        0 auto (= 3), 1 sub-warp CSR, 2 TMA-streamed CSR, 3 SELL-32
 *     "spmv_ctas_per_sm" grid of the SpMV kernels in CTAs per SM, 0 = default (96 for SELL)
 *     "cg_batch"         CG iterations per CUDA graph / between two polls of the device stop flag (default 16)
 *     "cg_persistent"    real Jacobi-PCG (CGSolver<double> + JacobiPrecond) as ONE persistent cooperative kernel, grid barriers instead
 *                        of three kernel launches per iteration: -1 automatic (default: matrices below 4 M rows), 0 off, 1 on
 *     "gmres_orth"       GMRES orthogonalisation of w = C A v_j against v_0..v_j.  1 (default): V^T w and the new row of L = strict lower
 *                        part of V^T V in ONE batched reduction (each v_k read once), H(0..j,j) from the unit triangular system
 *                        (I + L) h = V^T w, then w -= V h in one pass -- the coefficients modified Gram-Schmidt produces (identical in
 *                        exact arithmetic), 2 (j+1) instead of 4 (j+1) vector passes and 2 instead of j+2 reductions per step.
 *                        0: the reference's serial loop (linalg/cg.cpp:927-932).  On parallel vectors (peer-memory data path) the
 *                        2 (j+1) sums of a step travel in one exchange; the NCCL data path and Krylov spaces beyond 1023 vectors use 0.
 *     "cg_stream_hints"  0/1: the CG update kernel loads u, d, As, diagonal streaming and stores u, d evict-first (default 0; A/B)
 *     "cg_chunked"       0/1: the CG update kernel works on one contiguous chunk per CTA instead of grid-stride (default 0; A/B)
 *     "cg_fold_u"        0/1: the CG direction kernel also does u += al s (10 instead of 11 vector passes per iteration)
 *     "sell_variant"     inner-loop variant of the real SELL kernel (0 default; 1-6 kept for A/B, all bit-identical)
 *     "sell_pf_steps", "sell_pf_next"   L2 prefetch distances of the compressed loop (0 = off, default)
 *     "sell_c16"         0/1: use the 16-bit column offsets of real matrices (default 1; also read at creation)
 *     "timing"           0/1: record an event pair around every launch (ngsb_ctx_kernel_time)
 *   read when a matrix is created:
 *     "csr_keep"         the uploaded column / value arrays after the SELL copy exists: 1 keep, 0 release (rebuilt from the SELL copy
 *                        when CSR(), CreateTranspose, Reorder or the block-Jacobi constructor ask for them; the Jacobi
 *                        constructor reads the diagonal out of the SELL copy), -1 automatic (default) = release above 4 GiB
 *     "reorder"          internal dof reordering of square matrices (csrc/reorder.cu): -1 automatic (default), 0 off, 1 always.
 *                        The matrix keeps the caller's numbering at the interface; products and the fused solvers run on
 *                        P A P^T with P = Cuthill-McKee of the pattern (ngsb_csr_rcm).  Automatic = at least
 *                        "reorder_min_rows" rows (default 32768) and fewer than half of the natural 32-row slices fit for
 *                        16-bit column offsets -- the signature of netgen's entity-by-entity numbering
 *     "reorder_slot_order" 0/1 (default 0): the internal permutation = Cuthill-McKee composed with the SELL length sort (rows of the
 *                        Cuthill-McKee order sorted longest-first, stably, inside windows of sell_sigma rows), so that P A P^T is numbered
 *                        in slot order: no slot -> row table, y written in whole 256-byte runs.  Measured: no gain, kept for A/B
 *     "sell_c16_all"     0/1: 16-bit column offsets for Complex and Mat<3,3> matrices too (default 0; must still be on at launch)
 *     "sell_cap"         longest row part kept in a slice, 0 = max(64, 4 x mean row length) or the longest row when cheap
 *     "sell_sigma"       rows sorted by length inside windows of this many rows, -1 = automatic (65536 where natural slices would pad > 5 %), 0/1 = off
 *     "sell_schedule"    slices ordered by their smallest first column: 0 off, 1 automatic, 2 on
 *     "spmv_tile", "spmv_ncw", "spmv_stages", "spmv_subwarp"   CSR kernels, 0 = default
 *   read by ngsb_parmat_create:
 *     "dist_fused_push"  0/1: product + halo push in one kernel (default 1, see ngsb_parmat_create)
 *     "dist_overlap"     0/1: interface slices first, pushed while the interior slices are multiplied (default 0) */
int ngsb_ctx_set_option(ngsb_ctx *ctx, const char *name, long value);
/* device time in ms of the kernels recorded since the last reset for a class
 * ("spmv", "cgupdate", "all"); only meaningful when option "timing" is 1. */
int ngsb_ctx_kernel_time(ngsb_ctx *ctx, const char *klass, double *ms, uint64_t *launches);
int ngsb_ctx_kernel_time_reset(ngsb_ctx *ctx);

/* ---- vectors: replaces UnifiedVector (ngscuda/unifiedvector.hpp:8-98, .cpp:7-345) ---------
 * n_entries entries of `kind`; NGSB_BLOCK3 entries are Vec<3,double> (the reference's
 * UnifiedVector cannot hold these, ngscuda/unifiedvector.cpp:22-26). */
int ngsb_vec_create(ngsb_ctx *ctx, size_t n_entries, int kind, ngsb_vec **out);
int ngsb_vec_destroy(ngsb_vec *v);
int ngsb_vec_info(const ngsb_vec *v, size_t *n_entries, int *kind, size_t *n_scalars);
void *ngsb_vec_devptr(ngsb_vec *v);                      /* DevData(), unifiedvector.hpp:73 */
/* aliasing view of entries [begin,end): UnifiedVector::Range, unifiedvector.cpp:130-133.
 * The view keeps the parent's storage alive. */
int ngsb_vec_range(ngsb_vec *v, size_t begin, size_t end, ngsb_vec **view);
/* host<->device copies of whole entries: UpdateDevice/UpdateHost, unifiedvector.cpp:288-330.
 * d2h synchronises; h2d is stream-ordered and returns after the copy was issued from a
 * staging copy (the host buffer may be reused immediately). */
int ngsb_vec_h2d(ngsb_vec *v, const void *host, size_t first_entry, size_t n_entries);
int ngsb_vec_d2h(const ngsb_vec *v, void *host, size_t first_entry, size_t n_entries);
/* BaseVector::SetScalar / Scale / Set / Add: linalg/basevector.cpp:113-138, 75-111,
 * 146-199, 203-257; device versions ngscuda/unifiedvector.cpp:136-213. */
int ngsb_vec_set_scalar(ngsb_vec *x, const double s[2]);
int ngsb_vec_scale(ngsb_vec *x, const double s[2]);
int ngsb_vec_set(ngsb_vec *y, const double s[2], const ngsb_vec *x);      /* y  = s*x */
int ngsb_vec_axpy(ngsb_vec *y, const double s[2], const ngsb_vec *x);     /* y += s*x */
/* S_BaseVector<SCAL>::InnerProduct: linalg/basevector.cpp:1108-1159 (conjugate acts on
 * the argument y); device version ngscuda/unifiedvector.cpp:215-237 (cublasDdot). */
int ngsb_vec_dot(const ngsb_vec *x, const ngsb_vec *y, int conjugate, double out[2]);
/* BaseVector::L2Norm: linalg/basevector.cpp:41-73; ngscuda/unifiedvector.cpp:239-247. */
int ngsb_vec_nrm2(const ngsb_vec *x, double *out);

/* ---- device-resident scalars: replaces UnifiedScalar (ngscuda/unifiedvector.hpp:107-136)
 * and the BaseScalar hooks linalg/basevector.cpp:259-298, linalg/basescalar.hpp:17-29. */
int ngsb_scalar_create(ngsb_ctx *ctx, ngsb_scalar **out);
int ngsb_scalar_destroy(ngsb_scalar *s);
int ngsb_scalar_set(ngsb_scalar *s, const double v[2]);
int ngsb_scalar_get(const ngsb_scalar *s, double v[2]);                   /* synchronises */
/* out = a / b, out = -a, out = a   (the Div/Neg/Scal expression kernels,
 * ngscuda/unifiedvector.hpp:126-135, ngscuda/cuda_krylov.cpp:88-95) */
int ngsb_scalar_div(ngsb_scalar *out, const ngsb_scalar *a, const ngsb_scalar *b);
int ngsb_scalar_neg(ngsb_scalar *out, const ngsb_scalar *a);
int ngsb_scalar_copy(ngsb_scalar *out, const ngsb_scalar *a);
int ngsb_vec_dot_dev(const ngsb_vec *x, const ngsb_vec *y, int conjugate, ngsb_scalar *out);
int ngsb_vec_axpy_dev(ngsb_vec *y, const ngsb_scalar *s, const ngsb_vec *x);
int ngsb_vec_scale_dev(ngsb_vec *x, const ngsb_scalar *s);

/* ---- CSR matrices: replaces DevSparseMatrix (ngscuda/cuda_linalg.hpp:50-72,
 * ngscuda/cuda_linalg.cpp:187-316) for SparseMatrix<double>, SparseMatrix<Complex> and
 * SparseMatrix<Mat<3,3,double>> (linalg/sparsematrix.hpp:65-167, 340-542).
 * rowptr/col/val are exactly what SparseMatrix::CSR() returns
 * (linalg/python_linalg.cpp:121-138): uint64 row pointers (h+1), int32 columns sorted
 * ascending per row, values of `kind` (9 doubles row-major per entry for NGSB_BLOCK3). */
int ngsb_csr_create(ngsb_ctx *ctx, size_t height, size_t width, size_t nnz,
                    const uint64_t *rowptr, const int32_t *col, const void *val, int kind,
                    ngsb_csr **out);
/* same, adopting DEVICE arrays (copied device-to-device); for systems assembled on
 * the device.  rowptr is uint64 on the device as well. */
int ngsb_csr_create_from_device(ngsb_ctx *ctx, size_t height, size_t width, size_t nnz,
                                const uint64_t *d_rowptr, const int32_t *d_col, const void *d_val,
                                int kind, ngsb_csr **out);
int ngsb_csr_destroy(ngsb_csr *A);
int ngsb_csr_info(const ngsb_csr *A, size_t *height, size_t *width, size_t *nnz, int *kind);
/* SparseMatrix::MultAdd(double|Complex s, x, y): linalg/sparsematrix_impl.hpp:264-279,
 * 378-396;  DevSparseMatrix::MultAdd ngscuda/cuda_linalg.cpp:244-276.   y += s*A*x */
int ngsb_csr_multadd(const ngsb_csr *A, const double s[2], const ngsb_vec *x, ngsb_vec *y);
/* BaseMatrix::Mult (= SetZero + MultAdd(1)), linalg/basematrix.cpp:120-127.   y = A*x */
int ngsb_csr_mult(const ngsb_csr *A, const ngsb_vec *x, ngsb_vec *y);
/* SparseMatrix::Reorder(perm): linalg/sparsematrix_impl.hpp:762-783.
 * new(i, inv[c]) = old(perm[i], c); perm is a host array of `height` indices.  Done on the device. */
int ngsb_csr_reorder(const ngsb_csr *A, const uint64_t *perm, ngsb_csr **out);
/* The bandwidth-reducing permutation option "reorder" uses, in the form Reorder() takes (new row k = old row perm[k]).
 * The reference computes none (its dof numbering is netgen's, comp/h1hofespace.cpp:833-880); this is level-synchronous
 * Cuthill-McKee on the device; tests/ hold a serial restatement (plain C) that it equals bit for bit. */
int ngsb_csr_rcm(const ngsb_csr *A, uint64_t *perm);
/* whether products of A run on an internally reordered copy, its permutation (perm may be NULL), and the share of natural
 * 32-row slices fit for 16-bit column offsets that the automatic mode looked at (-1: not evaluated) */
int ngsb_csr_reorder_info(const ngsb_csr *A, int *reordered, uint64_t *perm, double *natural_c16_share);
/* copy the device CSR back (pattern parity checks): any pointer may be NULL */
int ngsb_csr_download(const ngsb_csr *A, uint64_t *rowptr, int32_t *col, void *val);
/* device layout diagnostics: padded entries of the SELL-32 copy the default SpMV kernel streams,
 * rows with an overflow part, and the per-row cap of the slices.  Any pointer may be NULL. */
int ngsb_csr_layout(const ngsb_csr *A, uint64_t *sell_entries, uint32_t *overflow_rows, uint32_t *cap);
/* bytes the default kernel streams per Mult with the layout as stored (padding included; real matrices
 * store 16-bit column offsets per slice where the columns of one entry step lie within 65535 of each
 * other: 10.125 instead of 12 bytes per entry); *c16_entries = padded entries in such slices */
int ngsb_csr_stream_bytes(const ngsb_csr *A, double *bytes, uint64_t *c16_entries);
/* Checkpoint wire format of a device system: the bytes SparseMatrix<TM>::DoArchive writes into ngcore's BinaryOutArchive
 * (linalg/sparsematrix_impl.hpp:443-452; raw little-endian: size_t size, width, nze; Array<size_t> firsti; Array<int> colnr;
 * Array<TM> data, each array as size_t count + elements).  What _write produces the reference's BinaryInArchive reads, and
 * _create_from_archive reads what the reference wrote. */
int ngsb_csr_archive_size(const ngsb_csr *A, size_t *bytes);
int ngsb_csr_archive_write(const ngsb_csr *A, void *buf, size_t capacity);
int ngsb_csr_create_from_archive(ngsb_ctx *ctx, const void *buf, size_t bytes, int kind, ngsb_csr **out);
/* device memory the matrix holds right now: row pointers + (if resident) the uploaded CSR arrays [+ permutation tables];
 * the SELL copy the products stream; whether the CSR arrays are resident (option csr_keep) */
int ngsb_csr_memory(const ngsb_csr *A, uint64_t *csr_bytes, uint64_t *sell_bytes, int *csr_resident);
/* algorithmic bytes of one Mult: nnz*(b*b*S+4) + 4*h + 2*N*S */
int ngsb_csr_mult_bytes(const ngsb_csr *A, double *bytes);

/* ---- Jacobi: replaces DevDiagonalMatrix built from JacobiPrecond<TM>::invdiag
 * (ngscuda/cuda_linalg.cpp:103-115, 321-366; linalg/jacobi.cpp:39-155).
 * invdiag: n entries of `kind` (9 doubles per entry for NGSB_BLOCK3); freebits: the
 * `inner` BitArray bytes (bit i -> byte i/8, bit i%8), or NULL. */
int ngsb_jacobi_create(ngsb_ctx *ctx, size_t n, const void *invdiag, int kind,
                       const uint8_t *freebits, ngsb_jacobi **out);
/* JacobiPrecond ctor on the device matrix (mat.CreateSmoother(freedofs),
 * linalg/python_linalg.cpp:1220-1236; linalg/jacobi.cpp:39-68). */
int ngsb_jacobi_create_from_csr(const ngsb_csr *A, const uint8_t *freebits, ngsb_jacobi **out);
int ngsb_jacobi_destroy(ngsb_jacobi *J);
int ngsb_jacobi_download(const ngsb_jacobi *J, void *invdiag);
int ngsb_jacobi_multadd(const ngsb_jacobi *J, const double s[2], const ngsb_vec *x, ngsb_vec *y);
int ngsb_jacobi_mult(const ngsb_jacobi *J, const ngsb_vec *x, ngsb_vec *y);

/* ---- transposes, symmetric storage, several right-hand sides ----
 * ngsb_csr_transpose: SparseMatrixTM::CreateTranspose(sorted=true) (linalg/sparsematrix.cpp; Python
 * `mat.CreateTranspose()`, linalg/python_linalg.cpp:172): a new device matrix, rows ascending.
 * ngsb_csr_multtransadd: SparseMatrix::MultTransAdd, y += s * A^T x (linalg/sparsematrix_impl.hpp:
 * 344-352; the reference scatters serially).  A^T is built on first use and cached in the handle;
 * x has Height() entries, y Width().
 * ngsb_csr_create_symmetric: a SparseMatrixSymmetric<TM> handed over as its stored lower triangle
 * (columns <= row, ascending; linalg/sparsematrix.hpp:760-835).  The full matrix is formed on the
 * device; Mult/MultAdd then equal SparseMatrixSymmetric::MultAdd (sparsematrix_impl.hpp:967-983).
 * ngsb_csr_multadd_multi: SparseMatrix<double>::MultAdd(FlatVector alpha, MultiVector x, MultiVector y)
 * (linalg/sparsematrix.cpp:2274-2351): y[k] += alpha[k] * A * x[k]; real matrices sweep four vectors
 * per pass over the matrix. */
int ngsb_csr_transpose(const ngsb_csr *A, ngsb_csr **out);
int ngsb_csr_multtransadd(const ngsb_csr *A, const double s[2], const ngsb_vec *x, ngsb_vec *y);
int ngsb_csr_create_symmetric(ngsb_ctx *ctx, size_t n, size_t nnz, const uint64_t *rowptr, const int32_t *col,
                              const void *val, int kind, ngsb_csr **out);
int ngsb_csr_multadd_multi(const ngsb_csr *A, size_t nvec, const double *alpha, const ngsb_vec *const *x,
                           ngsb_vec *const *y);

/* ---- block-Jacobi: replaces DevBlockJacobiMatrix (ngscuda/dev_blockjacobi.cpp:21-140) and the
 * BlockJacobiPrecond<double> constructor (linalg/blockjacobi.cpp:380-500; reached from Python by
 * mat.CreateBlockSmoother(blocks), linalg/python_linalg.cpp).  Block b owns the dofs
 * dofs[first[b] .. first[b+1]) (the reference's Table<int>, blocks may overlap); its matrix
 * A(block, block) is inverted on the device with the reference's T_CalcInverse
 * (basiclinalg/calcinverse.cpp:26-107).  TM = double only, like the reference device class.
 * MultAdd: y(block) += s * inv_b * x(block) summed over the blocks in ascending block order
 * (linalg/blockjacobi.cpp:594-634); transpose != 0: MultTransAdd (:637-681). */
int ngsb_blockjacobi_create(const ngsb_csr *A, size_t nblocks, const uint64_t *first, const int32_t *dofs,
                            ngsb_blockjacobi **out);
/* the same object from inverses computed elsewhere (the adapter hands over BlockJacobiPrecond::GetInverses(), which the
 * reference's constructor already built on the host -- what DevBlockJacobiMatrix's constructor copies,
 * ngscuda/dev_blockjacobi.cpp:95-113): blocks back to back, each row-major; n = vector length */
int ngsb_blockjacobi_create_from_inverses(ngsb_ctx *ctx, size_t n, size_t nblocks, const uint64_t *first, const int32_t *dofs,
                                          const double *inverses, ngsb_blockjacobi **out);
int ngsb_blockjacobi_destroy(ngsb_blockjacobi *J);
int ngsb_blockjacobi_info(const ngsb_blockjacobi *J, size_t *n, size_t *nblocks, size_t *maxbs, size_t *total,
                          size_t *matrix_entries);
/* the inverse blocks back to back, each row-major (BlockJacobiPrecond::GetInverses) */
int ngsb_blockjacobi_download(const ngsb_blockjacobi *J, double *inverses);
int ngsb_blockjacobi_multadd(const ngsb_blockjacobi *J, double s, const ngsb_vec *x, ngsb_vec *y, int transpose);
int ngsb_blockjacobi_mult(const ngsb_blockjacobi *J, const ngsb_vec *x, ngsb_vec *y, int transpose);

/* ---- Krylov solvers ----------------------------------------------------------------------
 * ngsb_cg_solve: CGSolver<IPTYPE>::Mult, linalg/cg.cpp:503-633 (same recurrences, same
 * stopping rule `n++ < maxsteps && Abs(wdn) > prec^2*Abs(wdn0)`, *steps = GetSteps()),
 * replacing DevCGSolver::Mult (ngscuda/cuda_krylov.cpp:19-203).  The loop runs on the
 * device: fused kernels, device-resident scalars and a device-side stop flag.
 * C may be NULL (w = d).  history (host, may be NULL) receives Abs(wdn) of the initial
 * residual and of every iteration, at most hist_cap values; *nhist = how many exist. */
int ngsb_cg_solve(const ngsb_csr *A, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *u,
                  double prec, int maxsteps, int ip_mode, int initialize,
                  int *steps, double *history, int hist_cap, int *nhist);
/* the same solve with HOST right-hand side / solution buffers (the call a script makes
 * through `gfu.vec.data = inv * f.vec`): copies f in, solves, copies u out, synchronises. */
int ngsb_cg_solve_host(const ngsb_csr *A, const ngsb_jacobi *C, const void *f_host, void *u_host,
                       double prec, int maxsteps, int ip_mode,
                       int *steps, double *history, int hist_cap, int *nhist);
/* GMRESSolver<IPTYPE>::Mult, linalg/cg.cpp:854-1022 (left-preconditioned, MGS, Givens,
 * no restart).  *steps = GetSteps(). */
int ngsb_gmres_solve(const ngsb_csr *A, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *x,
                     double prec, int maxsteps, int initialize,
                     int *steps, double *history, int hist_cap, int *nhist);

/* ---- distributed: ParallelDofs / ParallelMatrix / Cumulate on one GPU per process -------
 * linalg/paralleldofs.cpp:20-108, parallel/parallelvvector.cpp:247-358,
 * parallel/parallel_matrices.cpp:519-536.
 *
 * Data path.  Default ("peer memory"): every rank exports a mailbox and its interface
 * receive areas with CUDA IPC; the kernels of the solve loop store straight into the
 * neighbours' memory over NVLink and signal with sequence-numbered flags -- Cumulate and
 * the scalar all-reduces are part of the solver's own kernels, no library collective and
 * no host round trip per iteration (csrc/peer.cuh).  If the GPUs cannot map each other
 * the same calls run over NCCL (ncclSend/Recv + ncclAllReduce).  NGSB_COMM=nccl|p2p in the
 * environment forces one of them.
 *
 * Bootstrap (setup only).  Either an ncclUniqueId (128 bytes, from ngsb_comm_unique_id on
 * rank 0, distributed by the caller) -- then NCCL also carries the IPC handles -- or a
 * caller-supplied all-gather (the role MPI_Allgather on the NgMPI_Comm plays in
 * ParallelDofs, linalg/paralleldofs.cpp:29-44): it must copy `bytes_per_rank` bytes from
 * every rank's `send` into `recv` in rank order and return 0.  The callback is used only
 * during the create call it is passed to. */
typedef int (*ngsb_allgather_fn)(void *user, const void *send, void *recv, size_t bytes_per_rank);
int ngsb_comm_unique_id(void *uid128);
int ngsb_comm_create(ngsb_ctx *ctx, int nranks, int rank, const void *uid128, ngsb_comm **out);
/* p2p_mode: -1 auto, 0 NCCL only, 1 peer memory required.  uid128 may be NULL if allgather is given. */
int ngsb_comm_create_ex(ngsb_ctx *ctx, int nranks, int rank, const void *uid128, ngsb_allgather_fn allgather,
                        void *user, int p2p_mode, ngsb_comm **out);
int ngsb_comm_info(const ngsb_comm *comm, int *nranks, int *rank, int *peer_memory, int *has_nccl);
int ngsb_comm_destroy(ngsb_comm *comm);
/* local matrix + exchange tables: exchangedofs as a CSR table over ranks
 * (ex_first[nranks+1], ex_dofs ascending local dofs), exactly ParallelDofs::exchangedofs.
 * Collective.  The _ex form takes the bootstrap all-gather of a communicator without NCCL. */
int ngsb_parmat_create(ngsb_comm *comm, const ngsb_csr *local, const uint64_t *ex_first,
                       const int32_t *ex_dofs, ngsb_parmat **out);
int ngsb_parmat_create_ex(ngsb_comm *comm, const ngsb_csr *local, const uint64_t *ex_first,
                          const int32_t *ex_dofs, ngsb_allgather_fn allgather, void *user, ngsb_parmat **out);
int ngsb_parmat_destroy(ngsb_parmat *P);
int ngsb_parmat_info(const ngsb_parmat *P, int *peer_memory, int *n_neighbours, size_t *n_exchange,
                     size_t *n_interface);
/* option "dist_fused_push" (default 1; read by ngsb_parmat_create, peer-memory data path): the product kernel of the distributed
 * CG stores every interface row straight into the neighbours' receive areas (P2P stores over NVLink) as soon as its slice is
 * done, and its last block publishes the sequence flags and the rank's partial of <s, A s> -- product and halo push are ONE
 * kernel, the exchange overlaps the rest of the product; 0 = separate halo_push_kernel after the product.
 * option "dist_overlap" (set on the context before ngsb_parmat_create, peer-memory data path): the CG product runs the
 * SELL slices holding interface rows first, pushes them, and runs the interior slices while the values travel -- the
 * overlap the reference gets from MPI_Isend/Irecv around its local MultAdd (parallel/parallelvvector.cpp:452-475).
 * Reports whether the split is active and the two slice counts. */
int ngsb_parmat_overlap_info(const ngsb_parmat *P, int *enabled, size_t *interface_slices, size_t *interior_slices);
/* masterdofs bytes as the reference derives them (lowest rank owns), n entries */
int ngsb_parmat_masterdofs(const ngsb_parmat *P, uint8_t *ismaster);
/* JacobiPrecond of a ParallelMatrix: diagonal summed over the sharing ranks, then inverted
 * (AllReduceDofData(invdiag, SUM), linalg/jacobi.cpp:60-61); collective over the communicator */
int ngsb_parmat_jacobi_create(const ngsb_parmat *P, const uint8_t *freebits, ngsb_jacobi **out);
/* ParallelBaseVector::Cumulate on a DISTRIBUTED vector: neighbour exchange + add */
int ngsb_parmat_cumulate(const ngsb_parmat *P, ngsb_vec *v);
/* ParallelMatrix::MultAdd (C2D): x cumulated in, y distributed out */
int ngsb_parmat_mult(const ngsb_parmat *P, const ngsb_vec *x, ngsb_vec *y);
/* global inner product of a CUMULATED and a DISTRIBUTED vector (local dot + AllReduce),
 * or of two CUMULATED vectors (masked by master dofs), parallelvvector.cpp:289-358.
 * out = (re, im); conjugate (complex only) conjugates the argument y. */
int ngsb_parmat_dot(const ngsb_parmat *P, const ngsb_vec *x, const ngsb_vec *y, int both_cumulated,
                    int conjugate, double out[2]);
/* distributed Jacobi-PCG: f DISTRIBUTED in, u CUMULATED out; C holds the cumulated
 * inverse diagonal (linalg/jacobi.cpp:60-61).  ip_mode as in ngsb_cg_solve. */
int ngsb_parmat_cg_solve(const ngsb_parmat *P, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *u,
                         double prec, int maxsteps, int ip_mode, int *steps, double *history,
                         int hist_cap, int *nhist);
/* distributed GMRES (GMRESSolver<IPTYPE>::Mult on parallel vectors): f DISTRIBUTED in,
 * x CUMULATED out; every Krylov vector is CUMULATED, inner products are master-masked. */
int ngsb_parmat_gmres_solve(const ngsb_parmat *P, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *x,
                            double prec, int maxsteps, int *steps, double *history, int hist_cap,
                            int *nhist);

#ifdef __cplusplus
}
#endif
#endif /* NGSB200_H */
