/*
 * ngs_oracle.c -- CPU restatement of NGSolve's assembled-system solve hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ngs_oracle.h).  Plain C99 + OpenMP.
 * The arithmetic (operation order inside a row, the 16-chunk reductions, the
 * CG/GMRES recurrences and stopping rules) follows the reference line by line;
 * OpenMP only distributes independent rows / entries / chunks, which never
 * changes a result (the reference's TaskManager does the same).
 */
#include "ngs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_set_num_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

int orc_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* rows [lo,hi) of part t out of T, balanced by cost 1+rowlen like the reference's
 * `balance` Partitioning (linalg/sparsematrix.cpp:1259-1265). */
static void balanced_rows(size_t n, const uint64_t *firsti, int t, int T, size_t *lo, size_t *hi)
{
    uint64_t total = firsti[n] + n;
    uint64_t b0 = total * (uint64_t)t / (uint64_t)T, b1 = total * (uint64_t)(t + 1) / (uint64_t)T;
    size_t a = 0, b = n;
    while (a < b) { size_t m = (a + b) / 2; if (firsti[m] + m < b0) a = m + 1; else b = m; }
    *lo = a;
    a = *lo; b = n;
    while (a < b) { size_t m = (a + b) / 2; if (firsti[m] + m < b1) a = m + 1; else b = m; }
    *hi = (t == T - 1) ? n : a;
}

/* ===== SparseMatrix<TM>::MultAdd ==================================================== */
/* linalg/sparsematrix_impl.hpp:264-279 : fy(i) += s * RowTimesVector(i, fx)
 * linalg/sparsematrix.hpp:625-632      : sum = 0; for j in row: sum += data[j]*vec(colnr[j]) */

void orc_csr_multadd_d(size_t n, const uint64_t *firsti, const int32_t *colnr, const double *data,
                       double s, const double *x, double *y)
{
#pragma omp parallel
    {
        int T = 1, t = 0;
#ifdef _OPENMP
        T = omp_get_num_threads(); t = omp_get_thread_num();
#endif
        size_t lo, hi;
        balanced_rows(n, firsti, t, T, &lo, &hi);
        for (size_t i = lo; i < hi; i++) {
            double sum = 0.0;
            for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++)
                sum += data[j] * x[colnr[j]];
            y[i] += s * sum;
        }
    }
}

void orc_csr_multadd_z(size_t n, const uint64_t *firsti, const int32_t *colnr, const orc_cplx *data,
                       double s, const orc_cplx *x, orc_cplx *y)
{
#pragma omp parallel
    {
        int T = 1, t = 0;
#ifdef _OPENMP
        T = omp_get_num_threads(); t = omp_get_thread_num();
#endif
        size_t lo, hi;
        balanced_rows(n, firsti, t, T, &lo, &hi);
        for (size_t i = lo; i < hi; i++) {
            double sr = 0.0, si = 0.0; /* std::complex operator*: (ac-bd, ad+bc) */
            for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++) {
                double ar = creal(data[j]), ai = cimag(data[j]);
                double br = creal(x[colnr[j]]), bi = cimag(x[colnr[j]]);
                sr += ar * br - ai * bi;
                si += ar * bi + ai * br;
            }
            y[i] = (creal(y[i]) + s * sr) + (cimag(y[i]) + s * si) * I;
        }
    }
}

/* linalg/sparsematrix_impl.hpp:378-396 : serial loop, complex scale */
void orc_csr_multadd_zs(size_t n, const uint64_t *firsti, const int32_t *colnr, const orc_cplx *data,
                        double s_r, double s_i, const orc_cplx *x, orc_cplx *y)
{
    for (size_t i = 0; i < n; i++) {
        double sr = 0.0, si = 0.0;
        for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++) {
            double ar = creal(data[j]), ai = cimag(data[j]);
            double br = creal(x[colnr[j]]), bi = cimag(x[colnr[j]]);
            sr += ar * br - ai * bi;
            si += ar * bi + ai * br;
        }
        y[i] = (creal(y[i]) + (s_r * sr - s_i * si)) + (cimag(y[i]) + (s_r * si + s_i * sr)) * I;
    }
}

/* TM = Mat<3,3,double>, TV = Vec<3,double>: data[j]*vec is the dense 3x3 mat-vec
 * (row i: sum_k m(i,k)*v(k), k ascending), accumulated entry by entry. */
void orc_csr_multadd_b3(size_t n, const uint64_t *firsti, const int32_t *colnr, const double *data,
                        double s, const double *x, double *y)
{
#pragma omp parallel
    {
        int T = 1, t = 0;
#ifdef _OPENMP
        T = omp_get_num_threads(); t = omp_get_thread_num();
#endif
        size_t lo, hi;
        balanced_rows(n, firsti, t, T, &lo, &hi);
        for (size_t i = lo; i < hi; i++) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++) {
                const double *m = data + 9 * j;
                const double *v = x + 3 * (size_t)colnr[j];
                s0 += m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
                s1 += m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
                s2 += m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
            }
            y[3 * i + 0] += s * s0;
            y[3 * i + 1] += s * s1;
            y[3 * i + 2] += s * s2;
        }
    }
}

/* BaseMatrix::Mult, linalg/basematrix.cpp:120-127 : y.SetZero(); MultAdd(1, x, y) */
void orc_csr_mult(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr, const void *data,
                  const void *x, void *y)
{
    if (kind == 0) {
        orc_vec_set_scalar_d(n, (double *)y, 0.0);
        orc_csr_multadd_d(n, firsti, colnr, (const double *)data, 1.0, (const double *)x, (double *)y);
    } else if (kind == 1) {
        orc_vec_set_scalar_d(2 * n, (double *)y, 0.0);
        orc_csr_multadd_z(n, firsti, colnr, (const orc_cplx *)data, 1.0, (const orc_cplx *)x, (orc_cplx *)y);
    } else {
        orc_vec_set_scalar_d(3 * n, (double *)y, 0.0);
        orc_csr_multadd_b3(n, firsti, colnr, (const double *)data, 1.0, (const double *)x, (double *)y);
    }
}

/* ===== BaseVector updates ============================================================= */
/* SetScalar linalg/basevector.cpp:113-138 ; Scale :75-104 ; Set :146-186 ; Add :203-244 */

void orc_vec_set_scalar_d(size_t N, double *x, double s)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) x[i] = s;
}

void orc_vec_scale_d(size_t N, double *x, double s)
{
    if (s == 1) return; /* basevector.cpp:77 */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) x[i] *= s;
}

void orc_vec_set_d(size_t N, double *y, double s, const double *x)
{
    if (y == x && s == 1.0) return; /* basevector.cpp:155 */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) y[i] = s * x[i];
}

void orc_vec_add_d(size_t N, double *y, double s, const double *x)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) y[i] += s * x[i];
}

/* complex-scalar overloads are serial expression templates: basevector.cpp:107-111,
 * :188-199, :246-257 */
void orc_vec_scale_z(size_t N, orc_cplx *x, orc_cplx s)
{
    double sr = creal(s), si = cimag(s);
    for (size_t i = 0; i < N; i++) {
        double a = creal(x[i]), b = cimag(x[i]);
        x[i] = (a * sr - b * si) + (a * si + b * sr) * I;
    }
}

void orc_vec_set_z(size_t N, orc_cplx *y, orc_cplx s, const orc_cplx *x)
{
    double sr = creal(s), si = cimag(s);
    for (size_t i = 0; i < N; i++) {
        double a = creal(x[i]), b = cimag(x[i]);
        y[i] = (sr * a - si * b) + (sr * b + si * a) * I;
    }
}

void orc_vec_add_z(size_t N, orc_cplx *y, orc_cplx s, const orc_cplx *x)
{
    double sr = creal(s), si = cimag(s);
    for (size_t i = 0; i < N; i++) {
        double a = creal(x[i]), b = cimag(x[i]);
        y[i] = (creal(y[i]) + (sr * a - si * b)) + (cimag(y[i]) + (sr * b + si * a)) * I;
    }
}

/* ===== reductions ===================================================================== */
/* S_BaseVector<double>::InnerProduct, linalg/basevector.cpp:1126-1159: 16 fixed chunks
 * Range.Split(task,16) (netgen/libsrc/core/array.hpp:303-308), each summed sequentially
 * (basiclinalg/expr.hpp:1560-1573), then the 16 partials added in order. */
double orc_vec_inner_d(size_t N, const double *x, const double *y)
{
    double parts[16];
#pragma omp parallel for schedule(static)
    for (int t = 0; t < 16; t++) {
        size_t lo = (size_t)t * N / 16, hi = (size_t)(t + 1) * N / 16;
        double sum = 0.0;
        if (hi > lo) {
            sum = x[lo] * y[lo];
            for (size_t i = lo + 1; i < hi; i++) sum += x[i] * y[i];
        }
        parts[t] = sum;
    }
    double scal = 0;
    for (int t = 0; t < 16; t++) scal += parts[t];
    return scal;
}

/* S_BaseVector<Complex>::InnerProduct, linalg/basevector.cpp:1108-1122: serial,
 * conjugation on the argument. */
void orc_vec_inner_z(size_t N, const orc_cplx *x, const orc_cplx *y, int conjugate, double *out)
{
    double sr = 0.0, si = 0.0;
    for (size_t i = 0; i < N; i++) {
        double a = creal(x[i]), b = cimag(x[i]);
        double c = creal(y[i]), d = conjugate ? -cimag(y[i]) : cimag(y[i]);
        sr += a * c - b * d;
        si += a * d + b * c;
    }
    out[0] = sr;
    out[1] = si;
}

/* BaseVector::L2Norm, linalg/basevector.cpp:41-73: 16 chunks of L2Norm2 over FV<SCAL>. */
double orc_vec_l2norm(size_t N, const void *xv, int is_complex)
{
    double parts[16];
    const double *x = (const double *)xv;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < 16; t++) {
        size_t lo = (size_t)t * N / 16, hi = (size_t)(t + 1) * N / 16;
        double sum = 0.0;
        if (is_complex) {
            for (size_t i = lo; i < hi; i++) sum += x[2 * i] * x[2 * i] + x[2 * i + 1] * x[2 * i + 1];
        } else {
            for (size_t i = lo; i < hi; i++) sum += x[i] * x[i];
        }
        parts[t] = sum;
    }
    double sum = 0;
    for (int t = 0; t < 16; t++) sum += parts[t];
    return sqrt(sum);
}

/* ===== JacobiPrecond =================================================================== */

static int bit_test(const uint8_t *bits, size_t i) { return (bits[i >> 3] >> (i & 7)) & 1; }

/* position of (i,i) in row i, or -1 (rows are sorted ascending; linear scan is enough here) */
static int64_t diag_pos(const uint64_t *firsti, const int32_t *colnr, size_t i)
{
    for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++)
        if ((size_t)colnr[j] == i) return (int64_t)j;
    return -1;
}

/* T_CalcInverse, basiclinalg/calcinverse.cpp:26-107 (Gauss-Jordan with column pivoting),
 * reached from CalcInverse(Mat<3,3>&) via FlatMatrix (basiclinalg/calcinverse.hpp:63-68). */
static int calc_inverse_core(int n, double *inv, int *p, double *hv);
static int calc_inverse_n(int n, double *inv)
{
    int p8[8];
    double hv8[8];
    int *p = n <= 8 ? p8 : (int *)malloc((size_t)n * sizeof(int));
    double *hv = n <= 8 ? hv8 : (double *)malloc((size_t)n * sizeof(double));
    int rc = calc_inverse_core(n, inv, p, hv);
    if (n > 8) { free(p); free(hv); }
    return rc;
}

static int calc_inverse_core(int n, double *inv, int *p, double *hv)
{
    for (int j = 0; j < n; j++) p[j] = j;
    for (int j = 0; j < n; j++) {
        double maxval = fabs(inv[j * n + j]);
        int r = j;
        for (int i = j + 1; i < n; i++)
            if (fabs(inv[j * n + i]) > maxval) { r = i; maxval = fabs(inv[j * n + i]); }
        double rest = 0.0;
        for (int i = j + 1; i < n; i++) rest += fabs(inv[r * n + i]);
        if (maxval < 1e-20 * rest) return -1;
        if (r > j) {
            for (int k = 0; k < n; k++) { double tmp = inv[k * n + j]; inv[k * n + j] = inv[k * n + r]; inv[k * n + r] = tmp; }
            int tp = p[j]; p[j] = p[r]; p[r] = tp;
        }
        double hr = 1 / inv[j * n + j];
        for (int i = 0; i < n; i++) inv[j * n + i] = hr * inv[j * n + i];
        inv[j * n + j] = hr;
        for (int k = 0; k < n; k++)
            if (k != j) {
                double help = inv[n * k + j];
                double h = help * hr;
                for (int i = 0; i < n; i++) inv[n * k + i] -= help * inv[n * j + i];
                inv[k * n + j] = -h;
            }
    }
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < n; k++) hv[p[k]] = inv[k * n + i];
        for (int k = 0; k < n; k++) inv[k * n + i] = hv[k];
    }
    return 0;
}

/* JacobiPrecond ctor, linalg/jacobi.cpp:39-68: invdiag[i] = inverse(A(i,i)) for dofs in
 * `inner`, TM(0) otherwise.  A(i,i) of an absent position reads as zero
 * (SparseMatrixTM::operator() const, linalg/sparsematrix.hpp) -> 1/0 = inf for scalars. */
int orc_jacobi_setup(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr,
                     const void *data, const uint8_t *freebits, void *invdiag)
{
    int rc = 0;
    for (size_t i = 0; i < n; i++) {
        int in = !freebits || bit_test(freebits, i);
        int64_t p = diag_pos(firsti, colnr, i);
        if (kind == 0) {
            double *o = (double *)invdiag;
            if (!in) o[i] = 0.0;
            else o[i] = 1 / (p >= 0 ? ((const double *)data)[p] : 0.0);
        } else if (kind == 1) {
            orc_cplx *o = (orc_cplx *)invdiag;
            if (!in) o[i] = 0.0;
            else o[i] = 1.0 / (p >= 0 ? ((const orc_cplx *)data)[p] : 0.0);
        } else {
            double *o = (double *)invdiag + 9 * i;
            if (!in || p < 0) { for (int k = 0; k < 9; k++) o[k] = 0.0; if (in) rc = -2; }
            else {
                memcpy(o, (const double *)data + 9 * p, 9 * sizeof(double));
                if (calc_inverse_n(3, o) != 0) rc = -1;
            }
        }
    }
    return rc;
}

/* JacobiPrecond::MultAdd, linalg/jacobi.cpp:71-109 (complex :112-155):
 * fy(i) += s * (invdiag[i] * fx(i)) for i in inner. */
void orc_jacobi_multadd(int kind, size_t n, const void *invdiag, const uint8_t *freebits,
                        double s, const void *xv, void *yv)
{
    if (kind == 0) {
        const double *d = (const double *)invdiag, *x = (const double *)xv;
        double *y = (double *)yv;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++)
            if (!freebits || bit_test(freebits, i)) y[i] += s * (d[i] * x[i]);
    } else if (kind == 1) {
        const orc_cplx *d = (const orc_cplx *)invdiag, *x = (const orc_cplx *)xv;
        orc_cplx *y = (orc_cplx *)yv;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++)
            if (!freebits || bit_test(freebits, i)) {
                double a = creal(d[i]), b = cimag(d[i]), c = creal(x[i]), e = cimag(x[i]);
                double pr = a * c - b * e, pi = a * e + b * c;
                y[i] = (creal(y[i]) + s * pr) + (cimag(y[i]) + s * pi) * I;
            }
    } else {
        const double *d = (const double *)invdiag, *x = (const double *)xv;
        double *y = (double *)yv;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++)
            if (!freebits || bit_test(freebits, i)) {
                const double *m = d + 9 * i, *v = x + 3 * i;
                y[3 * i + 0] += s * (m[0] * v[0] + m[1] * v[1] + m[2] * v[2]);
                y[3 * i + 1] += s * (m[3] * v[0] + m[4] * v[1] + m[5] * v[2]);
                y[3 * i + 2] += s * (m[6] * v[0] + m[7] * v[1] + m[8] * v[2]);
            }
    }
}

/* ===== helpers for the solvers ========================================================= */

static size_t kind_scalars(int kind) { return kind == 0 ? 1 : (kind == 1 ? 2 : 3); }

static void op_mult(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr, const void *data,
                    const void *x, void *y)
{
    orc_csr_mult(kind, n, firsti, colnr, data, x, y);
}

/* w = C*d : BaseMatrix::Mult = SetZero + MultAdd(1) on the Jacobi operator, or w = d */
static void op_prec(int kind, size_t n, const void *invdiag, const uint8_t *freebits, const void *d, void *w)
{
    size_t N = n * kind_scalars(kind);
    if (!invdiag) { memcpy(w, d, N * sizeof(double)); return; }
    orc_vec_set_scalar_d(N, (double *)w, 0.0);
    orc_jacobi_multadd(kind, n, invdiag, freebits, 1.0, d, w);
}

static orc_cplx ip_generic(int kind, int ip_mode, size_t n, const void *a, const void *b)
{
    if (ip_mode == 0) return orc_vec_inner_d(n * kind_scalars(kind), (const double *)a, (const double *)b);
    double out[2];
    orc_vec_inner_z(n, (const orc_cplx *)a, (const orc_cplx *)b, ip_mode == 2, out);
    return out[0] + out[1] * I;
}

/* ===== CGSolver<IPTYPE>::Mult, linalg/cg.cpp:503-633 ================================== */
int orc_cg_solve(int kind, int ip_mode, size_t n, const uint64_t *firsti, const int32_t *colnr,
                 const void *data, const void *invdiag, const uint8_t *freebits,
                 const void *f, void *u, double prec, int maxsteps, int initialize,
                 double *history, int *nhist)
{
    size_t N = n * kind_scalars(kind);
    size_t bytes = N * sizeof(double);
    double *w = (double *)malloc(bytes), *s = (double *)malloc(bytes);
    double *d = (double *)malloc(bytes), *as = (double *)malloc(bytes);
    int nh = 0;
    int it = 0;
    orc_cplx al, be, wd, wdn, kss;
    double err;

    if (initialize) {
        memset(u, 0, bytes);                 /* u = 0.0 */
        memcpy(d, f, bytes);                 /* d = f   */
    } else {                                 /* d = f - A*u */
        op_mult(kind, n, firsti, colnr, data, u, as);
        memcpy(d, f, bytes);
        orc_vec_add_d(N, d, -1.0, as);
    }
    op_prec(kind, n, invdiag, freebits, d, w);
    memcpy(s, w, bytes);
    wdn = ip_generic(kind, ip_mode, n, w, d);
    if (history) history[nh] = cabs(wdn);
    nh++;
    if (wdn == 0.0) wdn = 1;
    err = prec * prec * cabs(wdn);

    while (it++ < maxsteps && cabs(wdn) > err) {
        op_mult(kind, n, firsti, colnr, data, s, as);      /* as = A*s */
        wd = wdn;
        kss = ip_generic(kind, ip_mode, n, s, as);
        if (kss == 0.0) break;
        al = wd / kss;
        if (ip_mode == 0) {
            orc_vec_add_d(N, (double *)u, creal(al), s);    /* u += al*s */
            orc_vec_add_d(N, d, -creal(al), as);            /* d -= al*as */
        } else {
            orc_vec_add_z(n, (orc_cplx *)u, al, (const orc_cplx *)s);
            orc_vec_add_z(n, (orc_cplx *)d, -al, (const orc_cplx *)as);
        }
        op_prec(kind, n, invdiag, freebits, d, w);           /* w = C*d */
        wdn = ip_generic(kind, ip_mode, n, d, w);
        be = wdn / wd;
        if (ip_mode == 0) {
            orc_vec_scale_d(N, s, creal(be));               /* s *= be */
            orc_vec_add_d(N, s, 1.0, w);                    /* s += w  */
        } else {
            orc_vec_scale_z(n, (orc_cplx *)s, be);
            orc_vec_add_d(N, s, 1.0, w);
        }
        if (history && nh <= maxsteps) history[nh] = cabs(wdn);
        nh++;
    }
    if (nhist) *nhist = nh;
    free(w); free(s); free(d); free(as);
    return it;
}

/* ===== GMRESSolver<IPTYPE>::Mult, linalg/cg.cpp:854-1022 ================================= */
/* IPTYPE double -> real; IPTYPE Complex -> bilinear inner product (no conjugation),
 * Givens with plain squares, exactly as written in the reference. */
int orc_gmres_solve(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr,
                    const void *data, const void *invdiag, const uint8_t *freebits,
                    const void *f, void *x, double prec, int maxsteps, int initialize,
                    double *history, int *nhist)
{
    int cplx = (kind == 1);
    int ip_mode = cplx ? 1 : 0;
    size_t N = n * kind_scalars(kind);
    size_t bytes = N * sizeof(double);
    double *v = (double *)malloc(bytes), *av = (double *)malloc(bytes), *r = (double *)malloc(bytes);
    double *w = (double *)malloc(bytes), *hv = (double *)malloc(bytes);
    double **vi = (double **)calloc((size_t)maxsteps, sizeof(double *));
    size_t ms = (size_t)maxsteps;
    orc_cplx *h = (orc_cplx *)calloc((ms + 1) * ms, sizeof(orc_cplx));
    orc_cplx *gammai = (orc_cplx *)calloc(ms + 1, sizeof(orc_cplx));
    orc_cplx *ci = (orc_cplx *)calloc(ms + 1, sizeof(orc_cplx));
    orc_cplx *si = (orc_cplx *)calloc(ms + 1, sizeof(orc_cplx));
    orc_cplx *y = (orc_cplx *)calloc(ms + 1, sizeof(orc_cplx));
#define H(i, j) h[(size_t)(i) * ms + (size_t)(j)]
    int nh = 0;

    if (initialize) {
        memset(x, 0, bytes);
        memcpy(r, f, bytes);
    } else {
        op_mult(kind, n, firsti, colnr, data, x, av);
        memcpy(r, f, bytes);
        orc_vec_add_d(N, r, -1.0, av);
    }
    if (invdiag) {
        op_prec(kind, n, invdiag, freebits, r, hv);
        memcpy(r, hv, bytes);
    }
    double norm = orc_vec_l2norm(cplx ? n : N, r, cplx);
    {   /* v = (1.0/sqrt(<r,r>)) * r */
        orc_cplx rr = ip_generic(kind, ip_mode, n, r, r);
        if (cplx) orc_vec_set_z(n, (orc_cplx *)v, 1.0 / csqrt(rr), (const orc_cplx *)r);
        else orc_vec_set_d(N, v, 1.0 / sqrt(creal(rr)), r);
    }
    gammai[0] = norm;
    if (history) history[nh] = norm;
    nh++;
    double err = prec * fabs(norm);

    int j = -1;
    while (j++ < maxsteps - 2 && norm > err) {
        vi[j] = (double *)malloc(bytes);
        memcpy(vi[j], v, bytes);
        op_mult(kind, n, firsti, colnr, data, v, av);
        if (invdiag) {
            op_prec(kind, n, invdiag, freebits, av, hv);
            memcpy(av, hv, bytes);
        }
        memcpy(w, av, bytes);
        for (int i = 0; i <= j; i++) {
            H(i, j) = ip_generic(kind, ip_mode, n, vi[i], w);
            if (cplx) orc_vec_add_z(n, (orc_cplx *)w, -H(i, j), (const orc_cplx *)vi[i]);
            else orc_vec_add_d(N, w, -creal(H(i, j)), vi[i]);
        }
        {
            orc_cplx ww = ip_generic(kind, ip_mode, n, w, w);
            H(j + 1, j) = cplx ? csqrt(ww) : (orc_cplx)sqrt(creal(ww));
            if (cplx) orc_vec_set_z(n, (orc_cplx *)v, 1. / H(j + 1, j), (const orc_cplx *)w);
            else orc_vec_set_d(N, v, 1. / creal(H(j + 1, j)), w);
        }
        for (int i = 0; i < j; i++) {
            orc_cplx hi = H(i, j), hip = H(i + 1, j);
            H(i, j) = ci[i + 1] * hi + si[i + 1] * hip;
            H(i + 1, j) = si[i + 1] * hi - ci[i + 1] * hip;
        }
        orc_cplx beta = cplx ? csqrt(H(j, j) * H(j, j) + H(j + 1, j) * H(j + 1, j))
                             : (orc_cplx)sqrt(creal(H(j, j)) * creal(H(j, j)) + creal(H(j + 1, j)) * creal(H(j + 1, j)));
        si[j + 1] = H(j + 1, j) / beta;
        ci[j + 1] = H(j, j) / beta;
        H(j, j) = beta;
        gammai[j + 1] = si[j + 1] * gammai[j];
        gammai[j] = ci[j + 1] * gammai[j];
        norm = cabs(gammai[j]);   /* reference: fabs(gammai(j)) -> std::abs for Complex */
        if (history && nh <= maxsteps) history[nh] = norm;
        nh++;
    }
    j--;
    for (int i = j; i >= 0; i--) {
        orc_cplx sum = gammai[i];
        for (int k = i + 1; k <= j; k++) sum -= H(i, k) * y[k];
        y[i] = sum / H(i, i);
    }
    for (int i = 0; i <= j; i++) {
        if (cplx) orc_vec_add_z(n, (orc_cplx *)x, y[i], (const orc_cplx *)vi[i]);
        else orc_vec_add_d(N, (double *)x, creal(y[i]), vi[i]);
    }
#undef H
    if (nhist) *nhist = nh;
    for (int i = 0; i < maxsteps; i++) free(vi[i]);
    free(vi); free(h); free(gammai); free(ci); free(si); free(y);
    free(v); free(av); free(r); free(w); free(hv);
    return j;
}

/* ===== SparseMatrix::Reorder, linalg/sparsematrix_impl.hpp:762-783 ===================== */
static int cmp_pair(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

void orc_csr_reorder(int kind, size_t n, const uint64_t *firsti, const int32_t *colnr, const void *data,
                     const uint64_t *reorder, uint64_t *nfirsti, int32_t *ncolnr, void *ndata)
{
    size_t es = (kind == 0 ? 1 : (kind == 1 ? 2 : 9)) * sizeof(double);
    uint64_t *inv = (uint64_t *)malloc(n * sizeof(uint64_t));
    for (size_t i = 0; i < n; i++) inv[reorder[i]] = i;
    nfirsti[0] = 0;
    for (size_t i = 0; i < n; i++)
        nfirsti[i + 1] = nfirsti[i] + (firsti[reorder[i] + 1] - firsti[reorder[i]]);
    for (size_t i = 0; i < n; i++) {
        size_t old = reorder[i];
        size_t len = firsti[old + 1] - firsti[old];
        /* (new col, old position) pairs sorted by new col: CreatePosition keeps rows sorted */
        int32_t *pairs = (int32_t *)malloc(2 * len * sizeof(int32_t) + 8);
        for (size_t k = 0; k < len; k++) {
            pairs[2 * k] = (int32_t)inv[colnr[firsti[old] + k]];
            pairs[2 * k + 1] = (int32_t)k;
        }
        qsort(pairs, len, 2 * sizeof(int32_t), cmp_pair);
        for (size_t k = 0; k < len; k++) {
            ncolnr[nfirsti[i] + k] = pairs[2 * k];
            memcpy((char *)ndata + (nfirsti[i] + k) * es,
                   (const char *)data + (firsti[old] + (size_t)pairs[2 * k + 1]) * es, es);
        }
        free(pairs);
    }
    free(inv);
}

/* ===== ParallelDofs ctor, linalg/paralleldofs.cpp:20-108 ================================ */
void orc_pardofs_build(int ntasks, int id, size_t ndof, const uint64_t *dp_first, const int32_t *dp_data,
                       uint64_t *ex_first, int32_t *ex_data, uint8_t *ismaster)
{
    uint64_t *cnt = (uint64_t *)calloc((size_t)ntasks, sizeof(uint64_t));
    for (size_t i = 0; i < ndof; i++)
        for (uint64_t k = dp_first[i]; k < dp_first[i + 1]; k++) cnt[dp_data[k]]++;
    ex_first[0] = 0;
    for (int p = 0; p < ntasks; p++) ex_first[p + 1] = ex_first[p] + cnt[p];
    memset(cnt, 0, (size_t)ntasks * sizeof(uint64_t));
    for (size_t i = 0; i < ndof; i++)           /* ascending local dof order per neighbour */
        for (uint64_t k = dp_first[i]; k < dp_first[i + 1]; k++) {
            int d = dp_data[k];
            ex_data[ex_first[d] + cnt[d]++] = (int32_t)i;
        }
    for (size_t i = 0; i < ndof; i++) ismaster[i] = 1;
    for (int p = 0; p < id; p++)                /* shared with any lower rank -> not master */
        for (uint64_t k = ex_first[p]; k < ex_first[p + 1]; k++) ismaster[ex_data[k]] = 0;
    free(cnt);
}


/* ---- BlockJacobiPrecond<double> ---------------------------------------------------------------------------------- */
/* ctor, linalg/blockjacobi.cpp:380-500: block b = dofs[bfirst[b] .. bfirst[b+1]); blockmat(j,k) = A(block[j], block[k])
 * (absent positions read as zero), then CalcInverse(blockmat).  The reference switches to LAPACK for blocks of 100 and
 * more dofs (basiclinalg/calcinverse.cpp:188-190); this restatement uses T_CalcInverse (:26-107) for every size.
 * inverses: blocks back to back, row-major.  Returns 0, or -1 - b if block b is singular. */
int orc_blockjacobi_setup(size_t n, const uint64_t *firsti, const int32_t *colnr, const double *data, size_t nblocks,
                          const uint64_t *bfirst, const int32_t *bdofs, double *inverses)
{
    (void)n;
    size_t off = 0;
    for (size_t b = 0; b < nblocks; b++) {
        const int bs = (int)(bfirst[b + 1] - bfirst[b]);
        const int32_t *blk = bdofs + bfirst[b];
        double *m = inverses + off;
        for (int j = 0; j < bs; j++)
            for (int k = 0; k < bs; k++) {
                double v = 0.0;
                for (uint64_t e = firsti[blk[j]]; e < firsti[blk[j] + 1]; e++)
                    if (colnr[e] == blk[k]) { v = data[e]; break; }
                m[(size_t)j * bs + k] = v;
            }
        if (bs > 0 && calc_inverse_n(bs, m) != 0) return -1 - (int)b;
        off += (size_t)bs * bs;
    }
    return 0;
}

/* MultAdd / MultTransAdd, linalg/blockjacobi.cpp:594-681: y(block) += s * (inv_b | inv_b^T) * x(block).  The reference
 * walks the blocks colour by colour; blocks are visited here in ascending order (the order only matters for the
 * rounding of dofs that sit in several blocks). */
void orc_blockjacobi_multadd(size_t nblocks, const uint64_t *bfirst, const int32_t *bdofs, const double *inverses, double s,
                             const double *x, double *y, int transpose)
{
    size_t off = 0;
    for (size_t b = 0; b < nblocks; b++) {
        const int bs = (int)(bfirst[b + 1] - bfirst[b]);
        const int32_t *blk = bdofs + bfirst[b];
        const double *m = inverses + off;
        double *hy = (double *)malloc((size_t)(bs > 0 ? bs : 1) * sizeof(double));
        for (int r = 0; r < bs; r++) {
            double sum = 0.0;
            for (int c = 0; c < bs; c++) sum += (transpose ? m[(size_t)c * bs + r] : m[(size_t)r * bs + c]) * x[blk[c]];
            hy[r] = sum;
        }
        for (int r = 0; r < bs; r++) y[blk[r]] += s * hy[r];
        free(hy);
        off += (size_t)bs * bs;
    }
}

/* ---- SparseMatrix::MultTransAdd, linalg/sparsematrix_impl.hpp:344-352 + AddRowTransToVector (linalg/sparsematrix.hpp):
 * serial scatter, row by row: y(col[j]) += Trans(data[j]) * (s * x(i)).  kind 0 / 1 (plain transpose, no conjugation) / 3 */
void orc_csr_multtransadd(int kind, size_t h, const uint64_t *firsti, const int32_t *colnr, const void *data, double s,
                          const void *xv, void *yv)
{
    if (kind == 0) {
        const double *d = (const double *)data, *x = (const double *)xv;
        double *y = (double *)yv;
        for (size_t i = 0; i < h; i++) {
            const double el = s * x[i];
            for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++) y[colnr[j]] += d[j] * el;
        }
    } else if (kind == 1) {
        const orc_cplx *d = (const orc_cplx *)data, *x = (const orc_cplx *)xv;
        orc_cplx *y = (orc_cplx *)yv;
        for (size_t i = 0; i < h; i++) {
            const orc_cplx el = s * x[i];
            for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++) y[colnr[j]] += d[j] * el;
        }
    } else {
        const double *d = (const double *)data, *x = (const double *)xv;
        double *y = (double *)yv;
        for (size_t i = 0; i < h; i++) {
            const double e0 = s * x[3 * i], e1 = s * x[3 * i + 1], e2 = s * x[3 * i + 2];
            for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++) {
                const double *m = d + 9 * j;
                double *yy = y + 3 * (size_t)colnr[j];
                yy[0] += m[0] * e0 + m[3] * e1 + m[6] * e2;
                yy[1] += m[1] * e0 + m[4] * e1 + m[7] * e2;
                yy[2] += m[2] * e0 + m[5] * e1 + m[8] * e2;
            }
        }
    }
}

/* ---- SparseMatrixSymmetric<double>::MultAdd, linalg/sparsematrix_impl.hpp:967-983: lower triangle stored (columns
 * <= row, diagonal last); y(i) += s * row(i).x ; then the strict lower part of row i is scattered transposed. */
void orc_csrsym_multadd_d(size_t n, const uint64_t *firsti, const int32_t *colnr, const double *data, double s,
                          const double *x, double *y)
{
    for (size_t i = 0; i < n; i++) {
        double sum = 0.0;
        for (uint64_t j = firsti[i]; j < firsti[i + 1]; j++) sum += data[j] * x[colnr[j]];
        y[i] += s * sum;
        const double el = s * x[i];
        uint64_t last = firsti[i + 1];
        if (last > firsti[i] && (size_t)colnr[last - 1] == i) last--;       /* AddRowTransToVectorNoDiag */
        for (uint64_t j = firsti[i]; j < last; j++) y[colnr[j]] += data[j] * el;
    }
}

/* ===== Cuthill-McKee dof ordering (own specification, no reference counterpart) =========================
 * The reference has SparseMatrix::Reorder(perm) (linalg/sparsematrix_impl.hpp:762-783) but computes no bandwidth-reducing
 * permutation itself; the device library computes one (csrc/reorder.cu) and this is its serial statement, so that the
 * permutation -- integer work -- can be checked bit for bit.  Graph = the rows' column lists (diagonal ignored).
 *   degree(i) = min(row length, 2^20 - 1)
 *   components in the order of their lowest-numbered dof; after `max_components` components the remaining dofs are
 *     appended in ascending order
 *   root of a component (George-Liu): r = lowest dof; repeat at most 8 times: x = dof of the last BFS level of r with the
 *     smallest (degree, index); if the BFS from x has more levels than the one from r, r = x, else stop
 *   Cuthill-McKee from r: parents in order, the unvisited neighbours of a parent appended sorted by (degree, index)
 *   perm = the whole sequence reversed; new row k = old row perm[k]  (the argument SparseMatrix::Reorder takes) */
#define ORC_RCM_DEGCAP ((1u << 20) - 1u)
static uint32_t rcm_deg(const uint64_t *firsti, size_t i)
{
    uint64_t l = firsti[i + 1] - firsti[i];
    return l > ORC_RCM_DEGCAP ? ORC_RCM_DEGCAP : (uint32_t)l;
}

/* plain BFS over dofs with tag != placed; returns the number of levels - 1 (eccentricity), fills queue, the last level
 * is queue[*last_a .. *total) */
static size_t rcm_bfs(size_t n, const uint64_t *firsti, const int32_t *colnr, const uint8_t *placed, uint32_t *seen, uint32_t stamp,
                      uint32_t root, uint32_t *queue, size_t *last_a, size_t *total)
{
    size_t a = 0, b = 1, ecc = 0;
    (void)n;
    queue[0] = root;
    seen[root] = stamp;
    for (;;) {
        size_t e = b;
        for (size_t q = a; q < b; q++) {
            uint32_t v = queue[q];
            for (uint64_t j = firsti[v]; j < firsti[v + 1]; j++) {
                uint32_t c = (uint32_t)colnr[j];
                if (c == v || placed[c] || seen[c] == stamp) continue;
                seen[c] = stamp;
                queue[e++] = c;
            }
        }
        if (e == b) break;
        a = b; b = e; ecc++;
    }
    *last_a = a; *total = b;
    return ecc;
}

typedef struct { uint32_t deg, id; } rcm_child;
static int rcm_cmp_child(const void *x, const void *y)
{
    const rcm_child *a = (const rcm_child *)x, *b = (const rcm_child *)y;
    if (a->deg != b->deg) return a->deg < b->deg ? -1 : 1;
    return (a->id > b->id) - (a->id < b->id);
}

void orc_rcm(size_t n, const uint64_t *firsti, const int32_t *colnr, int max_components, uint64_t *perm)
{
    uint8_t *placed = (uint8_t *)calloc(n ? n : 1, 1);
    uint32_t *seen = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    uint32_t *queue = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *order = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    size_t maxdeg = 1;
    for (size_t i = 0; i < n; i++) if (firsti[i + 1] - firsti[i] > maxdeg) maxdeg = firsti[i + 1] - firsti[i];
    rcm_child *ch = (rcm_child *)malloc(maxdeg * sizeof(rcm_child));
    size_t done = 0, scan = 0;
    uint32_t stamp = 0;
    int comps = 0;
    while (done < n) {
        while (placed[scan]) scan++;
        if (comps == max_components) {
            for (size_t i = scan; i < n; i++) if (!placed[i]) order[done++] = (uint32_t)i;
            break;
        }
        uint32_t r = (uint32_t)scan;
        size_t la, tot;
        size_t ecc = rcm_bfs(n, firsti, colnr, placed, seen, ++stamp, r, queue, &la, &tot);
        for (int it = 0; it < 8; it++) {
            uint32_t x = queue[la];
            for (size_t q = la; q < tot; q++) {
                uint32_t c = queue[q], dc = rcm_deg(firsti, c), dx = rcm_deg(firsti, x);
                if (dc < dx || (dc == dx && c < x)) x = c;
            }
            if (x == r) break;
            uint32_t *q2 = order + done;     /* scratch: the not yet used tail of order */
            size_t la2, tot2;
            size_t ecc2 = rcm_bfs(n, firsti, colnr, placed, seen, ++stamp, x, q2, &la2, &tot2);
            if (ecc2 <= ecc) break;
            r = x; ecc = ecc2;
            memcpy(queue, q2, tot2 * sizeof(uint32_t));
            la = la2; tot = tot2;
        }
        /* Cuthill-McKee from r */
        size_t head = done, tail = done;
        order[tail++] = r;
        placed[r] = 1;
        while (head < tail) {
            uint32_t v = order[head++];
            size_t k = 0;
            for (uint64_t j = firsti[v]; j < firsti[v + 1]; j++) {
                uint32_t c = (uint32_t)colnr[j];
                if (c == v || placed[c]) continue;
                placed[c] = 1;
                ch[k].deg = rcm_deg(firsti, c);
                ch[k].id = c;
                k++;
            }
            qsort(ch, k, sizeof(rcm_child), rcm_cmp_child);
            for (size_t q = 0; q < k; q++) order[tail++] = ch[q].id;
        }
        done = tail;
        comps++;
    }
    for (size_t k = 0; k < n; k++) perm[k] = order[n - 1 - k];
    free(placed); free(seen); free(queue); free(order); free(ch);
}
