// jacobi.cu -- diagonal (Jacobi) preconditioner on the device.
//
// Replaces DevDiagonalMatrix built from JacobiPrecond<TM>::invdiag
// (ngscuda/cuda_linalg.cpp:103-115, 321-366) and restates the JacobiPrecond constructor
// (linalg/jacobi.cpp:39-68) on the device for TM = double, Complex, Mat<3,3,double>,
// including the `inner` (freedofs) bit mask that the reference device path drops.
#include "jacobi.cuh"

namespace ngsb {

__device__ __forceinline__ bool bit_test(const uint8_t *bits, uint64_t i) { return (bits[i >> 3] >> (i & 7)) & 1; }

// y (+)= s * (invdiag .* x) on masked entries; ACC=false writes 0 to masked-out entries
// (BaseMatrix::Mult = SetZero + MultAdd, linalg/basematrix.cpp:120-127)
template <int KIND, bool ACC>
__global__ void __launch_bounds__(256) jacobi_apply_kernel(const double *__restrict__ invdiag, const uint8_t *__restrict__ bits,
                                                          const double *__restrict__ x, double *__restrict__ y, uint64_t n,
                                                          double sr, double si)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool in = bits == nullptr || bit_test(bits, i);
        if (KIND == NGSB_REAL) {
            if (in) { double v = sr * (invdiag[i] * x[i]); y[i] = ACC ? y[i] + v : v; }
            else if (!ACC) y[i] = 0.0;
        } else if (KIND == NGSB_COMPLEX) {
            double2 *y2 = reinterpret_cast<double2 *>(y);
            if (in) {
                double2 d = reinterpret_cast<const double2 *>(invdiag)[i];
                double2 xv = reinterpret_cast<const double2 *>(x)[i];
                double pr = d.x * xv.x - d.y * xv.y, pi = d.x * xv.y + d.y * xv.x;
                double vr = sr * pr - si * pi, vi = sr * pi + si * pr;
                if (ACC) { double2 o = y2[i]; vr += o.x; vi += o.y; }
                y2[i] = make_double2(vr, vi);
            } else if (!ACC) y2[i] = make_double2(0.0, 0.0);
        } else {
            double *yy = y + 3 * i;
            if (in) {
                const double *m = invdiag + 9 * i;
                const double *v = x + 3 * i;
                double x0 = v[0], x1 = v[1], x2 = v[2];
                double r0 = sr * (m[0] * x0 + m[1] * x1 + m[2] * x2);
                double r1 = sr * (m[3] * x0 + m[4] * x1 + m[5] * x2);
                double r2 = sr * (m[6] * x0 + m[7] * x1 + m[8] * x2);
                if (ACC) { r0 += yy[0]; r1 += yy[1]; r2 += yy[2]; }
                yy[0] = r0; yy[1] = r1; yy[2] = r2;
            } else if (!ACC) { yy[0] = 0.0; yy[1] = 0.0; yy[2] = 0.0; }
        }
    }
}

// T_CalcInverse for n = 3 (basiclinalg/calcinverse.cpp:26-107), in place
__device__ int calc_inverse3(double *inv)
{
    const int n = 3;
    int p[3] = {0, 1, 2};
    double hv[3];
    for (int j = 0; j < n; j++) {
        double maxval = fabs(inv[j * n + j]);
        int r = j;
        for (int i = j + 1; i < n; i++)
            if (fabs(inv[j * n + i]) > maxval) { r = i; maxval = fabs(inv[j * n + i]); }
        double rest = 0.0;
        for (int i = j + 1; i < n; i++) rest += fabs(inv[r * n + i]);
        if (maxval < 1e-20 * rest) return -1;
        if (r > j) {
            for (int k = 0; k < n; k++) { double t = inv[k * n + j]; inv[k * n + j] = inv[k * n + r]; inv[k * n + r] = t; }
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

// JacobiPrecond ctor in two phases (linalg/jacobi.cpp:49-67):
//   extract: invdiag[i] = A(i,i) for i in inner, TM(0) otherwise   (absent position reads as 0)
//   [distributed: AllReduceDofData(invdiag, SUM), jacobi.cpp:60-61 -- done by the caller]
//   invert : CalcInverse(invdiag[i]) for i in inner
template <int KIND>
__global__ void __launch_bounds__(256) jacobi_extract_kernel(const uint64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                            const double *__restrict__ val, const uint8_t *__restrict__ bits,
                                                            double *__restrict__ invdiag, uint64_t n, int *__restrict__ status)
{
    constexpr int MS = KIND == NGSB_REAL ? 1 : (KIND == NGSB_COMPLEX ? 2 : 9);
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool in = bits == nullptr || bit_test(bits, i);
        // binary search of column i in the sorted row
        uint64_t lo = rowptr[i], hi = rowptr[i + 1];
        int64_t pos = -1;
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            int c = col[mid];
            if ((uint64_t)c == i) { pos = (int64_t)mid; break; }
            if ((uint64_t)c < i) lo = mid + 1; else hi = mid;
        }
        if (KIND == NGSB_BLOCK3 && in && pos < 0) atomicExch(status, 2);
#pragma unroll
        for (int k = 0; k < MS; k++) invdiag[MS * i + k] = (in && pos >= 0) ? val[MS * pos + k] : 0.0;
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) jacobi_invert_kernel(const uint8_t *__restrict__ bits, double *__restrict__ invdiag, uint64_t n,
                                                           int *__restrict__ status)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!(bits == nullptr || bit_test(bits, i))) continue;
        if (KIND == NGSB_REAL) {
            invdiag[i] = 1 / invdiag[i];
        } else if (KIND == NGSB_COMPLEX) {
            double2 a = reinterpret_cast<double2 *>(invdiag)[i], o;
            if (a.y == 0.0) { o.x = 1.0 / a.x; o.y = 0.0; }
            else {
                double den = a.x * a.x + a.y * a.y;
                o.x = a.x / den;
                o.y = -a.y / den;
            }
            reinterpret_cast<double2 *>(invdiag)[i] = o;
        } else {
            double m[9];
            for (int k = 0; k < 9; k++) m[k] = invdiag[9 * i + k];
            if (calc_inverse3(m) != 0) atomicExch(status, 1);
            for (int k = 0; k < 9; k++) invdiag[9 * i + k] = m[k];
        }
    }
}

int jacobi_alloc(ngsb_ctx *ctx, size_t n, int kind, const uint8_t *freebits, ngsb_jacobi **out);

static int grid_for_n(ngsb_ctx *ctx, uint64_t n)
{
    uint64_t blocks = (n + 255) / 256;
    uint64_t cap = (uint64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int jacobi_apply(const ngsb_jacobi *J, double sr, double si, const double *x, double *y, bool accumulate)
{
    ngsb_ctx *ctx = J->ctx;
    if (J->n == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_VEC);
    int grid = grid_for_n(ctx, J->n);
#define LAUNCH(K)                                                                                                          \
    if (accumulate) jacobi_apply_kernel<K, true><<<grid, 256, 0, ctx->stream>>>(J->d_invdiag, J->d_bits, x, y, J->n, sr, si); \
    else jacobi_apply_kernel<K, false><<<grid, 256, 0, ctx->stream>>>(J->d_invdiag, J->d_bits, x, y, J->n, sr, si);
    if (J->kind == NGSB_REAL) { LAUNCH(NGSB_REAL) }
    else if (J->kind == NGSB_COMPLEX) { LAUNCH(NGSB_COMPLEX) }
    else { LAUNCH(NGSB_BLOCK3) }
#undef LAUNCH
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

} // namespace ngsb

using namespace ngsb;

int ngsb::jacobi_alloc(ngsb_ctx *ctx, size_t n, int kind, const uint8_t *freebits, ngsb_jacobi **out)
{
    ngsb_jacobi *J = new ngsb_jacobi();
    J->uid = ngsb::next_uid();
    J->ctx = ctx;
    J->n = n;
    J->kind = kind;
    size_t bytes = n * kind_matscalars(kind) * sizeof(double);
    NGSB_CUDA(cudaMalloc(&J->d_invdiag, bytes ? bytes : 16));
    if (freebits) {
        size_t nb = (n + 7) / 8;
        NGSB_CUDA(cudaMalloc(&J->d_bits, nb ? nb : 1));
        if (nb) NGSB_CUDA(cudaMemcpyAsync(J->d_bits, freebits, nb, cudaMemcpyHostToDevice, ctx->stream));
        NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *out = J;
    return NGSB_OK;
}

extern "C" int ngsb_jacobi_create(ngsb_ctx *ctx, size_t n, const void *invdiag, int kind, const uint8_t *freebits,
                                  ngsb_jacobi **out)
{
    NGSB_REQUIRE(ctx && out && (invdiag || n == 0), "ngsb_jacobi_create: NULL argument");
    NGSB_REQUIRE(kind_valid(kind), "ngsb_jacobi_create: bad kind %d", kind);
    NGSB_CUDA(cudaSetDevice(ctx->device));
    ngsb_jacobi *J = nullptr;
    NGSB_TRY(jacobi_alloc(ctx, n, kind, freebits, &J));
    size_t bytes = n * kind_matscalars(kind) * sizeof(double);
    if (bytes) NGSB_CUDA(cudaMemcpyAsync(J->d_invdiag, invdiag, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = J;
    return NGSB_OK;
}

namespace ngsb {
// phase 1 / phase 2 of the JacobiPrecond ctor on an allocated J (shared with dist.cu)
int jacobi_extract(const ngsb_csr *A, ngsb_jacobi *J, int *d_status)
{
    ngsb_ctx *ctx = A->ctx;
    if (A->h == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_OTHER);
    if (A->csr_released) return sell_extract_diag(A, J->d_bits, J->d_invdiag, d_status);     // no rebuild for the diagonal (csrview.cu)
    int grid = grid_for_n(ctx, A->h);
    if (A->kind == NGSB_REAL) jacobi_extract_kernel<NGSB_REAL><<<grid, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, J->d_bits, J->d_invdiag, A->h, d_status);
    else if (A->kind == NGSB_COMPLEX) jacobi_extract_kernel<NGSB_COMPLEX><<<grid, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, J->d_bits, J->d_invdiag, A->h, d_status);
    else jacobi_extract_kernel<NGSB_BLOCK3><<<grid, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, J->d_bits, J->d_invdiag, A->h, d_status);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

int jacobi_invert(ngsb_jacobi *J, int *d_status)
{
    ngsb_ctx *ctx = J->ctx;
    if (J->n == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_OTHER);
    int grid = grid_for_n(ctx, J->n);
    if (J->kind == NGSB_REAL) jacobi_invert_kernel<NGSB_REAL><<<grid, 256, 0, ctx->stream>>>(J->d_bits, J->d_invdiag, J->n, d_status);
    else if (J->kind == NGSB_COMPLEX) jacobi_invert_kernel<NGSB_COMPLEX><<<grid, 256, 0, ctx->stream>>>(J->d_bits, J->d_invdiag, J->n, d_status);
    else jacobi_invert_kernel<NGSB_BLOCK3><<<grid, 256, 0, ctx->stream>>>(J->d_bits, J->d_invdiag, J->n, d_status);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

// cumulate: optional hook between the phases (distributed: neighbour exchange + add of the diagonal)
int jacobi_build(const ngsb_csr *A, const uint8_t *freebits, int (*cumulate)(void *, double *, int), void *cum_arg, ngsb_jacobi **out)
{
    ngsb_ctx *ctx = A->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    ngsb_jacobi *J = nullptr;
    NGSB_TRY(jacobi_alloc(ctx, A->h, A->kind, freebits, &J));
    int *d_status = nullptr;
    NGSB_CUDA(cudaMalloc(&d_status, sizeof(int)));
    NGSB_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int), ctx->stream));
    int rc = jacobi_extract(A, J, d_status);
    if (rc == NGSB_OK && cumulate) rc = cumulate(cum_arg, J->d_invdiag, (int)kind_matscalars(A->kind));
    if (rc == NGSB_OK) rc = jacobi_invert(J, d_status);
    int status = 0;
    if (rc == NGSB_OK) {
        cudaError_t e = cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("JacobiPrecond: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; }
    }
    cudaFree(d_status);
    if (rc == NGSB_OK && status == 1) {   // reference: Exception("Inverse matrix: Matrix singular"), calcinverse.cpp:58
        set_error("Inverse matrix: Matrix singular");
        rc = NGSB_ERR_INVALID;
    }
    if (rc != NGSB_OK) { ngsb_jacobi_destroy(J); return rc; }
    *out = J;
    return NGSB_OK;
}
} // namespace ngsb

namespace ngsb {
int jacobi_for_inner(const ngsb_jacobi *J, const ngsb_csr *A, const ngsb_jacobi **out)
{
    NGSB_REQUIRE(A->inner && A->d_perm, "jacobi_for_inner: matrix is not reordered");
    if (J->permuted && J->permuted_for == A->uid) { *out = J->permuted; return NGSB_OK; }
    if (J->permuted) { ngsb_jacobi_destroy(J->permuted); J->permuted = nullptr; }
    ngsb_ctx *ctx = J->ctx;
    ngsb_jacobi *P = new ngsb_jacobi();
    P->ctx = ctx; P->n = J->n; P->kind = J->kind; P->uid = next_uid();
    const size_t ms = kind_matscalars(J->kind);
    cudaError_t e = cudaMalloc(&P->d_invdiag, std::max<size_t>(1, J->n * ms) * sizeof(double));
    if (e == cudaSuccess && J->d_bits) e = cudaMalloc(&P->d_bits, (J->n + 7) / 8 + 8);
    if (e != cudaSuccess) { ngsb_jacobi_destroy(P); set_error("jacobi_for_inner: cudaMalloc failed: %s", cudaGetErrorString(e)); return NGSB_ERR_NOMEM; }
    int rc = NGSB_OK;
    if (ms == 9) {
        // 3x3 blocks: three gathers of Vec<3> rows do not apply; move the nine doubles as three interleaved triples
        for (int k = 0; k < 3 && rc == NGSB_OK; k++) rc = launch_perm_gather_strided(ctx, J->d_invdiag + 3 * k, A->d_perm, J->n, 3, 9, P->d_invdiag + 3 * k);
    } else rc = launch_perm_gather(ctx, J->d_invdiag, A->d_perm, J->n, (int)ms, P->d_invdiag);
    if (rc == NGSB_OK && J->d_bits) {
        cudaMemsetAsync(P->d_bits, 0, (J->n + 7) / 8 + 8, ctx->stream);
        rc = launch_perm_bits(ctx, J->d_bits, A->d_perm, J->n, P->d_bits);
    }
    if (rc != NGSB_OK) { ngsb_jacobi_destroy(P); return rc; }
    J->permuted = P;
    J->permuted_for = A->uid;
    *out = P;
    return NGSB_OK;
}
} // namespace ngsb

extern "C" int ngsb_jacobi_create_from_csr(const ngsb_csr *A, const uint8_t *freebits, ngsb_jacobi **out)
{
    NGSB_REQUIRE(A && out, "ngsb_jacobi_create_from_csr: NULL argument");
    NGSB_REQUIRE(A->h == A->w, "ngsb_jacobi_create_from_csr: matrix must be square");
    return jacobi_build(A, freebits, nullptr, nullptr, out);
}

extern "C" int ngsb_jacobi_destroy(ngsb_jacobi *J)
{
    if (!J) return NGSB_OK;
    cudaSetDevice(J->ctx->device);
    cudaStreamSynchronize(J->ctx->stream);
    cudaFree(J->d_invdiag);
    cudaFree(J->d_bits);
    if (J->permuted) ngsb_jacobi_destroy(J->permuted);
    delete J;
    return NGSB_OK;
}

extern "C" int ngsb_jacobi_download(const ngsb_jacobi *J, void *invdiag)
{
    NGSB_REQUIRE(J && invdiag, "ngsb_jacobi_download: NULL argument");
    NGSB_CUDA(cudaSetDevice(J->ctx->device));
    size_t bytes = J->n * kind_matscalars(J->kind) * sizeof(double);
    if (bytes) NGSB_CUDA(cudaMemcpyAsync(invdiag, J->d_invdiag, bytes, cudaMemcpyDeviceToHost, J->ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(J->ctx->stream));
    return NGSB_OK;
}

static int check_jac_args(const ngsb_jacobi *J, const ngsb_vec *x, const ngsb_vec *y, const char *who)
{
    NGSB_REQUIRE(J && x && y, "%s: NULL argument", who);
    NGSB_REQUIRE(x->ctx == J->ctx && y->ctx == J->ctx, "%s: objects belong to different contexts", who);
    NGSB_REQUIRE(x->kind == J->kind && y->kind == J->kind, "%s: vector kind does not match preconditioner kind %d", who, J->kind);
    NGSB_REQUIRE(x->n == J->n && y->n == J->n, "%s: size of preconditioner = %zu, x = %zu, y = %zu", who, J->n, x->n, y->n);
    return NGSB_OK;
}

extern "C" int ngsb_jacobi_multadd(const ngsb_jacobi *J, const double s[2], const ngsb_vec *x, ngsb_vec *y)
{
    NGSB_TRY(check_jac_args(J, x, y, "JacobiPrecond::MultAdd"));
    NGSB_REQUIRE(s, "ngsb_jacobi_multadd: s is NULL");
    NGSB_REQUIRE(J->kind == NGSB_COMPLEX || s[1] == 0.0, "MultAdd complex called for real JacobiPrecond");   // jacobi.cpp:154
    NGSB_CUDA(cudaSetDevice(J->ctx->device));
    return jacobi_apply(J, s[0], s[1], x->d, y->d, true);
}

extern "C" int ngsb_jacobi_mult(const ngsb_jacobi *J, const ngsb_vec *x, ngsb_vec *y)
{
    NGSB_TRY(check_jac_args(J, x, y, "JacobiPrecond::Mult"));
    NGSB_REQUIRE(x->d != y->d, "JacobiPrecond::Mult: x and y must differ");
    NGSB_CUDA(cudaSetDevice(J->ctx->device));
    return jacobi_apply(J, 1.0, 0.0, x->d, y->d, false);
}
