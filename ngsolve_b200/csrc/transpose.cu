// transpose.cu -- SparseMatrix::CreateTranspose / MultTransAdd and symmetric (lower-triangle) storage on the device.
//
//  * ngsb_csr_transpose: SparseMatrixTM::CreateTranspose(sorted = true) (linalg/sparsematrix.cpp).  A stable radix sort
//    of the entries by column keeps, inside every row of A^T, the ascending order of the original rows.
//  * ngsb_csr_multtransadd: SparseMatrix::MultTransAdd (linalg/sparsematrix_impl.hpp:344-352) is a serial scatter
//    y(col) += Trans(val) * (s x(row)) on the CPU.  Here A^T is built once (cached in the matrix handle) and the product
//    runs through the same SELL SpMV as MultAdd: y(c) += s * sum_i Trans(A(i,c)) x(i), rows i ascending -- the
//    reference's order of contributions to y(c), summed before the scale instead of one by one (rounding only).
//  * ngsb_csr_create_symmetric: SparseMatrixSymmetric<TM> stores the lower triangle, diagonal last in its row
//    (linalg/sparsematrix.hpp:760-835) and multiplies with RowTimesVector + AddRowTransToVectorNoDiag
//    (linalg/sparsematrix_impl.hpp:967-983).  On the device the triangle is expanded once to the full CSR
//    (row i = stored row i ++ the strict upper part taken from the transposed triangle), so the bandwidth-optimal
//    SpMV kernels apply unchanged and no atomics are needed.
#include "spmv.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace ngsb {

int device_scan_u64(ngsb_ctx *ctx, uint64_t *d_a, uint64_t n);
int csr_adopt_device(ngsb_ctx *ctx, size_t h, size_t w, size_t nnz, uint64_t *d_rowptr, int32_t *d_col, double *d_val, int kind,
                     ngsb_csr **out, bool allow_reorder = true);

__global__ void __launch_bounds__(256) iota_u32_kernel(uint32_t *a, uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = (uint32_t)i;
}

// t_rowptr[c] = first position of key >= c in the sorted keys (c = 0 .. w)
__global__ void __launch_bounds__(256) lower_bound_kernel(const uint32_t *__restrict__ keys, uint64_t nnz, uint64_t *__restrict__ t_rowptr, uint64_t w)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= w; c += stride) {
        uint64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (keys[mid] < c) lo = mid + 1; else hi = mid;
        }
        t_rowptr[c] = lo;
    }
}

// entry k of A^T comes from entry e = perm[k] of A: column = row of e (binary search in rowptr), value transposed
template <int KIND>
__global__ void __launch_bounds__(256) transpose_fill_kernel(const uint64_t *__restrict__ rowptr, uint64_t h, const double *__restrict__ val,
                                                            const uint32_t *__restrict__ perm, uint64_t nnz, int32_t *__restrict__ t_col,
                                                            double *__restrict__ t_val)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const uint64_t e = perm[k];
        uint64_t lo = 0, hi = h;                 // last row with rowptr[row] <= e
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (rowptr[mid] <= e) lo = mid; else hi = mid;
        }
        t_col[k] = (int32_t)lo;
        if (KIND == NGSB_REAL) t_val[k] = val[e];
        else if (KIND == NGSB_COMPLEX) reinterpret_cast<double2 *>(t_val)[k] = reinterpret_cast<const double2 *>(val)[e];
        else {
            const double *m = val + 9 * e;
            double *t = t_val + 9 * k;
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) t[3 * a + b] = m[3 * b + a];
        }
    }
}

// raw transpose of device CSR arrays; the outputs are allocated here with the 16-entry slack csr_adopt_device wants
static int transpose_raw(ngsb_ctx *ctx, size_t h, size_t w, size_t nnz, int kind, const uint64_t *d_rowptr, const int32_t *d_col,
                         const double *d_val, uint64_t **t_rowptr, int32_t **t_col, double **t_val)
{
    NGSB_REQUIRE(nnz < (1ull << 32), "CreateTranspose: more than 2^32 entries are not supported");
    const size_t ms = kind_matscalars(kind), slack = 16;
    *t_rowptr = nullptr; *t_col = nullptr; *t_val = nullptr;
    uint32_t *keys_out = nullptr, *perm_in = nullptr, *perm_out = nullptr;
    void *tmp = nullptr;
    int rc = NGSB_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == NGSB_OK) { set_error("CreateTranspose: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; } };
    cu(cudaMalloc(t_rowptr, (w + 1) * sizeof(uint64_t)));
    cu(cudaMalloc(t_col, (nnz + slack) * sizeof(int32_t)));
    cu(cudaMalloc(t_val, (nnz + slack) * ms * sizeof(double)));
    cu(cudaMalloc(&keys_out, std::max<size_t>(1, nnz) * sizeof(uint32_t)));
    cu(cudaMalloc(&perm_in, std::max<size_t>(1, nnz) * sizeof(uint32_t)));
    cu(cudaMalloc(&perm_out, std::max<size_t>(1, nnz) * sizeof(uint32_t)));
    if (rc == NGSB_OK) {
        cu(cudaMemsetAsync(*t_col + nnz, 0, slack * sizeof(int32_t), ctx->stream));
        cu(cudaMemsetAsync(*t_val + nnz * ms, 0, slack * ms * sizeof(double), ctx->stream));
    }
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((nnz + 255) / 256, (size_t)ctx->sm_count * 32));
    if (rc == NGSB_OK && nnz) {
        iota_u32_kernel<<<grid, 256, 0, ctx->stream>>>(perm_in, nnz);
        int bits = 1;
        while (bits < 32 && (1ull << bits) < w) bits++;
        size_t tmp_bytes = 0;
        const uint32_t *keys_in = reinterpret_cast<const uint32_t *>(d_col);      // columns are non-negative
        cu(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, perm_in, perm_out, (uint64_t)nnz, 0, bits, ctx->stream));
        cu(cudaMalloc(&tmp, std::max<size_t>(1, tmp_bytes)));
        if (rc == NGSB_OK)
            cu(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, perm_in, perm_out, (uint64_t)nnz, 0, bits, ctx->stream));
    }
    if (rc == NGSB_OK) {
        const unsigned g2 = (unsigned)std::max<size_t>(1, std::min<size_t>((w + 256) / 256, (size_t)ctx->sm_count * 32));
        lower_bound_kernel<<<g2, 256, 0, ctx->stream>>>(keys_out, nnz, *t_rowptr, w);
        if (nnz) {
            if (kind == NGSB_REAL) transpose_fill_kernel<NGSB_REAL><<<grid, 256, 0, ctx->stream>>>(d_rowptr, h, d_val, perm_out, nnz, *t_col, *t_val);
            else if (kind == NGSB_COMPLEX) transpose_fill_kernel<NGSB_COMPLEX><<<grid, 256, 0, ctx->stream>>>(d_rowptr, h, d_val, perm_out, nnz, *t_col, *t_val);
            else transpose_fill_kernel<NGSB_BLOCK3><<<grid, 256, 0, ctx->stream>>>(d_rowptr, h, d_val, perm_out, nnz, *t_col, *t_val);
        }
        ctx->launches += 4;
        cu(cudaGetLastError());
        cu(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(keys_out); cudaFree(perm_in); cudaFree(perm_out); cudaFree(tmp);
    if (rc != NGSB_OK) { cudaFree(*t_rowptr); cudaFree(*t_col); cudaFree(*t_val); *t_rowptr = nullptr; *t_col = nullptr; *t_val = nullptr; }
    return rc;
}

// ---- symmetric storage -> full CSR --------------------------------------------------------------------------------
__device__ __forceinline__ bool has_diag_last(const uint64_t *rp, const int32_t *col, uint64_t i)
{
    return rp[i + 1] > rp[i] && (uint64_t)col[rp[i + 1] - 1] == i;
}

// cnt[i+1] = entries of full row i; the transposed triangle's row i starts with the diagonal when it is stored
__global__ void __launch_bounds__(256) sym_count_kernel(const uint64_t *__restrict__ l_rp, const int32_t *__restrict__ l_col,
                                                       const uint64_t *__restrict__ t_rp, uint64_t n, uint64_t *__restrict__ cnt)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        cnt[i + 1] = (l_rp[i + 1] - l_rp[i]) + (t_rp[i + 1] - t_rp[i]) - (has_diag_last(l_rp, l_col, i) ? 1 : 0);
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[0] = 0;
}

__global__ void __launch_bounds__(256) sym_fill_kernel(const uint64_t *__restrict__ l_rp, const int32_t *__restrict__ l_col, const double *__restrict__ l_val,
                                                      const uint64_t *__restrict__ t_rp, const int32_t *__restrict__ t_col, const double *__restrict__ t_val,
                                                      uint64_t n, int ms, const uint64_t *__restrict__ f_rp, int32_t *__restrict__ f_col, double *__restrict__ f_val)
{
    // one warp per row
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
        const uint64_t l0 = l_rp[i], nl = l_rp[i + 1] - l0;
        const uint64_t skip = has_diag_last(l_rp, l_col, i) ? 1 : 0;
        const uint64_t t0 = t_rp[i] + skip, nt = t_rp[i + 1] - t0;
        const uint64_t o = f_rp[i];
        for (uint64_t k = lane; k < nl; k += 32) {
            f_col[o + k] = l_col[l0 + k];
            for (int c = 0; c < ms; c++) f_val[(o + k) * ms + c] = l_val[(l0 + k) * ms + c];
        }
        for (uint64_t k = lane; k < nt; k += 32) {
            f_col[o + nl + k] = t_col[t0 + k];
            for (int c = 0; c < ms; c++) f_val[(o + nl + k) * ms + c] = t_val[(t0 + k) * ms + c];
        }
    }
}

} // namespace ngsb

using namespace ngsb;

extern "C" int ngsb_csr_transpose(const ngsb_csr *A, ngsb_csr **out)
{
    NGSB_REQUIRE(A && out, "ngsb_csr_transpose: NULL argument");
    ngsb_ctx *ctx = A->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    NGSB_REQUIRE(A->h < (1ull << 31), "CreateTranspose: height exceeds 32-bit column indices");
    uint64_t *t_rp = nullptr;
    int32_t *t_col = nullptr;
    double *t_val = nullptr;
    NGSB_TRY(csr_ensure(A));
    NGSB_TRY(transpose_raw(ctx, A->h, A->w, A->nnz, A->kind, A->d_rowptr, A->d_col, A->d_val, &t_rp, &t_col, &t_val));
    int rc = csr_adopt_device(ctx, A->w, A->h, A->nnz, t_rp, t_col, t_val, A->kind, out);
    if (rc != NGSB_OK) { cudaFree(t_rp); cudaFree(t_col); cudaFree(t_val); }
    return rc;
}

extern "C" int ngsb_csr_multtransadd(const ngsb_csr *A, const double s[2], const ngsb_vec *x, ngsb_vec *y)
{
    NGSB_REQUIRE(A && s && x && y, "SparseMatrix::MultTransAdd: NULL argument");
    NGSB_REQUIRE(x->ctx == A->ctx && y->ctx == A->ctx, "SparseMatrix::MultTransAdd: objects belong to different contexts");
    NGSB_REQUIRE(x->kind == A->kind && y->kind == A->kind, "SparseMatrix::MultTransAdd: vector kind does not match matrix kind %d", A->kind);
    NGSB_REQUIRE(x->n == A->h, "SparseMatrix::MultTransAdd: height of matrix = %zu != size of x = %zu", A->h, x->n);
    NGSB_REQUIRE(y->n == A->w, "SparseMatrix::MultTransAdd: width of matrix = %zu != size of y = %zu", A->w, y->n);
    NGSB_REQUIRE(A->kind == NGSB_COMPLEX || s[1] == 0.0, "MultTransAdd(complex) called for real matrix");
    const double *xb = x->d, *xe = x->d + x->nscal, *yb = y->d, *ye = y->d + y->nscal;
    NGSB_REQUIRE(xe <= yb || ye <= xb || x->nscal == 0 || y->nscal == 0, "SparseMatrix::MultTransAdd: x and y must not overlap");
    NGSB_CUDA(cudaSetDevice(A->ctx->device));
    if (!A->transposed) {
        ngsb_csr *T = nullptr;
        NGSB_TRY(ngsb_csr_transpose(A, &T));
        const_cast<ngsb_csr *>(A)->transposed = T;      // cache: device matrices are immutable after construction
    }
    SpmvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = A->transposed; a.x = x->d; a.y = y->d; a.sr = s[0]; a.si = s[1]; a.accumulate = true; a.epi = EPI_NONE;
    return spmv_launch(a);
}

extern "C" int ngsb_csr_create_symmetric(ngsb_ctx *ctx, size_t n, size_t nnz, const uint64_t *rowptr, const int32_t *col, const void *val,
                                         int kind, ngsb_csr **out)
{
    NGSB_REQUIRE(ctx && rowptr && out && (nnz == 0 || (col && val)), "ngsb_csr_create_symmetric: NULL argument");
    NGSB_REQUIRE(kind_valid(kind), "ngsb_csr_create_symmetric: bad kind %d", kind);
    NGSB_REQUIRE(rowptr[0] == 0 && rowptr[n] == nnz, "ngsb_csr_create_symmetric: rowptr inconsistent with nnz");
    NGSB_REQUIRE(n < (1ull << 31), "ngsb_csr_create_symmetric: dimension exceeds 32-bit column indices");
    for (size_t i = 0; i < n; i++) {
        NGSB_REQUIRE(rowptr[i] <= rowptr[i + 1], "ngsb_csr_create_symmetric: rowptr not monotone at row %zu", i);
        for (uint64_t j = rowptr[i]; j < rowptr[i + 1]; j++)
            NGSB_REQUIRE(col[j] >= 0 && (size_t)col[j] <= i, "SparseMatrixSymmetric: entry (%zu,%d) is not in the lower triangle", i, col[j]);
    }
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const size_t ms = kind_matscalars(kind), slack = 16;
    uint64_t *l_rp = nullptr, *t_rp = nullptr, *f_rp = nullptr;
    int32_t *l_col = nullptr, *t_col = nullptr, *f_col = nullptr;
    double *l_val = nullptr, *t_val = nullptr, *f_val = nullptr;
    int rc = NGSB_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == NGSB_OK) { set_error("ngsb_csr_create_symmetric: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; } };
    cu(cudaMalloc(&l_rp, (n + 1) * sizeof(uint64_t)));
    cu(cudaMalloc(&l_col, std::max<size_t>(1, nnz) * sizeof(int32_t)));
    cu(cudaMalloc(&l_val, std::max<size_t>(1, nnz) * ms * sizeof(double)));
    cu(cudaMalloc(&f_rp, (n + 1) * sizeof(uint64_t)));
    if (rc == NGSB_OK) {
        cu(cudaMemcpyAsync(l_rp, rowptr, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        if (nnz) {
            cu(cudaMemcpyAsync(l_col, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
            cu(cudaMemcpyAsync(l_val, val, nnz * ms * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    if (rc == NGSB_OK) rc = transpose_raw(ctx, n, n, nnz, kind, l_rp, l_col, l_val, &t_rp, &t_col, &t_val);
    uint64_t full_nnz = 0;
    if (rc == NGSB_OK) {
        const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 32));
        sym_count_kernel<<<grid, 256, 0, ctx->stream>>>(l_rp, l_col, t_rp, n, f_rp);
        cu(cudaGetLastError());
        if (rc == NGSB_OK) rc = device_scan_u64(ctx, f_rp, n + 1);
        if (rc == NGSB_OK) {
            cu(cudaMemcpyAsync(&full_nnz, f_rp + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
            cu(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (rc == NGSB_OK) {
        cu(cudaMalloc(&f_col, (full_nnz + slack) * sizeof(int32_t)));
        cu(cudaMalloc(&f_val, (full_nnz + slack) * ms * sizeof(double)));
    }
    if (rc == NGSB_OK) {
        cu(cudaMemsetAsync(f_col + full_nnz, 0, slack * sizeof(int32_t), ctx->stream));
        cu(cudaMemsetAsync(f_val + full_nnz * ms, 0, slack * ms * sizeof(double), ctx->stream));
        const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((n * 32 + 255) / 256, (size_t)ctx->sm_count * 32));
        sym_fill_kernel<<<grid, 256, 0, ctx->stream>>>(l_rp, l_col, l_val, t_rp, t_col, t_val, n, (int)ms, f_rp, f_col, f_val);
        ctx->launches += 2;
        cu(cudaGetLastError());
        cu(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(l_rp); cudaFree(l_col); cudaFree(l_val); cudaFree(t_rp); cudaFree(t_col); cudaFree(t_val);
    if (rc == NGSB_OK) {
        rc = csr_adopt_device(ctx, n, n, (size_t)full_nnz, f_rp, f_col, f_val, kind, out);
        if (rc == NGSB_OK) return NGSB_OK;
    }
    cudaFree(f_rp); cudaFree(f_col); cudaFree(f_val);
    return rc;
}
