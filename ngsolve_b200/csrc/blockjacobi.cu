// blockjacobi.cu -- block-Jacobi (additive Schwarz) preconditioner on the device.
//
// Replaces DevBlockJacobiMatrix (ngscuda/dev_blockjacobi.cpp:21-140) and restates the BlockJacobiPrecond<double>
// constructor (linalg/blockjacobi.cpp:380-500: block matrices A(block, block), CalcInverse of each) and
// MultAdd / MultTransAdd (linalg/blockjacobi.cpp:594-681) for TM = double.
//
// Layout in HBM: the inverse blocks are stored back to back, block b at doubles [moff[b], moff[b] + bs*bs), COLUMN-major
// (element (r, c) at c*bs + r): in the apply kernel lane r of a warp reads consecutive addresses for a fixed column c
// (coalesced stream of the 8*bs*bs bytes that bound the kernel), and x(block[c]) is one broadcast load per column.
// The reference device kernel accumulates with atomicAdd (overlapping blocks); here the block results go to a scratch
// array (one slot per (block, row)) and a second kernel adds, for every dof, its slots in ascending block order:
// deterministic, no atomics, and 16 extra bytes per block row against 8*bs bytes of matrix per block row.
#include "spmv.cuh"

struct ngsb_blockjacobi {
    ngsb_ctx *ctx = nullptr;
    size_t n = 0;             // dofs (vector length)
    size_t nblocks = 0;
    size_t total = 0;         // sum of block sizes
    size_t mtotal = 0;        // sum of squares
    uint32_t maxbs = 0;
    uint64_t *d_first = nullptr;     // nblocks+1: offsets into d_dofs / scratch
    uint64_t *d_moff = nullptr;      // nblocks+1: offsets into d_inv
    int32_t *d_dofs = nullptr;       // total
    double *d_inv = nullptr;         // mtotal, column-major blocks
    double *d_tmp = nullptr;         // total: block results
    uint64_t *d_dfirst = nullptr;    // n+1: dof -> its slots
    uint64_t *d_dslot = nullptr;     // total: slot indices, ascending block order per dof
};

namespace ngsb {

// ---- constructor kernels ------------------------------------------------------------------------------------------

// one warp per (block, block row j): scatter the entries of matrix row block[j] whose column lies in the block.
// Row-major during construction (T_CalcInverse works on rows), transposed to column-major at the end.
__global__ void __launch_bounds__(256) bj_extract_kernel(const uint64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                        const double *__restrict__ val, const uint64_t *__restrict__ first,
                                                        const uint64_t *__restrict__ moff, const int32_t *__restrict__ dofs,
                                                        const uint32_t *__restrict__ slot_block, size_t total, double *__restrict__ M)
{
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= total) return;
    const uint32_t b = slot_block[warp];
    const uint64_t f0 = first[b];
    const uint32_t bs = (uint32_t)(first[b + 1] - f0);
    const uint32_t j = (uint32_t)(warp - f0);
    const int32_t *bd = dofs + f0;
    double *row = M + moff[b] + (size_t)j * bs;
    const int32_t r = bd[j];
    const uint64_t e0 = rowptr[r], e1 = rowptr[r + 1];
    for (uint64_t e = e0; e < e1; e++) {
        const int32_t c = col[e];
        const double v = val[e];
        for (uint32_t k = lane; k < bs; k += 32)
            if (bd[k] == c) row[k] = v;
    }
}

// T_CalcInverse (basiclinalg/calcinverse.cpp:26-107) of one block per CTA, in place, row-major, same pivoting rule
// (largest |inv(j,i)|, i >= j, first maximum), same update order.  status[b] = 1 when the reference would throw
// "Inverse matrix: Matrix singular".  Finally the block is rewritten column-major.
__global__ void __launch_bounds__(128) bj_invert_kernel(const uint64_t *__restrict__ first, const uint64_t *__restrict__ moff,
                                                       double *__restrict__ M, int *__restrict__ perm_ws, double *__restrict__ col_ws,
                                                       uint32_t maxbs, int *__restrict__ status, size_t nblocks)
{
    __shared__ double s_val[128];
    __shared__ int s_idx[128];
    __shared__ double s_hr;
    __shared__ int s_r;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (size_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const int n = (int)(first[b + 1] - first[b]);
        if (n == 0) continue;
        double *inv = M + moff[b];
        int *p = perm_ws + (size_t)blockIdx.x * maxbs;
        double *hv = col_ws + (size_t)blockIdx.x * maxbs;
        for (int j = tid; j < n; j += nt) p[j] = j;
        __syncthreads();
        bool singular = false;
        for (int j = 0; j < n; j++) {
            // pivot search along row j
            double best = -1.0;
            int bi = n;
            for (int i = j + tid; i < n; i += nt) {
                double a = fabs(inv[(size_t)j * n + i]);
                if (a > best) { best = a; bi = i; }
            }
            s_val[tid] = best; s_idx[tid] = bi;
            __syncthreads();
            for (int o = nt >> 1; o > 0; o >>= 1) {
                if (tid < o) {
                    double a = s_val[tid + o]; int ai = s_idx[tid + o];
                    if (a > s_val[tid] || (a == s_val[tid] && ai < s_idx[tid])) { s_val[tid] = a; s_idx[tid] = ai; }
                }
                __syncthreads();
            }
            const int r = s_idx[0];
            const double maxval = s_val[0];
            __syncthreads();
            // rest = sum_{i>j} |inv(r,i)|
            double part = 0.0;
            for (int i = j + 1 + tid; i < n; i += nt) part += fabs(inv[(size_t)r * n + i]);
            s_val[tid] = part;
            __syncthreads();
            for (int o = nt >> 1; o > 0; o >>= 1) {
                if (tid < o) s_val[tid] += s_val[tid + o];
                __syncthreads();
            }
            if (maxval < 1e-20 * s_val[0] || maxval == 0.0) singular = true;
            __syncthreads();
            if (singular) break;
            if (r > j) {
                for (int k = tid; k < n; k += nt) {
                    double t = inv[(size_t)k * n + j];
                    inv[(size_t)k * n + j] = inv[(size_t)k * n + r];
                    inv[(size_t)k * n + r] = t;
                }
                if (tid == 0) { int t = p[j]; p[j] = p[r]; p[r] = t; }
            }
            __syncthreads();
            if (tid == 0) s_hr = 1.0 / inv[(size_t)j * n + j];
            __syncthreads();
            const double hr = s_hr;
            for (int i = tid; i < n; i += nt) inv[(size_t)j * n + i] = hr * inv[(size_t)j * n + i];
            __syncthreads();
            if (tid == 0) inv[(size_t)j * n + j] = hr;
            // help(k) = inv(k,j) must be read before anybody overwrites column j
            for (int k = tid; k < n; k += nt) hv[k] = inv[(size_t)k * n + j];
            __syncthreads();
            for (size_t t = tid; t < (size_t)n * n; t += nt) {
                const int k = (int)(t / n), i = (int)(t - (size_t)k * n);
                if (k == j) continue;
                const double help = hv[k];
                if (i == j) inv[t] = -(help * hr);
                else inv[t] -= help * inv[(size_t)j * n + i];
            }
            __syncthreads();
        }
        if (singular) {
            if (tid == 0) status[b] = 1;
            __syncthreads();
            continue;
        }
        // row exchange: column i of the result, hv(p[k]) = inv(k,i)
        for (int i = 0; i < n; i++) {
            for (int k = tid; k < n; k += nt) hv[p[k]] = inv[(size_t)k * n + i];
            __syncthreads();
            for (int k = tid; k < n; k += nt) inv[(size_t)k * n + i] = hv[k];
            __syncthreads();
        }
        // in-place transpose to column-major
        for (size_t t = tid; t < (size_t)n * n; t += nt) {
            const int k = (int)(t / n), i = (int)(t - (size_t)k * n);
            if (i > k) { double a = inv[t]; inv[t] = inv[(size_t)i * n + k]; inv[(size_t)i * n + k] = a; }
        }
        __syncthreads();
    }
}

// ---- apply ---------------------------------------------------------------------------------------------------------

// tmp(block, r) = sum_c inv(r,c) x(block[c]); TRANS: sum_c inv(c,r) x(block[c]).  One warp per block (grid-stride).
template <bool TRANS>
__global__ void __launch_bounds__(256) bj_apply_kernel(const uint64_t *__restrict__ first, const uint64_t *__restrict__ moff,
                                                      const int32_t *__restrict__ dofs, const double *__restrict__ inv,
                                                      const double *__restrict__ x, double *__restrict__ tmp, size_t nblocks)
{
    const int lane = threadIdx.x & 31;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nblocks; b += nwarps) {
        const uint64_t f0 = first[b];
        const uint32_t bs = (uint32_t)(first[b + 1] - f0);
        const int32_t *bd = dofs + f0;
        const double *m = inv + moff[b];
        if (!TRANS) {
            for (uint32_t r0 = 0; r0 < bs; r0 += 32) {
                const uint32_t r = r0 + lane;
                double acc = 0.0;
                if (r < bs)
                    for (uint32_t c = 0; c < bs; c++) acc += m[(size_t)c * bs + r] * x[bd[c]];
                if (r < bs) tmp[f0 + r] = acc;
            }
        } else {
            // column-major storage: row r of the transpose is contiguous -> lanes over c, warp reduction in lane order
            for (uint32_t r = 0; r < bs; r++) {
                double acc = 0.0;
                for (uint32_t c = lane; c < bs; c += 32) acc += m[(size_t)r * bs + c] * x[bd[c]];
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) tmp[f0 + r] = acc;
            }
        }
    }
}

// y(d) (+)= s * sum of the slots of dof d, ascending block order
template <bool ACC>
__global__ void __launch_bounds__(256) bj_gather_kernel(const uint64_t *__restrict__ dfirst, const uint64_t *__restrict__ dslot,
                                                       const double *__restrict__ tmp, double *__restrict__ y, size_t n, double s)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x; d < n; d += stride) {
        double acc = ACC ? y[d] : 0.0;
        for (uint64_t k = dfirst[d]; k < dfirst[d + 1]; k++) acc += s * tmp[dslot[k]];
        y[d] = acc;
    }
}

} // namespace ngsb

using namespace ngsb;

static void bj_free(ngsb_blockjacobi *J)
{
    if (!J) return;
    cudaFree(J->d_first); cudaFree(J->d_moff); cudaFree(J->d_dofs); cudaFree(J->d_inv); cudaFree(J->d_tmp);
    cudaFree(J->d_dfirst); cudaFree(J->d_dslot);
    delete J;
}

// A != NULL: build the block matrices from the device matrix and invert them here; A == NULL: `inverses` (host, blocks back
// to back, row-major) were computed by the caller (the reference's BlockJacobiPrecond::GetInverses) and are uploaded.
static int bj_build(ngsb_ctx *ctx, size_t ndof, const ngsb_csr *A, const double *inverses, size_t nblocks, const uint64_t *first,
                    const int32_t *dofs, ngsb_blockjacobi **out)
{
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const size_t total = nblocks ? (size_t)first[nblocks] : 0;
    NGSB_REQUIRE(nblocks == 0 || first[0] == 0, "BlockJacobiPrecond: block table must start at 0");
    std::vector<uint64_t> moff(nblocks + 1, 0), dfirst(ndof + 1, 0), dslot(total);
    std::vector<uint32_t> slot_block(total);
    uint32_t maxbs = 0;
    for (size_t b = 0; b < nblocks; b++) {
        NGSB_REQUIRE(first[b + 1] >= first[b], "BlockJacobiPrecond: block table offsets must not decrease (block %zu)", b);
        const uint64_t bs = first[b + 1] - first[b];
        NGSB_REQUIRE(bs < (1u << 15), "BlockJacobiPrecond: block %zu has %llu dofs (limit 32767)", b, (unsigned long long)bs);
        moff[b + 1] = moff[b] + bs * bs;
        if (bs > maxbs) maxbs = (uint32_t)bs;
        for (uint64_t k = first[b]; k < first[b + 1]; k++) {
            NGSB_REQUIRE(dofs[k] >= 0 && (size_t)dofs[k] < ndof, "BlockJacobiPrecond: dof %d of block %zu out of range [0,%zu)", dofs[k], b, ndof);
            dfirst[(size_t)dofs[k] + 1]++;
            slot_block[k] = (uint32_t)b;
        }
    }
    for (size_t d = 0; d < ndof; d++) dfirst[d + 1] += dfirst[d];
    {
        std::vector<uint64_t> fill(dfirst.begin(), dfirst.end() - 1);
        for (size_t k = 0; k < total; k++) dslot[fill[(size_t)dofs[k]]++] = k;      // ascending slot = ascending block per dof
    }
    ngsb_blockjacobi *J = new ngsb_blockjacobi();
    J->ctx = ctx; J->n = ndof; J->nblocks = nblocks; J->total = total; J->mtotal = (size_t)moff[nblocks]; J->maxbs = maxbs;
    uint32_t *d_slot_block = nullptr;
    int *d_status = nullptr, *d_perm = nullptr;
    double *d_colws = nullptr;
    int rc = NGSB_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == NGSB_OK) { set_error("ngsb_blockjacobi_create: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; } };
    const unsigned inv_grid = (unsigned)std::max<size_t>(1, std::min<size_t>(nblocks, (size_t)ctx->sm_count * 8));
    cu(cudaMalloc(&J->d_first, (nblocks + 1) * sizeof(uint64_t)));
    cu(cudaMalloc(&J->d_moff, (nblocks + 1) * sizeof(uint64_t)));
    cu(cudaMalloc(&J->d_dofs, std::max<size_t>(1, total) * sizeof(int32_t)));
    cu(cudaMalloc(&J->d_inv, std::max<size_t>(1, J->mtotal) * sizeof(double)));
    cu(cudaMalloc(&J->d_tmp, std::max<size_t>(1, total) * sizeof(double)));
    cu(cudaMalloc(&J->d_dfirst, (ndof + 1) * sizeof(uint64_t)));
    cu(cudaMalloc(&J->d_dslot, std::max<size_t>(1, total) * sizeof(uint64_t)));
    cu(cudaMalloc(&d_slot_block, std::max<size_t>(1, total) * sizeof(uint32_t)));
    cu(cudaMalloc(&d_status, std::max<size_t>(1, nblocks) * sizeof(int)));
    cu(cudaMalloc(&d_perm, (size_t)inv_grid * std::max<uint32_t>(1, maxbs) * sizeof(int)));
    cu(cudaMalloc(&d_colws, (size_t)inv_grid * std::max<uint32_t>(1, maxbs) * sizeof(double)));
    if (rc == NGSB_OK) {
        uint64_t zero = 0;
        cu(cudaMemcpyAsync(J->d_first, nblocks ? first : &zero, (nblocks + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        cu(cudaMemcpyAsync(J->d_moff, moff.data(), (nblocks + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        cu(cudaMemcpyAsync(J->d_dfirst, dfirst.data(), (ndof + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        if (total) {
            cu(cudaMemcpyAsync(J->d_dofs, dofs, total * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
            cu(cudaMemcpyAsync(J->d_dslot, dslot.data(), total * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
            cu(cudaMemcpyAsync(d_slot_block, slot_block.data(), total * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        }
        cu(cudaMemsetAsync(J->d_inv, 0, std::max<size_t>(1, J->mtotal) * sizeof(double), ctx->stream));
        cu(cudaMemsetAsync(d_status, 0, std::max<size_t>(1, nblocks) * sizeof(int), ctx->stream));
    }
    if (rc == NGSB_OK && total && !A) {
        // inverses from the host: row-major -> the column-major device layout
        std::vector<double> cm((size_t)moff[nblocks]);
        for (size_t b = 0; b < nblocks; b++) {
            const size_t bs = (size_t)(first[b + 1] - first[b]), off = (size_t)moff[b];
            for (size_t r = 0; r < bs; r++)
                for (size_t c = 0; c < bs; c++) cm[off + c * bs + r] = inverses[off + r * bs + c];
        }
        cu(cudaMemcpyAsync(J->d_inv, cm.data(), cm.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        cu(cudaStreamSynchronize(ctx->stream));
    } else if (rc == NGSB_OK && total && (rc = csr_ensure(A)) == NGSB_OK) {
        const size_t threads = total * 32;
        bj_extract_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, J->d_first, J->d_moff,
                                                                                      J->d_dofs, d_slot_block, total, J->d_inv);
        bj_invert_kernel<<<inv_grid, 128, 0, ctx->stream>>>(J->d_first, J->d_moff, J->d_inv, d_perm, d_colws, std::max<uint32_t>(1, maxbs),
                                                            d_status, nblocks);
        ctx->launches += 2;
        cu(cudaGetLastError());
        std::vector<int> status(nblocks);
        cu(cudaMemcpyAsync(status.data(), d_status, nblocks * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        cu(cudaStreamSynchronize(ctx->stream));
        if (rc == NGSB_OK)
            for (size_t b = 0; b < nblocks; b++)
                if (status[b]) { set_error("Inverse matrix: Matrix singular (block %zu of the block-Jacobi table)", b); rc = NGSB_ERR_INVALID; break; }
    } else {
        cu(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(d_slot_block); cudaFree(d_status); cudaFree(d_perm); cudaFree(d_colws);
    if (rc != NGSB_OK) { bj_free(J); return rc; }
    *out = J;
    return NGSB_OK;
}

extern "C" int ngsb_blockjacobi_create(const ngsb_csr *A, size_t nblocks, const uint64_t *first, const int32_t *dofs,
                                       ngsb_blockjacobi **out)
{
    NGSB_REQUIRE(A && out && (nblocks == 0 || (first && dofs)), "ngsb_blockjacobi_create: NULL argument");
    NGSB_REQUIRE(A->kind == NGSB_REAL, "BlockJacobiPrecond: only TM = double is supported on the device (as in ngscuda/dev_blockjacobi.cpp:38)");
    NGSB_REQUIRE(A->h == A->w, "BlockJacobiPrecond: matrix must be square (%zu x %zu)", A->h, A->w);
    return bj_build(A->ctx, A->h, A, nullptr, nblocks, first, dofs, out);
}

extern "C" int ngsb_blockjacobi_create_from_inverses(ngsb_ctx *ctx, size_t n, size_t nblocks, const uint64_t *first, const int32_t *dofs,
                                                     const double *inverses, ngsb_blockjacobi **out)
{
    NGSB_REQUIRE(ctx && out && (nblocks == 0 || (first && dofs)), "ngsb_blockjacobi_create_from_inverses: NULL argument");
    NGSB_REQUIRE(inverses || nblocks == 0 || first[nblocks] == 0, "ngsb_blockjacobi_create_from_inverses: inverses is NULL");
    return bj_build(ctx, n, nullptr, inverses, nblocks, first, dofs, out);
}

extern "C" int ngsb_blockjacobi_destroy(ngsb_blockjacobi *J)
{
    if (J) cudaSetDevice(J->ctx->device);
    bj_free(J);
    return NGSB_OK;
}

extern "C" int ngsb_blockjacobi_info(const ngsb_blockjacobi *J, size_t *n, size_t *nblocks, size_t *maxbs, size_t *total, size_t *matrix_entries)
{
    NGSB_REQUIRE(J, "ngsb_blockjacobi_info: NULL argument");
    if (n) *n = J->n;
    if (nblocks) *nblocks = J->nblocks;
    if (maxbs) *maxbs = J->maxbs;
    if (total) *total = J->total;
    if (matrix_entries) *matrix_entries = J->mtotal;
    return NGSB_OK;
}

// inverses of all blocks back to back, each ROW-major (the reference's FlatMatrix layout)
extern "C" int ngsb_blockjacobi_download(const ngsb_blockjacobi *J, double *inverses)
{
    NGSB_REQUIRE(J && (inverses || J->mtotal == 0), "ngsb_blockjacobi_download: NULL argument");
    if (J->mtotal == 0) return NGSB_OK;
    NGSB_CUDA(cudaSetDevice(J->ctx->device));
    std::vector<double> cm(J->mtotal);
    std::vector<uint64_t> first(J->nblocks + 1);
    NGSB_CUDA(cudaMemcpyAsync(cm.data(), J->d_inv, J->mtotal * sizeof(double), cudaMemcpyDeviceToHost, J->ctx->stream));
    NGSB_CUDA(cudaMemcpyAsync(first.data(), J->d_first, (J->nblocks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, J->ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(J->ctx->stream));
    size_t off = 0;
    for (size_t b = 0; b < J->nblocks; b++) {
        const size_t bs = (size_t)(first[b + 1] - first[b]);
        for (size_t r = 0; r < bs; r++)
            for (size_t c = 0; c < bs; c++) inverses[off + r * bs + c] = cm[off + c * bs + r];
        off += bs * bs;
    }
    return NGSB_OK;
}

static int bj_apply(const ngsb_blockjacobi *J, double s, const ngsb_vec *x, ngsb_vec *y, int transpose, bool accumulate, const char *who)
{
    NGSB_REQUIRE(J && x && y, "%s: NULL argument", who);
    NGSB_REQUIRE(x->kind == NGSB_REAL && y->kind == NGSB_REAL, "%s: block-Jacobi works on real vectors", who);
    NGSB_REQUIRE(x->n == J->n && y->n == J->n, "%s: vector sizes %zu / %zu do not fit the preconditioner (%zu)", who, x->n, y->n, J->n);
    NGSB_REQUIRE(x->ctx == J->ctx && y->ctx == J->ctx, "%s: operands live on different contexts", who);
    NGSB_REQUIRE(x->d != y->d, "%s: x and y must not alias", who);
    ngsb_ctx *ctx = J->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    SpanGuard g(ctx, KC_OTHER);
    if (J->nblocks) {
        const size_t warps = std::min<size_t>(J->nblocks, (size_t)ctx->sm_count * 64);
        const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
        if (transpose) bj_apply_kernel<true><<<grid, 256, 0, ctx->stream>>>(J->d_first, J->d_moff, J->d_dofs, J->d_inv, x->d, J->d_tmp, J->nblocks);
        else bj_apply_kernel<false><<<grid, 256, 0, ctx->stream>>>(J->d_first, J->d_moff, J->d_dofs, J->d_inv, x->d, J->d_tmp, J->nblocks);
        ctx->launches++;
    }
    const unsigned ggrid = (unsigned)std::max<size_t>(1, std::min<size_t>((J->n + 255) / 256, (size_t)ctx->sm_count * 16));
    if (accumulate) bj_gather_kernel<true><<<ggrid, 256, 0, ctx->stream>>>(J->d_dfirst, J->d_dslot, J->d_tmp, y->d, J->n, s);
    else bj_gather_kernel<false><<<ggrid, 256, 0, ctx->stream>>>(J->d_dfirst, J->d_dslot, J->d_tmp, y->d, J->n, s);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

extern "C" int ngsb_blockjacobi_multadd(const ngsb_blockjacobi *J, double s, const ngsb_vec *x, ngsb_vec *y, int transpose)
{
    return bj_apply(J, s, x, y, transpose, true, transpose ? "BlockJacobiPrecond::MultTransAdd" : "BlockJacobiPrecond::MultAdd");
}

extern "C" int ngsb_blockjacobi_mult(const ngsb_blockjacobi *J, const ngsb_vec *x, ngsb_vec *y, int transpose)
{
    return bj_apply(J, 1.0, x, y, transpose, false, transpose ? "BlockJacobiPrecond::MultTrans" : "BlockJacobiPrecond::Mult");
}
