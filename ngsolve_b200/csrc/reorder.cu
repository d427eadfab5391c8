// reorder.cu -- dof reordering on the device.
//
// 1. SparseMatrix::Reorder(perm) (linalg/sparsematrix_impl.hpp:762-783): new(i, inv[c]) = old(perm[i], c), rows of the
//    new matrix ascending in the new column numbers.  Integer/byte work, done entirely on the device (row lengths ->
//    scan -> one warp per new row: map the columns, rank them inside the row, move the values).
// 2. A computed bandwidth-reducing permutation (the reference has none): level-synchronous Cuthill-McKee.  netgen/NGSolve
//    number dofs entity by entity (vertices | edges | faces | cells, refined levels appended; comp/h1hofespace.cpp:833-880),
//    so a row's columns are spread over the whole vector: mean |i-j| = 0.29 n.  The SpMV then fetches x
//    lines from HBM many times and the 16-bit column compression never applies.  After Cuthill-McKee the gathers of the
//    resident slices fall into a window of O(n^(2/3)) entries that lives in L2.
//    The ordering has a serial specification (the test suite holds it in plain C) that is reproduced here bit for bit:
//      degree(i) = min(row length, 2^20-1); components by lowest dof, at most `max_components`, the rest appended ascending;
//      root by George-Liu (<= 8 trial BFS); level k+1 sorted by (position of the first parent, degree, index); reversed.
//    One BFS level = one expand kernel (warp per frontier dof, atomicMax stamps, atomicMin of the parent position) + two
//    stable radix sorts of the new level.
// 3. The internal use: ngsb_csr_create with option "reorder" builds inner = P A P^T (SELL copy only) and keeps perm / iperm;
//    products gather x through perm and write y through the slot -> user-row table, the fused solvers run entirely in the
//    permuted numbering (krylov.cu).
#include "spmv.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

namespace ngsb {

int device_scan_u64(ngsb_ctx *ctx, uint64_t *d_a, uint64_t n);   // vec.cu
int csr_adopt_device(ngsb_ctx *ctx, size_t h, size_t w, size_t nnz, uint64_t *d_rowptr, int32_t *d_col, double *d_val, int kind,
                     ngsb_csr **out, bool allow_reorder);

static const uint32_t RCM_PLACED = 0xffffffffu;
static const uint32_t RCM_DEGCAP = (1u << 20) - 1u;

// ------------------------------------------------------------------------------------------
// Cuthill-McKee
// ------------------------------------------------------------------------------------------
// one BFS level: frontier = queue[a, b), newly discovered dofs are appended behind b (unordered)
template <bool FINAL>
__global__ void __launch_bounds__(256) rcm_expand_kernel(const uint64_t *__restrict__ rowptr, const int32_t *__restrict__ col, uint32_t n,
                                                        uint32_t *__restrict__ queue, uint32_t a, uint32_t b, uint32_t stamp_base,
                                                        uint32_t stamp_new, uint32_t *__restrict__ tag, uint32_t *__restrict__ minpar,
                                                        uint32_t *__restrict__ count)
{
    const uint32_t w = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (a + w >= b) return;
    const uint32_t p = a + w;
    const uint32_t node = queue[p];
    const uint64_t e = rowptr[node + 1];
    for (uint64_t j = rowptr[node] + lane; j < e; j += 32) {
        const uint32_t c = (uint32_t)col[j];
        if (c == node || c >= n) continue;
        const uint32_t seen = tag[c];        // stamps only grow: a stale read can only look older than it is
        if (seen >= stamp_base && seen != stamp_new) continue;                   // earlier level of this BFS, or placed
        const uint32_t old = atomicMax(&tag[c], stamp_new);
        if (old < stamp_base) queue[b + atomicAdd(count, 1u)] = c;              // first discovery in this BFS
        if (FINAL && (old < stamp_base || old == stamp_new)) atomicMin(&minpar[c], p);
    }
}

__global__ void __launch_bounds__(256) rcm_key_kernel(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ ids, uint32_t m,
                                                     const uint32_t *__restrict__ minpar, uint32_t a, uint64_t *__restrict__ key)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const uint32_t id = ids[t];
    const uint64_t l = rowptr[id + 1] - rowptr[id];
    const uint32_t deg = l > RCM_DEGCAP ? RCM_DEGCAP : (uint32_t)l;
    key[t] = ((uint64_t)(minpar[id] - a) << 20) | deg;
}

// dof of queue[a, b) with the smallest (degree, index)
__global__ void __launch_bounds__(256) rcm_mindeg_kernel(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ queue, uint32_t a, uint32_t b,
                                                        unsigned long long *__restrict__ best)
{
    const uint32_t t = a + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b) return;
    const uint32_t id = queue[t];
    const uint64_t l = rowptr[id + 1] - rowptr[id];
    const uint32_t deg = l > RCM_DEGCAP ? RCM_DEGCAP : (uint32_t)l;
    atomicMin(best, ((unsigned long long)deg << 32) | id);
}

__global__ void __launch_bounds__(256) rcm_first_unplaced_kernel(const uint32_t *__restrict__ tag, uint32_t from, uint32_t n, uint32_t *__restrict__ out)
{
    const uint32_t t = from + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && tag[t] != RCM_PLACED) atomicMin(out, t);
}

__global__ void __launch_bounds__(256) rcm_mark_kernel(const uint32_t *__restrict__ list, uint32_t m, uint32_t *__restrict__ tag, uint32_t v)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) tag[list[t]] = v;
}

__global__ void __launch_bounds__(256) rcm_clear_kernel(uint32_t *__restrict__ tag, uint32_t n)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && tag[t] != RCM_PLACED) tag[t] = 0;
}

__global__ void __launch_bounds__(256) rcm_set1_kernel(uint32_t *queue, uint32_t root, uint32_t *tag, uint32_t stamp)
{
    queue[0] = root;
    tag[root] = stamp;
}

struct RcmUnplaced {
    const uint32_t *tag;
    __device__ bool operator()(const uint32_t &i) const { return tag[i] != RCM_PLACED; }
};

__global__ void __launch_bounds__(256) rcm_iota_kernel(uint32_t *a, uint32_t n)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] = t;
}

// perm[k] = order[n-1-k], iperm[perm[k]] = k
__global__ void __launch_bounds__(256) rcm_reverse_kernel(const uint32_t *__restrict__ order, uint32_t n, uint32_t *__restrict__ perm, uint32_t *__restrict__ iperm)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t o = order[n - 1 - k];
    perm[k] = o;
    iperm[o] = k;
}

struct RcmWork {
    ngsb_ctx *ctx;
    uint32_t n;
    const uint64_t *rowptr;
    const int32_t *col;
    uint32_t *tag = nullptr, *minpar = nullptr, *queue = nullptr, *ids2 = nullptr, *count = nullptr;
    uint64_t *key = nullptr, *key2 = nullptr;
    unsigned long long *best = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    uint32_t stamp = 1;        // next free stamp value
    uint64_t launches = 0;
    ~RcmWork()
    {
        cudaFree(tag); cudaFree(minpar); cudaFree(queue); cudaFree(ids2); cudaFree(count); cudaFree(key); cudaFree(key2); cudaFree(best); cudaFree(cub_tmp);
    }
};

// BFS from `root` over the dofs that are not placed.  queue: where the visit order goes (FINAL: the Cuthill-McKee order).
// Returns the eccentricity, the last level [*last_a, *total) of queue.
template <bool FINAL>
static int rcm_bfs(RcmWork &W, uint32_t root, uint32_t *queue, uint32_t *ecc, uint32_t *last_a, uint32_t *total)
{
    ngsb_ctx *ctx = W.ctx;
    cudaStream_t st = ctx->stream;
    if (W.stamp > 0xf0000000u - W.n - 2) {     // stamps exhausted: forget the trial marks
        rcm_clear_kernel<<<(W.n + 255) / 256, 256, 0, st>>>(W.tag, W.n);
        W.stamp = 1;
    }
    const uint32_t base = W.stamp;
    rcm_set1_kernel<<<1, 1, 0, st>>>(queue, root, W.tag, base);
    uint32_t a = 0, b = 1, L = 0;
    uint32_t *h_count = reinterpret_cast<uint32_t *>(ctx->h_pinned);
    for (;;) {
        NGSB_CUDA(cudaMemsetAsync(W.count, 0, sizeof(uint32_t), st));
        const uint64_t threads = (uint64_t)(b - a) * 32;
        rcm_expand_kernel<FINAL><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(W.rowptr, W.col, W.n, queue, a, b, base, base + L + 1, W.tag, W.minpar, W.count);
        NGSB_CUDA(cudaMemcpyAsync(h_count, W.count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        NGSB_CUDA(cudaStreamSynchronize(st));
        W.launches += 1;
        const uint32_t m = *h_count;
        if (m == 0) break;
        if (FINAL && m > 1) {
            // order the new level by (first parent, degree, index): stable sort by index, then by the 52-bit key
            size_t bytes = W.cub_bytes;
            NGSB_CUDA(cub::DeviceRadixSort::SortKeys(W.cub_tmp, bytes, queue + b, W.ids2, (int)m, 0, 32, st));
            rcm_key_kernel<<<(m + 255) / 256, 256, 0, st>>>(W.rowptr, W.ids2, m, W.minpar, a, W.key);
            int bits = 21;
            while (bits < 52 && ((uint64_t)(b - a) >> (bits - 20)) != 0) bits++;
            bytes = W.cub_bytes;
            NGSB_CUDA(cub::DeviceRadixSort::SortPairs(W.cub_tmp, bytes, W.key, W.key2, W.ids2, queue + b, (int)m, 0, bits, st));
            W.launches += 8;
        }
        a = b; b += m; L++;
    }
    W.stamp = base + L + 2;
    *ecc = L; *last_a = a; *total = b;
    return NGSB_OK;
}

// new -> old permutation (d_perm) and its inverse (d_iperm), n entries each (device, caller-allocated)
int rcm_device(ngsb_ctx *ctx, size_t n_, const uint64_t *d_rowptr, const int32_t *d_col, int max_components, uint32_t *d_perm, uint32_t *d_iperm)
{
    NvtxRange nv("Cuthill-McKee ordering");
    NGSB_REQUIRE(n_ < 0xf0000000ull / 2, "rcm: too many rows");
    const uint32_t n = (uint32_t)n_;
    if (n == 0) return NGSB_OK;
    cudaStream_t st = ctx->stream;
    RcmWork W;
    W.ctx = ctx; W.n = n; W.rowptr = d_rowptr; W.col = d_col;
    uint32_t *order = nullptr;
    NGSB_CUDA(cudaMalloc(&W.tag, (size_t)n * 4));
    NGSB_CUDA(cudaMalloc(&W.minpar, (size_t)n * 4));
    NGSB_CUDA(cudaMalloc(&W.queue, (size_t)n * 4));
    NGSB_CUDA(cudaMalloc(&W.ids2, (size_t)n * 4));
    NGSB_CUDA(cudaMalloc(&W.key, (size_t)n * 8));
    NGSB_CUDA(cudaMalloc(&W.key2, (size_t)n * 8));
    NGSB_CUDA(cudaMalloc(&W.count, 16));
    NGSB_CUDA(cudaMalloc(&W.best, 16));
    NGSB_CUDA(cudaMalloc(&order, (size_t)n * 4));
    struct OrderGuard { uint32_t *p; ~OrderGuard() { cudaFree(p); } } og{order};
    {
        size_t b1 = 0, b2 = 0, b3 = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, b1, W.queue, W.ids2, (int)n, 0, 32, st);
        cub::DeviceRadixSort::SortPairs(nullptr, b2, W.key, W.key2, W.ids2, W.queue, (int)n, 0, 52, st);
        RcmUnplaced sel{W.tag};
        cub::DeviceSelect::If(nullptr, b3, W.ids2, W.queue, W.count, (int)n, sel, st);
        W.cub_bytes = std::max(b1, std::max(b2, b3)) + 256;
        NGSB_CUDA(cudaMalloc(&W.cub_tmp, W.cub_bytes));
    }
    NGSB_CUDA(cudaMemsetAsync(W.tag, 0, (size_t)n * 4, st));
    NGSB_CUDA(cudaMemsetAsync(W.minpar, 0xff, (size_t)n * 4, st));
    uint32_t *h_u32 = reinterpret_cast<uint32_t *>(ctx->h_pinned);
    unsigned long long *h_u64 = reinterpret_cast<unsigned long long *>(ctx->h_pinned) + 4;
    uint32_t done = 0, scan = 0;
    int comps = 0;
    while (done < n) {
        // lowest dof not placed yet
        NGSB_CUDA(cudaMemsetAsync(W.count, 0xff, sizeof(uint32_t), st));
        rcm_first_unplaced_kernel<<<(n - scan + 255) / 256, 256, 0, st>>>(W.tag, scan, n, W.count);
        NGSB_CUDA(cudaMemcpyAsync(h_u32, W.count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        NGSB_CUDA(cudaStreamSynchronize(st));
        scan = *h_u32;
        NGSB_REQUIRE(scan < n, "rcm: internal error (no unplaced dof left)");
        if (comps == max_components) {
            // the rest in ascending order
            rcm_iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(W.ids2, n);
            RcmUnplaced sel{W.tag};
            size_t bytes = W.cub_bytes;
            NGSB_CUDA(cub::DeviceSelect::If(W.cub_tmp, bytes, W.ids2, order + done, W.count, (int)n, sel, st));
            NGSB_CUDA(cudaMemcpyAsync(h_u32, W.count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            NGSB_CUDA(cudaStreamSynchronize(st));
            NGSB_REQUIRE(*h_u32 == n - done, "rcm: internal error (%u dofs left, %u selected)", n - done, *h_u32);
            done = n;
            break;
        }
        uint32_t r = scan, ecc = 0, la = 0, tot = 0;
        NGSB_TRY(rcm_bfs<false>(W, r, W.queue, &ecc, &la, &tot));
        for (int it = 0; it < 8; it++) {
            NGSB_CUDA(cudaMemsetAsync(W.best, 0xff, sizeof(unsigned long long), st));
            rcm_mindeg_kernel<<<(tot - la + 255) / 256, 256, 0, st>>>(d_rowptr, W.queue, la, tot, W.best);
            NGSB_CUDA(cudaMemcpyAsync(h_u64, W.best, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            NGSB_CUDA(cudaStreamSynchronize(st));
            const uint32_t x = (uint32_t)(*h_u64 & 0xffffffffull);
            if (x == r) break;
            uint32_t ecc2 = 0, la2 = 0, tot2 = 0;
            uint32_t *q2 = order + done;      // scratch: the unused tail of `order`
            NGSB_TRY(rcm_bfs<false>(W, x, q2, &ecc2, &la2, &tot2));
            if (ecc2 <= ecc) break;
            r = x; ecc = ecc2; la = la2; tot = tot2;
            NGSB_CUDA(cudaMemcpyAsync(W.queue, q2, (size_t)tot2 * 4, cudaMemcpyDeviceToDevice, st));
        }
        NGSB_TRY(rcm_bfs<true>(W, r, order + done, &ecc, &la, &tot));
        rcm_mark_kernel<<<(tot + 255) / 256, 256, 0, st>>>(order + done, tot, W.tag, RCM_PLACED);
        done += tot;
        comps++;
    }
    rcm_reverse_kernel<<<(n + 255) / 256, 256, 0, st>>>(order, n, d_perm, d_iperm);
    NGSB_CUDA(cudaGetLastError());
    NGSB_CUDA(cudaStreamSynchronize(st));
    ctx->launches += W.launches;
    return NGSB_OK;
}

// ------------------------------------------------------------------------------------------
// Reorder on the device
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) perm_rowlen_kernel(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ perm, uint64_t n, uint64_t *__restrict__ nrp)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) nrp[0] = 0;
    if (i < n) { const uint32_t o = perm[i]; nrp[i + 1] = rowptr[o + 1] - rowptr[o]; }
}

static constexpr int PERM_SMEM_ROW = 1024;      // rows up to this length are ranked out of shared memory

// one warp per new row: columns mapped through iperm, ranked inside the row (all distinct), values moved
__global__ void __launch_bounds__(256) perm_fill_kernel(const uint64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ val,
                                                       const uint32_t *__restrict__ perm, const uint32_t *__restrict__ iperm, uint64_t n, int ms,
                                                       const uint64_t *__restrict__ nrp, int32_t *__restrict__ ncol, double *__restrict__ nval)
{
    __shared__ int32_t buf[8][PERM_SMEM_ROW];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t i = (uint64_t)blockIdx.x * 8 + wid;
    if (i >= n) return;
    const uint32_t o = perm[i];
    const uint64_t src = rowptr[o], dst = nrp[i];
    const uint32_t len = (uint32_t)(rowptr[o + 1] - src);
    if (len <= PERM_SMEM_ROW) {
        for (uint32_t k = lane; k < len; k += 32) buf[wid][k] = (int32_t)iperm[col[src + k]];
        __syncwarp();
        for (uint32_t k = lane; k < len; k += 32) {
            const int32_t c = buf[wid][k];
            uint32_t rank = 0;
            for (uint32_t q = 0; q < len; q++) rank += buf[wid][q] < c;
            ncol[dst + rank] = c;
            for (int z = 0; z < ms; z++) nval[(dst + rank) * ms + z] = val[(src + k) * ms + z];
        }
    } else {
        // long row (rare): rank straight from global memory
        for (uint32_t k = lane; k < len; k += 32) {
            const int32_t c = (int32_t)iperm[col[src + k]];
            uint32_t rank = 0;
            for (uint32_t q = 0; q < len; q++) rank += (int32_t)iperm[col[src + q]] < c;
            ncol[dst + rank] = c;
            for (int z = 0; z < ms; z++) nval[(dst + rank) * ms + z] = val[(src + k) * ms + z];
        }
    }
}

__global__ void __launch_bounds__(256) perm_check_kernel(const uint32_t *__restrict__ perm, uint64_t n, uint32_t *__restrict__ hits, uint32_t *__restrict__ bad)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = perm[i];
    if (p >= n) { atomicAdd(bad, 1u); return; }
    if (atomicAdd(&hits[p], 1u) != 0) atomicAdd(bad, 1u);
}

__global__ void __launch_bounds__(256) perm_invert_kernel(const uint32_t *__restrict__ perm, uint64_t n, uint32_t *__restrict__ iperm)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) iperm[perm[i]] = (uint32_t)i;
}

// P A P^T as fresh device CSR arrays (16 entries of zeroed slack behind nnz, like alloc_csr)
int csr_permute_device(const ngsb_csr *A, const uint32_t *d_perm, const uint32_t *d_iperm, uint64_t **o_rowptr, int32_t **o_col, double **o_val)
{
    ngsb_ctx *ctx = A->ctx;
    const size_t n = A->h, ms = kind_matscalars(A->kind), slack = 16;
    NGSB_TRY(csr_ensure(A));
    uint64_t *nrp = nullptr;
    int32_t *ncol = nullptr;
    double *nval = nullptr;
    cudaError_t e = cudaMalloc(&nrp, (n + 1) * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&ncol, (A->nnz + slack) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&nval, (A->nnz + slack) * ms * sizeof(double));
    if (e != cudaSuccess) {
        cudaFree(nrp); cudaFree(ncol); cudaFree(nval);
        set_error("Reorder: cudaMalloc failed: %s", cudaGetErrorString(e));
        return NGSB_ERR_NOMEM;
    }
    cudaMemsetAsync(ncol + A->nnz, 0, slack * sizeof(int32_t), ctx->stream);
    cudaMemsetAsync(nval + A->nnz * ms, 0, slack * ms * sizeof(double), ctx->stream);
    perm_rowlen_kernel<<<(unsigned)((n + 256) / 256), 256, 0, ctx->stream>>>(A->d_rowptr, d_perm, n, nrp);
    int rc = device_scan_u64(ctx, nrp, n + 1);
    if (rc == NGSB_OK && n > 0) {
        perm_fill_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, d_perm, d_iperm, n, (int)ms, nrp, ncol, nval);
        if (cudaGetLastError() != cudaSuccess) rc = NGSB_ERR_CUDA;
    }
    if (rc == NGSB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = NGSB_ERR_CUDA;
    if (rc != NGSB_OK) { cudaFree(nrp); cudaFree(ncol); cudaFree(nval); set_error("Reorder: device kernels failed"); return rc; }
    ctx->launches += 3;
    *o_rowptr = nrp; *o_col = ncol; *o_val = nval;
    return NGSB_OK;
}

// share of the rows living in 32-row slices (natural order) whose every entry step spans < 65536 columns: the slices the
// 16-bit column offsets apply to.  Low on entity-by-entity numberings of unstructured meshes.
__global__ void __launch_bounds__(256) c16_estimate_kernel(const uint64_t *__restrict__ rowptr, const int32_t *__restrict__ col, uint64_t n, uint32_t nslices,
                                                          unsigned long long *__restrict__ out)
{
    const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const uint64_t row = s * 32 + lane;
    uint64_t a = 0;
    uint32_t len = 0;
    if (row < n) { a = rowptr[row]; len = (uint32_t)min((uint64_t)4096, rowptr[row + 1] - a); }
    uint32_t w = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    bool ok = true;
    int32_t last = len ? col[a + len - 1] : 0;
    const bool have = len > 0;
    for (uint32_t j = 0; j < w && ok; j++) {
        int32_t c = j < len ? col[a + j] : last;
        int32_t mn = have ? c : 0x7fffffff, mx = have ? c : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (mx - mn > 65535) ok = false;
    }
    if (lane == 0 && ok) atomicAdd(out, 1ull);
}

int c16_share_estimate(const ngsb_csr *A, double *share)
{
    ngsb_ctx *ctx = A->ctx;
    const uint32_t ns = (uint32_t)((A->h + 31) / 32);
    *share = 1.0;
    if (ns == 0) return NGSB_OK;
    unsigned long long *d = nullptr, h = 0;
    NGSB_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream);
    c16_estimate_kernel<<<(unsigned)(((uint64_t)ns * 32 + 255) / 256), 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->h, ns, d);
    cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    NGSB_CUDA(e);
    ctx->launches += 1;
    *share = (double)h / (double)ns;
    return NGSB_OK;
}

// gather / scatter of vectors through the permutation: out[i] = in[perm[i]] (entries of es doubles)
__global__ void __launch_bounds__(256) perm_gather_kernel(const double *__restrict__ in, const uint32_t *__restrict__ perm, uint64_t n, int es, double *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t o = perm[i];
        if (es == 1) out[i] = in[o];
        else if (es == 2) reinterpret_cast<double2 *>(out)[i] = reinterpret_cast<const double2 *>(in)[o];
        else { out[3 * i] = in[3 * o]; out[3 * i + 1] = in[3 * o + 1]; out[3 * i + 2] = in[3 * o + 2]; }
    }
}

int launch_perm_gather(ngsb_ctx *ctx, const double *in, const uint32_t *perm, size_t n, int es, double *out)
{
    if (n == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_VEC);
    uint64_t blocks = std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
    perm_gather_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, perm, n, es, out);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

__global__ void __launch_bounds__(256) perm_gather_strided_kernel(const double *__restrict__ in, const uint32_t *__restrict__ perm, uint64_t n, int es, int stride,
                                                                 double *__restrict__ out)
{
    const uint64_t gs = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const uint64_t o = perm[i];
        for (int k = 0; k < es; k++) out[i * stride + k] = in[o * stride + k];
    }
}

int launch_perm_gather_strided(ngsb_ctx *ctx, const double *in, const uint32_t *perm, size_t n, int es, int stride, double *out)
{
    if (n == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_VEC);
    uint64_t blocks = std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
    perm_gather_strided_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, perm, n, es, stride, out);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

// bits of a BitArray through the permutation: out bit i = in bit perm[i]
__global__ void __launch_bounds__(256) perm_bits_kernel(const uint8_t *__restrict__ in, const uint32_t *__restrict__ perm, uint64_t n, uint8_t *__restrict__ out)
{
    const uint64_t byte = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (byte * 8 >= n) return;
    unsigned v = 0;
    for (int k = 0; k < 8; k++) {
        const uint64_t i = byte * 8 + k;
        if (i < n) { const uint32_t o = perm[i]; v |= ((in[o >> 3] >> (o & 7)) & 1u) << k; }
    }
    out[byte] = (uint8_t)v;
}

int launch_perm_bits(ngsb_ctx *ctx, const uint8_t *in, const uint32_t *perm, size_t n, uint8_t *out)
{
    if (n == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_OTHER);
    const uint64_t bytes = (n + 7) / 8;
    perm_bits_kernel<<<(unsigned)((bytes + 255) / 256), 256, 0, ctx->stream>>>(in, perm, n, out);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

// slot -> user row: row_user[slot] = perm[row_of[slot]] (padding slots stay 0xffffffff)
__global__ void __launch_bounds__(256) perm_rowuser_kernel(const uint32_t *__restrict__ row_of, const uint32_t *__restrict__ perm, uint64_t nslots, uint32_t *__restrict__ row_user)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nslots) return;
    const uint32_t r = row_of[t];
    row_user[t] = r == 0xffffffffu ? r : perm[r];
}

// Rows of the Cuthill-McKee order sorted by length inside windows of `sigma` rows (the SELL-C-sigma sort, longest first, stable),
// composed INTO the permutation: the inner matrix is then numbered in slot order, its slices need no slot -> row table, y and
// the fused dot's vector are written / read as whole 256-byte runs instead of 32 scattered doubles per slice, and x is gathered
// in the same numbering.  key = (window, 2^20 - min(len, 2^20)).
__global__ void __launch_bounds__(256) perm_lenkey_kernel(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ perm, uint64_t n, uint32_t sigma,
                                                         uint64_t *__restrict__ key, uint32_t *__restrict__ id)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t o = perm[j];
    const uint64_t l = rowptr[o + 1] - rowptr[o];
    const uint32_t len = l > RCM_DEGCAP ? RCM_DEGCAP : (uint32_t)l;
    key[j] = ((j / sigma) << 21) | (uint64_t)((1u << 20) - len);
    id[j] = (uint32_t)j;
}

__global__ void __launch_bounds__(256) perm_compose_kernel(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ order, uint64_t n,
                                                          uint32_t *__restrict__ perm2, uint32_t *__restrict__ iperm2)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t o = perm[order[j]];
    perm2[j] = o;
    iperm2[o] = (uint32_t)j;
}

// d_perm / d_iperm (Cuthill-McKee) -> composed with the length sort, in place
int perm_compose_length_sort(ngsb_ctx *ctx, const uint64_t *d_rowptr, size_t n, uint32_t sigma, uint32_t *d_perm, uint32_t *d_iperm)
{
    if (sigma <= 1 || n == 0) return NGSB_OK;
    uint64_t *k1 = nullptr, *k2 = nullptr;
    uint32_t *id = nullptr, *order = nullptr, *p2 = nullptr;
    void *tmp = nullptr;
    size_t bytes = 0;
    cudaError_t e = cudaMalloc(&k1, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&k2, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&id, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&order, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p2, n * 4);
    int bits = 22;
    while (bits < 64 && ((uint64_t)(n / sigma) >> (bits - 21)) != 0) bits++;
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, k1, k2, id, order, (int)n, 0, bits, ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, bytes);
    if (e == cudaSuccess) {
        const unsigned grid = (unsigned)((n + 255) / 256);
        perm_lenkey_kernel<<<grid, 256, 0, ctx->stream>>>(d_rowptr, d_perm, n, sigma, k1, id);
        e = cub::DeviceRadixSort::SortPairs(tmp, bytes, k1, k2, id, order, (int)n, 0, bits, ctx->stream);
        if (e == cudaSuccess) {
            perm_compose_kernel<<<grid, 256, 0, ctx->stream>>>(d_perm, order, n, p2, d_iperm);
            e = cudaMemcpyAsync(d_perm, p2, n * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(k1); cudaFree(k2); cudaFree(id); cudaFree(order); cudaFree(p2); cudaFree(tmp);
    if (e != cudaSuccess) { set_error("reorder (length sort): %s", cudaGetErrorString(e)); return NGSB_ERR_CUDA; }
    ctx->launches += 8;
    return NGSB_OK;
}

// build A->inner = P A P^T (SELL only) from a device permutation that A takes ownership of
int csr_attach_inner(ngsb_csr *A, uint32_t *d_perm, uint32_t *d_iperm)
{
    ngsb_ctx *ctx = A->ctx;
    uint64_t *nrp = nullptr;
    int32_t *ncol = nullptr;
    double *nval = nullptr;
    NGSB_TRY(csr_permute_device(A, d_perm, d_iperm, &nrp, &ncol, &nval));
    // large matrices: the uploaded arrays go before the SELL copy of P A P^T is built (they come back from that copy on
    // demand, csrview.cu) -- uploaded CSR + permuted CSR + SELL copy of a 100 M-dof system do not fit 180 GB together
    if (csr_release_wanted(A)) csr_release(A);
    ngsb_csr *in = nullptr;
    int rc = csr_adopt_device(ctx, A->h, A->w, A->nnz, nrp, ncol, nval, A->kind, &in, false);
    if (rc != NGSB_OK) { cudaFree(nrp); cudaFree(ncol); cudaFree(nval); return rc; }
    // the permuted CSR is never handed out: only its SELL copy (and the row pointers) stay
    cudaFree(in->d_col); in->d_col = nullptr;
    cudaFree(in->d_val); in->d_val = nullptr;
    cudaFree(in->d_blocks); in->d_blocks = nullptr;
    cudaFree(in->d_rowoff); in->d_rowoff = nullptr;
    in->csr_released = true;
    const uint64_t nslots = (uint64_t)in->nslices * 32;
    cudaError_t e = cudaMalloc(&in->d_row_user, std::max<uint64_t>(32, nslots) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&in->d_xperm, std::max<size_t>(2, A->w * kind_scalars(A->kind)) * sizeof(double));
    if (e != cudaSuccess) { ngsb_csr_destroy(in); set_error("reorder: cudaMalloc failed: %s", cudaGetErrorString(e)); return NGSB_ERR_NOMEM; }
    if (nslots) perm_rowuser_kernel<<<(unsigned)((nslots + 255) / 256), 256, 0, ctx->stream>>>(in->d_row_of, d_perm, nslots, in->d_row_user);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ngsb_csr_destroy(in); set_error("reorder: %s", cudaGetErrorString(e)); return NGSB_ERR_CUDA; }
    A->inner = in;
    A->d_perm = d_perm;
    A->d_iperm = d_iperm;
    return NGSB_OK;
}

// decision + construction, called at the end of matrix creation (spmv.cu finish_create).  *made = whether an inner
// matrix exists afterwards (then the caller skips the SELL copy of the outer matrix).
int csr_maybe_reorder(ngsb_csr *A, bool *made)
{
    ngsb_ctx *ctx = A->ctx;
    *made = false;
    const long mode = ctx->reorder;
    if (mode == 0 || A->h != A->w || A->h < 2 || A->nnz == 0) return NGSB_OK;
    if (mode < 0) {
        // automatic: only where the numbering is visibly non-local (few slices qualify for 16-bit column offsets) and
        // the vector is too large for the gathers to be served by L1 anyway
        if (A->h < (size_t)ctx->reorder_min_rows) return NGSB_OK;
        double share = 1.0;
        NGSB_TRY(c16_share_estimate(A, &share));
        A->natural_c16_share = share;
        if (share >= 0.5) return NGSB_OK;
    }
    uint32_t *d_perm = nullptr, *d_iperm = nullptr;
    NGSB_CUDA(cudaMalloc(&d_perm, A->h * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&d_iperm, A->h * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(d_perm); set_error("reorder: cudaMalloc failed"); return NGSB_ERR_NOMEM; }
    int rc = rcm_device(ctx, A->h, A->d_rowptr, A->d_col, 64, d_perm, d_iperm);
    // ... composed with the SELL length sort, so that the inner matrix is numbered in slot order (no row table, coalesced y)
    if (rc == NGSB_OK && ctx->reorder_slot_order)
        rc = perm_compose_length_sort(ctx, A->d_rowptr, A->h, ctx->sell_sigma >= 0 ? (uint32_t)ctx->sell_sigma : 65536u, d_perm, d_iperm);
    if (rc == NGSB_OK) rc = csr_attach_inner(A, d_perm, d_iperm);
    if (rc != NGSB_OK) { cudaFree(d_perm); cudaFree(d_iperm); return rc; }
    *made = true;
    return NGSB_OK;
}

} // namespace ngsb

using namespace ngsb;

// the permutation the library would use for this matrix (Cuthill-McKee of its pattern), as SparseMatrix::Reorder's argument
extern "C" int ngsb_csr_rcm(const ngsb_csr *A, uint64_t *perm)
{
    NGSB_REQUIRE(A && perm, "ngsb_csr_rcm: NULL argument");
    NGSB_REQUIRE(A->h == A->w, "ngsb_csr_rcm: matrix must be square");
    NGSB_TRY(csr_ensure(A));
    ngsb_ctx *ctx = A->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = A->h;
    if (n == 0) return NGSB_OK;
    uint32_t *d_perm = nullptr, *d_iperm = nullptr;
    NGSB_CUDA(cudaMalloc(&d_perm, n * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&d_iperm, n * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(d_perm); set_error("ngsb_csr_rcm: cudaMalloc failed"); return NGSB_ERR_NOMEM; }
    int rc = rcm_device(ctx, n, A->d_rowptr, A->d_col, 64, d_perm, d_iperm);
    std::vector<uint32_t> h(n);
    if (rc == NGSB_OK && cudaMemcpy(h.data(), d_perm, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("ngsb_csr_rcm: copy failed"); rc = NGSB_ERR_CUDA; }
    cudaFree(d_perm); cudaFree(d_iperm);
    if (rc == NGSB_OK) for (size_t i = 0; i < n; i++) perm[i] = h[i];
    return rc;
}

// SparseMatrix::Reorder (linalg/sparsematrix_impl.hpp:762-783) on the device
extern "C" int ngsb_csr_reorder(const ngsb_csr *A, const uint64_t *perm, ngsb_csr **out)
{
    NGSB_REQUIRE(A && perm && out, "ngsb_csr_reorder: NULL argument");
    NGSB_REQUIRE(A->h == A->w, "ngsb_csr_reorder: matrix must be square");
    ngsb_ctx *ctx = A->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = A->h;
    std::vector<uint32_t> h(n);
    for (size_t i = 0; i < n; i++) {
        NGSB_REQUIRE(perm[i] < n, "ngsb_csr_reorder: perm is not a permutation (entry %zu)", i);
        h[i] = (uint32_t)perm[i];
    }
    uint32_t *d_perm = nullptr, *d_iperm = nullptr, *d_chk = nullptr;
    NGSB_CUDA(cudaMalloc(&d_perm, std::max<size_t>(1, n) * 4));
    NGSB_CUDA(cudaMalloc(&d_iperm, std::max<size_t>(1, n) * 4));
    NGSB_CUDA(cudaMalloc(&d_chk, (n + 1) * 4));
    auto freeall = [&]() { cudaFree(d_perm); cudaFree(d_iperm); cudaFree(d_chk); };
    uint32_t bad = 0;
    cudaMemcpyAsync(d_perm, h.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(d_chk, 0, (n + 1) * 4, ctx->stream);
    if (n) {
        perm_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_perm, n, d_chk, d_chk + n);
        perm_invert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_perm, n, d_iperm);
    }
    cudaMemcpyAsync(&bad, d_chk + n, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { freeall(); set_error("ngsb_csr_reorder: %s", cudaGetErrorString(e)); return NGSB_ERR_CUDA; }
    if (bad) { freeall(); set_error("ngsb_csr_reorder: perm is not a permutation (%u repeated or out-of-range entries)", bad); return NGSB_ERR_INVALID; }
    uint64_t *nrp = nullptr;
    int32_t *ncol = nullptr;
    double *nval = nullptr;
    int rc = csr_permute_device(A, d_perm, d_iperm, &nrp, &ncol, &nval);
    freeall();
    if (rc != NGSB_OK) return rc;
    rc = csr_adopt_device(ctx, n, n, A->nnz, nrp, ncol, nval, A->kind, out, true);
    if (rc != NGSB_OK) { cudaFree(nrp); cudaFree(ncol); cudaFree(nval); }
    return rc;
}

// diagnostics of the internal reordering: whether the products run on P A P^T, the permutation (new -> old), the share
// of natural 32-row slices that qualified for 16-bit column offsets (the automatic criterion; -1 if not evaluated)
extern "C" int ngsb_csr_reorder_info(const ngsb_csr *A, int *reordered, uint64_t *perm, double *natural_c16_share)
{
    NGSB_REQUIRE(A, "ngsb_csr_reorder_info: A is NULL");
    if (reordered) *reordered = A->inner != nullptr;
    if (natural_c16_share) *natural_c16_share = A->natural_c16_share;
    if (perm && A->inner) {
        std::vector<uint32_t> h(A->h);
        NGSB_CUDA(cudaMemcpy(h.data(), A->d_perm, A->h * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < A->h; i++) perm[i] = h[i];
    }
    return NGSB_OK;
}
