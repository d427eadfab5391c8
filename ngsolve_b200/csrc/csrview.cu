// csrview.cu -- the CSR arrays of a device matrix on demand.
//
// SparseMatrix::CSR() (linalg/python_linalg.cpp:121-138) hands out the matrix as uploaded, and a few setup paths read it
// (block-Jacobi extraction, CreateTranspose, Reorder), but the products only ever stream the SELL copy.  Holding both costs
// 65 + 77 GB for the 111 M-dof system.  With option "csr_keep" = 0 (or -1 = automatic for matrices above 4 GiB) the column /
// value arrays are released once the SELL copy exists -- the row pointers stay -- and are rebuilt from the SELL copy (own
// or, for an internally reordered matrix, the permuted one) the first time something asks for them: bit-identical to the
// upload, because the SELL copy keeps every row's entries in storage order and the permutation is a bijection.
// The Jacobi constructor does not need the rebuild at all: the diagonal is picked out of the SELL copy directly.
#include "spmv.cuh"

namespace ngsb {

// slot_of[row] = slot, pos_of[slice] = schedule position
__global__ void __launch_bounds__(256) sellview_invert_kernel(const uint32_t *__restrict__ row_of, uint64_t nslots, const uint32_t *__restrict__ slice_src,
                                                             uint32_t nslices, uint32_t *__restrict__ slot_of, uint32_t *__restrict__ pos_of)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nslots) { const uint32_t r = row_of[t]; if (r != 0xffffffffu) slot_of[r] = (uint32_t)t; }
    if (t < nslices) pos_of[slice_src[t]] = (uint32_t)t;
}

struct SellView {
    const uint64_t *slice_off;
    const int32_t *scol;
    const double *sval;
    uint32_t cap;
    const uint32_t *ovf_slot;      // ascending
    const uint64_t *ovf_ptr;
    const int32_t *ovf_col;
    const double *ovf_val;
    uint32_t novf;
};

// entry j of the row living in `slot` of the slice scheduled at position t
template <int KIND>
__device__ __forceinline__ void sellview_entry(const SellView &S, uint64_t off, uint32_t lane, uint32_t slot, uint32_t j, int32_t *c, double *v /* MS doubles */)
{
    constexpr int MS = KIND == NGSB_REAL ? 1 : (KIND == NGSB_COMPLEX ? 2 : 9);
    if (j < S.cap) {
        if (KIND == NGSB_REAL) {
            const uint64_t pos = off + ((uint64_t)(j >> 1) * 32 + lane) * 2 + (j & 1);
            *c = S.scol[pos];
            v[0] = S.sval[pos];
        } else if (KIND == NGSB_COMPLEX) {
            const uint64_t pos = off + (uint64_t)j * 32 + lane;
            *c = S.scol[pos];
            v[0] = S.sval[2 * pos]; v[1] = S.sval[2 * pos + 1];
        } else {
            *c = S.scol[off + (uint64_t)j * 32 + lane];
#pragma unroll
            for (int k = 0; k < 9; k++) v[k] = S.sval[off * 9 + ((uint64_t)j * 9 + k) * 32 + lane];
        }
        return;
    }
    // overflow part: binary search of the slot in the ascending list
    uint32_t lo = 0, hi = S.novf;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (S.ovf_slot[mid] < slot) lo = mid + 1; else hi = mid; }
    const uint64_t q = S.ovf_ptr[lo] + (j - S.cap);
    *c = S.ovf_col[q];
    for (int k = 0; k < MS; k++) v[k] = S.ovf_val[q * MS + k];
}

static constexpr int VIEW_SMEM_ROW = 1024;

// one warp per USER row: entries out of the SELL copy back into CSR order
template <int KIND>
__global__ void __launch_bounds__(256) sell_to_csr_kernel(SellView S, const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ pos_of,
                                                         const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ iperm,
                                                         uint64_t n, int32_t *__restrict__ col, double *__restrict__ val)
{
    constexpr int MS = KIND == NGSB_REAL ? 1 : (KIND == NGSB_COMPLEX ? 2 : 9);
    __shared__ int32_t buf[8][VIEW_SMEM_ROW];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t r = (uint64_t)blockIdx.x * 8 + wid;
    if (r >= n) return;
    const uint64_t a = rowptr[r];
    const uint32_t len = (uint32_t)(rowptr[r + 1] - a);
    const uint32_t i = iperm ? iperm[r] : (uint32_t)r;       // row of the SELL copy
    const uint32_t slot = slot_of[i];
    const uint64_t off = S.slice_off[pos_of[slot >> 5]];
    const uint32_t sl = slot & 31;
    if (perm == nullptr) {
        for (uint32_t j = lane; j < len; j += 32) {
            int32_t c; double v[MS];
            sellview_entry<KIND>(S, off, sl, slot, j, &c, v);
            col[a + j] = c;
            for (int k = 0; k < MS; k++) val[(a + j) * MS + k] = v[k];
        }
        return;
    }
    // reordered: the copy holds the row sorted by NEW column; the upload was sorted by the caller's column = perm[new]
    if (len <= VIEW_SMEM_ROW) {
        for (uint32_t j = lane; j < len; j += 32) {
            int32_t c; double v[MS];
            sellview_entry<KIND>(S, off, sl, slot, j, &c, v);
            buf[wid][j] = (int32_t)perm[c];
        }
        __syncwarp();
        for (uint32_t j = lane; j < len; j += 32) {
            int32_t c; double v[MS];
            sellview_entry<KIND>(S, off, sl, slot, j, &c, v);
            const int32_t cu = buf[wid][j];
            uint32_t rank = 0;
            for (uint32_t q = 0; q < len; q++) rank += buf[wid][q] < cu;
            col[a + rank] = cu;
            for (int k = 0; k < MS; k++) val[(a + rank) * MS + k] = v[k];
        }
    } else {
        for (uint32_t j = lane; j < len; j += 32) {
            int32_t c; double v[MS];
            sellview_entry<KIND>(S, off, sl, slot, j, &c, v);
            const int32_t cu = (int32_t)perm[c];
            uint32_t rank = 0;
            for (uint32_t q = 0; q < len; q++) {
                int32_t c2; double v2[MS];
                sellview_entry<KIND>(S, off, sl, slot, q, &c2, v2);
                rank += (int32_t)perm[c2] < cu;
            }
            col[a + rank] = cu;
            for (int k = 0; k < MS; k++) val[(a + rank) * MS + k] = v[k];
        }
    }
}

// diagonal entries straight out of the SELL copy: one warp per scheduled slice, lane = row; the first entry whose column is
// the row itself (padding repeats the row's last column with value 0 and comes after every real entry)
template <int KIND>
__global__ void __launch_bounds__(256) sell_diag_kernel(SellView S, const uint32_t *__restrict__ slice_src, const uint32_t *__restrict__ row_of,
                                                       const uint32_t *__restrict__ row_user, const uint64_t *__restrict__ rowptr_s, uint32_t nslices,
                                                       const uint8_t *__restrict__ bits, double *__restrict__ diag, int *__restrict__ status)
{
    constexpr int MS = KIND == NGSB_REAL ? 1 : (KIND == NGSB_COMPLEX ? 2 : 9);
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (t >= nslices) return;
    const uint32_t slot = slice_src[t] * 32 + lane;
    const uint32_t row = row_of[slot];
    if (row == 0xffffffffu) return;
    const uint32_t ru = row_user ? row_user[slot] : row;
    const bool in = bits == nullptr || ((bits[ru >> 3] >> (ru & 7)) & 1);
    const uint64_t off = S.slice_off[t];
    const uint32_t len = (uint32_t)(rowptr_s[row + 1] - rowptr_s[row]);
    double d[MS];
    for (int k = 0; k < MS; k++) d[k] = 0.0;
    bool found = false;
    if (in)
        for (uint32_t j = 0; j < len && !found; j++) {
            int32_t c; double v[MS];
            sellview_entry<KIND>(S, off, lane, slot, j, &c, v);
            if ((uint32_t)c == row) { found = true; for (int k = 0; k < MS; k++) d[k] = v[k]; }
        }
    if (KIND == NGSB_BLOCK3 && in && !found) atomicExch(status, 2);
    for (int k = 0; k < MS; k++) diag[(size_t)MS * ru + k] = d[k];
}

static SellView make_view(const ngsb_csr *S)
{
    SellView v;
    v.slice_off = S->d_slice_off; v.scol = S->d_scol; v.sval = S->d_sval; v.cap = S->sell_cap;
    v.ovf_slot = S->d_ovf_slot; v.ovf_ptr = S->d_ovf_ptr; v.ovf_col = S->d_ovf_col; v.ovf_val = S->d_ovf_val; v.novf = S->novf;
    return v;
}

// the matrix whose SELL copy represents A
static const ngsb_csr *sell_holder(const ngsb_csr *A) { return A->inner ? A->inner : A; }

int sell_extract_diag(const ngsb_csr *A, const uint8_t *d_bits, double *d_diag, int *d_status)
{
    const ngsb_csr *S = sell_holder(A);
    ngsb_ctx *ctx = A->ctx;
    if (S->nslices == 0) return NGSB_OK;
    const SellView v = make_view(S);
    const unsigned grid = (unsigned)(((uint64_t)S->nslices * 32 + 255) / 256);
    const uint32_t *ru = A->inner ? S->d_row_user : nullptr;
    if (A->kind == NGSB_REAL) sell_diag_kernel<NGSB_REAL><<<grid, 256, 0, ctx->stream>>>(v, S->d_slice_src, S->d_row_of, ru, S->d_rowptr, S->nslices, d_bits, d_diag, d_status);
    else if (A->kind == NGSB_COMPLEX) sell_diag_kernel<NGSB_COMPLEX><<<grid, 256, 0, ctx->stream>>>(v, S->d_slice_src, S->d_row_of, ru, S->d_rowptr, S->nslices, d_bits, d_diag, d_status);
    else sell_diag_kernel<NGSB_BLOCK3><<<grid, 256, 0, ctx->stream>>>(v, S->d_slice_src, S->d_row_of, ru, S->d_rowptr, S->nslices, d_bits, d_diag, d_status);
    NGSB_CUDA(cudaGetLastError());
    ctx->launches++;
    return NGSB_OK;
}

// free the column / value arrays (and the structures of the legacy CSR kernels); the row pointers stay
void csr_release(ngsb_csr *A)
{
    if (A->csr_released || A->nnz == 0) return;
    cudaStreamSynchronize(A->ctx->stream);
    cudaFree(A->d_col); A->d_col = nullptr;
    cudaFree(A->d_val); A->d_val = nullptr;
    cudaFree(A->d_blocks); A->d_blocks = nullptr;
    cudaFree(A->d_rowoff); A->d_rowoff = nullptr;
    cudaFree(A->d_longrows); A->d_longrows = nullptr;
    A->nblocks = 0; A->nlong = 0;
    A->csr_released = true;
}

bool csr_release_wanted(const ngsb_csr *A)
{
    const long mode = A->ctx->csr_keep;
    if (mode == 1) return false;
    if (mode == 0) return true;
    const double bytes = (double)A->nnz * (4.0 + 8.0 * (double)kind_matscalars(A->kind));
    return bytes >= 4.0 * 1024.0 * 1024.0 * 1024.0;
}

// make d_col / d_val exist again (no-op when they do)
int csr_ensure(const ngsb_csr *Ac)
{
    ngsb_csr *A = const_cast<ngsb_csr *>(Ac);
    if (!A->csr_released) return NGSB_OK;
    const ngsb_csr *S = sell_holder(A);
    NGSB_REQUIRE(S->d_sval != nullptr && S->d_scol != nullptr, "the CSR arrays of this matrix were released and it has no SELL copy to rebuild them from");
    ngsb_ctx *ctx = A->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    NvtxRange nv("CSR rebuild from SELL");
    const size_t ms = kind_matscalars(A->kind), slack = 16;
    int32_t *col = nullptr;
    double *val = nullptr;
    uint32_t *slot_of = nullptr, *pos_of = nullptr;
    const uint64_t nslots = (uint64_t)S->nslices * 32;
    cudaError_t e = cudaMalloc(&col, (A->nnz + slack) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&val, (A->nnz + slack) * ms * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&slot_of, std::max<size_t>(1, A->h) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&pos_of, std::max<size_t>(1, S->nslices) * sizeof(uint32_t));
    if (e != cudaSuccess) {
        cudaFree(col); cudaFree(val); cudaFree(slot_of); cudaFree(pos_of);
        set_error("CSR rebuild: cudaMalloc failed: %s (option csr_keep = 1 keeps the uploaded arrays)", cudaGetErrorString(e));
        return NGSB_ERR_NOMEM;
    }
    cudaMemsetAsync(col + A->nnz, 0, slack * sizeof(int32_t), ctx->stream);
    cudaMemsetAsync(val + A->nnz * ms, 0, slack * ms * sizeof(double), ctx->stream);
    sellview_invert_kernel<<<(unsigned)((std::max<uint64_t>(nslots, S->nslices) + 255) / 256), 256, 0, ctx->stream>>>(S->d_row_of, nslots, S->d_slice_src, S->nslices, slot_of, pos_of);
    const SellView v = make_view(S);
    const unsigned grid = (unsigned)((A->h + 7) / 8);
    const uint32_t *perm = A->inner ? A->d_perm : nullptr, *iperm = A->inner ? A->d_iperm : nullptr;
    if (A->kind == NGSB_REAL) sell_to_csr_kernel<NGSB_REAL><<<grid, 256, 0, ctx->stream>>>(v, slot_of, pos_of, A->d_rowptr, perm, iperm, A->h, col, val);
    else if (A->kind == NGSB_COMPLEX) sell_to_csr_kernel<NGSB_COMPLEX><<<grid, 256, 0, ctx->stream>>>(v, slot_of, pos_of, A->d_rowptr, perm, iperm, A->h, col, val);
    else sell_to_csr_kernel<NGSB_BLOCK3><<<grid, 256, 0, ctx->stream>>>(v, slot_of, pos_of, A->d_rowptr, perm, iperm, A->h, col, val);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(slot_of); cudaFree(pos_of);
    if (e != cudaSuccess) { cudaFree(col); cudaFree(val); set_error("CSR rebuild: %s", cudaGetErrorString(e)); return NGSB_ERR_CUDA; }
    ctx->launches += 2;
    A->d_col = col; A->d_val = val;
    A->csr_released = false;
    return NGSB_OK;
}

} // namespace ngsb

using namespace ngsb;

// memory the matrix holds on the device right now (bytes): CSR arrays, SELL copy (own or the reordered one), tables
extern "C" int ngsb_csr_memory(const ngsb_csr *A, uint64_t *csr_bytes, uint64_t *sell_bytes, int *csr_resident)
{
    NGSB_REQUIRE(A, "ngsb_csr_memory: A is NULL");
    const size_t ms = kind_matscalars(A->kind);
    const ngsb_csr *S = A->inner ? A->inner : A;
    uint64_t c = (A->h + 1) * 8, s = 0;
    if (!A->csr_released) c += A->nnz * (4 + 8 * ms);
    if (S != A) c += (S->h + 1) * 8 + 2 * A->h * 4;       // inner row pointers, perm, iperm
    s += S->sell_entries * (4 + 8 * ms) + ((uint64_t)S->nslices + 1) * 8 + (uint64_t)S->nslices * 4 + (uint64_t)S->nslices * 32 * 4;
    if (S->d_scol16) s += S->sell_entries * 2 + S->sell_entries / 32 * 4 + S->nslices;
    if (S->d_row_user) s += (uint64_t)S->nslices * 32 * 4 + A->w * kind_scalars(A->kind) * 8;
    if (csr_bytes) *csr_bytes = c;
    if (sell_bytes) *sell_bytes = s;
    if (csr_resident) *csr_resident = A->csr_released ? 0 : 1;
    return NGSB_OK;
}

// ---- checkpoint wire format: SparseMatrix<TM>::DoArchive (linalg/sparsematrix_impl.hpp:443-452) into ngcore's BinaryOutArchive
// (raw little-endian scalars; Array<T> = size_t count + elements):
//   size_t size, width, nze | size_t n = size + 1, size_t firsti[n] | size_t nze, int colnr[nze] | size_t nze, TM data[nze]
// A device system written here is read by the reference's BinaryInArchive + DoArchive and vice versa.
static size_t archive_bytes(size_t h, size_t nnz, size_t ms) { return 24 + 8 + 8 * (h + 1) + 8 + 4 * nnz + 8 + 8 * ms * nnz; }

extern "C" int ngsb_csr_archive_size(const ngsb_csr *A, size_t *bytes)
{
    NGSB_REQUIRE(A && bytes, "ngsb_csr_archive_size: NULL argument");
    *bytes = archive_bytes(A->h, A->nnz, kind_matscalars(A->kind));
    return NGSB_OK;
}

extern "C" int ngsb_csr_archive_write(const ngsb_csr *A, void *buf, size_t capacity)
{
    NGSB_REQUIRE(A && buf, "ngsb_csr_archive_write: NULL argument");
    const size_t ms = kind_matscalars(A->kind), need = archive_bytes(A->h, A->nnz, ms);
    NGSB_REQUIRE(capacity >= need, "ngsb_csr_archive_write: buffer of %zu bytes, %zu needed", capacity, need);
    char *p = (char *)buf;
    auto put = [&](uint64_t v) { memcpy(p, &v, 8); p += 8; };
    put(A->h); put(A->w); put(A->nnz);
    put(A->h + 1);
    uint64_t *rp = (uint64_t *)p; p += 8 * (A->h + 1);
    put(A->nnz);
    int32_t *col = (int32_t *)p; p += 4 * A->nnz;
    put(A->nnz);
    // the three array sections are not 8-byte aligned in general: download into aligned scratch only when needed
    if (((uintptr_t)p & 7) == 0 && ((uintptr_t)col & 3) == 0 && ((uintptr_t)rp & 7) == 0) return ngsb_csr_download(A, rp, col, p);
    std::vector<uint64_t> trp(A->h + 1);
    std::vector<int32_t> tcol(A->nnz);
    std::vector<double> tval(A->nnz * ms);
    NGSB_TRY(ngsb_csr_download(A, trp.data(), tcol.data(), tval.data()));
    memcpy(rp, trp.data(), 8 * (A->h + 1));
    memcpy(col, tcol.data(), 4 * A->nnz);
    memcpy(p, tval.data(), 8 * ms * A->nnz);
    return NGSB_OK;
}

extern "C" int ngsb_csr_create_from_archive(ngsb_ctx *ctx, const void *buf, size_t bytes, int kind, ngsb_csr **out)
{
    NGSB_REQUIRE(ctx && buf && out, "ngsb_csr_create_from_archive: NULL argument");
    NGSB_REQUIRE(kind_valid(kind), "ngsb_csr_create_from_archive: bad kind %d", kind);
    NGSB_REQUIRE(bytes >= 48, "ngsb_csr_create_from_archive: truncated archive");
    const char *p = (const char *)buf;
    auto get = [&]() { uint64_t v; memcpy(&v, p, 8); p += 8; return v; };
    const uint64_t h = get(), w = get(), nnz = get(), n1 = get();
    const size_t ms = kind_matscalars(kind);
    NGSB_REQUIRE(n1 == h + 1 && bytes == archive_bytes(h, nnz, ms), "ngsb_csr_create_from_archive: %zu bytes do not hold a %llu x %llu matrix with %llu entries of kind %d",
                 bytes, (unsigned long long)h, (unsigned long long)w, (unsigned long long)nnz, kind);
    std::vector<uint64_t> rp(h + 1);
    memcpy(rp.data(), p, 8 * (h + 1)); p += 8 * (h + 1);
    NGSB_REQUIRE(get() == nnz, "ngsb_csr_create_from_archive: column count does not match nze");
    std::vector<int32_t> col(nnz);
    memcpy(col.data(), p, 4 * nnz); p += 4 * nnz;
    NGSB_REQUIRE(get() == nnz, "ngsb_csr_create_from_archive: value count does not match nze");
    std::vector<double> val(nnz * ms);
    memcpy(val.data(), p, 8 * ms * nnz);
    return ngsb_csr_create(ctx, h, w, nnz, rp.data(), col.data(), val.data(), kind, out);
}
