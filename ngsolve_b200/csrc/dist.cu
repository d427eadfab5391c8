// dist.cu -- one GPU per process: ParallelDofs exchange tables, Cumulate, ParallelMatrix
// (C2D) and the distributed Jacobi-PCG.
//
// Follows the reference's MPI split: every rank holds the dofs of its
// sub-domain including duplicated interface dofs; exchangedofs[p] lists the local dofs
// shared with rank p in ascending order, the lowest sharing rank is the master
// (linalg/paralleldofs.cpp:46-66).  A vector is DISTRIBUTED (true value = sum over
// sharers) or CUMULATED (every sharer holds the sum); Cumulate = neighbour exchange + add
// (parallel/parallelvvector.cpp:247-272, 536-549); ParallelMatrix::MultAdd (C2D) = local
// SpMV on a cumulated input giving a distributed output (parallel/parallel_matrices.cpp:
// 519-536); inner products are local (masked by master dofs when both operands are
// cumulated) + all-reduce (parallelvvector.cpp:289-331).
//
// The reference moves the interface values with MPI_Isend/Irecv on indexed datatypes from
// host memory.  Here the interface values are packed by a kernel into one send buffer,
// moved GPU-to-GPU over NVLink (ncclSend/ncclRecv in one group, on the context's stream),
// and added by one unpack kernel that visits every interface dof once and adds the
// received copies in ascending rank order (deterministic, unlike the reference's
// WaitAny order).  NCCL is bound at run time (dlopen) so that the single-GPU path of the
// library has no NCCL dependency.
#include "krylov.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>

struct ngsb_comm {
    ngsb_ctx *ctx = nullptr;
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    double *d_red = nullptr;      // 2 doubles: all-reduce slot
};

struct ngsb_parmat {
    ngsb_comm *comm = nullptr;
    const ngsb_csr *local = nullptr;
    size_t n = 0;                 // local dofs
    int es = 1;                   // scalars per entry
    std::vector<int> peers;       // neighbour ranks (ascending)
    std::vector<size_t> peer_off; // offset (in dofs) of each neighbour's slice in the packed lists
    size_t nex = 0;               // total exchange entries (sum over neighbours)
    int32_t *d_exdofs = nullptr;  // nex local dof indices, neighbour-major
    double *d_send = nullptr, *d_recv = nullptr;   // nex * es doubles each
    // dof-major view for the deterministic add: interface dof k (nif of them) has the
    // received copies d_recv[if_pos[if_first[k] .. if_first[k+1])] in ascending rank order
    size_t nif = 0;
    int32_t *d_if_dof = nullptr;
    uint32_t *d_if_first = nullptr;
    uint32_t *d_if_pos = nullptr;
    uint8_t *d_master = nullptr;  // n bytes
    std::vector<uint8_t> h_master;
};

namespace ngsb {

// ---- NCCL binding -------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.handle) return NGSB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { set_error("NCCL not found: %s", dlerror()); return NGSB_ERR_COMM; }
#define NGSB_SYM(field, name)                                                           \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                         \
    if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return NGSB_ERR_COMM; }
    NGSB_SYM(GetUniqueId, "ncclGetUniqueId")
    NGSB_SYM(CommInitRank, "ncclCommInitRank")
    NGSB_SYM(CommDestroy, "ncclCommDestroy")
    NGSB_SYM(AllReduce, "ncclAllReduce")
    NGSB_SYM(Send, "ncclSend")
    NGSB_SYM(Recv, "ncclRecv")
    NGSB_SYM(GroupStart, "ncclGroupStart")
    NGSB_SYM(GroupEnd, "ncclGroupEnd")
    NGSB_SYM(GetErrorString, "ncclGetErrorString")
#undef NGSB_SYM
    g_nccl.handle = h;
    return NGSB_OK;
}

#define NGSB_NCCL(call)                                                                          \
    do {                                                                                         \
        ncclResult_t r__ = (call);                                                               \
        if (r__ != ncclSuccess) {                                                                \
            ngsb::set_error("%s failed: %s", #call, ngsb::g_nccl.GetErrorString(r__));           \
            return NGSB_ERR_COMM;                                                                \
        }                                                                                        \
    } while (0)

// ---- kernels --------------------------------------------------------------------------------
template <int ES>
__global__ void __launch_bounds__(256) pack_kernel(const double *__restrict__ v, const int32_t *__restrict__ exdofs,
                                                  double *__restrict__ send, size_t nex)
{
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nex) return;
    const size_t dof = (size_t)exdofs[k];
#pragma unroll
    for (int c = 0; c < ES; c++) send[ES * k + c] = v[ES * dof + c];
}

// AddRecvValues for all neighbours at once, one thread per interface dof
template <int ES>
__global__ void __launch_bounds__(256) unpack_add_kernel(double *__restrict__ v, const int32_t *__restrict__ if_dof,
                                                        const uint32_t *__restrict__ if_first, const uint32_t *__restrict__ if_pos,
                                                        const double *__restrict__ recv, size_t nif)
{
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nif) return;
    const size_t dof = (size_t)if_dof[k];
    double acc[ES];
#pragma unroll
    for (int c = 0; c < ES; c++) acc[c] = v[ES * dof + c];
    for (uint32_t q = if_first[k]; q < if_first[k + 1]; q++) {
        const size_t p = if_pos[q];
#pragma unroll
        for (int c = 0; c < ES; c++) acc[c] += recv[ES * p + c];
    }
#pragma unroll
    for (int c = 0; c < ES; c++) v[ES * dof + c] = acc[c];
}

// run-time entry size versions (setup only: the 9-double diagonal blocks of the Jacobi ctor)
__global__ void pack_rt_kernel(const double *__restrict__ v, const int32_t *__restrict__ exdofs, double *__restrict__ send, size_t nex, int es)
{
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nex) return;
    const size_t dof = (size_t)exdofs[k];
    for (int c = 0; c < es; c++) send[es * k + c] = v[es * dof + c];
}

__global__ void unpack_add_rt_kernel(double *__restrict__ v, const int32_t *__restrict__ if_dof, const uint32_t *__restrict__ if_first,
                                     const uint32_t *__restrict__ if_pos, const double *__restrict__ recv, size_t nif, int es)
{
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nif) return;
    const size_t dof = (size_t)if_dof[k];
    for (int c = 0; c < es; c++) {
        double acc = v[es * dof + c];
        for (uint32_t q = if_first[k]; q < if_first[k + 1]; q++) acc += recv[es * (size_t)if_pos[q] + c];
        v[es * dof + c] = acc;
    }
}

static int all_reduce2(ngsb_comm *comm, double *d_buf)
{
    if (comm->nranks == 1) return NGSB_OK;
    comm->ctx->launches++;
    NGSB_NCCL(g_nccl.AllReduce(d_buf, d_buf, 2, ncclDouble, ncclSum, comm->comm, comm->ctx->stream));
    return NGSB_OK;
}

static int cumulate_raw(const ngsb_parmat *P, double *v)
{
    ngsb_comm *comm = P->comm;
    ngsb_ctx *ctx = comm->ctx;
    if (P->nex == 0 || comm->nranks == 1) return NGSB_OK;
    const int es = P->es;
    {
        SpanGuard g(ctx, KC_OTHER);
        unsigned grid = (unsigned)((P->nex + 255) / 256);
        if (es == 1) pack_kernel<1><<<grid, 256, 0, ctx->stream>>>(v, P->d_exdofs, P->d_send, P->nex);
        else if (es == 2) pack_kernel<2><<<grid, 256, 0, ctx->stream>>>(v, P->d_exdofs, P->d_send, P->nex);
        else pack_kernel<3><<<grid, 256, 0, ctx->stream>>>(v, P->d_exdofs, P->d_send, P->nex);
        NGSB_CUDA(cudaGetLastError());
    }
    ctx->launches++;
    NGSB_NCCL(g_nccl.GroupStart());
    for (size_t q = 0; q < P->peers.size(); q++) {
        const size_t off = P->peer_off[q] * es, cnt = (P->peer_off[q + 1] - P->peer_off[q]) * es;
        NGSB_NCCL(g_nccl.Send(P->d_send + off, cnt, ncclDouble, P->peers[q], comm->comm, ctx->stream));
        NGSB_NCCL(g_nccl.Recv(P->d_recv + off, cnt, ncclDouble, P->peers[q], comm->comm, ctx->stream));
    }
    NGSB_NCCL(g_nccl.GroupEnd());
    {
        SpanGuard g(ctx, KC_OTHER);
        unsigned grid = (unsigned)((P->nif + 255) / 256);
        if (es == 1) unpack_add_kernel<1><<<grid, 256, 0, ctx->stream>>>(v, P->d_if_dof, P->d_if_first, P->d_if_pos, P->d_recv, P->nif);
        else if (es == 2) unpack_add_kernel<2><<<grid, 256, 0, ctx->stream>>>(v, P->d_if_dof, P->d_if_first, P->d_if_pos, P->d_recv, P->nif);
        else unpack_add_kernel<3><<<grid, 256, 0, ctx->stream>>>(v, P->d_if_dof, P->d_if_first, P->d_if_pos, P->d_recv, P->nif);
        NGSB_CUDA(cudaGetLastError());
    }
    return NGSB_OK;
}

__global__ void mask_zero_kernel(double *v, const uint8_t *master, size_t n, int es)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (!master[i])
            for (int c = 0; c < es; c++) v[es * i + c] = 0.0;
}
int launch_mask_zero(ngsb_ctx *ctx, double *v, const uint8_t *master, size_t n, int es)
{
    if (n == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_VEC);
    size_t blocks = std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
    mask_zero_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(v, master, n, es);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

// Cumulate of an array with `es` doubles per dof through temporary buffers (setup path)
static int cumulate_any(void *arg, double *v, int es)
{
    const ngsb_parmat *P = (const ngsb_parmat *)arg;
    ngsb_comm *comm = P->comm;
    ngsb_ctx *ctx = comm->ctx;
    if (P->nex == 0 || comm->nranks == 1) return NGSB_OK;
    double *send = nullptr, *recv = nullptr;
    NGSB_CUDA(cudaMalloc(&send, P->nex * es * sizeof(double)));
    NGSB_CUDA(cudaMalloc(&recv, P->nex * es * sizeof(double)));
    pack_rt_kernel<<<(unsigned)((P->nex + 255) / 256), 256, 0, ctx->stream>>>(v, P->d_exdofs, send, P->nex, es);
    ctx->launches += 3;
    int rc = NGSB_OK;
    do {
        ncclResult_t r = g_nccl.GroupStart();
        for (size_t q = 0; q < P->peers.size() && r == ncclSuccess; q++) {
            const size_t off = P->peer_off[q] * es, cnt = (P->peer_off[q + 1] - P->peer_off[q]) * es;
            r = g_nccl.Send(send + off, cnt, ncclDouble, P->peers[q], comm->comm, ctx->stream);
            if (r == ncclSuccess) r = g_nccl.Recv(recv + off, cnt, ncclDouble, P->peers[q], comm->comm, ctx->stream);
        }
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) { set_error("Cumulate(diagonal): %s", g_nccl.GetErrorString(r)); rc = NGSB_ERR_COMM; }
    } while (0);
    if (rc == NGSB_OK) {
        unpack_add_rt_kernel<<<(unsigned)((P->nif + 255) / 256), 256, 0, ctx->stream>>>(v, P->d_if_dof, P->d_if_first, P->d_if_pos, recv, P->nif, es);
        if (cudaGetLastError() != cudaSuccess) rc = NGSB_ERR_CUDA;
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(send);
    cudaFree(recv);
    return rc;
}

} // namespace ngsb

using namespace ngsb;

// JacobiPrecond on a ParallelMatrix: the diagonal is summed over the sharing ranks before it is
// inverted (paralleldofs->AllReduceDofData(invdiag, SUM), linalg/jacobi.cpp:60-61)
extern "C" int ngsb_parmat_jacobi_create(const ngsb_parmat *P, const uint8_t *freebits, ngsb_jacobi **out)
{
    NGSB_REQUIRE(P && out, "ngsb_parmat_jacobi_create: NULL argument");
    return jacobi_build(P->local, freebits, cumulate_any, (void *)P, out);
}

extern "C" int ngsb_comm_unique_id(void *uid128)
{
    NGSB_REQUIRE(uid128, "ngsb_comm_unique_id: NULL argument");
    NGSB_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NGSB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(uid128, &id, sizeof(id));
    return NGSB_OK;
}

extern "C" int ngsb_comm_create(ngsb_ctx *ctx, int nranks, int rank, const void *uid128, ngsb_comm **out)
{
    NGSB_REQUIRE(ctx && out, "ngsb_comm_create: NULL argument");
    NGSB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "ngsb_comm_create: bad rank %d of %d", rank, nranks);
    NGSB_CUDA(cudaSetDevice(ctx->device));
    ngsb_comm *c = new ngsb_comm();
    c->ctx = ctx;
    c->nranks = nranks;
    c->rank = rank;
    if (nranks > 1) {
        NGSB_REQUIRE(uid128, "ngsb_comm_create: uid is NULL");
        NGSB_TRY(nccl_load());
        ncclUniqueId id;
        memcpy(&id, uid128, sizeof(id));
        NGSB_NCCL(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    }
    NGSB_CUDA(cudaMalloc(&c->d_red, 4 * sizeof(double)));
    *out = c;
    return NGSB_OK;
}

extern "C" int ngsb_comm_destroy(ngsb_comm *c)
{
    if (!c) return NGSB_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->d_red);
    delete c;
    return NGSB_OK;
}

extern "C" int ngsb_parmat_create(ngsb_comm *comm, const ngsb_csr *local, const uint64_t *ex_first, const int32_t *ex_dofs,
                                  ngsb_parmat **out)
{
    NGSB_REQUIRE(comm && local && ex_first && out, "ngsb_parmat_create: NULL argument");
    NGSB_REQUIRE(local->ctx == comm->ctx, "ngsb_parmat_create: matrix and communicator belong to different contexts");
    NGSB_REQUIRE(local->h == local->w, "ngsb_parmat_create: local matrix must be square");
    ngsb_ctx *ctx = comm->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const int np = comm->nranks;
    const size_t n = local->h;
    NGSB_REQUIRE(ex_first[0] == 0, "ngsb_parmat_create: ex_first[0] must be 0");
    NGSB_REQUIRE(ex_first[comm->rank + 1] == ex_first[comm->rank], "ngsb_parmat_create: a rank does not exchange with itself");
    const size_t nex = ex_first[np];
    NGSB_REQUIRE(nex == 0 || ex_dofs, "ngsb_parmat_create: ex_dofs is NULL");
    NGSB_REQUIRE(nex < (1ull << 32), "ngsb_parmat_create: too many exchange dofs");
    for (int p = 0; p < np; p++) {
        NGSB_REQUIRE(ex_first[p] <= ex_first[p + 1], "ngsb_parmat_create: ex_first not monotone");
        for (uint64_t k = ex_first[p]; k < ex_first[p + 1]; k++) {
            NGSB_REQUIRE(ex_dofs[k] >= 0 && (size_t)ex_dofs[k] < n, "ngsb_parmat_create: exchange dof out of range");
            // ascending local order per neighbour: the contract both sides pair entries by
            NGSB_REQUIRE(k == ex_first[p] || ex_dofs[k] > ex_dofs[k - 1], "ngsb_parmat_create: exchangedofs[%d] not ascending", p);
        }
    }
    ngsb_parmat *P = new ngsb_parmat();
    P->comm = comm;
    P->local = local;
    P->n = n;
    P->es = (int)kind_scalars(local->kind);
    P->nex = nex;
    P->peer_off.push_back(0);
    for (int p = 0; p < np; p++)
        if (ex_first[p + 1] > ex_first[p]) {
            P->peers.push_back(p);
            P->peer_off.push_back(ex_first[p + 1]);
        }
    // the packed lists are neighbour-major in ascending rank order == ex_dofs itself
    // master dofs: cleared for dofs shared with any lower rank (paralleldofs.cpp:61-66)
    P->h_master.assign(n, 1);
    for (int p = 0; p < comm->rank; p++)
        for (uint64_t k = ex_first[p]; k < ex_first[p + 1]; k++) P->h_master[ex_dofs[k]] = 0;
    // dof-major view
    std::vector<std::pair<int32_t, uint32_t>> pairs(nex);
    for (size_t k = 0; k < nex; k++) pairs[k] = std::make_pair(ex_dofs[k], (uint32_t)k);
    std::sort(pairs.begin(), pairs.end());   // by dof, then by position == by rank
    std::vector<int32_t> if_dof;
    std::vector<uint32_t> if_first, if_pos(nex);
    for (size_t k = 0; k < nex; k++) {
        if (k == 0 || pairs[k].first != pairs[k - 1].first) { if_dof.push_back(pairs[k].first); if_first.push_back((uint32_t)k); }
        if_pos[k] = pairs[k].second;
    }
    if_first.push_back((uint32_t)nex);
    P->nif = if_dof.size();
    auto up = [&](void **dst, const void *src, size_t bytes) -> int {
        NGSB_CUDA(cudaMalloc(dst, bytes ? bytes : 16));
        if (bytes) NGSB_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return NGSB_OK;
    };
    int rc = NGSB_OK;
    if (rc == NGSB_OK) rc = up((void **)&P->d_exdofs, ex_dofs, nex * sizeof(int32_t));
    if (rc == NGSB_OK) rc = up((void **)&P->d_if_dof, if_dof.data(), if_dof.size() * sizeof(int32_t));
    if (rc == NGSB_OK) rc = up((void **)&P->d_if_first, if_first.data(), if_first.size() * sizeof(uint32_t));
    if (rc == NGSB_OK) rc = up((void **)&P->d_if_pos, if_pos.data(), if_pos.size() * sizeof(uint32_t));
    if (rc == NGSB_OK) rc = up((void **)&P->d_master, P->h_master.data(), n);
    if (rc == NGSB_OK && cudaMalloc(&P->d_send, std::max<size_t>(16, nex * P->es * sizeof(double))) != cudaSuccess) rc = NGSB_ERR_NOMEM;
    if (rc == NGSB_OK && cudaMalloc(&P->d_recv, std::max<size_t>(16, nex * P->es * sizeof(double))) != cudaSuccess) rc = NGSB_ERR_NOMEM;
    cudaStreamSynchronize(ctx->stream);
    if (rc != NGSB_OK) { ngsb_parmat_destroy(P); return rc; }
    *out = P;
    return NGSB_OK;
}

extern "C" int ngsb_parmat_destroy(ngsb_parmat *P)
{
    if (!P) return NGSB_OK;
    cudaSetDevice(P->comm->ctx->device);
    cudaStreamSynchronize(P->comm->ctx->stream);
    cudaFree(P->d_exdofs); cudaFree(P->d_send); cudaFree(P->d_recv);
    cudaFree(P->d_if_dof); cudaFree(P->d_if_first); cudaFree(P->d_if_pos); cudaFree(P->d_master);
    delete P;
    return NGSB_OK;
}

extern "C" int ngsb_parmat_masterdofs(const ngsb_parmat *P, uint8_t *ismaster)
{
    NGSB_REQUIRE(P && ismaster, "ngsb_parmat_masterdofs: NULL argument");
    memcpy(ismaster, P->h_master.data(), P->n);
    return NGSB_OK;
}

static int check_pvec(const ngsb_parmat *P, const ngsb_vec *v, const char *who)
{
    NGSB_REQUIRE(P && v, "%s: NULL argument", who);
    NGSB_REQUIRE(v->ctx == P->comm->ctx, "%s: vector belongs to a different context", who);
    NGSB_REQUIRE(v->n == P->n && v->kind == P->local->kind, "%s: vector does not match the parallel matrix", who);
    return NGSB_OK;
}

extern "C" int ngsb_parmat_cumulate(const ngsb_parmat *P, ngsb_vec *v)
{
    NGSB_TRY(check_pvec(P, v, "ParallelBaseVector::Cumulate"));
    NGSB_CUDA(cudaSetDevice(P->comm->ctx->device));
    return cumulate_raw(P, v->d);
}

extern "C" int ngsb_parmat_mult(const ngsb_parmat *P, const ngsb_vec *x, ngsb_vec *y)
{
    NGSB_TRY(check_pvec(P, x, "ParallelMatrix::Mult"));
    NGSB_TRY(check_pvec(P, y, "ParallelMatrix::Mult"));
    return ngsb_csr_mult(P->local, x, y);
}

extern "C" int ngsb_parmat_dot(const ngsb_parmat *P, const ngsb_vec *x, const ngsb_vec *y, int both_cumulated, double *out)
{
    NGSB_TRY(check_pvec(P, x, "ParallelBaseVector::InnerProduct"));
    NGSB_TRY(check_pvec(P, y, "ParallelBaseVector::InnerProduct"));
    NGSB_REQUIRE(out, "ngsb_parmat_dot: out is NULL");
    NGSB_REQUIRE(P->local->kind != NGSB_COMPLEX, "ngsb_parmat_dot: complex not supported here");
    ngsb_comm *comm = P->comm;
    ngsb_ctx *ctx = comm->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    double *tmp = nullptr;
    if (both_cumulated) {
        // both CUMULATED: the reference masks by master dofs (entry size 1) or Distribute()s one
        // operand (parallelvvector.cpp:302-322); both equal "zero the non-master copies of x"
        NGSB_TRY(ws_get_buf(ctx, x->nscal + 2, &tmp));
        NGSB_CUDA(cudaMemcpyAsync(tmp, x->d, x->nscal * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        int rc0 = launch_mask_zero(ctx, tmp, P->d_master, P->n, P->es);
        if (rc0 != NGSB_OK) { ws_put_buf(ctx, x->nscal + 2, tmp); return rc0; }
    }
    int rc = launch_dot(ctx, both_cumulated ? tmp : x->d, y->d, x->nscal, 0, comm->d_red);
    if (rc == NGSB_OK) rc = all_reduce2(comm, comm->d_red);
    if (rc == NGSB_OK) {
        cudaError_t e = cudaMemcpyAsync(ctx->h_pinned, comm->d_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("ngsb_parmat_dot: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; }
        else *out = ctx->h_pinned[0];
    }
    if (tmp) ws_put_buf(ctx, x->nscal + 2, tmp);
    return rc;
}


// Distributed Jacobi-PCG, the reference's sequence of statuses:
//   d = f (DISTRIBUTED) -> Jacobi cumulates d; w, s CUMULATED; wdn = masked <w,d> + all-reduce
//   loop: as = A s (DISTRIBUTED); kss = <s,as> local + all-reduce; u += al s;
//         d -= al as cumulates as (the one neighbour exchange of the iteration);
//         w = C d; wdn = masked <d,w> + all-reduce; s = be s + w.
extern "C" int ngsb_parmat_cg_solve(const ngsb_parmat *P, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *u, double prec,
                                    int maxsteps, int *steps, double *history, int hist_cap, int *nhist)
{
    NGSB_TRY(check_pvec(P, f, "CGSolver::Mult(parallel)"));
    NGSB_TRY(check_pvec(P, u, "CGSolver::Mult(parallel)"));
    const ngsb_csr *A = P->local;
    NGSB_REQUIRE(A->kind != NGSB_COMPLEX, "ngsb_parmat_cg_solve: complex systems are not supported by the distributed CG");
    NGSB_REQUIRE(!C || (C->n == A->h && C->kind == A->kind && C->ctx == A->ctx), "ngsb_parmat_cg_solve: preconditioner does not match");
    NGSB_REQUIRE(maxsteps >= 0 && f->d != u->d, "ngsb_parmat_cg_solve: bad arguments");
    ngsb_comm *comm = P->comm;
    ngsb_ctx *ctx = comm->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const size_t nscal = A->h * kind_scalars(A->kind);
    if (hist_cap < 0 || !history) hist_cap = 0;
    CgState *d_state = nullptr, *hs = nullptr;
    double *d_hist = nullptr;
    NGSB_TRY(ws_state(ctx, (size_t)hist_cap, &d_state, &hs, &d_hist));
    double *w = nullptr, *s = nullptr, *d = nullptr, *as = nullptr;
    NGSB_TRY(ws_get_buf(ctx, nscal, &s));
    NGSB_TRY(ws_get_buf(ctx, nscal, &d));
    NGSB_TRY(ws_get_buf(ctx, nscal, &as));
    if (C) NGSB_TRY(ws_get_buf(ctx, nscal, &w));

    memset(hs, 0, sizeof(CgState));
    hs->prec2 = prec * prec;
    hs->maxsteps = maxsteps;
    hs->hist_cap = hist_cap;
    hs->cplx = 0;
    int rc = NGSB_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == NGSB_OK) { set_error("parallel CG: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; } };
    cu(cudaMemcpyAsync(d_state, hs, sizeof(CgState), cudaMemcpyHostToDevice, ctx->stream));

    CgVecs v;
    memset(&v, 0, sizeof(v));
    v.u = u->d; v.d = d; v.w = w; v.s = s; v.as = as; v.f = f->d;
    v.invdiag = C ? C->d_invdiag : nullptr;
    v.bits = C ? C->d_bits : nullptr;
    v.master = P->d_master;
    v.dot_out = comm->d_red;
    v.n = A->h;
    v.state = d_state;
    v.hist = d_hist;
    v.partials = ctx->d_partials;
    v.counter = ctx->d_counter;
    v.ip_mode = NGSB_IP_REAL;

    // u = 0; d = f, cumulated by the Jacobi application (linalg/jacobi.cpp:78)
    cu(cudaMemsetAsync(u->d, 0, nscal * sizeof(double), ctx->stream));
    cu(cudaMemcpyAsync(d, f->d, nscal * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (rc == NGSB_OK) rc = cumulate_raw(P, d);
    v.f = d;                       // init kernel reads f, writes d: same values
    if (rc == NGSB_OK) rc = cg_launch_fused(ctx, A->kind, 0, v, 0);
    if (rc == NGSB_OK) rc = all_reduce2(comm, comm->d_red);
    if (rc == NGSB_OK) rc = cg_launch_finalize(ctx, 0, d_state, comm->d_red, d_hist);

    const long batch = ctx->cg_batch;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cu(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    cu(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    const long max_batches = ((long)maxsteps + batch - 1) / batch + 1;
    long enq = 0;
    bool finished = false;
    while (rc == NGSB_OK && !finished) {
        if (enq < max_batches) {
            for (long k = 0; k < batch && rc == NGSB_OK; k++) {
                SpmvArgs a;
                memset(&a, 0, sizeof(a));
                a.A = A; a.x = s; a.y = as; a.sr = 1.0; a.accumulate = false;
                a.epi = EPI_DOT_OUT; a.dotvec = s; a.dot_out = comm->d_red; a.state = d_state;
                rc = spmv_launch(a);                                                   // as = A s, local <s,as>
                if (rc == NGSB_OK) rc = all_reduce2(comm, comm->d_red);
                if (rc == NGSB_OK) rc = cg_launch_finalize(ctx, 1, d_state, comm->d_red, d_hist);
                if (rc == NGSB_OK) rc = cumulate_raw(P, as);                            // as -> CUMULATED
                if (rc == NGSB_OK) rc = cg_launch_fused(ctx, A->kind, 1, v, 0);        // u, d, w, masked <d,w>
                if (rc == NGSB_OK) rc = all_reduce2(comm, comm->d_red);
                if (rc == NGSB_OK) rc = cg_launch_finalize(ctx, 2, d_state, comm->d_red, d_hist);
                if (rc == NGSB_OK) rc = cg_launch_dir(ctx, A->kind, v);                // s = be s + w
            }
            if (rc != NGSB_OK) break;
        }
        cu(cudaMemcpyAsync(&hs[1 + (enq & 1)], d_state, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream));
        cu(cudaEventRecord(ev[enq & 1], ctx->stream));
        if (enq > 0) {
            cu(cudaEventSynchronize(ev[(enq - 1) & 1]));
            if (hs[1 + ((enq - 1) & 1)].done) finished = true;
        }
        enq++;
        if (!finished && enq > max_batches + 1) { set_error("parallel CG: device loop did not terminate"); rc = NGSB_ERR_CUDA; }
    }
    cu(cudaStreamSynchronize(ctx->stream));
    if (ev[0]) cudaEventDestroy(ev[0]);
    if (ev[1]) cudaEventDestroy(ev[1]);
    if (rc == NGSB_OK) {
        cu(cudaMemcpyAsync(&hs[3], d_state, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream));
        cu(cudaStreamSynchronize(ctx->stream));
        if (steps) *steps = hs[3].n;
        if (nhist) *nhist = hs[3].nhist;
        int ncopy = hs[3].nhist < hist_cap ? hs[3].nhist : hist_cap;
        if (history && ncopy > 0) {
            cu(cudaMemcpyAsync(history, d_hist, ncopy * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            cu(cudaStreamSynchronize(ctx->stream));
        }
    }
    ws_put_buf(ctx, nscal, s);
    ws_put_buf(ctx, nscal, d);
    ws_put_buf(ctx, nscal, as);
    if (w) ws_put_buf(ctx, nscal, w);
    return rc;
}
