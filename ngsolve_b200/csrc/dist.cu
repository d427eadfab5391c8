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
#include "peer.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>

struct ngsb_comm {
    ngsb_ctx *ctx = nullptr;
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;    // NULL: no NCCL (bootstrap by callback, peer-memory data path only)
    int (*boot_allgather)(void *, const void *, void *, size_t) = nullptr;
    void *boot_user = nullptr;
    double *d_red = nullptr;      // 2 doubles: local dot / all-reduce slot
    // peer-memory mode (peer.cuh)
    bool p2p = false;
    char *mailbox = nullptr;                      // local, exported with CUDA IPC
    char *peer_mailbox[NGSB_MAX_RANKS] = {};      // mapped mailboxes of the other ranks
    PeerReduce *d_R = nullptr;                    // device copy of the reduce descriptor
    int *d_err = nullptr;                         // inside the mailbox
};

struct ngsb_parmat {
    ngsb_comm *comm = nullptr;
    const ngsb_csr *local = nullptr;
    size_t n = 0;                 // local dofs
    int es = 1;                   // scalars per entry
    int esmax = 1;                // scalars per matrix value: what the Jacobi constructor cumulates
    std::vector<int> peers;       // neighbour ranks (ascending)
    std::vector<size_t> peer_off; // offset (in dofs) of each neighbour's slice in the packed lists
    size_t nex = 0;               // total exchange entries (sum over neighbours)
    int32_t *d_exdofs = nullptr;  // nex local dof indices, neighbour-major
    double *d_send = nullptr, *d_recv = nullptr;   // nex * es doubles each (NCCL mode)
    // dof-major view for the deterministic add: interface dof k (nif of them) has the
    // received copies d_recv[if_pos[if_first[k] .. if_first[k+1])] in ascending rank order
    size_t nif = 0;
    int32_t *d_if_dof = nullptr;
    uint32_t *d_if_first = nullptr;
    uint32_t *d_if_pos = nullptr;
    uint32_t *d_if_nlow = nullptr;    // per interface dof: how many of its sharers have a lower rank than this one
    uint8_t *d_master = nullptr;  // n bytes
    std::vector<uint8_t> h_master;
    // peer-memory mode
    bool p2p = false;
    char *halo_mem = nullptr;                     // local, exported: flags | receive area (2 parities)
    char *peer_halo[NGSB_MAX_RANKS] = {};         // mapped, per neighbour index
    char *d_local = nullptr;                      // local only: sequence number + CTA counters
    PeerHalo *d_H = nullptr;
    // interface-first split of the local product (option "dist_overlap", peer-memory mode): scheduled SELL slices that
    // hold at least one interface row / all the others, both in ascending schedule order
    bool overlap = false;
    uint32_t *d_list_bnd = nullptr, *d_list_int = nullptr;
    uint32_t n_bnd = 0, n_int = 0;
    double *d_red_b = nullptr;                    // <s, A s> over the interface slices (2 doubles)
    // product + exchange in one kernel (option "dist_fused_push", default on in peer-memory mode): per scheduled slice the
    // record of its interface rows (sell.cu, SellPush)
    bool fused_push = false;
    uint32_t *d_slice_if = nullptr;
    int32_t *d_lane_if = nullptr;
    ngsb::SellPush *d_push_desc = nullptr;        // device copy of the descriptor
    uint32_t *d_slice_src_flagged = nullptr;      // the local matrix' schedule, bit 31 = slice holds interface rows
    // cached CUDA graph of one batch of CG iterations (peer-memory mode)
    cudaGraphExec_t graph_exec = nullptr;
    const void *g_key[9] = {};
    long g_batch = 0;
    uint64_t g_launches = 0;                      // kernels in one batch of the captured graph
};

namespace ngsb {

// mailbox layout (bytes)
static const size_t MB_SLOTS = 0;                 // PeerSlot[2][NGSB_MAX_RANKS]
static const size_t MB_SEQ = 4096;                // unsigned long long
static const size_t MB_ERR = 4096 + 64;           // int
static const size_t MB_MAGIC = 8192;              // unsigned long long, checked through the mapping
static const size_t MB_VEC = 65536;               // double[2][NGSB_MAX_RANKS][NGSB_PEER_VEC_LEN]: vector all-reduce area (1 MiB)
static_assert(MB_VEC + 2 * NGSB_MAX_RANKS * NGSB_PEER_VEC_LEN * sizeof(double) <= (2u << 20), "vector area must fit the mailbox");
static const size_t MB_BYTES = 2u << 20;          // own 2 MiB block: the IPC handle maps exactly this allocation
// halo block layout
static const size_t HB_FLAGS = 0;                 // unsigned long long[2][NGSB_MAX_RANKS]
static const size_t HB_MAGIC = 1024;
static const size_t HB_DATA = 4096;

// ---- NCCL binding -------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    NGSB_SYM(AllGather, "ncclAllGather")
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

// out-of-band all-gather of `bytes` per rank (setup only): the caller's callback, else NCCL
static int boot_allgather(ngsb_comm *c, const void *send, void *recv, size_t bytes)
{
    if (c->nranks == 1) { memcpy(recv, send, bytes); return NGSB_OK; }
    if (c->boot_allgather) {
        int rc = c->boot_allgather(c->boot_user, send, recv, bytes);
        if (rc != 0) { set_error("bootstrap all-gather callback failed (%d)", rc); return NGSB_ERR_COMM; }
        return NGSB_OK;
    }
    NGSB_REQUIRE(c->comm, "communicator has neither NCCL nor a bootstrap all-gather");
    ngsb_ctx *ctx = c->ctx;
    char *d_s = nullptr, *d_r = nullptr;
    NGSB_CUDA(cudaMalloc(&d_s, bytes));
    NGSB_CUDA(cudaMalloc(&d_r, bytes * c->nranks));
    NGSB_CUDA(cudaMemcpyAsync(d_s, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = g_nccl.AllGather(d_s, d_r, bytes, ncclChar, c->comm, ctx->stream);
    cudaError_t e = cudaMemcpyAsync(recv, d_r, bytes * c->nranks, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_s);
    cudaFree(d_r);
    if (r != ncclSuccess) { set_error("ncclAllGather (bootstrap) failed: %s", g_nccl.GetErrorString(r)); return NGSB_ERR_COMM; }
    if (e != cudaSuccess) { set_error("bootstrap all-gather: %s", cudaGetErrorString(e)); return NGSB_ERR_CUDA; }
    return NGSB_OK;
}

// all ranks agree: true only if `mine` is true everywhere
static int boot_all_ok(ngsb_comm *c, bool mine, bool *all)
{
    std::vector<unsigned char> buf(c->nranks);
    unsigned char m = mine ? 1 : 0;
    NGSB_TRY(boot_allgather(c, &m, buf.data(), 1));
    *all = true;
    for (int p = 0; p < c->nranks; p++) *all = *all && buf[p] != 0;
    return NGSB_OK;
}

// export `mem` (its own cudaMalloc block, magic already written at `magic_off`), gather every rank's handle,
// map the blocks of the ranks in `want` (NULL: all) and check the magic through the mapping.
// ok = false (no error) when this rank could not export / map; collective.
static int ipc_exchange(ngsb_comm *c, char *mem, size_t magic_off, const std::vector<int> *want, char **mapped, bool *ok)
{
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    bool good = mem != nullptr && cudaIpcGetMemHandle(&mine, mem) == cudaSuccess;
    if (!good) cudaGetLastError();
    std::vector<cudaIpcMemHandle_t> all(c->nranks);
    NGSB_TRY(boot_allgather(c, &mine, all.data(), sizeof(mine)));
    bool everyone = false;
    NGSB_TRY(boot_all_ok(c, good, &everyone));
    if (everyone) {
        for (int p = 0; p < c->nranks && good; p++) {
            if (p == c->rank) { mapped[p] = mem; continue; }
            if (want && std::find(want->begin(), want->end(), p) == want->end()) continue;
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; break; }
            mapped[p] = (char *)ptr;
            unsigned long long magic = 0;
            if (cudaMemcpy(&magic, mapped[p] + magic_off, sizeof(magic), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); good = false; break; }
            if (magic != (0x4e47534232303000ull | (unsigned long long)p)) good = false;
        }
        NGSB_TRY(boot_all_ok(c, good, &everyone));
    }
    if (!everyone)
        for (int p = 0; p < c->nranks; p++)
            if (p != c->rank && mapped[p]) { cudaIpcCloseMemHandle(mapped[p]); mapped[p] = nullptr; }
    *ok = everyone;
    return NGSB_OK;
}

static int write_magic(ngsb_ctx *ctx, char *mem, size_t off, int rank)
{
    unsigned long long magic = 0x4e47534232303000ull | (unsigned long long)rank;
    NGSB_CUDA(cudaMemcpy(mem + off, &magic, sizeof(magic), cudaMemcpyHostToDevice));
    return NGSB_OK;
}

// ---- kernels --------------------------------------------------------------------------------
// NCCL mode: gather the interface values into one send buffer ...
__global__ void __launch_bounds__(256) pack_kernel(const double *__restrict__ v, const int32_t *__restrict__ exdofs, double *__restrict__ send,
                                                  size_t nex, int es)
{
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nex) return;
    const size_t dof = (size_t)exdofs[k];
    for (int c = 0; c < es; c++) send[es * k + c] = v[es * dof + c];
}

// ... and AddRecvValues for all neighbours at once, one thread per interface dof, copies added in ascending rank order
// The copies of one dof are summed in ascending rank order WITH this rank's own value at its place in that order, so that
// every sharer computes the bit-identical sum (the reference gets the same guarantee for its Jacobi diagonal by reducing at
// the master, own value first, then the distant procs ascending, and scattering the result: linalg/paralleldofs.hpp:213-334).
__device__ __forceinline__ double sum_copies(double own, const double *__restrict__ recv, const uint32_t *__restrict__ if_pos, uint32_t a, uint32_t m,
                                             uint32_t b, int es, int c)
{
    double acc;
    if (m > a) {
        acc = __ldcg(recv + (size_t)es * if_pos[a] + c);
        for (uint32_t q = a + 1; q < m; q++) acc += __ldcg(recv + (size_t)es * if_pos[q] + c);
        acc += own;
    } else acc = own;
    for (uint32_t q = m; q < b; q++) acc += __ldcg(recv + (size_t)es * if_pos[q] + c);
    return acc;
}

__global__ void __launch_bounds__(256) unpack_add_kernel(double *__restrict__ v, const int32_t *__restrict__ if_dof,
                                                        const uint32_t *__restrict__ if_first, const uint32_t *__restrict__ if_pos,
                                                        const uint32_t *__restrict__ if_nlow, const double *__restrict__ recv, size_t nif, int es)
{
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nif) return;
    const size_t dof = (size_t)if_dof[k];
    const uint32_t a = if_first[k], b = if_first[k + 1], m = a + if_nlow[k];
    for (int c = 0; c < es; c++) v[es * dof + c] = sum_copies(v[es * dof + c], recv, if_pos, a, m, b, es, c);
}

// Peer-memory mode, first half of Cumulate: store my interface values into the neighbours' receive areas (P2P stores
// over NVLink), then the last CTA publishes the sequence number to every neighbour.  Thread (0,0) also publishes this
// rank's partial of the pending scalar reduction (`red`), so that it travels while the exchange is in flight.
__global__ void __launch_bounds__(256) halo_push_kernel(const PeerHalo *__restrict__ H, const double *__restrict__ v,
                                                       const int32_t *__restrict__ exdofs, size_t nex, int es, const CgState *st,
                                                       const PeerReduce *R, const double *red)
{
    if (st && st->done) return;
    const unsigned long long s = *(volatile unsigned long long *)H->seq + 1;
    const unsigned long long par = s & 1;
    if (R && blockIdx.x == 0 && threadIdx.x < 32) pr_push_warp(*R, red[0], red[1]);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < nex; k += stride) {
        int q = 0;
        while (k >= H->peer_off[q + 1]) q++;
        const size_t dof = (size_t)exdofs[k];
        double *dst = H->peer_recv[q] + par * H->peer_stride[q] + (H->peer_my_off[q] + (k - H->peer_off[q])) * (size_t)es;
        for (int c = 0; c < es; c++) dst[c] = v[es * dof + c];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&H->counter[0], 1u);
        if (t == gridDim.x - 1) {
            H->counter[0] = 0;
            __threadfence_system();
            for (int q = 0; q < H->npeers; q++) st_release_sys(H->peer_flags[q] + par * NGSB_MAX_RANKS + H->rank, s);
        }
    }
}

// second half: wait for the neighbours' flags, add the received copies (ascending rank order per dof).  The last CTA
// completes the exchange and, inside the CG loop (fin >= 1), finishes kss = <s, A s>: all-reduce + al = wd / kss.
// fin = 2 (interface-first split): the local dot is complete only after the interior product that ran behind the push,
// so this kernel publishes the partial (`red`) itself before it waits.
__global__ void __launch_bounds__(256) halo_unpack_kernel(const PeerHalo *__restrict__ H, double *__restrict__ v,
                                                         const int32_t *__restrict__ if_dof, const uint32_t *__restrict__ if_first,
                                                         const uint32_t *__restrict__ if_pos, const uint32_t *__restrict__ if_nlow, size_t nif, int es,
                                                         CgState *st, const PeerReduce *R, int fin, const double *red)
{
    if (st && st->done) return;
    const unsigned long long s = *(volatile unsigned long long *)H->seq + 1;
    const unsigned long long par = s & 1;
    if (fin == 2 && R && blockIdx.x == 0 && threadIdx.x < 32) pr_push_warp(*R, red[0], red[1]);
    if ((int)threadIdx.x < H->npeers) peer_wait(H->flags + par * NGSB_MAX_RANKS + H->peer_rank[threadIdx.x], s, H->err);
    __syncthreads();
    const double *recv = H->recv + par * H->stride;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < nif; k += stride) {
        const size_t dof = (size_t)if_dof[k];
        const uint32_t a = if_first[k], b = if_first[k + 1], m = a + if_nlow[k];
        for (int c = 0; c < es; c++) v[es * dof + c] = sum_copies(v[es * dof + c], recv, if_pos, a, m, b, es, c);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&H->counter[1], 1u);
        if (t == gridDim.x - 1) {
            H->counter[1] = 0;
            *(volatile unsigned long long *)H->seq = s;
            if (fin >= 1 && R && st) {
                double2 tot = pr_wait_sum(*R);
                cg_finalize_kss(st, tot);
            }
        }
    }
}

// buf(re,im) <- sum over ranks, one thread (API-level inner products)
__global__ void peer_allreduce_kernel(const PeerReduce *R, double *buf) { pr_allreduce_inplace(*R, buf); }

static int all_reduce2(ngsb_comm *comm, double *d_buf)
{
    if (comm->nranks == 1) return NGSB_OK;
    comm->ctx->launches++;
    if (comm->p2p) {
        peer_allreduce_kernel<<<1, 1, 0, comm->ctx->stream>>>(comm->d_R, d_buf);
        NGSB_CUDA(cudaGetLastError());
        return NGSB_OK;
    }
    NGSB_NCCL(g_nccl.AllReduce(d_buf, d_buf, 2, ncclDouble, ncclSum, comm->comm, comm->ctx->stream));
    return NGSB_OK;
}

static unsigned halo_grid(const ngsb_ctx *ctx, size_t n)
{
    size_t b = (n + 255) / 256, cap = (size_t)ctx->sm_count * 2;
    return (unsigned)std::max<size_t>(1, std::min(b, cap));
}

// Cumulate of a raw local array with `es` doubles per dof (es <= P->esmax).  Inside the CG loop `st`, `R`, `red`, `fin`
// fuse the kss reduction into the two kernels (peer-memory mode only).
static int cumulate_es(const ngsb_parmat *P, double *v, int es, CgState *st = nullptr, const double *red = nullptr, int fin = 0)
{
    ngsb_comm *comm = P->comm;
    ngsb_ctx *ctx = comm->ctx;
    if (comm->nranks == 1) return NGSB_OK;
    if (P->p2p) {
        // no early-out on nex == 0: a rank without neighbours still takes part in the fused reduction
        SpanGuard g(ctx, KC_OTHER);
        const PeerReduce *R = fin ? comm->d_R : nullptr;
        halo_push_kernel<<<halo_grid(ctx, P->nex), 256, 0, ctx->stream>>>(P->d_H, v, P->d_exdofs, P->nex, es, st, R, red);
        halo_unpack_kernel<<<halo_grid(ctx, P->nif), 256, 0, ctx->stream>>>(P->d_H, v, P->d_if_dof, P->d_if_first, P->d_if_pos, P->d_if_nlow, P->nif, es, st, R, fin, red);
        ctx->launches += 1;          // two kernels, one counted by the SpanGuard
        NGSB_CUDA(cudaGetLastError());
        return NGSB_OK;
    }
    if (P->nex == 0) return NGSB_OK;
    double *send = P->d_send, *recv = P->d_recv;
    const bool own = es > P->es;         // setup path (diagonal blocks): temporary buffers
    if (own) {
        NGSB_CUDA(cudaMalloc(&send, P->nex * es * sizeof(double)));
        NGSB_CUDA(cudaMalloc(&recv, P->nex * es * sizeof(double)));
    }
    int rc = NGSB_OK;
    {
        SpanGuard g(ctx, KC_OTHER);
        pack_kernel<<<(unsigned)((P->nex + 255) / 256), 256, 0, ctx->stream>>>(v, P->d_exdofs, send, P->nex, es);
        if (cudaGetLastError() != cudaSuccess) rc = NGSB_ERR_CUDA;
    }
    ctx->launches += 3;
    if (rc == NGSB_OK) {
        ncclResult_t r = g_nccl.GroupStart();
        for (size_t q = 0; q < P->peers.size() && r == ncclSuccess; q++) {
            const size_t off = P->peer_off[q] * es, cnt = (P->peer_off[q + 1] - P->peer_off[q]) * es;
            r = g_nccl.Send(send + off, cnt, ncclDouble, P->peers[q], comm->comm, ctx->stream);
            if (r == ncclSuccess) r = g_nccl.Recv(recv + off, cnt, ncclDouble, P->peers[q], comm->comm, ctx->stream);
        }
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) { set_error("Cumulate: %s", g_nccl.GetErrorString(r)); rc = NGSB_ERR_COMM; }
    }
    if (rc == NGSB_OK) {
        SpanGuard g(ctx, KC_OTHER);
        unpack_add_kernel<<<(unsigned)((P->nif + 255) / 256), 256, 0, ctx->stream>>>(v, P->d_if_dof, P->d_if_first, P->d_if_pos, P->d_if_nlow, recv, P->nif, es);
        if (cudaGetLastError() != cudaSuccess) rc = NGSB_ERR_CUDA;
    }
    if (own) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(send);
        cudaFree(recv);
    }
    return rc;
}

static int cumulate_raw(const ngsb_parmat *P, double *v) { return cumulate_es(P, v, P->es); }

// the two halves of the peer-memory Cumulate as separate calls (interface-first split of the CG product): the push
// carries no scalar, the unpack publishes `red` itself (fin = 2) and finishes kss
static int halo_push_only(const ngsb_parmat *P, const double *v, CgState *st)
{
    ngsb_ctx *ctx = P->comm->ctx;
    SpanGuard g(ctx, KC_OTHER);
    halo_push_kernel<<<halo_grid(ctx, P->nex), 256, 0, ctx->stream>>>(P->d_H, v, P->d_exdofs, P->nex, P->es, st, nullptr, nullptr);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}
static int halo_unpack_fin(const ngsb_parmat *P, double *v, CgState *st, const double *red)
{
    ngsb_ctx *ctx = P->comm->ctx;
    SpanGuard g(ctx, KC_OTHER);
    halo_unpack_kernel<<<halo_grid(ctx, P->nif), 256, 0, ctx->stream>>>(P->d_H, v, P->d_if_dof, P->d_if_first, P->d_if_pos, P->d_if_nlow, P->nif, P->es, st,
                                                                        P->comm->d_R, 2, red);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

__global__ void mask_zero_kernel(double *v, const uint8_t *master, size_t n, int es)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (!master[i])
            for (int c = 0; c < es; c++) v[es * i + c] = 0.0;
}

// Cumulate of the diagonal (setup path of the Jacobi constructor), `es` = doubles per matrix value
static int cumulate_any(void *arg, double *v, int es)
{
    const ngsb_parmat *P = (const ngsb_parmat *)arg;
    NGSB_REQUIRE(es <= P->esmax, "Cumulate(diagonal): %d doubles per entry exceed the exchange buffers (%d)", es, P->esmax);
    int rc = cumulate_es(P, v, es);
    if (rc == NGSB_OK && cudaStreamSynchronize(P->comm->ctx->stream) != cudaSuccess) rc = NGSB_ERR_CUDA;
    return rc;
}

// a wait in a peer-memory kernel timed out (a rank died or the ranks disagree on the sequence of collectives)
static int check_peer_error(ngsb_comm *comm, const char *who)
{
    if (!comm->p2p) return NGSB_OK;
    int err = 0;
    NGSB_CUDA(cudaMemcpyAsync(&err, comm->d_err, sizeof(int), cudaMemcpyDeviceToHost, comm->ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(comm->ctx->stream));
    if (err) { set_error("%s: a peer-memory wait timed out (rank %d of %d)", who, comm->rank, comm->nranks); return NGSB_ERR_COMM; }
    return NGSB_OK;
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

// set up the mailbox and map everybody else's; collective.  On return c->p2p tells whether ALL ranks succeeded.
static int comm_setup_p2p(ngsb_comm *c)
{
    ngsb_ctx *ctx = c->ctx;
    bool good = c->nranks <= NGSB_MAX_RANKS;
    if (good && cudaMalloc(&c->mailbox, MB_BYTES) != cudaSuccess) { cudaGetLastError(); c->mailbox = nullptr; good = false; }
    if (good) {
        NGSB_CUDA(cudaMemset(c->mailbox, 0, MB_BYTES));
        NGSB_TRY(write_magic(ctx, c->mailbox, MB_MAGIC, c->rank));
    }
    bool ok = false;
    NGSB_TRY(ipc_exchange(c, good ? c->mailbox : nullptr, MB_MAGIC, nullptr, c->peer_mailbox, &ok));
    if (!ok) {
        if (c->mailbox) { cudaFree(c->mailbox); c->mailbox = nullptr; }
        c->p2p = false;
        return NGSB_OK;
    }
    PeerReduce R;
    memset(&R, 0, sizeof(R));
    R.nranks = c->nranks;
    R.rank = c->rank;
    R.seq = (unsigned long long *)(c->mailbox + MB_SEQ);
    R.err = (int *)(c->mailbox + MB_ERR);
    R.mine = (PeerSlot *)(c->mailbox + MB_SLOTS);
    for (int p = 0; p < c->nranks; p++) R.theirs[p] = (PeerSlot *)(c->peer_mailbox[p] + MB_SLOTS);
    R.vec_mine = (double *)(c->mailbox + MB_VEC);
    for (int p = 0; p < c->nranks; p++) R.vec_theirs[p] = (double *)(c->peer_mailbox[p] + MB_VEC);
    c->d_err = R.err;
    NGSB_CUDA(cudaMalloc(&c->d_R, sizeof(PeerReduce)));
    NGSB_CUDA(cudaMemcpy(c->d_R, &R, sizeof(R), cudaMemcpyHostToDevice));
    c->p2p = true;
    return NGSB_OK;
}

// p2p_mode: -1 auto (peer memory if every rank can map every other rank, else NCCL), 0 NCCL only, 1 peer memory required.
// uid128 may be NULL when a bootstrap all-gather callback is given: then there is no NCCL communicator at all.
extern "C" int ngsb_comm_create_ex(ngsb_ctx *ctx, int nranks, int rank, const void *uid128,
                                   int (*allgather)(void *user, const void *send, void *recv, size_t bytes_per_rank), void *user,
                                   int p2p_mode, ngsb_comm **out)
{
    NGSB_REQUIRE(ctx && out, "ngsb_comm_create: NULL argument");
    NGSB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "ngsb_comm_create: bad rank %d of %d", rank, nranks);
    NGSB_REQUIRE(nranks == 1 || uid128 || allgather, "ngsb_comm_create: neither an ncclUniqueId nor a bootstrap all-gather given");
    NGSB_REQUIRE(p2p_mode >= -1 && p2p_mode <= 1, "ngsb_comm_create: p2p_mode must be -1, 0 or 1");
    NGSB_REQUIRE(p2p_mode != 0 || uid128 || nranks == 1, "ngsb_comm_create: NCCL-only mode needs an ncclUniqueId");
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const char *env = getenv("NGSB_COMM");
    if (env && !strcmp(env, "nccl") && uid128) p2p_mode = 0;
    if (env && !strcmp(env, "p2p")) p2p_mode = 1;
    ngsb_comm *c = new ngsb_comm();
    c->ctx = ctx;
    c->nranks = nranks;
    c->rank = rank;
    c->boot_allgather = allgather;
    c->boot_user = user;
    int rc = NGSB_OK;
    if (nranks > 1 && uid128) {
        rc = nccl_load();
        if (rc == NGSB_OK) {
            ncclUniqueId id;
            memcpy(&id, uid128, sizeof(id));
            ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
            if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); rc = NGSB_ERR_COMM; }
        }
    }
    if (rc == NGSB_OK && cudaMalloc(&c->d_red, 4 * sizeof(double)) != cudaSuccess) { set_error("ngsb_comm_create: out of memory"); rc = NGSB_ERR_NOMEM; }
    if (rc == NGSB_OK && nranks > 1 && p2p_mode != 0) {
        rc = comm_setup_p2p(c);
        if (rc == NGSB_OK && !c->p2p && (p2p_mode == 1 || !c->comm)) {
            set_error("ngsb_comm_create: peer memory (CUDA IPC between the ranks' GPUs) is not available%s", c->comm ? "" : " and there is no NCCL communicator");
            rc = NGSB_ERR_COMM;
        }
    }
    c->boot_allgather = nullptr;    // the callback is only valid during this call and ngsb_parmat_create_ex
    c->boot_user = nullptr;
    if (rc != NGSB_OK) { ngsb_comm_destroy(c); return rc; }
    *out = c;
    return NGSB_OK;
}

extern "C" int ngsb_comm_create(ngsb_ctx *ctx, int nranks, int rank, const void *uid128, ngsb_comm **out)
{
    NGSB_REQUIRE(nranks == 1 || uid128, "ngsb_comm_create: uid is NULL");
    return ngsb_comm_create_ex(ctx, nranks, rank, uid128, nullptr, nullptr, -1, out);
}

extern "C" int ngsb_comm_info(const ngsb_comm *c, int *nranks, int *rank, int *peer_memory, int *has_nccl)
{
    NGSB_REQUIRE(c, "ngsb_comm_info: comm is NULL");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (peer_memory) *peer_memory = c->p2p ? 1 : 0;
    if (has_nccl) *has_nccl = c->comm ? 1 : 0;
    return NGSB_OK;
}

extern "C" int ngsb_comm_destroy(ngsb_comm *c)
{
    if (!c) return NGSB_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    for (int p = 0; p < c->nranks && p < NGSB_MAX_RANKS; p++)
        if (p != c->rank && c->peer_mailbox[p]) cudaIpcCloseMemHandle(c->peer_mailbox[p]);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->d_R);
    cudaFree(c->mailbox);
    cudaFree(c->d_red);
    delete c;
    return NGSB_OK;
}

// receive areas + flags of the peer-memory Cumulate; collective.  P->p2p = all ranks succeeded.
static int parmat_setup_p2p(ngsb_parmat *P, const uint64_t *ex_first)
{
    ngsb_comm *c = P->comm;
    ngsb_ctx *ctx = c->ctx;
    const int np = c->nranks;
    const size_t stride = std::max<size_t>(2, P->nex * (size_t)P->esmax);           // doubles per parity
    size_t bytes = HB_DATA + 2 * stride * sizeof(double);
    bytes = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    bool good = (int)P->peers.size() < NGSB_MAX_RANKS;
    if (good && cudaMalloc(&P->halo_mem, bytes) != cudaSuccess) { cudaGetLastError(); P->halo_mem = nullptr; good = false; }
    if (good) {
        NGSB_CUDA(cudaMemset(P->halo_mem, 0, HB_DATA));
        NGSB_TRY(write_magic(ctx, P->halo_mem, HB_MAGIC, c->rank));
    }
    // everybody's exchange table offsets and strides: where my slice starts in each neighbour's packed list
    std::vector<uint64_t> mine(np + 2), all((size_t)np * (np + 2));
    for (int p = 0; p <= np; p++) mine[p] = ex_first[p];
    mine[np + 1] = stride;
    NGSB_TRY(boot_allgather(c, mine.data(), all.data(), mine.size() * sizeof(uint64_t)));
    bool ok = false;
    char *mapped[NGSB_MAX_RANKS] = {};
    NGSB_TRY(ipc_exchange(c, good ? P->halo_mem : nullptr, HB_MAGIC, &P->peers, mapped, &ok));
    if (!ok) {
        if (P->halo_mem) { cudaFree(P->halo_mem); P->halo_mem = nullptr; }
        P->p2p = false;
        return NGSB_OK;
    }
    NGSB_CUDA(cudaMalloc(&P->d_local, 256));
    NGSB_CUDA(cudaMemset(P->d_local, 0, 256));
    PeerHalo H;
    memset(&H, 0, sizeof(H));
    H.npeers = (int)P->peers.size();
    H.rank = c->rank;
    H.seq = (unsigned long long *)P->d_local;
    H.counter = (unsigned int *)(P->d_local + 64);
    H.err = c->d_err;
    H.recv = (double *)(P->halo_mem + HB_DATA);
    H.stride = stride;
    H.flags = (unsigned long long *)(P->halo_mem + HB_FLAGS);
    for (size_t q = 0; q < P->peers.size(); q++) {
        const int p = P->peers[q];
        const uint64_t *theirs = &all[(size_t)p * (np + 2)];
        P->peer_halo[q] = mapped[p];
        H.peer_rank[q] = p;
        H.peer_off[q] = (unsigned int)P->peer_off[q];
        H.peer_recv[q] = (double *)(mapped[p] + HB_DATA);
        H.peer_stride[q] = theirs[np + 1];
        H.peer_my_off[q] = theirs[c->rank];
        H.peer_flags[q] = (unsigned long long *)(mapped[p] + HB_FLAGS);
        // both sides must list the same number of shared dofs
        NGSB_REQUIRE(theirs[c->rank + 1] - theirs[c->rank] == P->peer_off[q + 1] - P->peer_off[q],
                     "ngsb_parmat_create: rank %d shares %llu dofs with rank %d, which lists %llu", c->rank,
                     (unsigned long long)(P->peer_off[q + 1] - P->peer_off[q]), p, (unsigned long long)(theirs[c->rank + 1] - theirs[c->rank]));
    }
    for (size_t q = P->peers.size(); q <= (size_t)NGSB_MAX_RANKS; q++) H.peer_off[q] = (unsigned int)P->nex;
    H.peer_off[P->peers.size()] = (unsigned int)P->nex;
    NGSB_CUDA(cudaMalloc(&P->d_H, sizeof(PeerHalo)));
    NGSB_CUDA(cudaMemcpy(P->d_H, &H, sizeof(H), cudaMemcpyHostToDevice));
    P->p2p = true;
    return NGSB_OK;
}

// option "dist_overlap": which scheduled SELL slices hold an interface row (they are multiplied first and pushed)
__global__ void __launch_bounds__(256) mark_interface_slices_kernel(const uint32_t *__restrict__ slice_src, const uint32_t *__restrict__ row_of,
                                                                   const uint8_t *__restrict__ is_if, uint32_t nslices, uint8_t *__restrict__ flag)
{
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= nslices) return;                      // uniform per warp
    const uint32_t row = row_of[(uint64_t)slice_src[t] * 32 + lane];
    const bool hit = row != 0xffffffffu && is_if[row] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) flag[t] = m ? 1 : 0;
}

// tables for the product + exchange kernel: which scheduled slices hold interface rows, and which interface dof sits in which lane
static int parmat_setup_fused_push(ngsb_parmat *P, const std::vector<int32_t> &if_dof)
{
    const ngsb_csr *A = P->local;
    ngsb_ctx *ctx = P->comm->ctx;
    const uint32_t ns = A->nslices;
    std::vector<uint32_t> slice_src(ns), row_of((size_t)ns * 32);
    NGSB_CUDA(cudaMemcpyAsync(slice_src.data(), A->d_slice_src, (size_t)ns * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    NGSB_CUDA(cudaMemcpyAsync(row_of.data(), A->d_row_of, (size_t)ns * 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> ifidx(P->n, -1);
    for (size_t k = 0; k < if_dof.size(); k++) ifidx[if_dof[k]] = (int32_t)k;
    std::vector<uint32_t> slice_if(ns, 0xffffffffu);
    std::vector<int32_t> lane_if;
    for (uint32_t t = 0; t < ns; t++) {
        const size_t base = (size_t)slice_src[t] * 32;
        bool any = false;
        for (int l = 0; l < 32 && !any; l++) any = row_of[base + l] != 0xffffffffu && ifidx[row_of[base + l]] >= 0;
        if (!any) continue;
        slice_if[t] = (uint32_t)(lane_if.size() / 32);
        slice_src[t] |= 0x80000000u;
        for (int l = 0; l < 32; l++) lane_if.push_back(row_of[base + l] != 0xffffffffu ? ifidx[row_of[base + l]] : -1);
    }
    NGSB_REQUIRE(ns < 0x80000000u, "parallel matrix: too many slices");
    NGSB_CUDA(cudaMalloc(&P->d_slice_src_flagged, std::max<size_t>(1, ns) * sizeof(uint32_t)));
    NGSB_CUDA(cudaMemcpyAsync(P->d_slice_src_flagged, slice_src.data(), (size_t)ns * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    NGSB_CUDA(cudaMalloc(&P->d_slice_if, std::max<size_t>(1, ns) * sizeof(uint32_t)));
    NGSB_CUDA(cudaMalloc(&P->d_lane_if, std::max<size_t>(32, lane_if.size()) * sizeof(int32_t)));
    NGSB_CUDA(cudaMemcpyAsync(P->d_slice_if, slice_if.data(), (size_t)ns * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (!lane_if.empty()) NGSB_CUDA(cudaMemcpyAsync(P->d_lane_if, lane_if.data(), lane_if.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    SellPush desc;
    desc.H = P->d_H;
    desc.R = P->comm->d_R;
    desc.slice_if = P->d_slice_if;
    desc.lane_if = P->d_lane_if;
    desc.if_first = P->d_if_first;
    desc.if_pos = P->d_if_pos;
    desc.es = P->es;
    NGSB_CUDA(cudaMalloc(&P->d_push_desc, sizeof(SellPush)));
    NGSB_CUDA(cudaMemcpy(P->d_push_desc, &desc, sizeof(SellPush), cudaMemcpyHostToDevice));
    P->fused_push = true;
    return NGSB_OK;
}

static int parmat_setup_overlap(ngsb_parmat *P, const std::vector<int32_t> &if_dof)
{
    const ngsb_csr *A = P->local;
    ngsb_ctx *ctx = P->comm->ctx;
    const uint32_t ns = A->nslices;
    std::vector<uint8_t> is_if(P->n, 0), flag(ns, 0);
    for (int32_t d : if_dof) is_if[d] = 1;
    uint8_t *d_is = nullptr, *d_flag = nullptr;
    NGSB_CUDA(cudaMalloc(&d_is, std::max<size_t>(16, P->n)));
    if (cudaMalloc(&d_flag, ns) != cudaSuccess) { cudaFree(d_is); set_error("ngsb_parmat_create: out of memory"); return NGSB_ERR_NOMEM; }
    cudaError_t e = cudaMemcpyAsync(d_is, is_if.data(), P->n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        mark_interface_slices_kernel<<<(unsigned)(((uint64_t)ns * 32 + 255) / 256), 256, 0, ctx->stream>>>(A->d_slice_src, A->d_row_of, d_is, ns, d_flag);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(flag.data(), d_flag, ns, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_is);
    cudaFree(d_flag);
    if (e != cudaSuccess) { set_error("ngsb_parmat_create (interface slices): %s", cudaGetErrorString(e)); return NGSB_ERR_CUDA; }
    std::vector<uint32_t> bnd, inner;
    for (uint32_t t = 0; t < ns; t++) (flag[t] ? bnd : inner).push_back(t);
    if (bnd.empty() || inner.empty()) return NGSB_OK;           // nothing to overlap: one launch as usual
    P->n_bnd = (uint32_t)bnd.size();
    P->n_int = (uint32_t)inner.size();
    NGSB_CUDA(cudaMalloc(&P->d_list_bnd, bnd.size() * sizeof(uint32_t)));
    NGSB_CUDA(cudaMalloc(&P->d_list_int, inner.size() * sizeof(uint32_t)));
    NGSB_CUDA(cudaMalloc(&P->d_red_b, 2 * sizeof(double)));
    NGSB_CUDA(cudaMemset(P->d_red_b, 0, 2 * sizeof(double)));
    NGSB_CUDA(cudaMemcpy(P->d_list_bnd, bnd.data(), bnd.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    NGSB_CUDA(cudaMemcpy(P->d_list_int, inner.data(), inner.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    P->overlap = true;
    return NGSB_OK;
}

extern "C" int ngsb_parmat_create_ex(ngsb_comm *comm, const ngsb_csr *local, const uint64_t *ex_first, const int32_t *ex_dofs,
                                     int (*allgather)(void *user, const void *send, void *recv, size_t bytes_per_rank), void *user,
                                     ngsb_parmat **out)
{
    NGSB_REQUIRE(comm && local && ex_first && out, "ngsb_parmat_create: NULL argument");
    NGSB_REQUIRE(local->ctx == comm->ctx, "ngsb_parmat_create: matrix and communicator belong to different contexts");
    NGSB_REQUIRE(local->h == local->w, "ngsb_parmat_create: local matrix must be square");
    NGSB_REQUIRE(local->inner == nullptr, "ngsb_parmat_create: the local matrix is internally reordered; create it with option reorder = 0 "
                 "(the exchange tables index the caller's numbering)");
    NGSB_REQUIRE(comm->comm || allgather || comm->nranks == 1, "ngsb_parmat_create: this communicator has no NCCL; pass the bootstrap all-gather");
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
    P->esmax = (int)kind_matscalars(local->kind);
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
    std::vector<uint32_t> if_first, if_pos(nex), if_nlow;
    for (size_t k = 0; k < nex; k++) {
        if (k == 0 || pairs[k].first != pairs[k - 1].first) { if_dof.push_back(pairs[k].first); if_first.push_back((uint32_t)k); if_nlow.push_back(0); }
        if_pos[k] = pairs[k].second;
        if (pairs[k].second < ex_first[comm->rank]) if_nlow.back()++;        // ex_dofs is rank-major: a position below my own range = a lower rank
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
    if (rc == NGSB_OK) rc = up((void **)&P->d_if_nlow, if_nlow.data(), if_nlow.size() * sizeof(uint32_t));
    if (rc == NGSB_OK) rc = up((void **)&P->d_master, P->h_master.data(), n);
    cudaStreamSynchronize(ctx->stream);
    if (rc == NGSB_OK && comm->p2p) {
        comm->boot_allgather = allgather;
        comm->boot_user = user;
        rc = parmat_setup_p2p(P, ex_first);
        comm->boot_allgather = nullptr;
        comm->boot_user = nullptr;
        if (rc == NGSB_OK && !P->p2p && !comm->comm) { set_error("ngsb_parmat_create: could not map the neighbours' receive areas and there is no NCCL communicator"); rc = NGSB_ERR_COMM; }
    }
    if (rc == NGSB_OK && P->p2p && np > 1 && ctx->dist_overlap && local->nslices > 0 && !if_dof.empty() &&
        (ctx->spmv_algo == 0 || ctx->spmv_algo == 3))
        rc = parmat_setup_overlap(P, if_dof);
    if (rc == NGSB_OK && P->p2p && np > 1 && !P->overlap && ctx->dist_fused_push && local->nslices > 0 && (ctx->spmv_algo == 0 || ctx->spmv_algo == 3))
        rc = parmat_setup_fused_push(P, if_dof);
    if (rc == NGSB_OK && !P->p2p && np > 1) {
        if (cudaMalloc(&P->d_send, std::max<size_t>(16, nex * P->es * sizeof(double))) != cudaSuccess) rc = NGSB_ERR_NOMEM;
        if (rc == NGSB_OK && cudaMalloc(&P->d_recv, std::max<size_t>(16, nex * P->es * sizeof(double))) != cudaSuccess) rc = NGSB_ERR_NOMEM;
        if (rc != NGSB_OK) set_error("ngsb_parmat_create: out of memory");
    }
    if (rc != NGSB_OK) { ngsb_parmat_destroy(P); return rc; }
    *out = P;
    return NGSB_OK;
}

extern "C" int ngsb_parmat_create(ngsb_comm *comm, const ngsb_csr *local, const uint64_t *ex_first, const int32_t *ex_dofs,
                                  ngsb_parmat **out)
{
    return ngsb_parmat_create_ex(comm, local, ex_first, ex_dofs, nullptr, nullptr, out);
}

extern "C" int ngsb_parmat_destroy(ngsb_parmat *P)
{
    if (!P) return NGSB_OK;
    cudaSetDevice(P->comm->ctx->device);
    cudaStreamSynchronize(P->comm->ctx->stream);
    if (P->graph_exec) cudaGraphExecDestroy(P->graph_exec);
    for (size_t q = 0; q < P->peers.size() && q < (size_t)NGSB_MAX_RANKS; q++)
        if (P->peer_halo[q]) cudaIpcCloseMemHandle(P->peer_halo[q]);
    cudaFree(P->d_H); cudaFree(P->d_local); cudaFree(P->halo_mem);
    cudaFree(P->d_exdofs); cudaFree(P->d_send); cudaFree(P->d_recv);
    cudaFree(P->d_if_dof); cudaFree(P->d_if_first); cudaFree(P->d_if_pos); cudaFree(P->d_if_nlow); cudaFree(P->d_master); cudaFree(P->d_slice_if); cudaFree(P->d_lane_if); cudaFree(P->d_push_desc); cudaFree(P->d_slice_src_flagged);
    cudaFree(P->d_list_bnd); cudaFree(P->d_list_int); cudaFree(P->d_red_b);
    delete P;
    return NGSB_OK;
}

extern "C" int ngsb_parmat_info(const ngsb_parmat *P, int *peer_memory, int *n_neighbours, size_t *n_exchange, size_t *n_interface)
{
    NGSB_REQUIRE(P, "ngsb_parmat_info: NULL argument");
    if (peer_memory) *peer_memory = P->p2p ? 1 : 0;
    if (n_neighbours) *n_neighbours = (int)P->peers.size();
    if (n_exchange) *n_exchange = P->nex;
    if (n_interface) *n_interface = P->nif;
    return NGSB_OK;
}

extern "C" int ngsb_parmat_overlap_info(const ngsb_parmat *P, int *enabled, size_t *interface_slices, size_t *interior_slices)
{
    NGSB_REQUIRE(P, "ngsb_parmat_overlap_info: NULL argument");
    if (enabled) *enabled = P->overlap ? 1 : 0;
    if (interface_slices) *interface_slices = P->n_bnd;
    if (interior_slices) *interior_slices = P->n_int;
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
    NGSB_TRY(cumulate_raw(P, v->d));
    return check_peer_error(P->comm, "ParallelBaseVector::Cumulate");
}

extern "C" int ngsb_parmat_mult(const ngsb_parmat *P, const ngsb_vec *x, ngsb_vec *y)
{
    NGSB_TRY(check_pvec(P, x, "ParallelMatrix::Mult"));
    NGSB_TRY(check_pvec(P, y, "ParallelMatrix::Mult"));
    return ngsb_csr_mult(P->local, x, y);
}

// out: (re, im); conjugate as in S_BaseVector<Complex>::InnerProduct (on the argument y)
extern "C" int ngsb_parmat_dot(const ngsb_parmat *P, const ngsb_vec *x, const ngsb_vec *y, int both_cumulated, int conjugate, double out[2])
{
    NGSB_TRY(check_pvec(P, x, "ParallelBaseVector::InnerProduct"));
    NGSB_TRY(check_pvec(P, y, "ParallelBaseVector::InnerProduct"));
    NGSB_REQUIRE(out, "ngsb_parmat_dot: out is NULL");
    ngsb_comm *comm = P->comm;
    ngsb_ctx *ctx = comm->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const bool cplx = P->local->kind == NGSB_COMPLEX;
    // both CUMULATED: the reference masks by master dofs (entry size 1) or Distribute()s one operand
    // (parallelvvector.cpp:302-322, 346-347); both equal "count every shared dof once, on its master"
    int rc = launch_dot_masked(ctx, x->d, y->d, cplx ? x->n : x->nscal, cplx ? (conjugate ? 2 : 1) : 0, comm->d_red,
                               both_cumulated ? P->d_master : nullptr, (unsigned)(cplx ? 1 : P->es));
    if (rc == NGSB_OK) rc = all_reduce2(comm, comm->d_red);
    if (rc == NGSB_OK) {
        cudaError_t e = cudaMemcpyAsync(ctx->h_pinned, comm->d_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("ngsb_parmat_dot: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; }
        else { out[0] = ctx->h_pinned[0]; out[1] = cplx ? ctx->h_pinned[1] : 0.0; }
    }
    if (rc == NGSB_OK) rc = check_peer_error(comm, "ParallelBaseVector::InnerProduct");
    return rc;
}

// one distributed CG iteration; peer-memory mode: 6 kernels, no library call, capturable
static int enqueue_par_iteration(const ngsb_parmat *P, const CgVecs &v, double *as)
{
    ngsb_comm *comm = P->comm;
    ngsb_ctx *ctx = comm->ctx;
    const ngsb_csr *A = P->local;
    SpmvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.x = v.s; a.y = as; a.sr = 1.0; a.accumulate = false;
    a.epi = EPI_DOT_OUT; a.dotvec = v.s; a.dot_conj = v.ip_mode == NGSB_IP_COMPLEX_CONJ; a.dot_out = comm->d_red; a.state = v.state;
    if (P->overlap) {
        // interface rows first, their values leave for the neighbours, interior rows while they travel; the neighbours'
        // values have landed long before the unpack kernel asks for them
        SpmvArgs b = a;
        b.slice_list = P->d_list_bnd; b.nlist = P->n_bnd; b.dot_out = P->d_red_b;
        NGSB_TRY(spmv_launch(b));                                               // as(interface slices), partial <s,as>
        NGSB_TRY(halo_push_only(P, as, v.state));
        a.slice_list = P->d_list_int; a.nlist = P->n_int; a.dot_add = P->d_red_b; a.skip_overflow = true;
        NGSB_TRY(spmv_launch(a));                                               // as(interior slices), local <s,as> complete
        NGSB_TRY(halo_unpack_fin(P, as, v.state, comm->d_red));                 // as -> CUMULATED; kss all-reduced, al = wd/kss
        NGSB_TRY(cg_launch_fused(ctx, A->kind, 1, v, 0));
        if (!v.R) NGSB_TRY(cg_launch_finalize(ctx, 2, v.state, comm->d_red, v.hist, comm->d_R));
        NGSB_TRY(cg_launch_dir(ctx, A->kind, v));
        return NGSB_OK;
    }
    if (P->fused_push) { a.push = P->d_push_desc; a.push_slice_src = P->d_slice_src_flagged; }                                  // interface rows leave for the neighbours from inside the product
    NGSB_TRY(spmv_launch(a));                                                   // as = A s (DISTRIBUTED), local <s,as>
    if (P->p2p) {
        if (P->fused_push) {
            // the product pushed the rows, the flags and the <s,As> partial itself: only the unpack half of Cumulate is left
            SpanGuard g(ctx, KC_OTHER);
            halo_unpack_kernel<<<halo_grid(ctx, P->nif), 256, 0, ctx->stream>>>(P->d_H, as, P->d_if_dof, P->d_if_first, P->d_if_pos, P->d_if_nlow, P->nif, P->es,
                                                                                v.state, comm->d_R, 1, comm->d_red);
            NGSB_CUDA(cudaGetLastError());
        } else
        NGSB_TRY(cumulate_es(P, as, P->es, v.state, comm->d_red, 1));           // as -> CUMULATED; kss all-reduced, al = wd/kss
        NGSB_TRY(cg_launch_fused(ctx, A->kind, 1, v, 0));                       // u, d, w, masked <d,w> (-> d_red, or all-reduced in place)
        if (!v.R) NGSB_TRY(cg_launch_finalize(ctx, 2, v.state, comm->d_red, v.hist, comm->d_R));   // all-reduce, be, loop condition
    } else {
        NGSB_TRY(all_reduce2(comm, comm->d_red));
        NGSB_TRY(cg_launch_finalize(ctx, 1, v.state, comm->d_red, v.hist));
        NGSB_TRY(cumulate_raw(P, as));
        NGSB_TRY(cg_launch_fused(ctx, A->kind, 1, v, 0));
        NGSB_TRY(all_reduce2(comm, comm->d_red));
        NGSB_TRY(cg_launch_finalize(ctx, 2, v.state, comm->d_red, v.hist));
    }
    NGSB_TRY(cg_launch_dir(ctx, A->kind, v));                                   // s = be s + w
    return NGSB_OK;
}

// Distributed Jacobi-PCG, the reference's sequence of statuses:
//   d = f (DISTRIBUTED) -> Jacobi cumulates d; w, s CUMULATED; wdn = masked <w,d> + all-reduce
//   loop: as = A s (DISTRIBUTED); kss = <s,as> local + all-reduce; u += al s;
//         d -= al as cumulates as (the one neighbour exchange of the iteration);
//         w = C d; wdn = masked <d,w> + all-reduce; s = be s + w.
extern "C" int ngsb_parmat_cg_solve(const ngsb_parmat *P, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *u, double prec,
                                    int maxsteps, int ip_mode, int *steps, double *history, int hist_cap, int *nhist)
{
    NGSB_TRY(check_pvec(P, f, "CGSolver::Mult(parallel)"));
    NGSB_TRY(check_pvec(P, u, "CGSolver::Mult(parallel)"));
    const ngsb_csr *A = P->local;
    NGSB_REQUIRE(ip_mode >= 0 && ip_mode <= 2 && (A->kind == NGSB_COMPLEX) == (ip_mode != NGSB_IP_REAL),
                 "CGSolver::Mult(parallel): ip_mode %d does not fit matrix kind %d", ip_mode, A->kind);
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
    hs->cplx = ip_mode != NGSB_IP_REAL;
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
    v.R = (P->p2p && ctx->dist_fused_push) ? comm->d_R : nullptr;      // the update kernel all-reduces <d,w> itself (krylov.cu)
    v.n = A->h;
    v.state = d_state;
    v.hist = d_hist;
    v.partials = ctx->d_partials;
    v.counter = ctx->d_counter;
    v.ip_mode = ip_mode;
    v.fold_u = ctx->cg_fold_u ? 1 : 0;
    v.chunked = ctx->cg_chunked ? 1 : 0;
    v.stream = ctx->cg_stream_hints ? 1 : 0;

    // u = 0; d = f, cumulated by the Jacobi application (linalg/jacobi.cpp:78)
    cu(cudaMemsetAsync(u->d, 0, nscal * sizeof(double), ctx->stream));
    cu(cudaMemcpyAsync(d, f->d, nscal * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (rc == NGSB_OK) rc = cumulate_raw(P, d);
    v.f = d;                       // init kernel reads f, writes d: same values
    if (rc == NGSB_OK) rc = cg_launch_fused(ctx, A->kind, 0, v, 0);
    if (rc == NGSB_OK && !P->p2p) rc = all_reduce2(comm, comm->d_red);
    if (rc == NGSB_OK && !v.R) rc = cg_launch_finalize(ctx, 0, d_state, comm->d_red, d_hist, P->p2p ? comm->d_R : nullptr);

    const long batch = ctx->cg_batch;
    // peer-memory mode has no library call inside the iteration: one CUDA graph per batch
    ngsb_parmat *Pm = const_cast<ngsb_parmat *>(P);
    const bool use_graph = P->p2p && !ctx->timing && getenv("NGSB_NO_CUDA_GRAPH") == nullptr && batch > 1;
    if (rc == NGSB_OK && use_graph) {
        // the preconditioner enters by its unique id, never by its address (a re-created object may get the same one back)
        const void *key[9] = {u->d, d, w, s, as, (const void *)(uintptr_t)(C ? C->uid : 0), (const void *)(intptr_t)ip_mode, d_hist,
                              (const void *)(intptr_t)(ctx->cg_fold_u | (ctx->cg_chunked << 1) | (ctx->sell_variant << 4) | (ctx->sell_c16 << 12) | (ctx->spmv_ctas_per_sm << 16))};
        const bool hit = Pm->graph_exec && Pm->g_batch == batch && memcmp(key, Pm->g_key, sizeof(key)) == 0;
        if (!hit) {
            cudaGraph_t graph = nullptr;
            const uint64_t launches_before = ctx->launches;
            cu(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            if (rc == NGSB_OK) {
                for (long k = 0; k < batch && rc == NGSB_OK; k++) rc = enqueue_par_iteration(P, v, as);
                cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
                Pm->g_launches = ctx->launches - launches_before;
                ctx->launches = launches_before;
                if (rc == NGSB_OK) cu(ce);
                if (rc == NGSB_OK && Pm->graph_exec) {      // same topology, other pointers: update in place
                    cudaGraphExecUpdateResultInfo info;
                    if (cudaGraphExecUpdate(Pm->graph_exec, graph, &info) != cudaSuccess) {
                        cudaGetLastError();
                        cudaGraphExecDestroy(Pm->graph_exec);
                        Pm->graph_exec = nullptr;
                    }
                }
                if (rc == NGSB_OK && !Pm->graph_exec) cu(cudaGraphInstantiate(&Pm->graph_exec, graph, 0));
                if (graph) cudaGraphDestroy(graph);
                if (rc == NGSB_OK) { memcpy(Pm->g_key, key, sizeof(key)); Pm->g_batch = batch; }
            }
        }
    }
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cu(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    cu(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    const long max_batches = ((long)maxsteps + batch - 1) / batch + 1;
    long enq = 0;
    bool finished = false;
    while (rc == NGSB_OK && !finished) {
        if (enq < max_batches) {
            if (use_graph) {
                cu(cudaGraphLaunch(Pm->graph_exec, ctx->stream));
                ctx->launches += Pm->g_launches;
            } else {
                for (long k = 0; k < batch && rc == NGSB_OK; k++) rc = enqueue_par_iteration(P, v, as);
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
    if (rc == NGSB_OK) rc = check_peer_error(comm, "CGSolver::Mult(parallel)");
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

// GMRESSolver<IPTYPE>::Mult on parallel vectors (linalg/cg.cpp:854-1022 with parallel/parallelvvector.cpp semantics):
// f DISTRIBUTED in, x CUMULATED out.
static int gm_allreduce(void *arg, double *d_buf) { return all_reduce2(((const ngsb_parmat *)arg)->comm, d_buf); }
static int gm_cumulate(void *arg, double *v) { return cumulate_raw((const ngsb_parmat *)arg, v); }

extern "C" int ngsb_parmat_gmres_solve(const ngsb_parmat *P, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *x, double prec,
                                       int maxsteps, int *steps, double *history, int hist_cap, int *nhist)
{
    NGSB_TRY(check_pvec(P, f, "GMRESSolver::Mult(parallel)"));
    NGSB_TRY(check_pvec(P, x, "GMRESSolver::Mult(parallel)"));
    ngsb_comm *comm = P->comm;
    GmresDist dist;
    dist.master = P->d_master;
    dist.R = (comm->nranks > 1 && P->p2p) ? comm->d_R : nullptr;
    dist.allreduce = (comm->nranks > 1 && !P->p2p) ? gm_allreduce : nullptr;
    dist.cumulate = gm_cumulate;
    dist.arg = (void *)P;
    NGSB_TRY(gmres_solve_impl(P->local, C, f, x, prec, maxsteps, 1, steps, history, hist_cap, nhist, &dist));
    return check_peer_error(comm, "GMRESSolver::Mult(parallel)");
}
