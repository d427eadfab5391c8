// peer.cuh -- collectives over peer memory (NVLink / NVSwitch P2P stores), device side.
//
// The reference's distributed vectors talk through MPI (parallel/parallelvvector.cpp:247-272,
// 452-475: ISend/IRecv of the interface values; paralleldofs->GetCommunicator().AllReduce for
// the scalars).  Here every rank exports a small mailbox and its interface receive buffers
// with CUDA IPC; the kernels of the solve loop write straight into the neighbours' memory and
// signal with sequence-numbered flags, so one CG iteration needs no library collective and no
// host involvement (and can be captured in a CUDA graph):
//
//  * PeerReduce: all-reduce (sum) of one (re,im) pair.  Every rank stores its partial into
//    slot [parity][own rank] of EVERY rank's mailbox, then reads the nranks slots of its own
//    mailbox in rank order -> the sum is bitwise identical on all ranks (they all take the same
//    `done` decision) and independent of arrival order.
//  * PeerHalo: Cumulate.  The push kernel gathers the interface values of neighbour q and stores
//    them into q's receive buffer at the place q's exchange table expects them; the last CTA
//    publishes flag[parity][own rank] = sequence number to every neighbour.  The unpack kernel
//    waits for the neighbours' flags and adds the received copies in ascending rank order.
//
// Two parities are enough: a rank can start exchange k+2 only after it finished k+1, which needs
// every neighbour's push k+1, which that neighbour issues after finishing its own unpack k.
#pragma once
#include <stdint.h>

#define NGSB_MAX_RANKS 16
#define NGSB_PEER_VEC_LEN 4096           // doubles per rank and parity in the vector all-reduce area

struct PeerSlot {
    double v[2];
    unsigned long long seq;
    unsigned long long pad;
};

struct PeerReduce {
    int nranks, rank;
    unsigned long long *seq;             // local: number of completed reductions
    int *err;                            // local: set to 1 when a wait timed out
    PeerSlot *mine;                      // local mailbox: [2][NGSB_MAX_RANKS]
    PeerSlot *theirs[NGSB_MAX_RANKS];    // theirs[p]: rank p's mailbox (mapped); theirs[rank] == mine
    double *vec_mine;                    // local vector area: [2][NGSB_MAX_RANKS][NGSB_PEER_VEC_LEN]
    double *vec_theirs[NGSB_MAX_RANKS];  // the same area of rank p (mapped); vec_theirs[rank] == vec_mine
};

struct PeerHalo {
    int npeers, rank;
    unsigned long long *seq;             // local: number of completed exchanges
    unsigned int *counter;               // local: [0] push CTAs done, [1] unpack CTAs done
    int *err;
    double *recv;                        // local receive area, parity p at recv + p*stride
    unsigned long long stride;           // doubles
    unsigned long long *flags;           // local: [2][NGSB_MAX_RANKS], written by the neighbours
    int peer_rank[NGSB_MAX_RANKS];
    unsigned int peer_off[NGSB_MAX_RANKS + 1];         // my packed exchange list, neighbour-major (dofs)
    double *peer_recv[NGSB_MAX_RANKS];                 // neighbour q's receive area (mapped)
    unsigned long long peer_stride[NGSB_MAX_RANKS];
    unsigned long long peer_my_off[NGSB_MAX_RANKS];    // where my slice starts in q's packed list (dofs)
    unsigned long long *peer_flags[NGSB_MAX_RANKS];
};

#ifdef __CUDACC__
namespace ngsb {

static const unsigned long long PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag == want; false (and *err = 1) after PEER_TIMEOUT_NS
__device__ __forceinline__ bool peer_wait(const unsigned long long *flag, unsigned long long want, int *err)
{
    if (ld_acquire_sys(flag) == want) return true;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        for (int k = 0; k < 64; k++)
            if (ld_acquire_sys(flag) == want) return true;
        if (*(volatile int *)err) return false;                    // somebody else already gave up
        if (global_timer_ns() - t0 > PEER_TIMEOUT_NS) { *(volatile int *)err = 1; return false; }
    }
}

// one thread: publish my partial of reduction (*seq + 1) to every rank
__device__ __forceinline__ void pr_push(const PeerReduce &R, double a, double b)
{
    const unsigned long long s = *(volatile unsigned long long *)R.seq + 1;
    for (int p = 0; p < R.nranks; p++) {
        PeerSlot *t = R.theirs[p] + (s & 1) * NGSB_MAX_RANKS + R.rank;
        *(volatile double *)&t->v[0] = a;
        *(volatile double *)&t->v[1] = b;
        st_release_sys(&t->seq, s);
    }
}

// a whole warp (all lanes hold a, b): lane p publishes the partial to rank p -- the nranks release stores travel over NVLink
// side by side, one round trip instead of nranks in a row (at 8 ranks the serial form costs about 20 us on the critical
// path of every reduction)
__device__ __forceinline__ void pr_push_warp(const PeerReduce &R, double a, double b)
{
    const unsigned long long s = *(volatile unsigned long long *)R.seq + 1;
    for (int p = threadIdx.x & 31; p < R.nranks; p += 32) {
        PeerSlot *t = R.theirs[p] + (s & 1) * NGSB_MAX_RANKS + R.rank;
        *(volatile double *)&t->v[0] = a;
        *(volatile double *)&t->v[1] = b;
        st_release_sys(&t->seq, s);
    }
}

// a whole warp: lane p waits for rank p's partial; the sum is formed in rank order (bitwise the one-thread result) and
// returned in every lane; lane 0 completes the reduction
__device__ __forceinline__ double2 pr_wait_sum_warp(const PeerReduce &R)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long s = *(volatile unsigned long long *)R.seq + 1;
    double pa = 0.0, pb = 0.0;
    if (lane < R.nranks) {
        const PeerSlot *t = R.mine + (s & 1) * NGSB_MAX_RANKS + lane;
        peer_wait(&t->seq, s, R.err);
        pa = *(volatile const double *)&t->v[0];
        pb = *(volatile const double *)&t->v[1];
    }
    double a = 0.0, b = 0.0;
    for (int p = 0; p < R.nranks; p++) {
        a += __shfl_sync(0xffffffffu, pa, p);
        b += __shfl_sync(0xffffffffu, pb, p);
    }
    __syncwarp();
    if (lane == 0) *(volatile unsigned long long *)R.seq = s;
    return make_double2(a, b);
}

// one thread: wait for all partials of reduction (*seq + 1), sum them in rank order, complete it
__device__ __forceinline__ double2 pr_wait_sum(const PeerReduce &R)
{
    const unsigned long long s = *(volatile unsigned long long *)R.seq + 1;
    double a = 0.0, b = 0.0;
    for (int p = 0; p < R.nranks; p++) {
        const PeerSlot *t = R.mine + (s & 1) * NGSB_MAX_RANKS + p;
        peer_wait(&t->seq, s, R.err);
        a += *(volatile const double *)&t->v[0];
        b += *(volatile const double *)&t->v[1];
    }
    *(volatile unsigned long long *)R.seq = s;
    return make_double2(a, b);
}

// a whole CTA (blockDim >= 32, every thread calls): buf[0..n) <- sum over ranks, n <= NGSB_PEER_VEC_LEN; buf is memory the
// whole CTA sees (shared or global).  Every rank stores its n values into slot [parity][own rank] of EVERY rank's vector area,
// publishes them with the flag of a scalar reduction (the stores are fenced before the CTA barrier that precedes the release)
// and, once all flags are there, sums the nranks copies in its own area in rank order: the same bits on every rank.  Counts as
// reduction number *seq + 1 like the scalar form, so the two kinds may be mixed freely as long as all ranks issue them in the
// same order.
__device__ __forceinline__ void pr_allreduce_vec_block(const PeerReduce &R, double *buf, int n)
{
    const unsigned long long s = *(volatile unsigned long long *)R.seq + 1;
    __syncthreads();                                       // everybody has read seq before warp 0 completes the reduction
    for (int p = 0; p < R.nranks; p++) {
        double *dst = R.vec_theirs[p] + ((s & 1) * NGSB_MAX_RANKS + R.rank) * (size_t)NGSB_PEER_VEC_LEN;
        for (int i = threadIdx.x; i < n; i += blockDim.x) *(volatile double *)&dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < 32) {
        pr_push_warp(R, 0.0, 0.0);
        pr_wait_sum_warp(R);
    }
    __syncthreads();
    const double *src = R.vec_mine + (s & 1) * NGSB_MAX_RANKS * (size_t)NGSB_PEER_VEC_LEN;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double a = 0.0;
        for (int p = 0; p < R.nranks; p++) a += *(volatile const double *)&src[p * (size_t)NGSB_PEER_VEC_LEN + i];
        buf[i] = a;
    }
    __syncthreads();
}

// one thread: buf(re,im) <- sum over ranks
__device__ __forceinline__ void pr_allreduce_inplace(const PeerReduce &R, double *buf)
{
    pr_push(R, buf[0], buf[1]);
    double2 t = pr_wait_sum(R);
    buf[0] = t.x;
    buf[1] = t.y;
}

} // namespace ngsb
#endif
