// krylov.cu -- device-resident CG and GMRES.
//
// CG: CGSolver<IPTYPE>::Mult (linalg/cg.cpp:503-633) with the same recurrences and stopping
// rule, replacing DevCGSolver::Mult (ngscuda/cuda_krylov.cpp:19-203: 12 unfused graph nodes
// per iteration).  One iteration here is three kernels:
//   A  spmv (+ fused kss = <s, A s>, last block: al = wd/kss)                 [spmv.cu]
//   B  u += al s ; d -= al as ; w = C d ; wdn = <d, w> ; last block: be, loop condition
//   C  s = be s + w
// All scalars and the loop counter live in a CgState on the device; every kernel returns
// at once when state->done is set, so the host enqueues iterations in batches (optionally
// as one CUDA graph per batch) and only polls the flag between batches.
//
// GMRES: GMRESSolver<IPTYPE>::Mult (linalg/cg.cpp:854-1022): left preconditioning, Givens
// rotations and the triangular solve in single-thread kernels on the device.  The
// orthogonalisation produces the coefficients of the reference's modified Gram-Schmidt loop
// with one batched reduction per step (gmres_dots_kernel and the comment above it); option
// gmres_orth = 0 runs the loop in the reference's order, each projection fused with the next
// inner product.
#include "krylov.cuh"
#include "peer.cuh"

#include <map>

namespace ngsb {

// ------------------------------------------------------------------------------------------
// workspace cache
// ------------------------------------------------------------------------------------------
struct Workspace {
    std::vector<std::pair<size_t, double *>> free_bufs;
    CgState *d_state = nullptr;
    CgState *h_state = nullptr;       // pinned
    double *d_hist = nullptr;
    size_t hist_cap = 0;
    cudaGraphExec_t graph_exec = nullptr;
    // key of the cached graph: the unique ids of matrix and preconditioner (never their addresses -- a destroyed and
    // re-created object may get the same address back), the vectors, and every option the captured launches read
    uint64_t g_A = 0, g_C = 0;
    double *g_ptrs[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    long g_opts[11] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
};

// buffers taken from the workspace go back on every exit path
struct BufLease {
    ngsb_ctx *ctx;
    std::vector<std::pair<size_t, double *>> held;
    explicit BufLease(ngsb_ctx *c) : ctx(c) {}
    int get(size_t nscal, double **out)
    {
        NGSB_TRY(ws_get_buf(ctx, nscal, out));
        held.emplace_back(nscal, *out);
        return NGSB_OK;
    }
    ~BufLease() { for (auto &b : held) ws_put_buf(ctx, b.first, b.second); }
};
struct EventPair {
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~EventPair() { if (ev[0]) cudaEventDestroy(ev[0]); if (ev[1]) cudaEventDestroy(ev[1]); }
};

static void ws_free(void *p)
{
    Workspace *ws = (Workspace *)p;
    for (auto &b : ws->free_bufs) cudaFree(b.second);
    if (ws->d_state) cudaFree(ws->d_state);
    if (ws->h_state) cudaFreeHost(ws->h_state);
    if (ws->d_hist) cudaFree(ws->d_hist);
    if (ws->graph_exec) cudaGraphExecDestroy(ws->graph_exec);
    delete ws;
}

static Workspace *get_ws(ngsb_ctx *ctx)
{
    if (!ctx->ws) {
        ctx->ws = new Workspace();
        ctx->ws_free = ws_free;
    }
    return (Workspace *)ctx->ws;
}

int ws_get_buf(ngsb_ctx *ctx, size_t nscal, double **out)
{
    Workspace *ws = get_ws(ctx);
    for (size_t i = 0; i < ws->free_bufs.size(); i++)
        if (ws->free_bufs[i].first == nscal) {
            *out = ws->free_bufs[i].second;
            ws->free_bufs.erase(ws->free_bufs.begin() + i);
            return NGSB_OK;
        }
    void *p = nullptr;
    size_t bytes = nscal * sizeof(double);
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        // out of memory: release every cached buffer and retry once
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        for (auto &b : ws->free_bufs) cudaFree(b.second);
        ws->free_bufs.clear();
        e = cudaMalloc(&p, bytes ? bytes : 16);
    }
    if (e != cudaSuccess) { set_error("solver workspace: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return NGSB_ERR_NOMEM; }
    *out = (double *)p;
    return NGSB_OK;
}

void ws_put_buf(ngsb_ctx *ctx, size_t nscal, double *p)
{
    if (p) get_ws(ctx)->free_bufs.emplace_back(nscal, p);
}

static int ws_state_impl(ngsb_ctx *ctx, size_t hist_cap)
{
    Workspace *ws = get_ws(ctx);
    if (!ws->d_state) {
        NGSB_CUDA(cudaMalloc(&ws->d_state, sizeof(CgState)));
        NGSB_CUDA(cudaMallocHost(&ws->h_state, 4 * sizeof(CgState)));
    }
    if (ws->hist_cap < hist_cap + 1) {
        if (ws->d_hist) { cudaStreamSynchronize(ctx->stream); cudaFree(ws->d_hist); }
        NGSB_CUDA(cudaMalloc(&ws->d_hist, (hist_cap + 1) * sizeof(double)));
        ws->hist_cap = hist_cap + 1;
        if (ws->graph_exec) { cudaGraphExecDestroy(ws->graph_exec); ws->graph_exec = nullptr; }
    }
    return NGSB_OK;
}

int ws_state(ngsb_ctx *ctx, size_t hist_cap, CgState **d_state, CgState **h_state, double **d_hist)
{
    NGSB_TRY(ws_state_impl(ctx, hist_cap));
    Workspace *ws = get_ws(ctx);
    *d_state = ws->d_state;
    *h_state = ws->h_state;
    *d_hist = ws->d_hist;
    return NGSB_OK;
}

// ------------------------------------------------------------------------------------------
// fused CG kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bit_test_k(const uint8_t *bits, uint64_t i) { return (bits[i >> 3] >> (i & 7)) & 1; }

__device__ __forceinline__ double warp_sum_k(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// grid-wide deterministic reduction finish; returns true on thread 0 of the last block (WARP: on all 32 lanes of its first
// warp, every lane holding the total -- for finishes that talk to the peers with one lane per rank)
template <bool WARP = false>
__device__ __forceinline__ bool grid_finish(double a, double b, double *partials, unsigned int *counter, double2 *total)
{
    __shared__ double red[64];
    __shared__ int s_last;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_sum_k(a);
    b = warp_sum_k(b);
    if (lane == 0) { red[wid] = a; red[32 + wid] = b; }
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        a = lane < nw ? red[lane] : 0.0;
        b = lane < nw ? red[32 + lane] : 0.0;
        a = warp_sum_k(a);
        b = warp_sum_k(b);
        if (lane == 0) {
            partials[2 * blockIdx.x] = a;
            partials[2 * blockIdx.x + 1] = b;
            __threadfence();
            unsigned int t = atomicAdd(counter, 1u);
            s_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return false;
    __threadfence();
    a = 0.0;
    b = 0.0;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += 32) {
        a += __ldcg(&partials[2 * k]);
        b += __ldcg(&partials[2 * k + 1]);
    }
    a = warp_sum_k(a);
    b = warp_sum_k(b);
    if (threadIdx.x == 0) *counter = 0;
    if (WARP || threadIdx.x == 0) {
        *total = make_double2(a, b);
        return true;
    }
    return false;
}

// w = C*d for one entry (masked Jacobi) -- returns w, or d when there is no preconditioner
template <int KIND>
__device__ __forceinline__ void prec_entry(const CgVecs &v, uint64_t i, const double *dn, double *wn)
{
    if (v.invdiag == nullptr) {
        wn[0] = dn[0];
        if (KIND != NGSB_REAL) wn[1] = dn[1];
        if (KIND == NGSB_BLOCK3) wn[2] = dn[2];
        return;
    }
    const bool in = v.bits == nullptr || bit_test_k(v.bits, i);
    if (KIND == NGSB_REAL) {
        wn[0] = in ? v.invdiag[i] * dn[0] : 0.0;
    } else if (KIND == NGSB_COMPLEX) {
        if (in) {
            double2 m = reinterpret_cast<const double2 *>(v.invdiag)[i];
            wn[0] = m.x * dn[0] - m.y * dn[1];
            wn[1] = m.x * dn[1] + m.y * dn[0];
        } else { wn[0] = 0.0; wn[1] = 0.0; }
    } else {
        if (in) {
            const double *m = v.invdiag + 9 * i;
            wn[0] = m[0] * dn[0] + m[1] * dn[1] + m[2] * dn[2];
            wn[1] = m[3] * dn[0] + m[4] * dn[1] + m[5] * dn[2];
            wn[2] = m[6] * dn[0] + m[7] * dn[1] + m[8] * dn[2];
        } else { wn[0] = 0.0; wn[1] = 0.0; wn[2] = 0.0; }
    }
}

// MODE 0: init      d = f (or f - as when !initialize, flag in `sub`), w = C d, s = w, <w,d>
// MODE 1: update    u += al s, d -= al as, w = C d, <d,w>
// FOLD (MODE 1):    without `u += al s` -- the direction kernel of the same iteration does it (cg_dir_kernel<., true>)
template <int KIND, int MODE, bool FOLD = false>
__global__ void __launch_bounds__(256) cg_fused_kernel(const CgVecs v, int sub)
{
    constexpr int ES = KIND == NGSB_REAL ? 1 : (KIND == NGSB_COMPLEX ? 2 : 3);
    CgState *st = v.state;
    if (MODE == 1 && st->done) {
        if (FOLD && blockIdx.x == 0 && threadIdx.x == 0) st->u_pending = 0;     // the owed update ran in the previous iteration
        return;
    }
    const double alr = MODE == 1 ? st->al[0] : 0.0;
    const double ali = MODE == 1 ? st->al[1] : 0.0;
    const bool conj = v.ip_mode == NGSB_IP_COMPLEX_CONJ;
    // Work split (fixed for a given n and grid -> deterministic sums).  Default: grid-stride, i.e. at any moment all CTAs
    // read neighbouring addresses of each of the eight vectors -- eight DRAM streams instead of eight per CTA (the chunked
    // split ran at 0.82 of the copy rate, the grid-stride direction kernel at 0.95; profiles/r1_ncu_launches_bench_n1_summary.txt).
    // v.chunked: one contiguous chunk per block, even length so that pairs never straddle chunks (kept for A/B).
    uint64_t lo, hi, first, step;
    if (v.chunked) {
        uint64_t per = ((v.n + gridDim.x - 1) / gridDim.x + 1) & ~(uint64_t)1;
        lo = (uint64_t)blockIdx.x * per;
        hi = lo + per < v.n ? lo + per : v.n;
        if (lo > hi) lo = hi;
        first = threadIdx.x; step = blockDim.x;
    } else {
        lo = 0; hi = v.n;
        first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; step = (uint64_t)gridDim.x * blockDim.x;
    }
    double accr = 0.0, acci = 0.0;
    if (KIND == NGSB_REAL && MODE == 1 && v.invdiag != nullptr &&
        ((reinterpret_cast<uintptr_t>(v.u) | reinterpret_cast<uintptr_t>(v.d) | reinterpret_cast<uintptr_t>(v.w) |
          reinterpret_cast<uintptr_t>(v.s) | reinterpret_cast<uintptr_t>(v.as) | reinterpret_cast<uintptr_t>(v.invdiag)) & 15) == 0) {
        // the hot case (real Jacobi-PCG): two entries per thread and step, 128-bit accesses, five streams in,
        // three out; identical arithmetic to the scalar loop below
        double acc1 = 0.0;
        const uint64_t hi2 = lo + ((hi - lo) & ~(uint64_t)1);
#pragma unroll 2
        for (uint64_t i = lo + 2 * first; i < hi2; i += 2 * step) {
            // v.stream (option cg_stream_hints): u, d, As and the diagonal are not touched again before the next iteration's
            // update -- load them streaming and store u, d evict-first, so that the L2 keeps w (read by the direction kernel
            // next) and s (gathered by the next product) instead of 2 x N dirty doubles nobody asks for
            const double2 a2 = v.stream ? __ldcs(reinterpret_cast<const double2 *>(v.as + i)) : *reinterpret_cast<const double2 *>(v.as + i);
            const double2 m2 = v.stream ? __ldcs(reinterpret_cast<const double2 *>(v.invdiag + i)) : *reinterpret_cast<const double2 *>(v.invdiag + i);
            double2 d2 = v.stream ? __ldcs(reinterpret_cast<const double2 *>(v.d + i)) : *reinterpret_cast<double2 *>(v.d + i);
            if (!FOLD) {
                const double2 s2 = *reinterpret_cast<const double2 *>(v.s + i);
                double2 u2 = v.stream ? __ldcs(reinterpret_cast<const double2 *>(v.u + i)) : *reinterpret_cast<double2 *>(v.u + i);
                u2.x += alr * s2.x; u2.y += alr * s2.y;
                if (v.stream) __stcs(reinterpret_cast<double2 *>(v.u + i), u2);
                else *reinterpret_cast<double2 *>(v.u + i) = u2;
            }
            d2.x -= alr * a2.x; d2.y -= alr * a2.y;
            unsigned bits = v.bits ? (unsigned)(v.bits[i >> 3] >> (i & 7)) : 3u;
            double2 w2;
            w2.x = (bits & 1u) ? m2.x * d2.x : 0.0;
            w2.y = (bits & 2u) ? m2.y * d2.y : 0.0;
            if (v.stream) __stcs(reinterpret_cast<double2 *>(v.d + i), d2);
            else *reinterpret_cast<double2 *>(v.d + i) = d2;
            *reinterpret_cast<double2 *>(v.w + i) = w2;
            if (v.master == nullptr || v.master[i]) accr = fma(d2.x, w2.x, accr);
            if (v.master == nullptr || v.master[i + 1]) acc1 = fma(d2.y, w2.y, acc1);
        }
        if (hi2 < hi && first == 0) {            // odd tail (of the last chunk / of the vector)
            const uint64_t i = hi2;
            if (!FOLD) v.u[i] += alr * v.s[i];
            const double dn = v.d[i] - alr * v.as[i];
            const double wn = (v.bits == nullptr || bit_test_k(v.bits, i)) ? v.invdiag[i] * dn : 0.0;
            v.d[i] = dn;
            v.w[i] = wn;
            if (v.master == nullptr || v.master[i]) accr = fma(dn, wn, accr);
        }
        accr += acc1;
        lo = hi;                                  // skip the generic loop
    }
    for (uint64_t i = lo + first; i < hi; i += step) {
        double dn[3], wn[3];
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < ES; c++) {
                dn[c] = v.f[ES * i + c];
                if (sub) dn[c] -= v.as[ES * i + c];
            }
        } else {
            if (KIND == NGSB_COMPLEX) {
                double ar = v.as[2 * i], ai = v.as[2 * i + 1];
                if (!FOLD) {
                    double sr = v.s[2 * i], si = v.s[2 * i + 1];
                    v.u[2 * i] += alr * sr - ali * si;
                    v.u[2 * i + 1] += alr * si + ali * sr;
                }
                dn[0] = v.d[2 * i] - (alr * ar - ali * ai);
                dn[1] = v.d[2 * i + 1] - (alr * ai + ali * ar);
            } else {
#pragma unroll
                for (int c = 0; c < ES; c++) {
                    if (!FOLD) v.u[ES * i + c] += alr * v.s[ES * i + c];
                    dn[c] = v.d[ES * i + c] - alr * v.as[ES * i + c];
                }
            }
        }
        prec_entry<KIND>(v, i, dn, wn);
#pragma unroll
        for (int c = 0; c < ES; c++) {
            v.d[ES * i + c] = dn[c];
            if (v.invdiag != nullptr) v.w[ES * i + c] = wn[c];
            if (MODE == 0) v.s[ES * i + c] = wn[c];
        }
        if (v.master != nullptr && !v.master[i]) continue;
        if (KIND == NGSB_COMPLEX) {
            // init: <w, d> (conj on d) ; update: <d, w> (conj on w)
            double xr = MODE == 0 ? wn[0] : dn[0], xi = MODE == 0 ? wn[1] : dn[1];
            double yr = MODE == 0 ? dn[0] : wn[0], yi = MODE == 0 ? dn[1] : wn[1];
            if (conj) yi = -yi;
            accr += xr * yr - xi * yi;
            acci += xr * yi + xi * yr;
        } else {
#pragma unroll
            for (int c = 0; c < ES; c++) accr = fma(dn[c], wn[c], accr);
        }
    }
    double2 total;
    if (grid_finish<true>(accr, acci, v.partials, v.counter, &total)) {
        if (v.R != nullptr) {
            // distributed: the first warp of the last block all-reduces the dot over the ranks itself, one lane per rank
            pr_push_warp(*v.R, total.x, total.y);
            total = pr_wait_sum_warp(*v.R);
        }
        if (threadIdx.x == 0) {
            if (v.R == nullptr && v.dot_out != nullptr) { v.dot_out[0] = total.x; v.dot_out[1] = total.y; }
            else if (MODE == 0) cg_finalize_init(st, total, v.hist);
            else cg_finalize_wdn(st, total, v.hist);
        }
    }
}

// s = be*s + w   (`s *= be; s += w`, linalg/cg.cpp:611-612: two roundings, kept)
// FOLD: also u += al*s with the s of this iteration (read before it is overwritten) -- the update the fused kernel left
// out; when the loop ended in this iteration (done && u_pending) only that update runs.
template <bool CPLX, bool FOLD>
__global__ void __launch_bounds__(256) cg_dir_kernel(double *__restrict__ s, const double *__restrict__ w, double *__restrict__ u, uint64_t N,
                                                    const CgState *__restrict__ st)
{
    const bool only_u = st->done != 0;
    if (only_u && !(FOLD && st->u_pending)) return;
    const double ber = st->be[0], bei = st->be[1];
    const double alr = st->al[0], ali = st->al[1];
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (CPLX) {
        double2 *s2 = reinterpret_cast<double2 *>(s);
        const double2 *w2 = reinterpret_cast<const double2 *>(w);
        for (; i < N; i += stride) {
            double2 a = s2[i];
            if (FOLD) {
                const double sr = a.x, si = a.y;
                u[2 * i] += alr * sr - ali * si;
                u[2 * i + 1] += alr * si + ali * sr;
                if (only_u) continue;
            }
            double2 b = w2[i];
            double pr = a.x * ber - a.y * bei, pi = a.x * bei + a.y * ber;
            s2[i] = make_double2(__dadd_rn(pr, b.x), __dadd_rn(pi, b.y));
        }
    } else {
        uint64_t n2 = N / 2;
        double2 *s2 = reinterpret_cast<double2 *>(s);
        double2 *u2p = reinterpret_cast<double2 *>(u);
        const double2 *w2 = reinterpret_cast<const double2 *>(w);
        for (uint64_t k = i; k < n2; k += stride) {
            double2 a = s2[k];
            if (FOLD) {
                double2 u2 = u2p[k];
                u2.x += alr * a.x; u2.y += alr * a.y;
                u2p[k] = u2;
                if (only_u) continue;
            }
            double2 b = w2[k];
            s2[k] = make_double2(__dadd_rn(__dmul_rn(a.x, ber), b.x), __dadd_rn(__dmul_rn(a.y, ber), b.y));
        }
        if (i == 0 && (N & 1)) {
            if (FOLD) u[N - 1] += alr * s[N - 1];
            if (!only_u) s[N - 1] = __dadd_rn(__dmul_rn(s[N - 1], ber), w[N - 1]);
        }
    }
}

static int reduce_grid(ngsb_ctx *ctx, uint64_t n)
{
    uint64_t blocks = (n + 2047) / 2048;
    uint64_t cap = (uint64_t)ctx->sm_count * 8;
    if (cap > (uint64_t)MAX_PARTIALS - 8) cap = MAX_PARTIALS - 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <int MODE>
static int launch_cg_fused(ngsb_ctx *ctx, int kind, const CgVecs &v, int sub)
{
    SpanGuard g(ctx, KC_CGUPDATE);
    int grid = reduce_grid(ctx, v.n);
    if (MODE == 1 && v.fold_u) {
        if (kind == NGSB_REAL) cg_fused_kernel<NGSB_REAL, 1, true><<<grid, 256, 0, ctx->stream>>>(v, sub);
        else if (kind == NGSB_COMPLEX) cg_fused_kernel<NGSB_COMPLEX, 1, true><<<grid, 256, 0, ctx->stream>>>(v, sub);
        else cg_fused_kernel<NGSB_BLOCK3, 1, true><<<grid, 256, 0, ctx->stream>>>(v, sub);
        NGSB_CUDA(cudaGetLastError());
        return NGSB_OK;
    }
    if (kind == NGSB_REAL) cg_fused_kernel<NGSB_REAL, MODE><<<grid, 256, 0, ctx->stream>>>(v, sub);
    else if (kind == NGSB_COMPLEX) cg_fused_kernel<NGSB_COMPLEX, MODE><<<grid, 256, 0, ctx->stream>>>(v, sub);
    else cg_fused_kernel<NGSB_BLOCK3, MODE><<<grid, 256, 0, ctx->stream>>>(v, sub);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

__global__ void cg_finalize_kernel(int which, CgState *st, double *dot, double *hist, const PeerReduce *R)
{
    if (which != 0 && st->done) return;
    if (R) pr_allreduce_inplace(*R, dot);        // distributed, peer-memory mode: sum of the ranks' partials
    double2 t = make_double2(dot[0], dot[1]);
    if (which == 0) cg_finalize_init(st, t, hist);
    else if (which == 1) cg_finalize_kss(st, t);
    else cg_finalize_wdn(st, t, hist);
}

int cg_launch_finalize(ngsb_ctx *ctx, int which, CgState *st, double *dot, double *hist, const PeerReduce *R)
{
    SpanGuard g(ctx, KC_OTHER);
    cg_finalize_kernel<<<1, 1, 0, ctx->stream>>>(which, st, dot, hist, R);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

int cg_launch_fused(ngsb_ctx *ctx, int kind, int mode, const CgVecs &v, int sub)
{
    return mode == 0 ? launch_cg_fused<0>(ctx, kind, v, sub) : launch_cg_fused<1>(ctx, kind, v, sub);
}

static int launch_cg_dir(ngsb_ctx *ctx, int kind, const CgVecs &v)
{
    SpanGuard g(ctx, KC_CGUPDATE);
    const bool cplx = kind == NGSB_COMPLEX;
    uint64_t N = cplx ? v.n : v.n * kind_scalars(kind);
    uint64_t blocks = (N + 1023) / 1024;
    uint64_t cap = (uint64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const double *w = v.invdiag ? v.w : v.d;
    if (v.fold_u) {
        if (cplx) cg_dir_kernel<true, true><<<(int)blocks, 256, 0, ctx->stream>>>(v.s, w, v.u, N, v.state);
        else cg_dir_kernel<false, true><<<(int)blocks, 256, 0, ctx->stream>>>(v.s, w, v.u, N, v.state);
    } else {
        if (cplx) cg_dir_kernel<true, false><<<(int)blocks, 256, 0, ctx->stream>>>(v.s, w, v.u, N, v.state);
        else cg_dir_kernel<false, false><<<(int)blocks, 256, 0, ctx->stream>>>(v.s, w, v.u, N, v.state);
    }
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

int cg_launch_dir(ngsb_ctx *ctx, int kind, const CgVecs &v) { return launch_cg_dir(ctx, kind, v); }

static int enqueue_iteration(ngsb_ctx *ctx, const ngsb_csr *A, const CgVecs &v, double *as)
{
    SpmvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.x = v.s; a.y = as; a.sr = 1.0; a.si = 0.0; a.accumulate = false;
    a.epi = EPI_CG_KSS; a.dotvec = v.s; a.dot_conj = v.ip_mode == NGSB_IP_COMPLEX_CONJ; a.state = v.state;
    NGSB_TRY(spmv_launch(a));
    NGSB_TRY(launch_cg_fused<1>(ctx, A->kind, v, 0));
    NGSB_TRY(launch_cg_dir(ctx, A->kind, v));
    return NGSB_OK;
}

static bool sell_path(const ngsb_ctx *ctx) { return ctx->spmv_algo == 0 || ctx->spmv_algo == 3; }

int cg_solve_device(const ngsb_csr *A, const ngsb_jacobi *C, const double *f, double *u, double prec, int maxsteps,
                    int ip_mode, int initialize, int *steps, double *history, int hist_cap, int *nhist)
{
    ngsb_ctx *ctx = A->ctx;
    Workspace *ws = get_ws(ctx);
    const size_t nscal = A->h * kind_scalars(A->kind);
    if (hist_cap < 0) hist_cap = 0;
    if (!history) hist_cap = 0;
    NGSB_TRY(ws_state_impl(ctx, (size_t)hist_cap));
    BufLease lease(ctx);
    // internally reordered matrix (reorder.cu): the whole loop runs in the permuted numbering -- f and the start value are
    // gathered once, every iteration multiplies P A P^T directly (no per-product gather), u is scattered back at the end
    const ngsb_csr *Auser = A;
    double *u_user = u;
    if (A->inner && sell_path(ctx)) {
        const int es = (int)kind_scalars(A->kind);
        double *fp = nullptr, *up = nullptr;
        NGSB_TRY(lease.get(nscal, &fp));
        NGSB_TRY(lease.get(nscal, &up));
        NGSB_TRY(launch_perm_gather(ctx, f, A->d_perm, A->h, es, fp));
        if (!initialize) NGSB_TRY(launch_perm_gather(ctx, u, A->d_perm, A->h, es, up));
        if (C) NGSB_TRY(jacobi_for_inner(C, A, &C));
        f = fp; u = up; A = A->inner;
    }
    double *w = nullptr, *s = nullptr, *d = nullptr, *as = nullptr;
    NGSB_TRY(lease.get(nscal, &s));
    NGSB_TRY(lease.get(nscal, &d));
    NGSB_TRY(lease.get(nscal, &as));
    if (C) NGSB_TRY(lease.get(nscal, &w));

    CgState *hs = ws->h_state;
    memset(hs, 0, sizeof(CgState));
    hs->prec2 = prec * prec;
    hs->maxsteps = maxsteps;
    hs->hist_cap = hist_cap;
    hs->cplx = ip_mode != NGSB_IP_REAL;
    hs->done = 0;
    NGSB_CUDA(cudaMemcpyAsync(ws->d_state, hs, sizeof(CgState), cudaMemcpyHostToDevice, ctx->stream));

    CgVecs v;
    memset(&v, 0, sizeof(v));
    v.u = u; v.d = d; v.w = w; v.s = s; v.as = as; v.f = f;
    v.invdiag = C ? C->d_invdiag : nullptr;
    v.bits = C ? C->d_bits : nullptr;
    v.n = A->h;
    v.state = ws->d_state;
    v.hist = ws->d_hist;
    v.partials = ctx->d_partials;
    v.counter = ctx->d_counter;
    v.ip_mode = ip_mode;
    v.fold_u = ctx->cg_fold_u ? 1 : 0;
    v.chunked = ctx->cg_chunked ? 1 : 0;
    v.stream = ctx->cg_stream_hints ? 1 : 0;

    int sub = 0;
    if (initialize) {
        NGSB_CUDA(cudaMemsetAsync(u, 0, nscal * sizeof(double), ctx->stream));   // u = 0.0
    } else {
        SpmvArgs a;
        memset(&a, 0, sizeof(a));
        a.A = A; a.x = u; a.y = as; a.sr = 1.0; a.accumulate = false; a.epi = EPI_NONE;
        NGSB_TRY(spmv_launch(a));
        sub = 1;                                                                // d = f - A*u
    }
    NGSB_TRY(launch_cg_fused<0>(ctx, A->kind, v, sub));

    // Small real systems: the whole loop as one persistent cooperative kernel (sell.cu, cg_persistent_kernel) -- grid barriers
    // instead of three kernel boundaries per iteration.  Automatic below 4 M rows (above, the wide grid of the product
    // kernel and its sliding x windows win); option cg_persistent 0 / 1 forces it off / on.
    {
        const bool want = ctx->cg_persistent == 1 || (ctx->cg_persistent < 0 && A->h < (4u << 20));
        if (want && !ctx->timing && sell_path(ctx) && cg_persistent_applicable(A, v)) {
            int rc = cg_persistent_launch(A, v, maxsteps + 1);
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (rc == NGSB_OK && e != cudaSuccess) { set_error("CG (persistent kernel): %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; }
            if (rc == NGSB_OK && Auser != A) rc = launch_perm_gather(ctx, u, Auser->d_iperm, Auser->h, (int)kind_scalars(A->kind), u_user);
            if (rc == NGSB_OK) {
                NGSB_CUDA(cudaMemcpyAsync(&hs[3], ws->d_state, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream));
                NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
                if (steps) *steps = hs[3].n;
                const int nh = hs[3].nhist;
                if (nhist) *nhist = nh;
                const int ncopy = nh < hist_cap ? nh : hist_cap;
                if (history && ncopy > 0) {
                    NGSB_CUDA(cudaMemcpyAsync(history, ws->d_hist, ncopy * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
                }
                ctx->launches += 1;
            }
            return rc;
        }
    }
    // iterate in batches; the stop flag is polled one batch behind the enqueue front
    const long batch = ctx->cg_batch;
    const bool use_graph = !ctx->timing && getenv("NGSB_NO_CUDA_GRAPH") == nullptr && batch > 1;
    if (use_graph) {
        double *key[6] = {u, d, w, s, as, (double *)f};
        const long opts[11] = {batch, (long)ip_mode, ctx->spmv_algo, ctx->spmv_ctas_per_sm, ctx->cg_fold_u, ctx->sell_variant, ctx->sell_c16,
                               ctx->sell_c16_all, ctx->sell_pf_steps, ctx->sell_pf_next, ctx->cg_chunked | (ctx->cg_stream_hints << 1)};
        bool hit = ws->graph_exec && ws->g_A == A->uid && ws->g_C == (C ? C->uid : 0) && memcmp(opts, ws->g_opts, sizeof(opts)) == 0 &&
                   memcmp(key, ws->g_ptrs, sizeof(key)) == 0;
        if (!hit) {
            cudaGraph_t graph = nullptr;
            uint64_t launches_before = ctx->launches;
            NGSB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            int rc = NGSB_OK;
            for (long k = 0; k < batch && rc == NGSB_OK; k++) rc = enqueue_iteration(ctx, A, v, as);
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
            ctx->launches = launches_before;
            if (rc != NGSB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            NGSB_CUDA(ce);
            // same topology with other pointers (a script's `inv * f` hands in a fresh result vector every time): update the
            // instantiated graph in place, which costs microseconds; instantiate only when that is refused
            bool updated = false;
            if (ws->graph_exec) {
                cudaGraphExecUpdateResultInfo info;
                if (cudaGraphExecUpdate(ws->graph_exec, graph, &info) == cudaSuccess) updated = true;
                else { cudaGetLastError(); cudaGraphExecDestroy(ws->graph_exec); ws->graph_exec = nullptr; }
            }
            if (!updated) NGSB_CUDA(cudaGraphInstantiate(&ws->graph_exec, graph, 0));
            cudaGraphDestroy(graph);
            ws->g_A = A->uid; ws->g_C = C ? C->uid : 0;
            memcpy(ws->g_opts, opts, sizeof(opts));
            memcpy(ws->g_ptrs, key, sizeof(key));
        }
    }

    EventPair evp;
    cudaEvent_t *ev = evp.ev;
    NGSB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    NGSB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    long enq = 0;          // batches enqueued
    bool finished = false;
    int rc = NGSB_OK;
    // upper bound of batches ever needed
    const long max_batches = ((long)maxsteps + batch - 1) / batch + 1;
    while (!finished) {
        if (enq < max_batches) {
            if (use_graph) {
                cudaError_t e = cudaGraphLaunch(ws->graph_exec, ctx->stream);
                if (e != cudaSuccess) { set_error("cudaGraphLaunch failed: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; break; }
                ctx->launches += 3 * batch;
            } else {
                for (long k = 0; k < batch && rc == NGSB_OK; k++) rc = enqueue_iteration(ctx, A, v, as);
                if (rc != NGSB_OK) break;
            }
        }
        cudaMemcpyAsync(&hs[1 + (enq & 1)], ws->d_state, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(ev[enq & 1], ctx->stream);
        if (enq > 0) {
            // look at the state copied after the previous batch
            cudaError_t e = cudaEventSynchronize(ev[(enq - 1) & 1]);
            if (e != cudaSuccess) { set_error("CG: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; break; }
            if (hs[1 + ((enq - 1) & 1)].done) finished = true;
        }
        enq++;
        if (!finished && enq > max_batches + 1) {
            set_error("CG: device loop did not terminate");
            rc = NGSB_ERR_CUDA;
            break;
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == NGSB_OK && e != cudaSuccess) { set_error("CG: %s", cudaGetErrorString(e)); rc = NGSB_ERR_CUDA; }
    if (rc == NGSB_OK && Auser != A)       // back into the caller's numbering: u_user[j] = u[iperm[j]]
        rc = launch_perm_gather(ctx, u, Auser->d_iperm, Auser->h, (int)kind_scalars(A->kind), u_user);
    if (rc == NGSB_OK) {
        NGSB_CUDA(cudaMemcpyAsync(&hs[3], ws->d_state, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream));
        NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (steps) *steps = hs[3].n;
        int nh = hs[3].nhist;
        if (nhist) *nhist = nh;
        int ncopy = nh < hist_cap ? nh : hist_cap;
        if (history && ncopy > 0) {
            NGSB_CUDA(cudaMemcpyAsync(history, ws->d_hist, ncopy * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    return rc;
}

// ------------------------------------------------------------------------------------------
// GMRES
// ------------------------------------------------------------------------------------------
struct GmresState {
    double norm, err, prec;
    int j;            // the reference's loop variable
    int done;
    int maxsteps;
    int cplx;
    int nhist, hist_cap;
    double tmp[2];    // reduction result slot
    double tmp2[2];
    double scale[2];  // scalar for the next "v = scale * w"
};

__device__ __forceinline__ double2 z_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 z_div(double2 a, double2 b)
{
    if (b.y == 0.0 && a.y == 0.0) return make_double2(a.x / b.x, 0.0);
    double den = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
}
__device__ __forceinline__ double2 z_sqrt(double2 z, int cplx)
{
    if (!cplx) return make_double2(sqrt(z.x), 0.0);
    double r = hypot(z.x, z.y);
    if (r == 0.0) return make_double2(0.0, 0.0);
    double re, im;
    if (z.x >= 0.0) { re = sqrt(0.5 * (r + z.x)); im = z.y / (2.0 * re); }
    else { im = sqrt(0.5 * (r - z.x)); if (z.y < 0.0) im = -im; re = z.y / (2.0 * im); }
    return make_double2(re, im);
}

// after norm2 = sum |r|^2 (tmp) and rr = <r,r> (tmp2): cg.cpp:889-903
__global__ void gmres_init_kernel(GmresState *st, double2 *gammai, double *hist, const PeerReduce *R)
{
    if (R) { pr_allreduce_inplace(*R, st->tmp); pr_allreduce_inplace(*R, st->tmp2); }
    st->norm = sqrt(st->tmp[0]);
    double2 rr = make_double2(st->tmp2[0], st->cplx ? st->tmp2[1] : 0.0);
    double2 sq = z_sqrt(rr, st->cplx);
    double2 sc = z_div(make_double2(1.0, 0.0), sq);
    st->scale[0] = sc.x;
    st->scale[1] = sc.y;
    gammai[0] = make_double2(st->norm, 0.0);
    if (st->hist_cap > 0) hist[0] = st->norm;
    st->nhist = 1;
    st->err = st->prec * fabs(st->norm);
    st->j = -1;
    st->done = 0;
    bool cont = (st->j++ < st->maxsteps - 2) && (st->norm > st->err);
    if (!cont) st->done = 1;
}

// h(i,j) = tmp; publish -h(i,j) as the next axpy scalar (kept in hcol[i])
__global__ void gmres_store_h_kernel(GmresState *st, double2 *h, int ms, int i, const PeerReduce *R)
{
    if (st->done) return;
    if (R) pr_allreduce_inplace(*R, st->tmp);
    int j = st->j;
    h[(size_t)i * ms + j] = make_double2(st->tmp[0], st->cplx ? st->tmp[1] : 0.0);
}

// after <w,w> (tmp): h(j+1,j), scale, Givens, norm, loop condition -- cg.cpp:934-962
__global__ void gmres_givens_kernel(GmresState *st, double2 *h, double2 *gammai, double2 *ci, double2 *si, int ms, double *hist,
                                    const PeerReduce *R)
{
    if (st->done) return;
    if (R) pr_allreduce_inplace(*R, st->tmp);
    const int j = st->j;
    const int cplx = st->cplx;
#define H(a, b) h[(size_t)(a) * ms + (b)]
    double2 ww = make_double2(st->tmp[0], cplx ? st->tmp[1] : 0.0);
    H(j + 1, j) = z_sqrt(ww, cplx);
    double2 sc = z_div(make_double2(1.0, 0.0), H(j + 1, j));
    st->scale[0] = sc.x;
    st->scale[1] = sc.y;
    for (int i = 0; i < j; i++) {
        double2 hi = H(i, j), hip = H(i + 1, j);
        double2 a = z_mul(ci[i + 1], hi), b = z_mul(si[i + 1], hip);
        double2 c = z_mul(si[i + 1], hi), d = z_mul(ci[i + 1], hip);
        H(i, j) = make_double2(a.x + b.x, a.y + b.y);
        H(i + 1, j) = make_double2(c.x - d.x, c.y - d.y);
    }
    double2 hjj = H(j, j), hj1 = H(j + 1, j);
    double2 q1 = z_mul(hjj, hjj), q2 = z_mul(hj1, hj1);
    double2 beta = z_sqrt(make_double2(q1.x + q2.x, q1.y + q2.y), cplx);
    si[j + 1] = z_div(hj1, beta);
    ci[j + 1] = z_div(hjj, beta);
    H(j, j) = beta;
    gammai[j + 1] = z_mul(si[j + 1], gammai[j]);
    gammai[j] = z_mul(ci[j + 1], gammai[j]);
    st->norm = cplx ? hypot(gammai[j].x, gammai[j].y) : fabs(gammai[j].x);
    if (st->nhist < st->hist_cap) hist[st->nhist] = st->norm;
    st->nhist++;
    bool cont = (st->j++ < st->maxsteps - 2) && (st->norm > st->err);
    if (!cont) st->done = 1;
#undef H
}

// back substitution (cg.cpp:964-973); y[i] for i <= jfinal, jfinal = j-1 after the loop
__global__ void gmres_backsolve_kernel(GmresState *st, const double2 *h, const double2 *gammai, double2 *y, int ms)
{
    int j = st->j - 1;
    for (int i = j; i >= 0; i--) {
        double2 sum = gammai[i];
        for (int k = i + 1; k <= j; k++) {
            double2 p = z_mul(h[(size_t)i * ms + k], y[k]);
            sum.x -= p.x;
            sum.y -= p.y;
        }
        y[i] = z_div(sum, h[(size_t)i * ms + i]);
    }
}

// w -= hprev * vprev (if vprev), then partial <vnext, w> (bilinear) -> tmp.  selfdot: vnext = w.
template <bool CPLX>
__global__ void __launch_bounds__(256) gmres_mgs_kernel(double *__restrict__ w, const double *__restrict__ vprev,
                                                       const double2 *__restrict__ hprev, const double *__restrict__ vnext,
                                                       int selfdot, uint64_t N, GmresState *st, double *partials,
                                                       unsigned int *counter, const uint8_t *__restrict__ master, unsigned mask_div)
{
    if (st->done) return;
    double hr = 0.0, hi = 0.0;
    if (vprev) { hr = hprev->x; hi = CPLX ? hprev->y : 0.0; }
    uint64_t per = (N + gridDim.x - 1) / gridDim.x;
    uint64_t lo = (uint64_t)blockIdx.x * per;
    uint64_t hi_ = lo + per < N ? lo + per : N;
    double ar = 0.0, ai = 0.0;
    if (CPLX) {
        // two entries per thread and step: six independent 16-byte loads in flight (the loop is pure streaming)
        double2 *w2 = reinterpret_cast<double2 *>(w);
        const double2 *p2 = reinterpret_cast<const double2 *>(vprev);
        const double2 *q2 = reinterpret_cast<const double2 *>(vnext);
        double br = 0.0, bi = 0.0;
        for (uint64_t i = lo + threadIdx.x; i < hi_; i += 2 * (uint64_t)blockDim.x) {
            const uint64_t k = i + blockDim.x;
            const bool two = k < hi_;
            double2 wa = w2[i], wb = two ? w2[k] : make_double2(0.0, 0.0);
            if (vprev) {
                const double2 pa = p2[i], pb = two ? p2[k] : make_double2(0.0, 0.0);
                wa.x -= hr * pa.x - hi * pa.y;
                wa.y -= hr * pa.y + hi * pa.x;
                wb.x -= hr * pb.x - hi * pb.y;
                wb.y -= hr * pb.y + hi * pb.x;
                w2[i] = wa;
                if (two) w2[k] = wb;
            }
            if (!master || master[i]) {      // masked inner product: master dofs only
                const double2 q = selfdot ? wa : q2[i];
                ar += q.x * wa.x - q.y * wa.y;
                ai += q.x * wa.y + q.y * wa.x;
            }
            if (two && (!master || master[k])) {
                const double2 q = selfdot ? wb : q2[k];
                br += q.x * wb.x - q.y * wb.y;
                bi += q.x * wb.y + q.y * wb.x;
            }
        }
        ar += br;
        ai += bi;
    } else {
        double br = 0.0;
        for (uint64_t i = lo + threadIdx.x; i < hi_; i += 2 * (uint64_t)blockDim.x) {
            const uint64_t k = i + blockDim.x;
            const bool two = k < hi_;
            double wa = w[i], wb = two ? w[k] : 0.0;
            if (vprev) {
                wa -= hr * vprev[i];
                w[i] = wa;
                if (two) { wb -= hr * vprev[k]; w[k] = wb; }
            }
            if (!master || master[i / mask_div]) ar = fma(selfdot ? wa : vnext[i], wa, ar);
            if (two && (!master || master[k / mask_div])) br = fma(selfdot ? wb : vnext[k], wb, br);
        }
        ar += br;
    }
    double2 total;
    if (grid_finish(ar, ai, partials, counter, &total)) {
        st->tmp[0] = total.x;
        st->tmp[1] = total.y;
    }
}

// x += y[i] * v_i
template <bool CPLX>
__global__ void __launch_bounds__(256) gmres_update_kernel(double *__restrict__ x, const double *__restrict__ vi, const double2 *__restrict__ yi,
                                                          uint64_t N, const GmresState *st, int i)
{
    if (i > st->j - 1) return;
    double yr = yi->x, yim = yi->y;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
        if (CPLX) {
            double2 a = reinterpret_cast<const double2 *>(vi)[k];
            double2 o = reinterpret_cast<double2 *>(x)[k];
            o.x += yr * a.x - yim * a.y;
            o.y += yr * a.y + yim * a.x;
            reinterpret_cast<double2 *>(x)[k] = o;
        } else x[k] += yr * vi[k];
    }
}


// ------------------------------------------------------------------------------------------
// GMRES, orthogonalisation with ONE batched reduction per step (option gmres_orth = 1)
//
// The reference's loop (cg.cpp:927-932) is modified Gram-Schmidt: h_i = <v_i, w>, w -= h_i v_i, one after the other -- j+1
// dependent reductions and, however the two operations are fused, two reads of every v_i plus a read and a write of w per
// projection (4 (j+1) + O(1) vector passes per step).  MGS applies (I - v_j v_j^T) ... (I - v_0 v_0^T) to w; written out,
// its coefficients solve the unit lower triangular system
//      (I + L) h = V^T w,      L_ik = <v_i, v_k>  (k < i),
// so the same h -- identical in exact arithmetic, and with the same O(eps * kappa) loss of orthogonality, since the
// triangular solve is carried out exactly instead of being truncated -- comes from
//   pass 1:  V^T w together with the new row of L (<v_k, v_j>, k < j): every v_k is read ONCE for both dots, the per-CTA
//            partials are summed in CTA order by the finish kernel (bitwise reproducible), which also runs the forward
//            substitution and stores H(0..j, j);
//   pass 2:  w -= sum_k h_k v_k in ascending k (each v_k read once, w read and written once) fused with <w, w>.
// That is 2 (j+1) + O(1) passes and two reductions per step, the byte model SURVEY 8(d) states for GMRES.
// All inner products are the reference's bilinear ones (no conjugation: GMRESSolver<Complex> uses S_InnerProduct).
// ------------------------------------------------------------------------------------------
__global__ void gmres_set_ptr_kernel(const double **tab, int idx, const double *p) { tab[idx] = p; }

// pass 1, one tile of T basis vectors: out[blockIdx.x][32] = per-CTA partials of (<v_k, w>, <v_k, v_j>), k = k0 .. k0+nk-1
template <bool CPLX>
__global__ void __launch_bounds__(256, CPLX ? 2 : 1) gmres_dots_kernel(const double *const *__restrict__ vtab, int k0, int nk, int jlast,
                                                        const double *__restrict__ w, uint64_t nscal, const GmresState *st,
                                                        double *__restrict__ out, const uint8_t *__restrict__ master, unsigned mask_div)
{
    constexpr int T = CPLX ? 8 : 16;          // T * (CPLX ? 4 : 2) = 32 accumulators per thread
    if (st->done) return;
    const double *p[T];
#pragma unroll
    for (int t = 0; t < T; t++) p[t] = vtab[k0 + (t < nk ? t : nk - 1)];     // short tile: repeat the last vector (same lines, L1 hits)
    const double *vj = vtab[jlast];
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0.0;
    const uint64_t ne = (nscal + 1) >> 1;       // pairs of doubles (one complex entry, or two real ones)
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t e = (uint64_t)blockIdx.x * 256 + threadIdx.x; e < ne; e += stride) {
        double2 a, b, v[T];
        if (2 * e + 1 < nscal) {
            a = *reinterpret_cast<const double2 *>(w + 2 * e);
            b = *reinterpret_cast<const double2 *>(vj + 2 * e);
#pragma unroll
            for (int t = 0; t < T; t++) v[t] = *reinterpret_cast<const double2 *>(p[t] + 2 * e);
        } else {                                 // odd real length: the last entry alone
            a = make_double2(w[2 * e], 0.0);
            b = make_double2(vj[2 * e], 0.0);
#pragma unroll
            for (int t = 0; t < T; t++) v[t] = make_double2(p[t][2 * e], 0.0);
        }
        if (master) {                            // parallel vectors: inner products over the master dofs only
            if (CPLX) {
                if (!master[e]) a = b = make_double2(0.0, 0.0);
            } else {
                if (!master[(2 * e) / mask_div]) a.x = b.x = 0.0;
                if (2 * e + 1 < nscal && !master[(2 * e + 1) / mask_div]) a.y = b.y = 0.0;
            }
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
            if (CPLX) {
                acc[4 * t + 0] = fma(-v[t].y, a.y, fma(v[t].x, a.x, acc[4 * t + 0]));
                acc[4 * t + 1] = fma(v[t].y, a.x, fma(v[t].x, a.y, acc[4 * t + 1]));
                acc[4 * t + 2] = fma(-v[t].y, b.y, fma(v[t].x, b.x, acc[4 * t + 2]));
                acc[4 * t + 3] = fma(v[t].y, b.x, fma(v[t].x, b.y, acc[4 * t + 3]));
            } else {
                acc[2 * t + 0] = fma(v[t].x, a.x, fma(v[t].y, a.y, acc[2 * t + 0]));
                acc[2 * t + 1] = fma(v[t].x, b.x, fma(v[t].y, b.y, acc[2 * t + 1]));
            }
        }
    }
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double mine = 0.0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const double s = warp_sum_k(acc[i]);
        if (lane == i) mine = s;                 // lane i keeps accumulator i
    }
    red[wid][lane] = mine;
    __syncthreads();
    if (wid == 0) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) s += red[q][lane];
        out[(size_t)blockIdx.x * 32 + lane] = s;
    }
}

// one CTA, blockDim >= nv: sums the partials of all tiles in CTA order, stores row j of L, solves (I + L) h = V^T w by forward
// substitution (thread k owns h_k; the subtractions run in ascending i, the order of the serial loop) and writes H(0..j, j)
__global__ void __launch_bounds__(1024) gmres_orth_finish_kernel(const GmresState *st, double2 *__restrict__ h, double2 *__restrict__ L,
                                                               const double *__restrict__ partials, int grid, int nv, int ms, int ldl, int cplx,
                                                               const PeerReduce *R)
{
    extern __shared__ double s_raw[];            // [nv][4]: <v_k,w> (re,im), <v_k,v_j> (re,im); then [nv] double2 h
    if (st->done) return;
    const int T = cplx ? 8 : 16, VALS = cplx ? 4 : 2;
    const int j = nv - 1;
    double2 *s_h = reinterpret_cast<double2 *>(s_raw + 4 * (size_t)nv);
    for (int idx = threadIdx.x; idx < nv * VALS; idx += blockDim.x) {
        const int k = idx / VALS, c = idx % VALS, tile = k / T, t = k % T;
        const double *p = partials + (size_t)tile * grid * 32 + t * VALS + c;
        double s = 0.0;
        for (int b = 0; b < grid; b++) s += p[(size_t)b * 32];
        if (cplx) s_raw[4 * k + c] = s;
        else { s_raw[4 * k + 2 * c] = s; s_raw[4 * k + 2 * c + 1] = 0.0; }
    }
    __syncthreads();
    if (R) pr_allreduce_vec_block(*R, s_raw, 4 * nv);        // distributed: all 2 (j+1) inner products in ONE exchange
    const int k = threadIdx.x;
    if (k < j) L[(size_t)j * ldl + k] = make_double2(s_raw[4 * k + 2], s_raw[4 * k + 3]);
    double2 mine = k <= j ? make_double2(s_raw[4 * k], s_raw[4 * k + 1]) : make_double2(0.0, 0.0);
    for (int i = 0; i < j; i++) {
        if (k == i) s_h[i] = mine;
        __syncthreads();
        if (k > i && k <= j) {
            const double2 l = k == j ? make_double2(s_raw[4 * i + 2], s_raw[4 * i + 3]) : L[(size_t)k * ldl + i];
            const double2 hi = s_h[i];
            mine.x -= l.x * hi.x - l.y * hi.y;
            mine.y -= l.x * hi.y + l.y * hi.x;
        }
    }
    if (k <= j) h[(size_t)k * ms + j] = mine;
}

// pass 2: w -= sum_k H(k,j) v_k (ascending k), fused with <w, w> -> st->tmp
template <bool CPLX>
__global__ void __launch_bounds__(256) gmres_project_kernel(double *__restrict__ w, const double *const *__restrict__ vtab,
                                                           const double2 *__restrict__ hcol, int ms, int nv, uint64_t nscal,
                                                           GmresState *st, double *partials, unsigned int *counter,
                                                           const uint8_t *__restrict__ master, unsigned mask_div)
{
    __shared__ double2 s_h[1024];
    __shared__ const double *s_p[1024];
    if (st->done) return;
    for (int k = threadIdx.x; k < nv; k += 256) { s_h[k] = hcol[(size_t)k * ms]; s_p[k] = vtab[k]; }
    __syncthreads();
    const uint64_t ne = (nscal + 1) >> 1;
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    double ar = 0.0, ai = 0.0;
    for (uint64_t e = (uint64_t)blockIdx.x * 256 + threadIdx.x; e < ne; e += stride) {
        const bool full = 2 * e + 1 < nscal;
        double2 a = full ? *reinterpret_cast<const double2 *>(w + 2 * e) : make_double2(w[2 * e], 0.0);
        int k = 0;
        if (full) {
            for (; k + 8 <= nv; k += 8) {
                double2 v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = *reinterpret_cast<const double2 *>(s_p[k + q] + 2 * e);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const double2 hk = s_h[k + q];
                    if (CPLX) {
                        a.x = fma(hk.y, v[q].y, fma(-hk.x, v[q].x, a.x));
                        a.y = fma(-hk.y, v[q].x, fma(-hk.x, v[q].y, a.y));
                    } else {
                        a.x = fma(-hk.x, v[q].x, a.x);
                        a.y = fma(-hk.x, v[q].y, a.y);
                    }
                }
            }
        }
        for (; k < nv; k++) {
            const double2 hk = s_h[k];
            const double2 v = full ? *reinterpret_cast<const double2 *>(s_p[k] + 2 * e) : make_double2(s_p[k][2 * e], 0.0);
            if (CPLX) {
                a.x = fma(hk.y, v.y, fma(-hk.x, v.x, a.x));
                a.y = fma(-hk.y, v.x, fma(-hk.x, v.y, a.y));
            } else {
                a.x = fma(-hk.x, v.x, a.x);
                a.y = fma(-hk.x, v.y, a.y);
            }
        }
        if (full) *reinterpret_cast<double2 *>(w + 2 * e) = a;
        else w[2 * e] = a.x;
        if (CPLX) {
            if (!master || master[e]) {
                ar += a.x * a.x - a.y * a.y;
                ai += a.x * a.y + a.y * a.x;
            }
        } else {
            if (master) {
                if (!master[(2 * e) / mask_div]) a.x = 0.0;
                if (full && !master[(2 * e + 1) / mask_div]) a.y = 0.0;
            }
            ar = fma(a.x, a.x, fma(a.y, a.y, ar));
        }
    }
    double2 total;
    if (grid_finish(ar, ai, partials, counter, &total)) {
        st->tmp[0] = total.x;
        st->tmp[1] = total.y;
    }
}

} // namespace ngsb

using namespace ngsb;

static int check_solver_args(const ngsb_csr *A, const ngsb_jacobi *C, const ngsb_vec *f, const ngsb_vec *u, const char *who)
{
    NGSB_REQUIRE(A && f && u, "%s: NULL argument", who);
    NGSB_REQUIRE(A->h == A->w, "%s: matrix must be square", who);
    NGSB_REQUIRE(f->ctx == A->ctx && u->ctx == A->ctx && (!C || C->ctx == A->ctx), "%s: objects belong to different contexts", who);
    NGSB_REQUIRE(f->kind == A->kind && u->kind == A->kind, "%s: vector kind does not match matrix kind", who);
    NGSB_REQUIRE(f->n == A->h && u->n == A->h, "%s: size of matrix = %zu, f = %zu, u = %zu", who, A->h, f->n, u->n);
    NGSB_REQUIRE(!C || (C->n == A->h && C->kind == A->kind), "%s: preconditioner does not match matrix", who);
    NGSB_REQUIRE(f->d != u->d, "%s: f and u must be different vectors", who);
    return NGSB_OK;
}

extern "C" int ngsb_cg_solve(const ngsb_csr *A, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *u, double prec, int maxsteps,
                             int ip_mode, int initialize, int *steps, double *history, int hist_cap, int *nhist)
{
    NvtxRange nv("CG solver");                      // the reference's Timer name, linalg/cg.cpp:511
    NGSB_TRY(check_solver_args(A, C, f, u, "CGSolver::Mult"));
    NGSB_REQUIRE(ip_mode >= 0 && ip_mode <= 2, "CGSolver::Mult: bad ip_mode %d", ip_mode);
    NGSB_REQUIRE((A->kind == NGSB_COMPLEX) == (ip_mode != NGSB_IP_REAL), "CGSolver::Mult: ip_mode %d does not fit matrix kind %d", ip_mode, A->kind);
    NGSB_REQUIRE(maxsteps >= 0, "CGSolver::Mult: maxsteps < 0");
    NGSB_CUDA(cudaSetDevice(A->ctx->device));
    return cg_solve_device(A, C, f->d, u->d, prec, maxsteps, ip_mode, initialize, steps, history, hist_cap, nhist);
}

extern "C" int ngsb_cg_solve_host(const ngsb_csr *A, const ngsb_jacobi *C, const void *f_host, void *u_host, double prec,
                                  int maxsteps, int ip_mode, int *steps, double *history, int hist_cap, int *nhist)
{
    NGSB_REQUIRE(A && f_host && u_host, "ngsb_cg_solve_host: NULL argument");
    ngsb_ctx *ctx = A->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const size_t nscal = A->h * kind_scalars(A->kind);
    double *f = nullptr, *u = nullptr;
    NGSB_TRY(ws_get_buf(ctx, nscal + 1, &f));   // +1: keep these apart from the solver's own buffers
    NGSB_TRY(ws_get_buf(ctx, nscal + 1, &u));
    ngsb_vec fv, uv;
    fv.ctx = ctx; fv.n = A->h; fv.kind = A->kind; fv.nscal = nscal; fv.d = f;
    uv = fv; uv.d = u;
    int rc = ngsb_vec_h2d(&fv, f_host, 0, A->h);
    if (rc == NGSB_OK) rc = ngsb_cg_solve(A, C, &fv, &uv, prec, maxsteps, ip_mode, 1, steps, history, hist_cap, nhist);
    if (rc == NGSB_OK) rc = ngsb_vec_d2h(&uv, u_host, 0, A->h);
    ws_put_buf(ctx, nscal + 1, f);
    ws_put_buf(ctx, nscal + 1, u);
    return rc;
}

extern "C" int ngsb_gmres_solve(const ngsb_csr *A, const ngsb_jacobi *C, const ngsb_vec *fvec, ngsb_vec *xvec, double prec,
                                int maxsteps, int initialize, int *steps, double *history, int hist_cap, int *nhist)
{
    return ngsb::gmres_solve_impl(A, C, fvec, xvec, prec, maxsteps, initialize, steps, history, hist_cap, nhist, nullptr);
}

// dist == NULL: one GPU.  Otherwise the reference's parallel-vector choreography: A v is DISTRIBUTED and is
// cumulated before the preconditioner, every Krylov vector is CUMULATED, inner products and the norm are
// restricted to master dofs and summed over the ranks (parallelvvector.cpp:289-358).
int ngsb::gmres_solve_impl(const ngsb_csr *A, const ngsb_jacobi *C, const ngsb_vec *fvec, ngsb_vec *xvec, double prec,
                           int maxsteps, int initialize, int *steps, double *history, int hist_cap, int *nhist, const GmresDist *dist)
{
    NvtxRange nv("GMRES solver");
    NGSB_TRY(check_solver_args(A, C, fvec, xvec, "GMRESSolver::Mult"));
    NGSB_REQUIRE(maxsteps >= 1, "GMRESSolver::Mult: maxsteps < 1");
    ngsb_ctx *ctx = A->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    if (A->inner && !dist && sell_path(ctx)) {
        // internally reordered matrix: the whole solve in the permuted numbering (see cg_solve_device)
        const size_t nsc = A->h * kind_scalars(A->kind);
        const int es = (int)kind_scalars(A->kind);
        BufLease lease(ctx);
        double *fp = nullptr, *xp = nullptr;
        NGSB_TRY(lease.get(nsc, &fp));
        NGSB_TRY(lease.get(nsc, &xp));
        NGSB_TRY(launch_perm_gather(ctx, fvec->d, A->d_perm, A->h, es, fp));
        if (!initialize) NGSB_TRY(launch_perm_gather(ctx, xvec->d, A->d_perm, A->h, es, xp));
        const ngsb_jacobi *Cp = nullptr;
        if (C) NGSB_TRY(jacobi_for_inner(C, A, &Cp));
        ngsb_vec fv = *fvec, xv = *xvec;
        fv.d = fp; fv.storage.reset(); xv.d = xp; xv.storage.reset();
        NGSB_TRY(gmres_solve_impl(A->inner, Cp, &fv, &xv, prec, maxsteps, initialize, steps, history, hist_cap, nhist, nullptr));
        return launch_perm_gather(ctx, xp, A->d_iperm, A->h, es, xvec->d);
    }
    const uint8_t *master = dist ? dist->master : nullptr;
    const PeerReduce *R = dist ? dist->R : nullptr;
    const unsigned mask_div = (unsigned)(A->kind == NGSB_BLOCK3 ? 3 : 1);
    // host-enqueued all-reduce of a device (re,im) pair (NCCL mode); peer-memory mode reduces inside the scalar kernels
    auto allred = [&](double *d_buf) -> int { return (dist && dist->allreduce) ? dist->allreduce(dist->arg, d_buf) : NGSB_OK; };
    auto cumulate = [&](double *v) -> int { return (dist && dist->cumulate) ? dist->cumulate(dist->arg, v) : NGSB_OK; };
    const bool cplx = A->kind == NGSB_COMPLEX;
    const size_t nscal = A->h * kind_scalars(A->kind);
    const uint64_t N = cplx ? A->h : nscal;        // reduction / update length in scalars of the IP type
    const int ms = maxsteps;
    if (hist_cap < 0 || !history) hist_cap = 0;

    // small device arrays
    GmresState *d_st = nullptr;
    double2 *d_h = nullptr, *d_gam = nullptr, *d_ci = nullptr, *d_si = nullptr, *d_y = nullptr;
    double *d_hist = nullptr, *d_scale = nullptr;
    const double **d_vtab = nullptr;      // one-reduction orthogonalisation: table of the basis vectors, L, per-CTA partials
    double2 *d_L = nullptr;
    double *d_dotp = nullptr;
    std::vector<double *> vi, chunks;     // basis vectors live in chunks of up to 8 (one cudaMalloc = one device sync)
    double *av = nullptr, *w = nullptr, *r = nullptr;
    int rc = NGSB_OK;
    GmresState hst;
    auto cleanup = [&]() {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_st); cudaFree(d_h); cudaFree(d_gam); cudaFree(d_ci); cudaFree(d_si); cudaFree(d_y); cudaFree(d_hist); cudaFree(d_scale);
        cudaFree(d_vtab); cudaFree(d_L); cudaFree(d_dotp);
        // the big buffers come from the stream-ordered pool: the next solve gets them back without mapping memory again
        for (auto p : chunks) cudaFreeAsync(p, ctx->stream);
        if (av) cudaFreeAsync(av, ctx->stream);
        if (w) cudaFreeAsync(w, ctx->stream);
        if (r) cudaFreeAsync(r, ctx->stream);
    };
#define GM_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e__)); cleanup(); return NGSB_ERR_CUDA; } \
    } while (0)
#define GM_TRY(call)                                      \
    do {                                                  \
        rc = (call);                                      \
        if (rc != NGSB_OK) { cleanup(); return rc; }      \
    } while (0)

    GM_CUDA(cudaMalloc(&d_st, sizeof(GmresState)));
    GM_CUDA(cudaMalloc(&d_h, sizeof(double2) * (size_t)(ms + 1) * ms));
    GM_CUDA(cudaMemsetAsync(d_h, 0, sizeof(double2) * (size_t)(ms + 1) * ms, ctx->stream));
    GM_CUDA(cudaMalloc(&d_gam, sizeof(double2) * (ms + 2)));
    GM_CUDA(cudaMalloc(&d_ci, sizeof(double2) * (ms + 2)));
    GM_CUDA(cudaMalloc(&d_si, sizeof(double2) * (ms + 2)));
    GM_CUDA(cudaMalloc(&d_y, sizeof(double2) * (ms + 2)));
    GM_CUDA(cudaMemsetAsync(d_gam, 0, sizeof(double2) * (ms + 2), ctx->stream));
    GM_CUDA(cudaMemsetAsync(d_ci, 0, sizeof(double2) * (ms + 2), ctx->stream));
    GM_CUDA(cudaMemsetAsync(d_si, 0, sizeof(double2) * (ms + 2), ctx->stream));
    GM_CUDA(cudaMemsetAsync(d_y, 0, sizeof(double2) * (ms + 2), ctx->stream));
    GM_CUDA(cudaMalloc(&d_hist, sizeof(double) * (hist_cap + 1)));
    GM_CUDA(cudaMalloc(&d_scale, sizeof(double) * 2));
    // gmres_orth = 1: V^T w in one batched reduction (see gmres_dots_kernel); distributed, the 2 (j+1) sums of a step travel
    // in one peer-memory exchange.  The NCCL data path keeps the serial form (its reductions are single (re,im) pairs), and
    // so do Krylov spaces beyond the finish kernel's block size.
    const bool batched = ctx->gmres_orth == 1 && !(dist && dist->allreduce) && ms + 1 <= 1024 && 4 * (ms + 1) <= NGSB_PEER_VEC_LEN + 4;
    const int tile = cplx ? 8 : 16;
    int dgrid = 1;
    if (batched) {
        int per_sm = 1;
        if (cplx) GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gmres_dots_kernel<true>, 256, 0));
        else GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gmres_dots_kernel<false>, 256, 0));
        dgrid = ctx->sm_count * std::max(1, std::min(per_sm, 4));
        const uint64_t need = (((uint64_t)(A->h * kind_scalars(A->kind)) / 2 + 256) / 256);
        if ((uint64_t)dgrid > need) dgrid = (int)std::max<uint64_t>(1, need);
        GM_CUDA(cudaMalloc(&d_vtab, sizeof(double *) * (size_t)(ms + 2)));
        GM_CUDA(cudaMalloc(&d_L, sizeof(double2) * (size_t)(ms + 1) * (ms + 1)));
        GM_CUDA(cudaMalloc(&d_dotp, sizeof(double) * 32 * (size_t)dgrid * ((ms + tile) / tile)));
    }
    const size_t vbytes = (nscal ? nscal : 2) * sizeof(double);
    {
        // keep up to a fifth of the device memory in the default pool between solves (cudaMalloc/cudaFree of the multi-GB
        // basis chunks cost tens of milliseconds each and drain the stream)
        cudaMemPool_t pool = nullptr;
        size_t free_b = 0, total_b = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            uint64_t thr = (uint64_t)total_b / 5;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaGetLastError();
    }
    GM_CUDA(cudaMallocAsync(&av, vbytes, ctx->stream));
    GM_CUDA(cudaMallocAsync(&w, vbytes, ctx->stream));
    GM_CUDA(cudaMallocAsync(&r, vbytes, ctx->stream));
    const size_t vstride = (vbytes + 255) & ~(size_t)255;
    const size_t per_chunk = std::max<size_t>(1, std::min<size_t>(8, ((size_t)4 << 30) / vstride));
    size_t chunk_used = per_chunk;
    auto new_basis_vector = [&](double **out) -> cudaError_t {
        if (chunk_used == per_chunk) {
            double *c = nullptr;
            const size_t want = std::min<size_t>(per_chunk, (size_t)ms + 1 > vi.size() ? (size_t)ms + 1 - vi.size() : 1);
            cudaError_t e = cudaMallocAsync(&c, vstride * std::max<size_t>(1, want), ctx->stream);
            if (e != cudaSuccess) return e;
            chunks.push_back(c);
            chunk_used = 0;
        }
        *out = (double *)((char *)chunks.back() + vstride * chunk_used++);
        return cudaSuccess;
    };

    memset(&hst, 0, sizeof(hst));
    hst.prec = prec;
    hst.maxsteps = maxsteps;
    hst.cplx = cplx ? 1 : 0;
    hst.hist_cap = hist_cap;
    GM_CUDA(cudaMemcpyAsync(d_st, &hst, sizeof(hst), cudaMemcpyHostToDevice, ctx->stream));
    GM_CUDA(cudaStreamSynchronize(ctx->stream));

    double *x = xvec->d;
    const double *f = fvec->d;
    double *d_tmp = (double *)((char *)d_st + offsetof(GmresState, tmp));
    double *d_tmp2 = (double *)((char *)d_st + offsetof(GmresState, tmp2));
    auto spmv = [&](const double *in, double *out) {
        SpmvArgs a;
        memset(&a, 0, sizeof(a));
        a.A = A; a.x = in; a.y = out; a.sr = 1.0; a.accumulate = false; a.epi = EPI_NONE;
        return spmv_launch(a);
    };
    const int rgrid = reduce_grid(ctx, N);
    auto mgs = [&](const double *vprev, const double2 *hprev, const double *vnext, int selfdot) {
        SpanGuard g(ctx, KC_VEC);
        if (cplx) gmres_mgs_kernel<true><<<rgrid, 256, 0, ctx->stream>>>(w, vprev, hprev, vnext, selfdot, N, d_st, ctx->d_partials, ctx->d_counter, master, mask_div);
        else gmres_mgs_kernel<false><<<rgrid, 256, 0, ctx->stream>>>(w, vprev, hprev, vnext, selfdot, N, d_st, ctx->d_partials, ctx->d_counter, master, mask_div);
        if (cudaGetLastError() != cudaSuccess) return NGSB_ERR_CUDA;
        return allred(d_tmp);
    };

    // r = f (or f - A x); r = C r
    if (initialize) {
        GM_CUDA(cudaMemsetAsync(x, 0, nscal * sizeof(double), ctx->stream));
        GM_CUDA(cudaMemcpyAsync(r, f, nscal * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        GM_TRY(spmv(x, av));
        GM_CUDA(cudaMemcpyAsync(r, f, nscal * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        GM_TRY(launch_axpby(ctx, r, av, N, -1.0, 0.0, cplx, true));
    }
    GM_TRY(cumulate(r));
    if (C) {
        GM_TRY(jacobi_apply(C, 1.0, 0.0, r, w, false));
        GM_CUDA(cudaMemcpyAsync(r, w, nscal * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    // norm = r.L2Norm(); v = 1/sqrt(<r,r>) r
    GM_TRY(launch_dot_masked(ctx, r, r, nscal, 3, d_tmp, master, cplx ? 2 : mask_div));
    GM_TRY(allred(d_tmp));
    GM_TRY(launch_dot_masked(ctx, r, r, N, cplx ? 1 : 0, d_tmp2, master, mask_div));
    GM_TRY(allred(d_tmp2));
    {
        SpanGuard g(ctx, KC_OTHER);
        gmres_init_kernel<<<1, 1, 0, ctx->stream>>>(d_st, d_gam, d_hist, R);
    }
    double *d_st_scale = (double *)((char *)d_st + offsetof(GmresState, scale));
    // v_0
    {
        double *v0 = nullptr;
        GM_CUDA(new_basis_vector(&v0));
        vi.push_back(v0);
        if (batched) gmres_set_ptr_kernel<<<1, 1, 0, ctx->stream>>>(d_vtab, 0, v0);
        GM_TRY(launch_axpby_dev(ctx, v0, r, N, d_st_scale, cplx, false, false));
    }

    int j = -1;
    bool done = false;
    GM_CUDA(cudaMemcpyAsync(&hst, d_st, sizeof(hst), cudaMemcpyDeviceToHost, ctx->stream));
    GM_CUDA(cudaStreamSynchronize(ctx->stream));
    done = hst.done != 0;
    // The host runs one step ahead of the device-side stopping rule: step j+1 is enqueued before the `done` flag of step j
    // has been read back (every kernel of the step that touches solver state returns at once when the flag is set), so the
    // stream never drains between steps.  The state is polled through two pinned slots.
    GmresState *slot = reinterpret_cast<GmresState *>(ctx->h_pinned);
    EventPair evp;             // destroyed on every exit path
    cudaEvent_t *ev = evp.ev;
    GM_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    GM_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int jh = hst.j, enq = 0;
    while (!done) {
        j = jh;      // current column (the device's st->j as long as the loop condition holds)
        double *v = vi[j];
        // w = C A v: the product lands in `av` and the preconditioner writes w (no preconditioner: straight into w)
        GM_TRY(spmv(v, C ? av : w));
        GM_TRY(cumulate(C ? av : w));
        if (C) GM_TRY(jacobi_apply(C, 1.0, 0.0, av, w, false));
        if (batched) {
            const int nv = j + 1;
            for (int k0 = 0, t = 0; k0 < nv; k0 += tile, t++) {
                SpanGuard g(ctx, KC_VEC);
                const int nk = std::min(tile, nv - k0);
                double *out = d_dotp + (size_t)t * dgrid * 32;
                if (cplx) gmres_dots_kernel<true><<<dgrid, 256, 0, ctx->stream>>>(d_vtab, k0, nk, j, w, nscal, d_st, out, master, mask_div);
                else gmres_dots_kernel<false><<<dgrid, 256, 0, ctx->stream>>>(d_vtab, k0, nk, j, w, nscal, d_st, out, master, mask_div);
            }
            {
                SpanGuard g(ctx, KC_OTHER);
                const int bs = std::max(64, (nv + 31) & ~31);
                gmres_orth_finish_kernel<<<1, bs, (size_t)nv * 48, ctx->stream>>>(d_st, d_h, d_L, d_dotp, dgrid, nv, ms, ms + 1, cplx ? 1 : 0, R);
            }
            SpanGuard g(ctx, KC_VEC);
            if (cplx) gmres_project_kernel<true><<<rgrid, 256, 0, ctx->stream>>>(w, d_vtab, d_h + j, ms, nv, nscal, d_st, ctx->d_partials, ctx->d_counter, master, mask_div);
            else gmres_project_kernel<false><<<rgrid, 256, 0, ctx->stream>>>(w, d_vtab, d_h + j, ms, nv, nscal, d_st, ctx->d_partials, ctx->d_counter, master, mask_div);
            GM_CUDA(cudaGetLastError());
        } else {
            // MGS: h(i,j) = <v_i, w>; w -= h(i,j) v_i  (projection i fused with inner product i+1)
            for (int i = 0; i <= j; i++) {
                GM_TRY(mgs(i > 0 ? vi[i - 1] : nullptr, i > 0 ? d_h + (size_t)(i - 1) * ms + j : nullptr, vi[i], 0));
                SpanGuard g(ctx, KC_OTHER);
                gmres_store_h_kernel<<<1, 1, 0, ctx->stream>>>(d_st, d_h, ms, i, R);
            }
            GM_TRY(mgs(vi[j], d_h + (size_t)j * ms + j, nullptr, 1));       // last projection + <w,w>
        }
        {
            SpanGuard g(ctx, KC_OTHER);
            gmres_givens_kernel<<<1, 1, 0, ctx->stream>>>(d_st, d_h, d_gam, d_ci, d_si, ms, d_hist, R);
        }
        // v_{j+1} = 1/h(j+1,j) * w  (always formed, like the reference)
        double *vn = nullptr;
        GM_CUDA(new_basis_vector(&vn));
        vi.push_back(vn);
        if (batched) gmres_set_ptr_kernel<<<1, 1, 0, ctx->stream>>>(d_vtab, (int)vi.size() - 1, vn);
        GM_TRY(launch_axpby_dev(ctx, vn, w, N, d_st_scale, cplx, false, false));
        GM_CUDA(cudaMemcpyAsync(&slot[enq & 1], d_st, sizeof(GmresState), cudaMemcpyDeviceToHost, ctx->stream));
        GM_CUDA(cudaEventRecord(ev[enq & 1], ctx->stream));
        if (enq > 0) {
            GM_CUDA(cudaEventSynchronize(ev[(enq - 1) & 1]));
            done = slot[(enq - 1) & 1].done != 0;
        }
        enq++;
        jh++;
        if (jh >= ms) {       // the last column the arrays can hold: no further speculation
            GM_CUDA(cudaStreamSynchronize(ctx->stream));
            break;
        }
    }
    GM_CUDA(cudaMemcpyAsync(&hst, d_st, sizeof(hst), cudaMemcpyDeviceToHost, ctx->stream));
    GM_CUDA(cudaStreamSynchronize(ctx->stream));
    // j-- ; back substitution ; x += y_i v_i
    {
        SpanGuard g(ctx, KC_OTHER);
        gmres_backsolve_kernel<<<1, 1, 0, ctx->stream>>>(d_st, d_h, d_gam, d_y, ms);
    }
    const int jfinal = hst.j - 1;
    for (int i = 0; i <= jfinal; i++) {
        SpanGuard g(ctx, KC_VEC);
        uint64_t blocks = (N + 1023) / 1024;
        uint64_t cap = (uint64_t)ctx->sm_count * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        if (cplx) gmres_update_kernel<true><<<(int)blocks, 256, 0, ctx->stream>>>(x, vi[i], d_y + i, N, d_st, i);
        else gmres_update_kernel<false><<<(int)blocks, 256, 0, ctx->stream>>>(x, vi[i], d_y + i, N, d_st, i);
    }
    GM_CUDA(cudaGetLastError());
    GM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (steps) *steps = jfinal;
    if (nhist) *nhist = hst.nhist;
    int ncopy = hst.nhist < hist_cap ? hst.nhist : hist_cap;
    if (history && ncopy > 0) {
        GM_CUDA(cudaMemcpyAsync(history, d_hist, ncopy * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        GM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    cleanup();
#undef GM_CUDA
#undef GM_TRY
    return NGSB_OK;
}
