// spmv.cu -- CSR sparse matrix-vector product for sm_100a.
//
// Replaces DevSparseMatrix (cusparseSpMV, ngscuda/cuda_linalg.cpp:187-316) behind
// SparseMatrix<TM>::MultAdd (linalg/sparsematrix_impl.hpp:264-279) for TM = double,
// Complex and Mat<3,3,double>.
//
// Two hand-written kernels:
//  * spmv_stream_kernel ("tma-stream", default): persistent CTAs; one producer lane
//    streams row blocks of (values | column indices | row offsets) into a 4-stage
//    shared-memory ring with 1-D bulk TMA copies (cp.async.bulk + mbarrier
//    complete_tx), so the 95 % of the traffic that is a pure stream never touches the
//    LSU/L1 path or registers.  Consumer warps work on sub-warp groups of W lanes per
//    row (W picked from the mean row length), gather x through the read-only path,
//    reduce with shuffles, and optionally fuse the <dotvec, A x> reduction (kss of CG)
//    with a deterministic last-block finish.
//  * spmv_subwarp_kernel ("subwarp"): classic vector-CSR with direct global loads; kept
//    as the simple cross-check and for A/B measurements (option spmv_algo = 1).
#include "spmv.cuh"

#include <algorithm>
#include <cuda.h>

namespace ngsb {

// ------------------------------------------------------------------------------------------
// configuration per entry kind
// ------------------------------------------------------------------------------------------
template <int KIND> struct KindCfg;
template <> struct KindCfg<NGSB_REAL>    { static constexpr int VB = 8;  static constexpr int XS = 1; };
template <> struct KindCfg<NGSB_COMPLEX> { static constexpr int VB = 16; static constexpr int XS = 2; };
template <> struct KindCfg<NGSB_BLOCK3>  { static constexpr int VB = 72; static constexpr int XS = 3; };

static constexpr int MAX_STAGES = 8;
static constexpr int RMAX = 504;           // max rows per block (row-offset slot holds RMAX+8 uint16)

// shared-memory layout of one stage, fixed per matrix at creation (tile = entries per stage)
struct StageLayout {
    int cols, roff, hdr, bytes;
    __host__ __device__ StageLayout(int tile, int vb)
    {
        cols = tile * vb;
        roff = cols + tile * 4;
        hdr = roff + (RMAX + 8) * 2;
        bytes = ((hdr + 16 + 127) / 128) * 128;
    }
};

static int default_tile(int kind)
{
    return kind == NGSB_REAL ? 2048 : (kind == NGSB_COMPLEX ? 1024 : 512);
}

static int kind_vb(int kind) { return kind == NGSB_REAL ? 8 : (kind == NGSB_COMPLEX ? 16 : 72); }

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk TMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // bounded spin: a protocol bug must trap, never hang the GPU
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int W> __device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------
struct SpmvParams {
    const SpmvBlock *blocks;
    const uint16_t *rowoff;
    const int32_t *col;
    const double *val;
    const uint64_t *rowptr;
    const double *x;
    double *y;
    double sr, si;
    int accumulate;
    int epi;
    const double *dotvec;
    int dot_conj;
    double *dot_out;
    CgState *state;
    double *partials;
    unsigned int *counter;
    uint32_t block_begin, block_end;
    uint64_t nrows;
    int tile, stages;
};

// store one finished row; returns this row's contribution to the fused dot in (dr, di)
template <int KIND>
__device__ __forceinline__ void finish_row(const SpmvParams &p, uint64_t row, double s0, double s1, double s2, double &dr,
                                           double &di)
{
    if (KIND == NGSB_REAL) {
        double r = p.sr * s0;
        if (p.accumulate) r += p.y[row];
        p.y[row] = r;
        if (p.epi) dr += p.dotvec[row] * r;
    } else if (KIND == NGSB_COMPLEX) {
        double2 *y2 = reinterpret_cast<double2 *>(p.y);
        double rr = p.sr * s0 - p.si * s1, ri = p.sr * s1 + p.si * s0;
        if (p.accumulate) { double2 o = y2[row]; rr += o.x; ri += o.y; }
        y2[row] = make_double2(rr, ri);
        if (p.epi) {
            double2 v = reinterpret_cast<const double2 *>(p.dotvec)[row];
            double ci = p.dot_conj ? -ri : ri;
            dr += v.x * rr - v.y * ci;
            di += v.x * ci + v.y * rr;
        }
    } else {
        double r0 = p.sr * s0, r1 = p.sr * s1, r2 = p.sr * s2;
        double *yy = p.y + 3 * row;
        if (p.accumulate) { r0 += yy[0]; r1 += yy[1]; r2 += yy[2]; }
        yy[0] = r0; yy[1] = r1; yy[2] = r2;
        if (p.epi) {
            const double *v = p.dotvec + 3 * row;
            dr += v[0] * r0 + v[1] * r1 + v[2] * r2;
        }
    }
}

// block-wide (sum of a, sum of b) -> thread 0
__device__ __forceinline__ double2 cta_sum2(double a, double b, double *red /* 2*32 doubles smem */)
{
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    if (lane == 0) { red[wid] = a; red[32 + wid] = b; }
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    if (wid == 0) {
        a = lane < nw ? red[lane] : 0.0;
        b = lane < nw ? red[32 + lane] : 0.0;
        a = warp_sum_d(a);
        b = warp_sum_d(b);
    }
    return make_double2(a, b);
}

// deterministic grid-wide finish of the fused dot + epilogue (scalar step of CG)
__device__ __forceinline__ void dot_epilogue(const SpmvParams &p, double dr, double di, double *red)
{
    double2 mine = cta_sum2(dr, di, red);
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        p.partials[2 * blockIdx.x] = mine.x;
        p.partials[2 * blockIdx.x + 1] = mine.y;
        __threadfence();
        unsigned int t = atomicAdd(p.counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += 32) {
        a += __ldcg(&p.partials[2 * k]);
        b += __ldcg(&p.partials[2 * k + 1]);
    }
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    if (threadIdx.x == 0) {
        *p.counter = 0;
        if (p.epi == EPI_DOT_OUT) { p.dot_out[0] = a; p.dot_out[1] = b; }
        else if (p.epi == EPI_CG_KSS) cg_finalize_kss(p.state, make_double2(a, b));
    }
}

// ------------------------------------------------------------------------------------------
// TMA-streamed kernel
// ------------------------------------------------------------------------------------------
template <int KIND, int W>
__global__ void __launch_bounds__(544) spmv_stream_kernel(const SpmvParams p)
{
    using C = KindCfg<KIND>;
    constexpr int G = 32 / W;   // rows per warp step
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double red[64];

    if (p.state != nullptr && p.state->done) return;   // uniform: CG already finished

    const StageLayout L(p.tile, C::VB);
    const int STAGES = p.stages;
    const int NCW = (int)(blockDim.x >> 5) - 1;        // consumer warps; the last warp is the producer
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + MAX_STAGES;
    unsigned char *stages = smem + 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t nb = p.block_end - p.block_begin;
    const uint32_t b0 = p.block_begin + (uint32_t)((uint64_t)blockIdx.x * nb / gridDim.x);
    const uint32_t b1 = p.block_begin + (uint32_t)((uint64_t)(blockIdx.x + 1) * nb / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        mbar_fence_init();
    }
    __syncthreads();

    double dr = 0.0, di = 0.0;

    if (warp == NCW) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int s = 0;
            uint32_t k = 0;
            for (uint32_t b = b0; b < b1; ++b) {
                if (k > 0) mbar_wait(&empty[s], (k & 1) ^ 1);
                const SpmvBlock d = p.blocks[b];
                unsigned char *st = stages + (size_t)s * L.bytes;
                uint32_t *hdr = reinterpret_cast<uint32_t *>(st + L.hdr);
                hdr[0] = d.row_begin;
                hdr[1] = d.nrows;
                hdr[2] = d.flags;
                if (d.flags & 1u) {
                    mbar_arrive(&full[s]);
                } else {
                    const uint32_t nwin = (uint32_t)d.nwin4 * 4u;
                    const uint32_t roff_bytes = (((uint32_t)d.nrows + 1u + 7u) & ~7u) * 2u;
                    const uint32_t bytes = nwin * (uint32_t)C::VB + nwin * 4u + roff_bytes;
                    mbar_arrive_expect_tx(&full[s], bytes);
                    if (nwin) {   // a block of empty rows has no entries to stream
                        tma_bulk_g2s(st, reinterpret_cast<const unsigned char *>(p.val) + d.nnz_base * (uint64_t)C::VB,
                                     nwin * (uint32_t)C::VB, &full[s]);
                        tma_bulk_g2s(st + L.cols, p.col + d.nnz_base, nwin * 4u, &full[s]);
                    }
                    tma_bulk_g2s(st + L.roff, p.rowoff + d.roff_base, roff_bytes, &full[s]);
                }
                if (++s == STAGES) { s = 0; ++k; }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int grp = lane / W;      // row slot inside the warp step
        const int gl = lane % W;       // lane inside the group
        uint32_t rr = 0;               // round-robin offset: groups handed out so far (mod NCW)
        int s = 0;
        uint32_t k = 0;
        for (uint32_t b = b0; b < b1; ++b) {
            mbar_wait(&full[s], k & 1);
            const unsigned char *st = stages + (size_t)s * L.bytes;
            const uint32_t *hdr = reinterpret_cast<const uint32_t *>(st + L.hdr);
            const uint32_t row_begin = hdr[0], nrows = hdr[1], flags = hdr[2];
            const uint32_t ngroups = (nrows + G - 1) / G;
            if (flags & 1u) {
                // long row: y[row] already final (long-row kernel ran first); only the dot part
                if (p.epi && (rr == (uint32_t)warp) && lane == 0) {
                    const uint64_t row = row_begin;
                    if (KIND == NGSB_REAL) dr += p.dotvec[row] * p.y[row];
                    else if (KIND == NGSB_COMPLEX) {
                        double2 v = reinterpret_cast<const double2 *>(p.dotvec)[row];
                        double2 r = reinterpret_cast<const double2 *>(p.y)[row];
                        double ci = p.dot_conj ? -r.y : r.y;
                        dr += v.x * r.x - v.y * ci;
                        di += v.x * ci + v.y * r.x;
                    } else {
                        const double *v = p.dotvec + 3 * row, *r = p.y + 3 * row;
                        dr += v[0] * r[0] + v[1] * r[1] + v[2] * r[2];
                    }
                }
            } else {
                const uint16_t *roff = reinterpret_cast<const uint16_t *>(st + L.roff);
                const int32_t *cs = reinterpret_cast<const int32_t *>(st + L.cols);
                uint32_t q = (uint32_t)warp >= rr ? (uint32_t)warp - rr : (uint32_t)warp + NCW - rr;
                for (; q < ngroups; q += NCW) {
                    const uint32_t rl = q * G + grp;
                    const bool valid = rl < nrows;
                    int o0 = 0, o1 = 0;
                    if (valid) { o0 = roff[rl]; o1 = roff[rl + 1]; }
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                    if (KIND == NGSB_REAL) {
                        const double *vs = reinterpret_cast<const double *>(st);
                        int j = o0 + gl;
                        for (; j + 3 * W < o1; j += 4 * W) {
                            int c0 = cs[j], c1 = cs[j + W], c2 = cs[j + 2 * W], c3 = cs[j + 3 * W];
                            double x0 = __ldg(p.x + c0), x1 = __ldg(p.x + c1), x2 = __ldg(p.x + c2), x3 = __ldg(p.x + c3);
                            s0 = fma(vs[j], x0, s0);
                            s1 = fma(vs[j + W], x1, s1);
                            s0 = fma(vs[j + 2 * W], x2, s0);
                            s1 = fma(vs[j + 3 * W], x3, s1);
                        }
                        for (; j < o1; j += W) s0 = fma(vs[j], __ldg(p.x + cs[j]), s0);
                        s0 = group_sum<W>(s0 + s1);
                        s1 = 0.0;
                    } else if (KIND == NGSB_COMPLEX) {
                        const double2 *vs = reinterpret_cast<const double2 *>(st);
                        const double2 *x2 = reinterpret_cast<const double2 *>(p.x);
                        int j = o0 + gl;
                        for (; j + W < o1; j += 2 * W) {
                            int c0 = cs[j], c1 = cs[j + W];
                            double2 xa = __ldg(x2 + c0), xb = __ldg(x2 + c1);
                            double2 va = vs[j], vb = vs[j + W];
                            s0 += va.x * xa.x - va.y * xa.y;
                            s1 += va.x * xa.y + va.y * xa.x;
                            s0 += vb.x * xb.x - vb.y * xb.y;
                            s1 += vb.x * xb.y + vb.y * xb.x;
                        }
                        for (; j < o1; j += W) {
                            double2 xa = __ldg(x2 + cs[j]);
                            double2 va = vs[j];
                            s0 += va.x * xa.x - va.y * xa.y;
                            s1 += va.x * xa.y + va.y * xa.x;
                        }
                        s0 = group_sum<W>(s0);
                        s1 = group_sum<W>(s1);
                    } else {
                        const double *vs = reinterpret_cast<const double *>(st);
                        for (int j = o0 + gl; j < o1; j += W) {
                            const double *m = vs + 9 * j;
                            const double *xv = p.x + 3 * (size_t)cs[j];
                            double x0 = __ldg(xv), x1 = __ldg(xv + 1), x2 = __ldg(xv + 2);
                            s0 += m[0] * x0 + m[1] * x1 + m[2] * x2;
                            s1 += m[3] * x0 + m[4] * x1 + m[5] * x2;
                            s2 += m[6] * x0 + m[7] * x1 + m[8] * x2;
                        }
                        s0 = group_sum<W>(s0);
                        s1 = group_sum<W>(s1);
                        s2 = group_sum<W>(s2);
                    }
                    if (valid && gl == 0) finish_row<KIND>(p, (uint64_t)row_begin + rl, s0, s1, s2, dr, di);
                }
            }
            rr = (rr + ((flags & 1u) ? 1u : ngroups)) % (uint32_t)NCW;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == STAGES) { s = 0; ++k; }
        }
    }

    if (p.epi) dot_epilogue(p, dr, di, red);
}

// ------------------------------------------------------------------------------------------
// long rows (more entries than one tile): one CTA per row
// ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) spmv_longrow_kernel(const SpmvParams p, const uint32_t *longrows)
{
    __shared__ double red[64];
    if (p.state != nullptr && p.state->done) return;
    const uint64_t row = longrows[blockIdx.x];
    const uint64_t a = p.rowptr[row], b = p.rowptr[row + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (uint64_t j = a + threadIdx.x; j < b; j += blockDim.x) {
        const int c = p.col[j];
        if (KIND == NGSB_REAL) s0 = fma(p.val[j], __ldg(p.x + c), s0);
        else if (KIND == NGSB_COMPLEX) {
            double2 v = reinterpret_cast<const double2 *>(p.val)[j];
            double2 xv = __ldg(reinterpret_cast<const double2 *>(p.x) + c);
            s0 += v.x * xv.x - v.y * xv.y;
            s1 += v.x * xv.y + v.y * xv.x;
        } else {
            const double *m = p.val + 9 * j;
            const double *xv = p.x + 3 * (size_t)c;
            double x0 = __ldg(xv), x1 = __ldg(xv + 1), x2 = __ldg(xv + 2);
            s0 += m[0] * x0 + m[1] * x1 + m[2] * x2;
            s1 += m[3] * x0 + m[4] * x1 + m[5] * x2;
            s2 += m[6] * x0 + m[7] * x1 + m[8] * x2;
        }
    }
    double2 t01 = cta_sum2(s0, s1, red);
    __syncthreads();
    double2 t2 = cta_sum2(s2, 0.0, red);
    if (threadIdx.x == 0) {
        double dr = 0.0, di = 0.0;
        SpmvParams q = p;
        q.epi = 0;   // the dot part of long rows is taken by the stream kernel
        finish_row<KIND>(q, row, t01.x, t01.y, t2.x, dr, di);
    }
}

// ------------------------------------------------------------------------------------------
// plain vector-CSR kernel (direct global loads), W lanes per row
// ------------------------------------------------------------------------------------------
template <int KIND, int W>
__global__ void __launch_bounds__(256) spmv_subwarp_kernel(const SpmvParams p)
{
    if (p.state != nullptr && p.state->done) return;
    const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / W;
    const int gl = threadIdx.x % W;
    const bool valid = gid < p.nrows;
    uint64_t a = 0, b = 0;
    if (valid) { a = p.rowptr[gid]; b = p.rowptr[gid + 1]; }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (KIND == NGSB_REAL) {
        for (uint64_t j = a + gl; j < b; j += W) s0 = fma(p.val[j], __ldg(p.x + p.col[j]), s0);
        s0 = group_sum<W>(s0);
    } else if (KIND == NGSB_COMPLEX) {
        const double2 *v2 = reinterpret_cast<const double2 *>(p.val);
        const double2 *x2 = reinterpret_cast<const double2 *>(p.x);
        for (uint64_t j = a + gl; j < b; j += W) {
            double2 v = v2[j];
            double2 xv = __ldg(x2 + p.col[j]);
            s0 += v.x * xv.x - v.y * xv.y;
            s1 += v.x * xv.y + v.y * xv.x;
        }
        s0 = group_sum<W>(s0);
        s1 = group_sum<W>(s1);
    } else {
        for (uint64_t j = a + gl; j < b; j += W) {
            const double *m = p.val + 9 * j;
            const double *xv = p.x + 3 * (size_t)p.col[j];
            double x0 = __ldg(xv), x1 = __ldg(xv + 1), x2 = __ldg(xv + 2);
            s0 += m[0] * x0 + m[1] * x1 + m[2] * x2;
            s1 += m[3] * x0 + m[4] * x1 + m[5] * x2;
            s2 += m[6] * x0 + m[7] * x1 + m[8] * x2;
        }
        s0 = group_sum<W>(s0);
        s1 = group_sum<W>(s1);
        s2 = group_sum<W>(s2);
    }
    if (valid && gl == 0) {
        double dr = 0.0, di = 0.0;
        SpmvParams q = p;
        q.epi = 0;
        finish_row<KIND>(q, gid, s0, s1, s2, dr, di);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int KIND, int W>
static int launch_stream(ngsb_ctx *ctx, const SpmvParams &p, int grid, int ncw)
{
    const StageLayout L(p.tile, KindCfg<KIND>::VB);
    const int smem = 128 + p.stages * L.bytes;
    static int configured[64] = {0};
    NGSB_REQUIRE(ctx->device < 64, "device index too large");
    if (configured[ctx->device] < smem) {
        NGSB_CUDA(cudaFuncSetAttribute(spmv_stream_kernel<KIND, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[ctx->device] = smem;
    }
    spmv_stream_kernel<KIND, W><<<grid, (ncw + 1) * 32, smem, ctx->stream>>>(p);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

template <int KIND>
static int dispatch_stream(ngsb_ctx *ctx, const SpmvParams &p, int grid, int W, int ncw)
{
    switch (W) {
    case 4: return launch_stream<KIND, 4>(ctx, p, grid, ncw);
    case 8: return launch_stream<KIND, 8>(ctx, p, grid, ncw);
    case 16: return launch_stream<KIND, 16>(ctx, p, grid, ncw);
    default: return launch_stream<KIND, 32>(ctx, p, grid, ncw);
    }
}

template <int KIND>
static int dispatch_subwarp(ngsb_ctx *ctx, const SpmvParams &p, int W)
{
    uint64_t threads = p.nrows * (uint64_t)W;
    uint64_t grid = (threads + 255) / 256;
    if (grid == 0) return NGSB_OK;
    NGSB_REQUIRE(grid < (1ull << 31), "matrix too large for the subwarp kernel grid");
    switch (W) {
    case 4: spmv_subwarp_kernel<KIND, 4><<<(unsigned)grid, 256, 0, ctx->stream>>>(p); break;
    case 8: spmv_subwarp_kernel<KIND, 8><<<(unsigned)grid, 256, 0, ctx->stream>>>(p); break;
    case 16: spmv_subwarp_kernel<KIND, 16><<<(unsigned)grid, 256, 0, ctx->stream>>>(p); break;
    default: spmv_subwarp_kernel<KIND, 32><<<(unsigned)grid, 256, 0, ctx->stream>>>(p); break;
    }
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

int launch_dot(ngsb_ctx *ctx, const double *x, const double *y, size_t N, int mode, double *d_out);

__global__ void cg_kss_from_dot_kernel(CgState *st, const double *dot)
{
    if (st->done) return;
    cg_finalize_kss(st, make_double2(dot[0], dot[1]));
}

int spmv_launch(const SpmvArgs &a)
{
    const ngsb_csr *A = a.A;
    ngsb_ctx *ctx = A->ctx;
    SpmvParams p;
    memset(&p, 0, sizeof(p));
    p.blocks = A->d_blocks;
    p.rowoff = A->d_rowoff;
    p.col = A->d_col;
    p.val = A->d_val;
    p.rowptr = A->d_rowptr;
    p.x = a.x;
    p.y = a.y;
    p.sr = a.sr;
    p.si = A->kind == NGSB_COMPLEX ? a.si : 0.0;
    p.accumulate = a.accumulate ? 1 : 0;
    p.epi = a.epi;
    p.dotvec = a.dotvec;
    p.dot_conj = a.dot_conj;
    p.dot_out = a.dot_out;
    p.state = a.state;
    p.partials = ctx->d_partials;
    p.counter = ctx->d_counter;
    p.nrows = A->h;
    p.block_begin = a.use_range ? a.block_begin : 0;
    p.block_end = a.use_range ? a.block_end : A->nblocks;
    p.tile = A->tile;
    p.stages = A->stages;

    if ((ctx->spmv_algo == 0 || ctx->spmv_algo == 3) && !a.use_range) {
        if (A->inner != nullptr) {
            // internally reordered matrix: x into the permuted numbering (one gather pass), the product on P A P^T, rows written
            // back through the slot -> caller-row table
            NGSB_REQUIRE(a.slice_list == nullptr, "SpMV: slice lists are not available on a reordered matrix");
            ngsb_csr *in = A->inner;
            NGSB_TRY(launch_perm_gather(ctx, a.x, A->d_perm, A->w, (int)kind_scalars(A->kind), in->d_xperm));
            SpmvArgs b = a;
            b.A = in; b.x = in->d_xperm; b.user_rows = true;
            return sell_launch(b);
        }
        return sell_launch(a);
    }
    NGSB_REQUIRE(!A->csr_released, "SpMV: the CSR kernels (spmv_algo 1/2) need the CSR arrays, which this matrix released (create it with option csr_keep = 1)");
    NGSB_REQUIRE(a.slice_list == nullptr, "SpMV: a slice list needs the SELL kernel (spmv_algo 0 or 3)");
    const bool stream = ctx->spmv_algo != 1;
    if (stream) {
        if (A->nlong > 0 && !a.use_range) {
            SpanGuard g(ctx, KC_SPMV);
            if (A->kind == NGSB_REAL) spmv_longrow_kernel<NGSB_REAL><<<A->nlong, 256, 0, ctx->stream>>>(p, A->d_longrows);
            else if (A->kind == NGSB_COMPLEX) spmv_longrow_kernel<NGSB_COMPLEX><<<A->nlong, 256, 0, ctx->stream>>>(p, A->d_longrows);
            else spmv_longrow_kernel<NGSB_BLOCK3><<<A->nlong, 256, 0, ctx->stream>>>(p, A->d_longrows);
            NGSB_CUDA(cudaGetLastError());
        }
        long cps = ctx->spmv_ctas_per_sm > 0 ? ctx->spmv_ctas_per_sm : A->ctas_per_sm;
        uint64_t nb = p.block_end - p.block_begin;
        uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)cps;
        if (grid > nb) grid = nb;
        if (grid < 1) grid = 1;
        SpanGuard g(ctx, KC_SPMV);
        if (A->kind == NGSB_REAL) return dispatch_stream<NGSB_REAL>(ctx, p, (int)grid, A->subwarp, A->ncw);
        if (A->kind == NGSB_COMPLEX) return dispatch_stream<NGSB_COMPLEX>(ctx, p, (int)grid, A->subwarp, A->ncw);
        return dispatch_stream<NGSB_BLOCK3>(ctx, p, (int)grid, A->subwarp, A->ncw);
    }
    // subwarp kernel: dot not fused
    {
        SpanGuard g(ctx, KC_SPMV);
        if (A->kind == NGSB_REAL) NGSB_TRY(dispatch_subwarp<NGSB_REAL>(ctx, p, A->subwarp));
        else if (A->kind == NGSB_COMPLEX) NGSB_TRY(dispatch_subwarp<NGSB_COMPLEX>(ctx, p, A->subwarp));
        else NGSB_TRY(dispatch_subwarp<NGSB_BLOCK3>(ctx, p, A->subwarp));
    }
    if (a.epi) {
        // <dotvec, y> as a separate deterministic reduction
        const bool cplx = A->kind == NGSB_COMPLEX;
        size_t N = cplx ? A->h : A->h * kind_scalars(A->kind);
        double *tmp = a.epi == EPI_DOT_OUT ? a.dot_out : ctx->d_partials + 2 * (MAX_PARTIALS - 2);
        NGSB_TRY(launch_dot(ctx, a.dotvec, a.y, N, cplx ? (a.dot_conj ? 2 : 1) : 0, tmp));
        if (a.epi == EPI_CG_KSS) {
            SpanGuard g(ctx, KC_OTHER);
            cg_kss_from_dot_kernel<<<1, 1, 0, ctx->stream>>>(a.state, tmp);
            NGSB_CUDA(cudaGetLastError());
        }
    }
    return NGSB_OK;
}

// ---- block builder --------------------------------------------------------------------------
static void build_blocks(const uint64_t *rowptr, size_t h, int tile, int NCW, int W, std::vector<SpmvBlock> &blocks,
                         std::vector<uint16_t> &rowoff, std::vector<uint32_t> &longrows)
{
    const uint64_t TILE = (uint64_t)tile;
    const int G = 32 / W;
    // rows per block: a multiple of the rows one CTA step covers, bounded by RMAX
    const size_t target_rows = std::min<size_t>(RMAX / (NCW * G) * (NCW * G), (size_t)NCW * G * 2);
    size_t r = 0;
    while (r < h) {
        const uint64_t base = rowptr[r] & ~(uint64_t)3;
        if (rowptr[r + 1] - base > TILE) {
            // a single row does not fit one tile
            SpmvBlock b;
            memset(&b, 0, sizeof(b));
            b.nnz_base = base;
            b.row_begin = (uint32_t)r;
            b.nrows = 1;
            b.flags = 1;
            b.roff_base = (uint32_t)rowoff.size();
            blocks.push_back(b);
            longrows.push_back((uint32_t)r);
            r++;
            continue;
        }
        size_t e = r + 1;
        while (e < h && e - r < target_rows && rowptr[e + 1] - base <= TILE) e++;
        SpmvBlock b;
        memset(&b, 0, sizeof(b));
        b.nnz_base = base;
        b.row_begin = (uint32_t)r;
        b.nrows = (uint16_t)(e - r);
        b.nwin4 = (uint16_t)((rowptr[e] - base + 3) / 4);
        b.flags = 0;
        b.roff_base = (uint32_t)rowoff.size();
        for (size_t i = r; i <= e; i++) rowoff.push_back((uint16_t)(rowptr[i] - base));
        while (rowoff.size() % 8) rowoff.push_back(0);
        blocks.push_back(b);
        r = e;
    }
}

static int pick_subwarp(int kind, double mean_row)
{
    if (kind == NGSB_BLOCK3) return mean_row <= 24 ? 8 : (mean_row <= 96 ? 16 : 32);
    if (mean_row <= 12) return 4;
    if (mean_row <= 64) return 8;
    if (mean_row <= 160) return 16;
    return 32;
}

static uint64_t g_uid = 0;
uint64_t next_uid() { return __atomic_add_fetch(&g_uid, 1, __ATOMIC_RELAXED); }

static int finish_create(ngsb_csr *A, const uint64_t *h_rowptr, bool allow_reorder = true)
{
    A->uid = next_uid();
    ngsb_ctx *ctx = A->ctx;
    A->mean_row = A->h ? (double)A->nnz / (double)A->h : 0.0;
    size_t mx = 0;
    for (size_t i = 0; i < A->h; i++) mx = std::max<size_t>(mx, h_rowptr[i + 1] - h_rowptr[i]);
    A->max_row = mx;
    A->subwarp = ctx->spmv_subwarp > 0 ? (int)ctx->spmv_subwarp : pick_subwarp(A->kind, A->mean_row);
    A->tile = ctx->spmv_tile > 0 ? (int)ctx->spmv_tile : default_tile(A->kind);
    A->ncw = ctx->spmv_ncw > 0 ? (int)ctx->spmv_ncw : 8;
    A->stages = ctx->spmv_stages > 0 ? (int)ctx->spmv_stages : 4;
    {
        // resident CTAs per SM the shared-memory footprint allows (227 KB usable)
        const StageLayout L(A->tile, kind_vb(A->kind));
        const int smem = 128 + A->stages * L.bytes + 1024;
        NGSB_REQUIRE(smem <= 227 * 1024, "spmv tile/stages do not fit shared memory");
        int fit = (227 * 1024) / smem;
        int by_threads = 2048 / ((A->ncw + 1) * 32);
        A->ctas_per_sm = std::max(1, std::min(fit, by_threads));
    }
    std::vector<SpmvBlock> blocks;
    std::vector<uint16_t> rowoff;
    std::vector<uint32_t> longrows;
    build_blocks(h_rowptr, A->h, A->tile, A->ncw, A->subwarp, blocks, rowoff, longrows);
    NGSB_REQUIRE(blocks.size() < (1ull << 32), "too many row blocks");
    A->nblocks = (uint32_t)blocks.size();
    A->nlong = (uint32_t)longrows.size();
    NGSB_CUDA(cudaMalloc(&A->d_blocks, std::max<size_t>(1, blocks.size()) * sizeof(SpmvBlock)));
    NGSB_CUDA(cudaMalloc(&A->d_rowoff, std::max<size_t>(8, rowoff.size()) * sizeof(uint16_t)));
    NGSB_CUDA(cudaMalloc(&A->d_longrows, std::max<size_t>(1, longrows.size()) * sizeof(uint32_t)));
    if (!blocks.empty()) NGSB_CUDA(cudaMemcpyAsync(A->d_blocks, blocks.data(), blocks.size() * sizeof(SpmvBlock), cudaMemcpyHostToDevice, ctx->stream));
    if (!rowoff.empty()) NGSB_CUDA(cudaMemcpyAsync(A->d_rowoff, rowoff.data(), rowoff.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    if (!longrows.empty()) NGSB_CUDA(cudaMemcpyAsync(A->d_longrows, longrows.data(), longrows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (allow_reorder) {
        // netgen-style numberings: multiply P A P^T instead (reorder.cu); the SELL copy then belongs to the inner matrix
        bool made = false;
        NGSB_TRY(csr_maybe_reorder(A, &made));
        if (made) {
            if (csr_release_wanted(A)) csr_release(A);
            return NGSB_OK;
        }
    }
    NGSB_TRY(sell_build(A, h_rowptr));
    // the products only stream the SELL copy: large matrices give the uploaded column / value arrays back (csrview.cu)
    if (allow_reorder && (ctx->spmv_algo == 0 || ctx->spmv_algo == 3) && csr_release_wanted(A)) csr_release(A);
    return NGSB_OK;
}

static int alloc_csr(ngsb_csr *A)
{
    const size_t ms = kind_matscalars(A->kind);
    const size_t slack = 16;   // bulk copies read whole 4-entry groups past the last entry
    NGSB_CUDA(cudaMalloc(&A->d_rowptr, (A->h + 1) * sizeof(uint64_t)));
    NGSB_CUDA(cudaMalloc(&A->d_col, (A->nnz + slack) * sizeof(int32_t)));
    NGSB_CUDA(cudaMalloc(&A->d_val, (A->nnz + slack) * ms * sizeof(double)));
    NGSB_CUDA(cudaMemsetAsync(A->d_col + A->nnz, 0, slack * sizeof(int32_t), A->ctx->stream));
    NGSB_CUDA(cudaMemsetAsync(A->d_val + A->nnz * ms, 0, slack * ms * sizeof(double), A->ctx->stream));
    return NGSB_OK;
}

static int validate_host_csr(size_t h, size_t w, size_t nnz, const uint64_t *rowptr, const int32_t *col)
{
    NGSB_REQUIRE(rowptr[0] == 0, "ngsb_csr_create: rowptr[0] must be 0");
    NGSB_REQUIRE(rowptr[h] == nnz, "ngsb_csr_create: rowptr[h]=%llu != nnz=%zu", (unsigned long long)rowptr[h], nnz);
    NGSB_REQUIRE(w < (1ull << 31) && h < (1ull << 32) - 1024, "ngsb_csr_create: dimensions exceed 32-bit column indices");
    for (size_t i = 0; i < h; i++) {
        NGSB_REQUIRE(rowptr[i] <= rowptr[i + 1], "ngsb_csr_create: rowptr not monotone at row %zu", i);
    }
    for (size_t j = 0; j < nnz; j++)
        NGSB_REQUIRE(col[j] >= 0 && (size_t)col[j] < w, "ngsb_csr_create: column index %d out of range at position %zu", col[j], j);
    return NGSB_OK;
}

// take ownership of device CSR arrays (allocated with >= 16 entries of zeroed slack behind nnz)
int csr_adopt_device(ngsb_ctx *ctx, size_t h, size_t w, size_t nnz, uint64_t *d_rowptr, int32_t *d_col, double *d_val, int kind,
                     ngsb_csr **out, bool allow_reorder)
{
    NGSB_REQUIRE(ctx && d_rowptr && d_col && d_val && out, "csr_adopt_device: NULL argument");
    NGSB_REQUIRE(w < (1ull << 31) && h < (1ull << 32) - 1024, "csr_adopt_device: dimensions exceed 32-bit indices");
    std::vector<uint64_t> h_rowptr(h + 1);
    NGSB_CUDA(cudaMemcpyAsync(h_rowptr.data(), d_rowptr, (h + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    NGSB_REQUIRE(h_rowptr[0] == 0 && h_rowptr[h] == nnz, "csr_adopt_device: rowptr inconsistent with nnz");
    ngsb_csr *A = new ngsb_csr();
    A->ctx = ctx; A->h = h; A->w = w; A->nnz = nnz; A->kind = kind;
    A->d_rowptr = d_rowptr; A->d_col = d_col; A->d_val = d_val;
    int rc = finish_create(A, h_rowptr.data(), allow_reorder);
    if (rc != NGSB_OK) { A->d_rowptr = nullptr; A->d_col = nullptr; A->d_val = nullptr; ngsb_csr_destroy(A); return rc; }
    *out = A;
    return NGSB_OK;
}

} // namespace ngsb

using namespace ngsb;

extern "C" int ngsb_csr_create(ngsb_ctx *ctx, size_t height, size_t width, size_t nnz, const uint64_t *rowptr,
                               const int32_t *col, const void *val, int kind, ngsb_csr **out)
{
    NGSB_REQUIRE(ctx && rowptr && out && (nnz == 0 || (col && val)), "ngsb_csr_create: NULL argument");
    NGSB_REQUIRE(kind_valid(kind), "ngsb_csr_create: bad kind %d", kind);
    NvtxRange nv("CreateDeviceMatrix (upload, reorder, SELL build)");
    NGSB_TRY(validate_host_csr(height, width, nnz, rowptr, col));
    NGSB_CUDA(cudaSetDevice(ctx->device));
    ngsb_csr *A = new ngsb_csr();
    A->ctx = ctx;
    A->h = height;
    A->w = width;
    A->nnz = nnz;
    A->kind = kind;
    int rc = alloc_csr(A);
    if (rc != NGSB_OK) { ngsb_csr_destroy(A); return rc; }
    const size_t ms = kind_matscalars(kind);
    NGSB_CUDA(cudaMemcpyAsync(A->d_rowptr, rowptr, (height + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        NGSB_CUDA(cudaMemcpyAsync(A->d_col, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        NGSB_CUDA(cudaMemcpyAsync(A->d_val, val, nnz * ms * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    rc = finish_create(A, rowptr);
    if (rc != NGSB_OK) { ngsb_csr_destroy(A); return rc; }
    *out = A;
    return NGSB_OK;
}

extern "C" int ngsb_csr_create_from_device(ngsb_ctx *ctx, size_t height, size_t width, size_t nnz, const uint64_t *d_rowptr,
                                           const int32_t *d_col, const void *d_val, int kind, ngsb_csr **out)
{
    NGSB_REQUIRE(ctx && d_rowptr && out && (nnz == 0 || (d_col && d_val)), "ngsb_csr_create_from_device: NULL argument");
    NGSB_REQUIRE(kind_valid(kind), "ngsb_csr_create_from_device: bad kind %d", kind);
    NGSB_REQUIRE(width < (1ull << 31) && height < (1ull << 32) - 1024, "ngsb_csr_create_from_device: dimensions exceed 32-bit indices");
    NGSB_CUDA(cudaSetDevice(ctx->device));
    ngsb_csr *A = new ngsb_csr();
    A->ctx = ctx;
    A->h = height;
    A->w = width;
    A->nnz = nnz;
    A->kind = kind;
    int rc = alloc_csr(A);
    if (rc != NGSB_OK) { ngsb_csr_destroy(A); return rc; }
    const size_t ms = kind_matscalars(kind);
    NGSB_CUDA(cudaMemcpyAsync(A->d_rowptr, d_rowptr, (height + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    if (nnz) {
        NGSB_CUDA(cudaMemcpyAsync(A->d_col, d_col, nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        NGSB_CUDA(cudaMemcpyAsync(A->d_val, d_val, nnz * ms * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    std::vector<uint64_t> h_rowptr(height + 1);
    NGSB_CUDA(cudaMemcpyAsync(h_rowptr.data(), A->d_rowptr, (height + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_rowptr[0] != 0 || h_rowptr[height] != nnz) {
        ngsb_csr_destroy(A);
        set_error("ngsb_csr_create_from_device: rowptr[0]/rowptr[h] inconsistent with nnz");
        return NGSB_ERR_INVALID;
    }
    rc = finish_create(A, h_rowptr.data());
    if (rc != NGSB_OK) { ngsb_csr_destroy(A); return rc; }
    *out = A;
    return NGSB_OK;
}

extern "C" int ngsb_csr_destroy(ngsb_csr *A)
{
    if (!A) return NGSB_OK;
    cudaSetDevice(A->ctx->device);
    cudaStreamSynchronize(A->ctx->stream);
    cudaFree(A->d_rowptr);
    cudaFree(A->d_col);
    cudaFree(A->d_val);
    cudaFree(A->d_blocks);
    cudaFree(A->d_rowoff);
    cudaFree(A->d_longrows);
    sell_free(A);
    if (A->transposed) ngsb_csr_destroy(A->transposed);
    if (A->inner) ngsb_csr_destroy(A->inner);
    cudaFree(A->d_perm); cudaFree(A->d_iperm); cudaFree(A->d_row_user); cudaFree(A->d_xperm);
    delete A;
    return NGSB_OK;
}

extern "C" int ngsb_csr_info(const ngsb_csr *A, size_t *height, size_t *width, size_t *nnz, int *kind)
{
    NGSB_REQUIRE(A, "ngsb_csr_info: A is NULL");
    if (height) *height = A->h;
    if (width) *width = A->w;
    if (nnz) *nnz = A->nnz;
    if (kind) *kind = A->kind;
    return NGSB_OK;
}

static int check_mult_args(const ngsb_csr *A, const ngsb_vec *x, const ngsb_vec *y, const char *who)
{
    NGSB_REQUIRE(A && x && y, "%s: NULL argument", who);
    NGSB_REQUIRE(x->ctx == A->ctx && y->ctx == A->ctx, "%s: objects belong to different contexts", who);
    NGSB_REQUIRE(x->kind == A->kind && y->kind == A->kind, "%s: vector kind does not match matrix kind %d", who, A->kind);
    // reference: BaseMatrix::Mult size checks (linalg/basematrix.cpp:373-382)
    NGSB_REQUIRE(x->n == A->w, "%s: width of matrix = %zu != size of x = %zu", who, A->w, x->n);
    NGSB_REQUIRE(y->n == A->h, "%s: height of matrix = %zu != size of y = %zu", who, A->h, y->n);
    const double *xb = x->d, *xe = x->d + x->nscal, *yb = y->d, *ye = y->d + y->nscal;
    NGSB_REQUIRE(xe <= yb || ye <= xb || x->nscal == 0 || y->nscal == 0, "%s: x and y must not overlap", who);
    return NGSB_OK;
}

extern "C" int ngsb_csr_multadd(const ngsb_csr *A, const double s[2], const ngsb_vec *x, ngsb_vec *y)
{
    NvtxRange nv("SparseMatrix::MultAdd");
    NGSB_TRY(check_mult_args(A, x, y, "SparseMatrix::MultAdd"));
    NGSB_REQUIRE(s, "ngsb_csr_multadd: s is NULL");
    NGSB_REQUIRE(A->kind == NGSB_COMPLEX || s[1] == 0.0, "MultAdd(complex) called for real matrix");   // sparsematrix_impl.hpp:394
    NGSB_CUDA(cudaSetDevice(A->ctx->device));
    SpmvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.x = x->d; a.y = y->d; a.sr = s[0]; a.si = s[1]; a.accumulate = true; a.epi = EPI_NONE;
    return spmv_launch(a);
}

extern "C" int ngsb_csr_mult(const ngsb_csr *A, const ngsb_vec *x, ngsb_vec *y)
{
    NvtxRange nv("SparseMatrix::Mult");
    NGSB_TRY(check_mult_args(A, x, y, "BaseMatrix::Mult"));
    NGSB_CUDA(cudaSetDevice(A->ctx->device));
    SpmvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.x = x->d; a.y = y->d; a.sr = 1.0; a.si = 0.0; a.accumulate = false; a.epi = EPI_NONE;
    return spmv_launch(a);
}

// SparseMatrix<double>::MultAdd(alpha, MultiVector x, MultiVector y), linalg/sparsematrix.cpp:2274-2351: groups of four
// vectors share one sweep over the matrix, the remainder goes vector by vector (the reference's own grouping)
extern "C" int ngsb_csr_multadd_multi(const ngsb_csr *A, size_t nvec, const double *alpha, const ngsb_vec *const *x, ngsb_vec *const *y)
{
    NGSB_REQUIRE(A && (nvec == 0 || (alpha && x && y)), "SparseMatrix::MultAdd(MultiVector): NULL argument");
    for (size_t k = 0; k < nvec; k++) {
        NGSB_TRY(check_mult_args(A, x[k], y[k], "SparseMatrix::MultAdd(MultiVector)"));
        for (size_t l = 0; l < nvec; l++) {
            const double *xb = x[l]->d, *xe = x[l]->d + x[l]->nscal, *yb = y[k]->d, *ye = y[k]->d + y[k]->nscal;
            NGSB_REQUIRE(xe <= yb || ye <= xb || x[l]->nscal == 0, "SparseMatrix::MultAdd(MultiVector): y[%zu] overlaps x[%zu]", k, l);
            if (l != k) {
                const double *zb = y[l]->d, *ze = y[l]->d + y[l]->nscal;
                NGSB_REQUIRE(ze <= yb || ye <= zb || y[l]->nscal == 0, "SparseMatrix::MultAdd(MultiVector): y[%zu] overlaps y[%zu]", k, l);
            }
        }
    }
    NGSB_CUDA(cudaSetDevice(A->ctx->device));
    size_t k = 0;
    const bool grouped = A->kind == NGSB_REAL && A->novf == 0 && A->ctx->spmv_algo == 0 && A->inner == nullptr;
    if (grouped)
        for (; k + 4 <= nvec; k += 4) {
            const double *xs[4] = {x[k]->d, x[k + 1]->d, x[k + 2]->d, x[k + 3]->d};
            double *ys[4] = {y[k]->d, y[k + 1]->d, y[k + 2]->d, y[k + 3]->d};
            NGSB_TRY(sell_launch_multi4(A, xs, ys, alpha + k));
        }
    for (; k < nvec; k++) {
        SpmvArgs a;
        memset(&a, 0, sizeof(a));
        a.A = A; a.x = x[k]->d; a.y = y[k]->d; a.sr = alpha[k]; a.si = 0.0; a.accumulate = true; a.epi = EPI_NONE;
        NGSB_TRY(spmv_launch(a));
    }
    return NGSB_OK;
}

extern "C" int ngsb_csr_download(const ngsb_csr *A, uint64_t *rowptr, int32_t *col, void *val)
{
    NGSB_REQUIRE(A, "ngsb_csr_download: A is NULL");
    if ((col || val) && A->nnz) NGSB_TRY(csr_ensure(A));
    ngsb_ctx *ctx = A->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    const size_t ms = kind_matscalars(A->kind);
    if (rowptr) NGSB_CUDA(cudaMemcpyAsync(rowptr, A->d_rowptr, (A->h + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (col && A->nnz) NGSB_CUDA(cudaMemcpyAsync(col, A->d_col, A->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (val && A->nnz) NGSB_CUDA(cudaMemcpyAsync(val, A->d_val, A->nnz * ms * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NGSB_OK;
}

extern "C" int ngsb_csr_layout(const ngsb_csr *A, uint64_t *sell_entries, uint32_t *overflow_rows, uint32_t *cap)
{
    NGSB_REQUIRE(A, "ngsb_csr_layout: A is NULL");
    if (A->inner) A = A->inner;
    if (sell_entries) *sell_entries = A->sell_entries;
    if (overflow_rows) *overflow_rows = A->novf;
    if (cap) *cap = A->sell_cap;
    return NGSB_OK;
}

// bytes the default (SELL) kernel streams per Mult with the layout as stored: padded entries, 16-bit column offsets
// where a slice allows them (8 + 2 + 4/32 bytes per entry instead of 12), slot -> row table, x once, y once
extern "C" int ngsb_csr_stream_bytes(const ngsb_csr *A, double *bytes, uint64_t *c16_entries)
{
    NGSB_REQUIRE(A && bytes, "ngsb_csr_stream_bytes: NULL argument");
    const double gather = A->inner ? (double)A->w * (A->kind == NGSB_COMPLEX ? 16.0 : (A->kind == NGSB_BLOCK3 ? 24.0 : 8.0)) * 2.0 : 0.0;   // x' = x[perm]
    if (A->inner) A = A->inner;
    const double S = A->kind == NGSB_COMPLEX ? 16.0 : 8.0;
    const double b = A->kind == NGSB_BLOCK3 ? 3.0 : 1.0;
    const double per_entry = b * b * S + 4.0;
    const double c16 = (double)A->sell_c16_entries, rest = (double)A->sell_entries - c16;
    *bytes = c16 * (b * b * S + 2.0 + 4.0 / 32.0) + rest * per_entry + 4.0 * (double)A->h + (double)(A->w + A->h) * b * S + gather;
    if (c16_entries) *c16_entries = A->sell_c16_entries;
    return NGSB_OK;
}

extern "C" int ngsb_csr_mult_bytes(const ngsb_csr *A, double *bytes)
{
    NGSB_REQUIRE(A && bytes, "ngsb_csr_mult_bytes: NULL argument");
    const double S = A->kind == NGSB_COMPLEX ? 16.0 : 8.0;
    const double b = A->kind == NGSB_BLOCK3 ? 3.0 : 1.0;
    *bytes = (double)A->nnz * (b * b * S + 4.0) + 4.0 * (double)A->h + (double)(A->h + A->w) * b * S;
    return NGSB_OK;
}
