// vec.cu -- context, device vectors, device scalars and the BaseVector kernels.
//
// Replaces ngscuda/unifiedvector.{hpp,cpp} (UnifiedVector, UnifiedScalar) and the
// CUDA_forall lambdas it launches (ngscuda/cuda_core.hpp:49-74: one thread per element,
// scalar 8-byte accesses, grid = n/256+1).  Here every update is a 128-bit vectorised
// grid-stride kernel sized to the SM count, and reductions are single-pass and
// deterministic (fixed chunking, last-block finish) with the result left on the device.
#include "common.cuh"

#include <stdarg.h>
#include <mutex>

namespace ngsb {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

SpanGuard::SpanGuard(ngsb_ctx *c, int klass) : ctx(c), idx(-1)
{
    ctx->launches++;
    if (!ctx->timing) return;
    TimedSpan sp;
    sp.klass = klass;
    auto get = [&]() {
        cudaEvent_t e;
        if (!ctx->event_pool.empty()) { e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    };
    sp.a = get();
    sp.b = get();
    cudaEventRecord(sp.a, ctx->stream);
    ctx->spans.push_back(sp);
    idx = (int)ctx->spans.size() - 1;
}

SpanGuard::~SpanGuard()
{
    if (idx >= 0) cudaEventRecord(ctx->spans[idx].b, ctx->stream);
}

int stage_reserve(ngsb_ctx *ctx, size_t bytes)
{
    if (ctx->h_stage_bytes >= bytes) return NGSB_OK;
    if (ctx->h_stage) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; }
    size_t want = bytes < (size_t(1) << 20) ? (size_t(1) << 20) : bytes;
    NGSB_CUDA(cudaMallocHost(&ctx->h_stage, want));
    ctx->h_stage_bytes = want;
    return NGSB_OK;
}

// ------------------------------------------------------------------------------------------
// elementwise kernels
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(256) fill_kernel(double *__restrict__ x, size_t N, double re, double im, int cplx,
                                                   int vec2)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec2) {
        // pairs: for complex (re,im), for real (re,re)
        double2 v = make_double2(re, cplx ? im : re);
        double2 *x2 = reinterpret_cast<double2 *>(x);
        size_t n2 = cplx ? N : N / 2;
        if (!cplx && (N & 1) && i == 0) x[N - 1] = re;
        for (; i < n2; i += stride) x2[i] = v;
    } else {
        for (; i < N; i += stride) x[i] = re;
    }
}

// y = s*x (+y).  HOSTS: scalar by value; DEV: scalar read from device memory.
template <bool CPLX, bool ACC, bool DEVS>
__global__ void __launch_bounds__(256) axpby_kernel(double *__restrict__ y, const double *__restrict__ x, size_t N,
                                                    double sr, double si, const double *__restrict__ ds, int neg)
{
    if (DEVS) {
        sr = ds[0];
        si = ds[1];
        if (neg) { sr = -sr; si = -si; }
    }
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (CPLX) {
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        double2 *y2 = reinterpret_cast<double2 *>(y);
        double2 s = make_double2(sr, si);
        for (; i < N; i += stride) {
            double2 p = cmul(s, x2[i]);
            if (ACC) { double2 o = y2[i]; p.x += o.x; p.y += o.y; }
            y2[i] = p;
        }
    } else {
        size_t n2 = N / 2;
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        double2 *y2 = reinterpret_cast<double2 *>(y);
        bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
        if (aligned) {
            for (size_t k = i; k < n2; k += stride) {
                double2 a = x2[k];
                double2 p;
                if (ACC) { double2 o = y2[k]; p.x = o.x + sr * a.x; p.y = o.y + sr * a.y; }
                else { p.x = sr * a.x; p.y = sr * a.y; }
                y2[k] = p;
            }
            if (i == 0 && (N & 1)) {
                size_t k = N - 1;
                y[k] = ACC ? y[k] + sr * x[k] : sr * x[k];
            }
        } else {
            for (size_t k = i; k < N; k += stride) y[k] = ACC ? y[k] + sr * x[k] : sr * x[k];
        }
    }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) scale_dev_kernel(double *__restrict__ x, size_t N, const double *__restrict__ ds)
{
    double sr = ds[0], si = ds[1];
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (CPLX) {
        double2 *x2 = reinterpret_cast<double2 *>(x);
        double2 s = make_double2(sr, si);
        for (; i < N; i += stride) x2[i] = cmul(x2[i], s);
    } else {
        for (; i < N; i += stride) x[i] *= sr;
    }
}

static inline int grid_for(ngsb_ctx *ctx, size_t work_items, int per_thread = 4)
{
    size_t blocks = (work_items + (size_t)256 * per_thread - 1) / ((size_t)256 * per_thread);
    size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int launch_fill(ngsb_ctx *ctx, double *x, size_t N, double re, double im, bool cplx)
{
    if (N == 0) return NGSB_OK;
    bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    SpanGuard g(ctx, KC_VEC);
    fill_kernel<<<grid_for(ctx, N), 256, 0, ctx->stream>>>(x, N, re, im, cplx ? 1 : 0, (aligned || cplx) ? 1 : 0);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

int launch_axpby(ngsb_ctx *ctx, double *y, const double *x, size_t N, double sr, double si, bool cplx, bool acc)
{
    if (N == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_VEC);
    int grid = grid_for(ctx, N);
    if (cplx) {
        if (acc) axpby_kernel<true, true, false><<<grid, 256, 0, ctx->stream>>>(y, x, N, sr, si, nullptr, 0);
        else axpby_kernel<true, false, false><<<grid, 256, 0, ctx->stream>>>(y, x, N, sr, si, nullptr, 0);
    } else {
        if (acc) axpby_kernel<false, true, false><<<grid, 256, 0, ctx->stream>>>(y, x, N, sr, si, nullptr, 0);
        else axpby_kernel<false, false, false><<<grid, 256, 0, ctx->stream>>>(y, x, N, sr, si, nullptr, 0);
    }
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

int launch_axpby_dev(ngsb_ctx *ctx, double *y, const double *x, size_t N, const double *ds, bool cplx, bool acc,
                     bool neg)
{
    if (N == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_VEC);
    int grid = grid_for(ctx, N);
    if (cplx) {
        if (acc) axpby_kernel<true, true, true><<<grid, 256, 0, ctx->stream>>>(y, x, N, 0, 0, ds, neg);
        else axpby_kernel<true, false, true><<<grid, 256, 0, ctx->stream>>>(y, x, N, 0, 0, ds, neg);
    } else {
        if (acc) axpby_kernel<false, true, true><<<grid, 256, 0, ctx->stream>>>(y, x, N, 0, 0, ds, neg);
        else axpby_kernel<false, false, true><<<grid, 256, 0, ctx->stream>>>(y, x, N, 0, 0, ds, neg);
    }
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

int launch_scale_dev(ngsb_ctx *ctx, double *x, size_t N, const double *ds, bool cplx)
{
    if (N == 0) return NGSB_OK;
    SpanGuard g(ctx, KC_VEC);
    int grid = grid_for(ctx, N);
    if (cplx) scale_dev_kernel<true><<<grid, 256, 0, ctx->stream>>>(x, N, ds);
    else scale_dev_kernel<false><<<grid, 256, 0, ctx->stream>>>(x, N, ds);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

// ------------------------------------------------------------------------------------------
// deterministic single-pass dot product
//   each CTA reduces one fixed contiguous chunk (thread-strided inside the chunk, then a
//   fixed shuffle/smem tree), writes partial[cta]; the last CTA to finish (atomic ticket)
//   adds the partials in index order.  Result depends only on (N, grid), never on timing.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of (a,b); result valid in thread 0
__device__ __forceinline__ double2 block_sum2(double a, double b)
{
    __shared__ double sa[32], sb[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { sa[wid] = a; sb[wid] = b; }
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    if (wid == 0) {
        a = lane < nw ? sa[lane] : 0.0;
        b = lane < nw ? sb[lane] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
    }
    return make_double2(a, b);
}

// last-block finish: returns true in thread 0 of the last CTA after summing partials in order
__device__ __forceinline__ bool finish_partials(double2 mine, double *partials, unsigned int *counter,
                                                double2 *total)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = mine.x;
        partials[2 * blockIdx.x + 1] = mine.y;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    // ordered sum by warp 0: lane l adds partials l, l+32, ... then fixed tree
    double a = 0.0, b = 0.0;
    if (threadIdx.x < 32) {
        __threadfence();
        for (unsigned int k = threadIdx.x; k < gridDim.x; k += 32) {
            a += __ldcg(&partials[2 * k]);
            b += __ldcg(&partials[2 * k + 1]);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (threadIdx.x == 0) {
            *total = make_double2(a, b);
            *counter = 0;
        }
    }
    return threadIdx.x == 0;
}

template <int MODE>
__global__ void __launch_bounds__(256) dot_kernel(const double *__restrict__ x, const double *__restrict__ y, size_t N,
                                                  double *__restrict__ partials, unsigned int *counter,
                                                  double *__restrict__ out)
{
    // chunk of this CTA
    size_t per = (N + gridDim.x - 1) / gridDim.x;
    per = (per + 1) & ~(size_t)1;
    size_t lo = (size_t)blockIdx.x * per;
    size_t hi = lo + per < N ? lo + per : N;
    double a = 0.0, b = 0.0;
    if (MODE == 0 || MODE == 3) {
        bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
        if (aligned) {
            const double2 *x2 = reinterpret_cast<const double2 *>(x);
            const double2 *y2 = reinterpret_cast<const double2 *>(y);
            size_t lo2 = lo / 2, hi2 = hi / 2;
            for (size_t k = lo2 + threadIdx.x; k < hi2; k += blockDim.x) {
                double2 u = x2[k];
                double2 v = (MODE == 3) ? u : y2[k];
                a = fma(u.x, v.x, a);
                b = fma(u.y, v.y, b);
            }
            if ((hi & 1) && hi == N && threadIdx.x == 0 && hi > lo) a = fma(x[N - 1], (MODE == 3 ? x[N - 1] : y[N - 1]), a);
        } else {
            for (size_t k = lo + threadIdx.x; k < hi; k += blockDim.x) a = fma(x[k], (MODE == 3 ? x[k] : y[k]), a);
        }
        a += b;
        b = 0.0;
    } else {
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        const double2 *y2 = reinterpret_cast<const double2 *>(y);
        for (size_t k = lo + threadIdx.x; k < hi; k += blockDim.x) {
            double2 u = x2[k], v = y2[k];
            if (MODE == 2) v.y = -v.y;
            a += u.x * v.x - u.y * v.y;
            b += u.x * v.y + u.y * v.x;
        }
    }
    double2 mine = block_sum2(a, b);
    double2 total;
    if (finish_partials(mine, partials, counter, &total)) {
        out[0] = total.x;
        out[1] = total.y;
    }
}

int launch_dot(ngsb_ctx *ctx, const double *x, const double *y, size_t N, int mode, double *d_out)
{
    SpanGuard g(ctx, KC_VEC);
    // complex modes count complex entries
    size_t blocks = (N + 4095) / 4096;
    size_t cap = (size_t)ctx->sm_count * 4;
    if (cap > (size_t)MAX_PARTIALS) cap = MAX_PARTIALS;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    int grid = (int)blocks;
    switch (mode) {
    case 0: dot_kernel<0><<<grid, 256, 0, ctx->stream>>>(x, y, N, ctx->d_partials, ctx->d_counter, d_out); break;
    case 1: dot_kernel<1><<<grid, 256, 0, ctx->stream>>>(x, y, N, ctx->d_partials, ctx->d_counter, d_out); break;
    case 2: dot_kernel<2><<<grid, 256, 0, ctx->stream>>>(x, y, N, ctx->d_partials, ctx->d_counter, d_out); break;
    default: dot_kernel<3><<<grid, 256, 0, ctx->stream>>>(x, x, N, ctx->d_partials, ctx->d_counter, d_out); break;
    }
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

// the same reduction restricted to entries whose owner byte is set (masked inner product of two
// CUMULATED parallel vectors, parallel/parallelvvector.cpp:305-314); mask index = i / mask_div
template <int MODE>
__global__ void __launch_bounds__(256) dot_masked_kernel(const double *__restrict__ x, const double *__restrict__ y, size_t N,
                                                         const uint8_t *__restrict__ mask, unsigned mask_div,
                                                         double *__restrict__ partials, unsigned int *counter, double *__restrict__ out)
{
    size_t per = (N + gridDim.x - 1) / gridDim.x;
    size_t lo = (size_t)blockIdx.x * per;
    size_t hi = lo + per < N ? lo + per : N;
    double a = 0.0, b = 0.0;
    for (size_t k = lo + threadIdx.x; k < hi; k += blockDim.x) {
        if (!mask[k / mask_div]) continue;
        if (MODE == 0 || MODE == 3) {
            a = fma(x[k], (MODE == 3 ? x[k] : y[k]), a);
        } else {
            double2 u = reinterpret_cast<const double2 *>(x)[k], v = reinterpret_cast<const double2 *>(y)[k];
            if (MODE == 2) v.y = -v.y;
            a += u.x * v.x - u.y * v.y;
            b += u.x * v.y + u.y * v.x;
        }
    }
    double2 mine = block_sum2(a, b);
    double2 total;
    if (finish_partials(mine, partials, counter, &total)) {
        out[0] = total.x;
        out[1] = total.y;
    }
}

int launch_dot_masked(ngsb_ctx *ctx, const double *x, const double *y, size_t N, int mode, double *d_out, const uint8_t *mask,
                      unsigned mask_div)
{
    if (!mask) return launch_dot(ctx, x, y, N, mode, d_out);
    SpanGuard g(ctx, KC_VEC);
    size_t blocks = (N + 4095) / 4096;
    size_t cap = (size_t)ctx->sm_count * 4;
    if (cap > (size_t)MAX_PARTIALS) cap = MAX_PARTIALS;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    int grid = (int)blocks;
    if (mask_div < 1) mask_div = 1;
    switch (mode) {
    case 0: dot_masked_kernel<0><<<grid, 256, 0, ctx->stream>>>(x, y, N, mask, mask_div, ctx->d_partials, ctx->d_counter, d_out); break;
    case 1: dot_masked_kernel<1><<<grid, 256, 0, ctx->stream>>>(x, y, N, mask, mask_div, ctx->d_partials, ctx->d_counter, d_out); break;
    case 2: dot_masked_kernel<2><<<grid, 256, 0, ctx->stream>>>(x, y, N, mask, mask_div, ctx->d_partials, ctx->d_counter, d_out); break;
    default: dot_masked_kernel<3><<<grid, 256, 0, ctx->stream>>>(x, x, N, mask, mask_div, ctx->d_partials, ctx->d_counter, d_out); break;
    }
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}


// in-place inclusive scan of rowlen[0..n] (three-phase, chunk per CTA)
__global__ void __launch_bounds__(256) scan_chunks_kernel(uint64_t *a, uint64_t n, uint64_t chunk, uint64_t *sums)
{
    __shared__ uint64_t sh[256];
    const uint64_t lo = (uint64_t)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    const uint64_t per = (chunk + 255) / 256;
    const uint64_t tlo = lo + threadIdx.x * per < hi ? lo + threadIdx.x * per : hi;
    const uint64_t thi = tlo + per < hi ? tlo + per : hi;
    uint64_t s = 0;
    for (uint64_t i = tlo; i < thi; i++) s += a[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int t = 0; t < 256; t++) { uint64_t v = sh[t]; sh[t] = run; run += v; }
        sums[blockIdx.x] = run;
    }
    __syncthreads();
    uint64_t run = sh[threadIdx.x];
    for (uint64_t i = tlo; i < thi; i++) { run += a[i]; a[i] = run; }
}

__global__ void scan_sums_kernel(uint64_t *sums, int nchunks)
{
    uint64_t run = 0;
    for (int i = 0; i < nchunks; i++) { uint64_t v = sums[i]; sums[i] = run; run += v; }
}

__global__ void __launch_bounds__(256) scan_add_kernel(uint64_t *a, uint64_t n, uint64_t chunk, const uint64_t *sums)
{
    const uint64_t lo = (uint64_t)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    const uint64_t add = sums[blockIdx.x];
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) a[i] += add;
}


int device_scan_u64(ngsb_ctx *ctx, uint64_t *d_a, uint64_t n)
{
    if (n == 0) return NGSB_OK;
    const uint64_t chunk = 1 << 16;
    const int nchunks = (int)((n + chunk - 1) / chunk);
    uint64_t *d_sums = nullptr;
    NGSB_CUDA(cudaMalloc(&d_sums, (size_t)nchunks * sizeof(uint64_t)));
    scan_chunks_kernel<<<nchunks, 256, 0, ctx->stream>>>(d_a, n, chunk, d_sums);
    scan_sums_kernel<<<1, 1, 0, ctx->stream>>>(d_sums, nchunks);
    scan_add_kernel<<<nchunks, 256, 0, ctx->stream>>>(d_a, n, chunk, d_sums);
    ctx->launches += 3;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_sums);
    if (e != cudaSuccess) { set_error("device_scan_u64: %s", cudaGetErrorString(e)); return NGSB_ERR_CUDA; }
    return NGSB_OK;
}

// scalar expression kernels (ngscuda/unifiedvector.hpp:126-135)
__global__ void scalar_op_kernel(double *out, const double *a, const double *b, int op)
{
    double ar = a[0], ai = a[1];
    if (op == 0) {   // a / b (complex division, same formula as std::complex for finite values)
        double br = b[0], bi = b[1];
        if (bi == 0.0 && ai == 0.0) { out[0] = ar / br; out[1] = 0.0; }
        else {
            double den = br * br + bi * bi;
            out[0] = (ar * br + ai * bi) / den;
            out[1] = (ai * br - ar * bi) / den;
        }
    } else if (op == 1) { out[0] = -ar; out[1] = -ai; }
    else { out[0] = ar; out[1] = ai; }
}

} // namespace ngsb

using namespace ngsb;

// ==========================================================================================
// C ABI
// ==========================================================================================

extern "C" const char *ngsb_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *ngsb_version(void) { return "ngsb200 0.1 (sm_100a)"; }

extern "C" int ngsb_ctx_create(int device, ngsb_ctx **out)
{
    NGSB_REQUIRE(out != nullptr, "ngsb_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("ngsb_ctx_create: no CUDA device available (%s); libngsb200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return NGSB_ERR_CUDA;
    }
    if (device < 0) {
        const char *env = getenv("NGS_CUDA_DEVICE_INDEX");   // ngscuda/cuda_ngstd.cpp:84-93
        device = env ? atoi(env) : 0;
    }
    NGSB_REQUIRE(device < ndev, "ngsb_ctx_create: device %d out of range (%d devices)", device, ndev);
    NGSB_CUDA(cudaSetDevice(device));
    ngsb_ctx *ctx = new ngsb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    NGSB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    NGSB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    NGSB_CUDA(cudaMalloc(&ctx->d_partials, sizeof(double) * 2 * MAX_PARTIALS));
    NGSB_CUDA(cudaMalloc(&ctx->d_counter, sizeof(unsigned int) * 16));
    NGSB_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int) * 16, ctx->stream));
    NGSB_CUDA(cudaMallocHost(&ctx->h_pinned, sizeof(double) * 4096));
    NGSB_CUDA(cudaEventCreateWithFlags(&ctx->stage_free, cudaEventDisableTiming));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx;
    return NGSB_OK;
}

extern "C" int ngsb_ctx_destroy(ngsb_ctx *ctx)
{
    if (!ctx) return NGSB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ws && ctx->ws_free) ctx->ws_free(ctx->ws);
    for (auto &sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_counter);
    cudaFreeHost(ctx->h_pinned);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaEventDestroy(ctx->stage_free);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return NGSB_OK;
}

extern "C" int ngsb_ctx_sync(ngsb_ctx *ctx)
{
    NGSB_REQUIRE(ctx, "ngsb_ctx_sync: ctx is NULL");
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NGSB_OK;
}

extern "C" int ngsb_ctx_device(const ngsb_ctx *ctx, int *device, int *sm_count)
{
    NGSB_REQUIRE(ctx, "ngsb_ctx_device: ctx is NULL");
    if (device) *device = ctx->device;
    if (sm_count) *sm_count = ctx->sm_count;
    return NGSB_OK;
}

extern "C" void *ngsb_ctx_stream(ngsb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int ngsb_ctx_launch_count(const ngsb_ctx *ctx, uint64_t *count)
{
    NGSB_REQUIRE(ctx && count, "ngsb_ctx_launch_count: NULL argument");
    *count = ctx->launches;
    return NGSB_OK;
}

extern "C" int ngsb_ctx_set_option(ngsb_ctx *ctx, const char *name, long value)
{
    NGSB_REQUIRE(ctx && name, "ngsb_ctx_set_option: NULL argument");
    if (!strcmp(name, "spmv_algo")) { NGSB_REQUIRE(value >= 0 && value <= 3, "spmv_algo must be 0 (auto = SELL), 1 (sub-warp CSR), 2 (TMA-streamed CSR), 3 (SELL)"); ctx->spmv_algo = value; }
    else if (!strcmp(name, "cg_batch")) { NGSB_REQUIRE(value >= 1 && value <= 4096, "cg_batch out of range"); ctx->cg_batch = value; }
    else if (!strcmp(name, "spmv_ctas_per_sm")) { NGSB_REQUIRE(value >= 0 && value <= 7000, "spmv_ctas_per_sm out of range"); ctx->spmv_ctas_per_sm = value; }
    else if (!strcmp(name, "timing")) { ctx->timing = value ? 1 : 0; }
    else if (!strcmp(name, "reorder")) { NGSB_REQUIRE(value >= -1 && value <= 1, "reorder must be -1 (automatic), 0 (off) or 1 (always)"); ctx->reorder = value; }
    else if (!strcmp(name, "csr_keep")) { NGSB_REQUIRE(value >= -1 && value <= 1, "csr_keep must be -1 (automatic), 0 (release) or 1 (keep)"); ctx->csr_keep = value; }
    else if (!strcmp(name, "reorder_slot_order")) { NGSB_REQUIRE(value == 0 || value == 1, "reorder_slot_order must be 0 or 1"); ctx->reorder_slot_order = value; }
    else if (!strcmp(name, "reorder_min_rows")) { NGSB_REQUIRE(value >= 0, "reorder_min_rows must be >= 0"); ctx->reorder_min_rows = value; }
    else if (!strcmp(name, "cg_persistent")) { NGSB_REQUIRE(value >= -1 && value <= 1, "cg_persistent must be -1 (automatic), 0 or 1"); ctx->cg_persistent = value; }
    else if (!strcmp(name, "gmres_orth")) { NGSB_REQUIRE(value == 0 || value == 1, "gmres_orth must be 0 (serial MGS) or 1 (one batched reduction)"); ctx->gmres_orth = value; }
    else if (!strcmp(name, "cg_stream_hints")) { NGSB_REQUIRE(value == 0 || value == 1, "cg_stream_hints must be 0 or 1"); ctx->cg_stream_hints = value; }
    else if (!strcmp(name, "cg_chunked")) { NGSB_REQUIRE(value == 0 || value == 1, "cg_chunked must be 0 or 1"); ctx->cg_chunked = value; }
    else if (!strcmp(name, "cg_fold_u")) { NGSB_REQUIRE(value == 0 || value == 1, "cg_fold_u must be 0 or 1"); ctx->cg_fold_u = value; }
    else if (!strcmp(name, "dist_fused_push")) { NGSB_REQUIRE(value == 0 || value == 1, "dist_fused_push must be 0 or 1"); ctx->dist_fused_push = value; }
    else if (!strcmp(name, "dist_overlap")) { NGSB_REQUIRE(value == 0 || value == 1, "dist_overlap must be 0 or 1"); ctx->dist_overlap = value; }
    else if (!strcmp(name, "sell_variant")) { NGSB_REQUIRE(value >= 0 && value <= 8, "sell_variant out of range"); ctx->sell_variant = value; }
    else if (!strcmp(name, "sell_schedule")) { NGSB_REQUIRE(value >= 0 && value <= 2, "sell_schedule must be 0 (off), 1 (auto), 2 (on)"); ctx->sell_schedule = value; }
    else if (!strcmp(name, "sell_pf_steps")) { NGSB_REQUIRE(value >= 0 && value <= 16, "sell_pf_steps out of range"); ctx->sell_pf_steps = value; }
    else if (!strcmp(name, "sell_pf_next")) { NGSB_REQUIRE(value >= 0 && value <= 64, "sell_pf_next out of range"); ctx->sell_pf_next = value; }
    else if (!strcmp(name, "sell_c16")) { NGSB_REQUIRE(value == 0 || value == 1, "sell_c16 must be 0 or 1"); ctx->sell_c16 = value; }
    else if (!strcmp(name, "sell_c16_all")) { NGSB_REQUIRE(value == 0 || value == 1, "sell_c16_all must be 0 or 1"); ctx->sell_c16_all = value; }
    else if (!strcmp(name, "sell_sigma")) { NGSB_REQUIRE(value >= -1 && value <= (1 << 24), "sell_sigma out of range"); ctx->sell_sigma = value; }
    else if (!strcmp(name, "sell_cap")) { NGSB_REQUIRE(value >= 0 && value <= (1 << 20), "sell_cap out of range"); ctx->sell_cap = value; }
    else if (!strcmp(name, "spmv_tile")) { NGSB_REQUIRE(value == 0 || (value >= 256 && value <= 8192 && value % 256 == 0), "spmv_tile out of range"); ctx->spmv_tile = value; }
    else if (!strcmp(name, "spmv_ncw")) { NGSB_REQUIRE(value >= 0 && value <= 16, "spmv_ncw out of range"); ctx->spmv_ncw = value; }
    else if (!strcmp(name, "spmv_stages")) { NGSB_REQUIRE(value >= 0 && value <= 8, "spmv_stages out of range"); ctx->spmv_stages = value; }
    else if (!strcmp(name, "spmv_subwarp")) { NGSB_REQUIRE(value == 0 || value == 4 || value == 8 || value == 16 || value == 32, "spmv_subwarp must be 0,4,8,16,32"); ctx->spmv_subwarp = value; }
    else { set_error("ngsb_ctx_set_option: unknown option '%s'", name); return NGSB_ERR_INVALID; }
    return NGSB_OK;
}

extern "C" int ngsb_ctx_kernel_time(ngsb_ctx *ctx, const char *klass, double *ms, uint64_t *launches)
{
    NGSB_REQUIRE(ctx && klass && ms, "ngsb_ctx_kernel_time: NULL argument");
    int want = -1;
    if (!strcmp(klass, "spmv")) want = KC_SPMV;
    else if (!strcmp(klass, "cgupdate")) want = KC_CGUPDATE;
    else if (!strcmp(klass, "vec")) want = KC_VEC;
    else if (!strcmp(klass, "other")) want = KC_OTHER;
    else if (!strcmp(klass, "all")) want = -1;
    else { set_error("ngsb_ctx_kernel_time: unknown class '%s'", klass); return NGSB_ERR_INVALID; }
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    double total = 0.0;
    uint64_t cnt = 0;
    for (auto &sp : ctx->spans) {
        if (want >= 0 && sp.klass != want) continue;
        float t = 0.f;
        NGSB_CUDA(cudaEventElapsedTime(&t, sp.a, sp.b));
        total += t;
        cnt++;
    }
    *ms = total;
    if (launches) *launches = cnt;
    return NGSB_OK;
}

extern "C" int ngsb_ctx_kernel_time_reset(ngsb_ctx *ctx)
{
    NGSB_REQUIRE(ctx, "ngsb_ctx_kernel_time_reset: ctx is NULL");
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &sp : ctx->spans) { ctx->event_pool.push_back(sp.a); ctx->event_pool.push_back(sp.b); }
    ctx->spans.clear();
    return NGSB_OK;
}

// ---- vectors ------------------------------------------------------------------------------

extern "C" int ngsb_vec_create(ngsb_ctx *ctx, size_t n_entries, int kind, ngsb_vec **out)
{
    NGSB_REQUIRE(ctx && out, "ngsb_vec_create: NULL argument");
    NGSB_REQUIRE(kind_valid(kind), "ngsb_vec_create: bad kind %d", kind);
    NGSB_CUDA(cudaSetDevice(ctx->device));
    ngsb_vec *v = new ngsb_vec();
    v->ctx = ctx;
    v->n = n_entries;
    v->kind = kind;
    v->nscal = n_entries * kind_scalars(kind);
    void *p = nullptr;
    size_t bytes = v->nscal * sizeof(double);
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        delete v;
        set_error("ngsb_vec_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return NGSB_ERR_NOMEM;
    }
    v->storage = std::shared_ptr<void>(p, [](void *q) { cudaFree(q); });
    v->d = (double *)p;
    // a fresh UnifiedVector is not initialised by the reference either; zero it for safety
    NGSB_CUDA(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    *out = v;
    return NGSB_OK;
}

extern "C" int ngsb_vec_destroy(ngsb_vec *v)
{
    if (!v) return NGSB_OK;
    // storage may still be in use by enqueued work: cudaFree synchronises implicitly
    delete v;
    return NGSB_OK;
}

extern "C" int ngsb_vec_info(const ngsb_vec *v, size_t *n_entries, int *kind, size_t *n_scalars)
{
    NGSB_REQUIRE(v, "ngsb_vec_info: v is NULL");
    if (n_entries) *n_entries = v->n;
    if (kind) *kind = v->kind;
    if (n_scalars) *n_scalars = v->nscal;
    return NGSB_OK;
}

extern "C" void *ngsb_vec_devptr(ngsb_vec *v) { return v ? (void *)v->d : nullptr; }

extern "C" int ngsb_vec_range(ngsb_vec *v, size_t begin, size_t end, ngsb_vec **view)
{
    NGSB_REQUIRE(v && view, "ngsb_vec_range: NULL argument");
    NGSB_REQUIRE(begin <= end && end <= v->n, "ngsb_vec_range: range [%zu,%zu) outside vector of size %zu", begin, end, v->n);
    ngsb_vec *r = new ngsb_vec();
    r->ctx = v->ctx;
    r->n = end - begin;
    r->kind = v->kind;
    r->nscal = r->n * kind_scalars(v->kind);
    r->d = v->d + begin * kind_scalars(v->kind);
    r->storage = v->storage;
    *view = r;
    return NGSB_OK;
}

extern "C" int ngsb_vec_h2d(ngsb_vec *v, const void *host, size_t first_entry, size_t n_entries)
{
    NGSB_REQUIRE(v && (host || n_entries == 0), "ngsb_vec_h2d: NULL argument");
    NGSB_REQUIRE(first_entry + n_entries <= v->n, "ngsb_vec_h2d: range [%zu,%zu) outside vector of size %zu",
                 first_entry, first_entry + n_entries, v->n);
    if (n_entries == 0) return NGSB_OK;
    ngsb_ctx *ctx = v->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    size_t ks = kind_scalars(v->kind);
    size_t bytes = n_entries * ks * sizeof(double);
    double *dst = v->d + first_entry * ks;
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        NGSB_CUDA(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        // caller-owned pinned memory: make the copy complete before returning (buffer reuse)
        NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        // pageable source: stage through pinned memory in chunks
        const size_t CH = size_t(64) << 20;
        NGSB_TRY(stage_reserve(ctx, bytes < CH ? bytes : CH));
        size_t off = 0;
        while (off < bytes) {
            size_t len = bytes - off < CH ? bytes - off : CH;
            NGSB_CUDA(cudaEventSynchronize(ctx->stage_free));
            memcpy(ctx->h_stage, (const char *)host + off, len);
            NGSB_CUDA(cudaMemcpyAsync((char *)dst + off, ctx->h_stage, len, cudaMemcpyHostToDevice, ctx->stream));
            NGSB_CUDA(cudaEventRecord(ctx->stage_free, ctx->stream));
            off += len;
        }
    }
    return NGSB_OK;
}

extern "C" int ngsb_vec_d2h(const ngsb_vec *v, void *host, size_t first_entry, size_t n_entries)
{
    NGSB_REQUIRE(v && (host || n_entries == 0), "ngsb_vec_d2h: NULL argument");
    NGSB_REQUIRE(first_entry + n_entries <= v->n, "ngsb_vec_d2h: range [%zu,%zu) outside vector of size %zu",
                 first_entry, first_entry + n_entries, v->n);
    ngsb_ctx *ctx = v->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    size_t ks = kind_scalars(v->kind);
    size_t bytes = n_entries * ks * sizeof(double);
    if (bytes) NGSB_CUDA(cudaMemcpyAsync(host, v->d + first_entry * ks, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NGSB_OK;
}

static int check_same(const ngsb_vec *a, const ngsb_vec *b, const char *who)
{
    NGSB_REQUIRE(a && b, "%s: NULL vector", who);
    NGSB_REQUIRE(a->ctx == b->ctx, "%s: vectors belong to different contexts", who);
    // reference: "BaseVector::Add: size of me = .. != size of other = .." (linalg/basevector.cpp:208-210)
    NGSB_REQUIRE(a->n == b->n, "%s: size of me = %zu != size of other = %zu", who, a->n, b->n);
    NGSB_REQUIRE(a->nscal == b->nscal, "%s: entry kinds differ (%d vs %d)", who, a->kind, b->kind);
    return NGSB_OK;
}

extern "C" int ngsb_vec_set_scalar(ngsb_vec *x, const double s[2])
{
    NGSB_REQUIRE(x && s, "ngsb_vec_set_scalar: NULL argument");
    NGSB_CUDA(cudaSetDevice(x->ctx->device));
    bool cplx = x->kind == NGSB_COMPLEX;
    return launch_fill(x->ctx, x->d, cplx ? x->n : x->nscal, s[0], s[1], cplx);
}

extern "C" int ngsb_vec_scale(ngsb_vec *x, const double s[2])
{
    NGSB_REQUIRE(x && s, "ngsb_vec_scale: NULL argument");
    bool cplx = x->kind == NGSB_COMPLEX;
    if (s[0] == 1.0 && (!cplx || s[1] == 0.0)) return NGSB_OK;   // linalg/basevector.cpp:77
    NGSB_CUDA(cudaSetDevice(x->ctx->device));
    return launch_axpby(x->ctx, x->d, x->d, cplx ? x->n : x->nscal, s[0], cplx ? s[1] : 0.0, cplx, false);
}

extern "C" int ngsb_vec_set(ngsb_vec *y, const double s[2], const ngsb_vec *x)
{
    NGSB_TRY(check_same(y, x, "BaseVector::Set"));
    NGSB_REQUIRE(s, "ngsb_vec_set: s is NULL");
    bool cplx = y->kind == NGSB_COMPLEX;
    if (y->d == x->d && s[0] == 1.0 && (!cplx || s[1] == 0.0)) return NGSB_OK;   // basevector.cpp:155
    NGSB_CUDA(cudaSetDevice(y->ctx->device));
    return launch_axpby(y->ctx, y->d, x->d, cplx ? y->n : y->nscal, s[0], cplx ? s[1] : 0.0, cplx, false);
}

extern "C" int ngsb_vec_axpy(ngsb_vec *y, const double s[2], const ngsb_vec *x)
{
    NGSB_TRY(check_same(y, x, "BaseVector::Add"));
    NGSB_REQUIRE(s, "ngsb_vec_axpy: s is NULL");
    bool cplx = y->kind == NGSB_COMPLEX;
    NGSB_CUDA(cudaSetDevice(y->ctx->device));
    return launch_axpby(y->ctx, y->d, x->d, cplx ? y->n : y->nscal, s[0], cplx ? s[1] : 0.0, cplx, true);
}

static int read_result(ngsb_ctx *ctx, const double *d_src, double out[2])
{
    NGSB_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_src, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    out[0] = ctx->h_pinned[0];
    out[1] = ctx->h_pinned[1];
    return NGSB_OK;
}

extern "C" int ngsb_vec_dot(const ngsb_vec *x, const ngsb_vec *y, int conjugate, double out[2])
{
    NGSB_TRY(check_same(x, y, "BaseVector::InnerProduct"));
    NGSB_REQUIRE(out, "ngsb_vec_dot: out is NULL");
    ngsb_ctx *ctx = x->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    bool cplx = x->kind == NGSB_COMPLEX;
    double *d_res = ctx->d_partials + 2 * (MAX_PARTIALS - 1);   // last slot is never a partial (grid < MAX)
    NGSB_TRY(launch_dot(ctx, x->d, y->d, cplx ? x->n : x->nscal, cplx ? (conjugate ? 2 : 1) : 0, d_res));
    NGSB_TRY(read_result(ctx, d_res, out));
    if (!cplx) out[1] = 0.0;
    return NGSB_OK;
}

extern "C" int ngsb_vec_nrm2(const ngsb_vec *x, double *out)
{
    NGSB_REQUIRE(x && out, "ngsb_vec_nrm2: NULL argument");
    ngsb_ctx *ctx = x->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    double *d_res = ctx->d_partials + 2 * (MAX_PARTIALS - 1);
    NGSB_TRY(launch_dot(ctx, x->d, x->d, x->nscal, 3, d_res));
    double r[2];
    NGSB_TRY(read_result(ctx, d_res, r));
    *out = sqrt(r[0]);
    return NGSB_OK;
}

// ---- device scalars -----------------------------------------------------------------------

extern "C" int ngsb_scalar_create(ngsb_ctx *ctx, ngsb_scalar **out)
{
    NGSB_REQUIRE(ctx && out, "ngsb_scalar_create: NULL argument");
    NGSB_CUDA(cudaSetDevice(ctx->device));
    ngsb_scalar *s = new ngsb_scalar();
    s->ctx = ctx;
    NGSB_CUDA(cudaMalloc(&s->d, 2 * sizeof(double)));
    NGSB_CUDA(cudaMemsetAsync(s->d, 0, 2 * sizeof(double), ctx->stream));
    *out = s;
    return NGSB_OK;
}

extern "C" int ngsb_scalar_destroy(ngsb_scalar *s)
{
    if (!s) return NGSB_OK;
    cudaFree(s->d);
    delete s;
    return NGSB_OK;
}

extern "C" int ngsb_scalar_set(ngsb_scalar *s, const double v[2])
{
    NGSB_REQUIRE(s && v, "ngsb_scalar_set: NULL argument");
    NGSB_CUDA(cudaSetDevice(s->ctx->device));
    return launch_fill(s->ctx, s->d, 1, v[0], v[1], true);
}

extern "C" int ngsb_scalar_get(const ngsb_scalar *s, double v[2])
{
    NGSB_REQUIRE(s && v, "ngsb_scalar_get: NULL argument");
    NGSB_CUDA(cudaSetDevice(s->ctx->device));
    return read_result(s->ctx, s->d, v);
}

static int scalar_op(ngsb_scalar *out, const ngsb_scalar *a, const ngsb_scalar *b, int op)
{
    NGSB_REQUIRE(out && a && (op != 0 || b), "ngsb_scalar op: NULL argument");
    ngsb_ctx *ctx = out->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    SpanGuard g(ctx, KC_OTHER);
    scalar_op_kernel<<<1, 1, 0, ctx->stream>>>(out->d, a->d, b ? b->d : a->d, op);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

extern "C" int ngsb_scalar_div(ngsb_scalar *out, const ngsb_scalar *a, const ngsb_scalar *b) { return scalar_op(out, a, b, 0); }
extern "C" int ngsb_scalar_neg(ngsb_scalar *out, const ngsb_scalar *a) { return scalar_op(out, a, nullptr, 1); }
extern "C" int ngsb_scalar_copy(ngsb_scalar *out, const ngsb_scalar *a) { return scalar_op(out, a, nullptr, 2); }

extern "C" int ngsb_vec_dot_dev(const ngsb_vec *x, const ngsb_vec *y, int conjugate, ngsb_scalar *out)
{
    NGSB_TRY(check_same(x, y, "BaseVector::InnerProduct"));
    NGSB_REQUIRE(out, "ngsb_vec_dot_dev: out is NULL");
    ngsb_ctx *ctx = x->ctx;
    NGSB_CUDA(cudaSetDevice(ctx->device));
    bool cplx = x->kind == NGSB_COMPLEX;
    return launch_dot(ctx, x->d, y->d, cplx ? x->n : x->nscal, cplx ? (conjugate ? 2 : 1) : 0, out->d);
}

extern "C" int ngsb_vec_axpy_dev(ngsb_vec *y, const ngsb_scalar *s, const ngsb_vec *x)
{
    NGSB_TRY(check_same(y, x, "BaseVector::Add"));
    NGSB_REQUIRE(s, "ngsb_vec_axpy_dev: s is NULL");
    bool cplx = y->kind == NGSB_COMPLEX;
    NGSB_CUDA(cudaSetDevice(y->ctx->device));
    return launch_axpby_dev(y->ctx, y->d, x->d, cplx ? y->n : y->nscal, s->d, cplx, true, false);
}

extern "C" int ngsb_vec_scale_dev(ngsb_vec *x, const ngsb_scalar *s)
{
    NGSB_REQUIRE(x && s, "ngsb_vec_scale_dev: NULL argument");
    bool cplx = x->kind == NGSB_COMPLEX;
    NGSB_CUDA(cudaSetDevice(x->ctx->device));
    return launch_scale_dev(x->ctx, x->d, cplx ? x->n : x->nscal, s->d, cplx);
}
