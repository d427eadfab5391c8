// common.cuh -- shared definitions of libngsb200 (internal, not part of the C ABI)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <memory>
#include <algorithm>

#include <nvtx3/nvToolsExt.h>     // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "../../include/ngsb200.h"

namespace ngsb {

void set_error(const char *fmt, ...);

#define NGSB_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ngsb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                            __LINE__);                                                         \
            return NGSB_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

#define NGSB_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            ngsb::set_error(__VA_ARGS__);  \
            return NGSB_ERR_INVALID;       \
        }                                  \
    } while (0)

#define NGSB_TRY(call)                 \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != NGSB_OK) return rc__; \
    } while (0)

// NVTX range named after the reference's Timer of the same region ("SparseMatrix::MultAdd", "CG solver", ...),
// so that a timeline of the library reads like the reference's own profile
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// kernel classes for the optional per-class device timing (option "timing")
enum KClass { KC_SPMV = 0, KC_CGUPDATE = 1, KC_VEC = 2, KC_OTHER = 3, KC_COUNT = 4 };

struct TimedSpan {
    cudaEvent_t a, b;
    int klass;
};

} // namespace ngsb

// device-resident scalar block used by the solvers (all (re,im) pairs)
struct ngsb_scalar {
    ngsb_ctx *ctx;
    double *d;      // 2 doubles on the device
};

struct ngsb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    // options
    long spmv_algo = 0;          // 0 auto, 1 subwarp, 2 tma-stream
    long cg_batch = 16;          // iterations enqueued between two polls of the stop flag
    long spmv_ctas_per_sm = 0;   // 0 = kernel default
    long sell_cap = 0;           // 0 = max(64, 4 * mean row length)
    long sell_schedule = 1;      // order slices by their smallest first column (locality of the x gathers)
    long sell_sigma = -1;        // rows sorted by length inside windows of sigma rows; -1 = 4096, 0/1 = off
    long sell_variant = 0;       // inner-loop variant of the real SELL kernel (tuning)
    long sell_pf_steps = 0;      // compressed slices: in-slice L2 prefetch distance (steps of 4 packets), 0 = off
    long sell_pf_next = 0;       // compressed slices: packets of the next slice prefetched at slice start, 0 = off
    long sell_c16 = 1;           // 16-bit column offsets where a slice allows it (read at matrix creation and at launch)
    long sell_c16_all = 0;       // the same for complex and 3x3-block matrices (18.1 instead of 20, 74.1 instead of 76 bytes per entry)
    long spmv_tile = 0, spmv_ncw = 0, spmv_stages = 0, spmv_subwarp = 0;   // 0 = default; read when a matrix is created
    long cg_persistent = -1;     // real Jacobi-PCG as one persistent cooperative kernel: -1 automatic (< 4 M rows), 0 off, 1 on
    long gmres_orth = 1;         // GMRES orthogonalisation: 1 = V^T w in one batched reduction + triangular solve (2 (j+1) vector passes per
                                 // step, same coefficients as MGS in exact arithmetic), 0 = the reference's serial modified Gram-Schmidt
    long cg_stream_hints = 0;    // CG update kernel: streaming loads of u, d, As, diagonal and evict-first stores of u, d (A/B)
    long cg_chunked = 0;         // CG update kernel: contiguous chunk per CTA instead of the grid-stride split (A/B)
    long cg_fold_u = 0;          // CG: `u += al s` in the direction kernel instead of the update kernel (10 vector passes, not 11)
    long dist_fused_push = 1;    // distributed CG, peer-memory path: interface rows are stored into the neighbours' receive areas by
                                 // the product kernel itself (no separate push kernel); read when a parallel matrix is created
    long dist_overlap = 0;       // distributed CG: interface slices first, push, interior slices while the values travel
                                 // (read when a parallel matrix is created; peer-memory data path only)
    long csr_keep = -1;          // CSR column/value arrays after the SELL copy exists: 1 keep, 0 release (rebuilt from the SELL copy on
                                 // demand), -1 automatic = release above 4 GiB (read when a matrix is created)
    long reorder = -1;           // internal Cuthill-McKee reordering of square matrices: 0 off, 1 always, -1 automatic
                                 // (read when a matrix is created)
    long reorder_slot_order = 0; // the internal permutation also absorbs the SELL length sort (inner matrix numbered in slot order).
                                 // Measured (profiles/r2_slot_order_*.json): same DRAM traffic, product 5 % slower (x gathers
                                 // scrambled inside the sort windows), CG loop +1 % / +0 % at 13.6 M / 108 M dofs -> off
    long reorder_min_rows = 32768;   // automatic mode: smaller matrices keep their numbering
    long timing = 0;
    // reduction workspace (partials + counters), pinned host scratch
    double *d_partials = nullptr;   // 2 * max_partials doubles
    unsigned int *d_counter = nullptr;
    double *h_pinned = nullptr;     // small pinned scratch (>= 4096 doubles)
    void *h_stage = nullptr;        // pinned staging for h2d/d2h
    size_t h_stage_bytes = 0;
    cudaEvent_t stage_free = nullptr;
    std::vector<ngsb::TimedSpan> spans;
    std::vector<cudaEvent_t> event_pool;
    // solver workspace cache (krylov.cu), released by ngsb_ctx_destroy
    void *ws = nullptr;
    void (*ws_free)(void *) = nullptr;
};

struct ngsb_vec {
    ngsb_ctx *ctx;
    size_t n;            // entries
    int kind;
    size_t nscal;        // doubles
    double *d;           // device pointer (may alias parent storage)
    std::shared_ptr<void> storage;   // owner of the allocation
};

namespace ngsb {

inline size_t kind_scalars(int kind) { return kind == NGSB_REAL ? 1 : (kind == NGSB_COMPLEX ? 2 : 3); }
inline size_t kind_matscalars(int kind) { return kind == NGSB_REAL ? 1 : (kind == NGSB_COMPLEX ? 2 : 9); }
inline bool kind_valid(int kind) { return kind == NGSB_REAL || kind == NGSB_COMPLEX || kind == NGSB_BLOCK3; }

static const int MAX_PARTIALS = 1 << 20;      // per-CTA partials of the fused dots (the product may run one 8-slice step per CTA)

// RAII-less helper: record a timed span around a launch when ctx->timing is on
struct SpanGuard {
    ngsb_ctx *ctx;
    int idx;
    SpanGuard(ngsb_ctx *c, int klass);
    ~SpanGuard();
};

int stage_reserve(ngsb_ctx *ctx, size_t bytes);
// in-place inclusive prefix sum of n uint64 on the context's stream
int device_scan_u64(ngsb_ctx *ctx, uint64_t *d_a, uint64_t n);

// ---- vector kernels (vec.cu), all enqueue on ctx->stream --------------------------------
int launch_fill(ngsb_ctx *ctx, double *x, size_t N, double re, double im, bool cplx);
// y = a*x (+ y if accumulate); complex if cplx (N complex entries), scalars from host
int launch_axpby(ngsb_ctx *ctx, double *y, const double *x, size_t N, double sr, double si, bool cplx,
                 bool accumulate);
// scalar taken from device memory (2 doubles); neg => use -s
int launch_axpby_dev(ngsb_ctx *ctx, double *y, const double *x, size_t N, const double *ds, bool cplx,
                     bool accumulate, bool neg);
int launch_scale_dev(ngsb_ctx *ctx, double *x, size_t N, const double *ds, bool cplx);
// deterministic dot: out (device, 2 doubles) = sum x_i * (conj? conj(y_i) : y_i)
// mode: 0 real, 1 complex bilinear, 2 complex conj(y), 3 real sum of squares of x (norm^2)
int launch_dot(ngsb_ctx *ctx, const double *x, const double *y, size_t N, int mode, double *d_out);
// the same over the entries whose mask byte (index i / mask_div) is set; mask == NULL: all
int launch_dot_masked(ngsb_ctx *ctx, const double *x, const double *y, size_t N, int mode, double *d_out, const uint8_t *mask,
                      unsigned mask_div);

} // namespace ngsb
