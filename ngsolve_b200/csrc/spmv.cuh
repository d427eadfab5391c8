// spmv.cuh -- internal interface of the CSR SpMV (spmv.cu), used by krylov.cu / dist.cu
#pragma once
#include "common.cuh"
#include "cg_state.cuh"

// one TMA-streamed row block: rows [row_begin, row_begin+nrows), whose entries live in
// the 4-entry-aligned window [nnz_base, nnz_base + 4*nwin4) of col/val.
struct SpmvBlock {
    uint64_t nnz_base;
    uint32_t row_begin;
    uint32_t roff_base;   // index (in uint16 units, multiple of 8) into the row-offset array
    uint16_t nrows;
    uint16_t nwin4;       // window length / 4 entries
    uint32_t flags;       // 1: long row (handled by the long-row kernel, only its dot part here)
    uint64_t pad;
};

struct ngsb_csr {
    ngsb_ctx *ctx = nullptr;
    size_t h = 0, w = 0, nnz = 0;
    int kind = 0;
    uint64_t *d_rowptr = nullptr;   // h+1
    int32_t *d_col = nullptr;       // nnz (+ slack)
    double *d_val = nullptr;        // nnz * matscalars (+ slack)
    // TMA-stream structures
    SpmvBlock *d_blocks = nullptr;
    uint16_t *d_rowoff = nullptr;
    uint32_t nblocks = 0;
    uint32_t *d_longrows = nullptr;
    uint32_t nlong = 0;
    int subwarp = 8;                // lanes per row
    int tile = 2048;                // entries per streamed stage
    int ncw = 8;                    // consumer warps per CTA
    int stages = 4;                 // ring depth
    int ctas_per_sm = 2;
    // SELL-32 copy (sell.cu): slices of 32 rows stored entry-major, default SpMV layout
    uint32_t nslices = 0;
    uint64_t sell_entries = 0;      // padded entries
    uint32_t sell_cap = 0;          // longest row part kept in a slice; the rest lives in the overflow CSR
    uint64_t *d_slice_off = nullptr;
    uint32_t *d_slice_src = nullptr;   // schedule: position t -> slice of the sigma-sorted row order
    uint32_t *d_row_of = nullptr;      // slot -> row (sigma-sorted), 0xffffffff = padding
    uint32_t *d_ovf_slot = nullptr;
    int32_t *d_scol = nullptr;
    double *d_sval = nullptr;
    uint16_t *d_scol16 = nullptr;      // 16-bit column offsets of the compressed slices (real matrices)
    int32_t *d_sbase = nullptr;        // per (slice, entry step) base column
    uint8_t *d_slice_c16 = nullptr;    // per scheduled slice: compressed?
    uint64_t sell_c16_entries = 0;     // padded entries living in compressed slices
    uint32_t novf = 0;
    uint32_t *d_ovf_rows = nullptr;
    uint64_t *d_ovf_ptr = nullptr;
    int32_t *d_slice_ovf = nullptr;
    int32_t *d_ovf_col = nullptr;
    double *d_ovf_val = nullptr;
    double *d_ovf_sum = nullptr;
    double mean_row = 0.0;
    size_t max_row = 0;
    ngsb_csr *transposed = nullptr;    // A^T, built by the first MultTransAdd (transpose.cu), owned
    // internal dof reordering (reorder.cu): products and fused solvers run on inner = P A P^T, which holds only its SELL copy
    ngsb_csr *inner = nullptr;         // owned; NULL: this matrix is multiplied as numbered by the caller
    uint32_t *d_perm = nullptr;        // new -> old: row i of inner = row d_perm[i] of this matrix
    uint32_t *d_iperm = nullptr;       // old -> new
    uint32_t *d_row_user = nullptr;    // (inner matrices) slot -> row in the CALLER's numbering, for y / the fused dot
    double *d_xperm = nullptr;         // (inner matrices) x gathered into the permuted numbering for one product
    bool row_identity = false;         // SELL rows sit in slot order (d_row_of is the identity): the kernels skip the table
    bool csr_released = false;         // d_col / d_val were freed (inner matrices)
    double natural_c16_share = -1.0;   // share of natural slices fit for 16-bit column offsets (automatic reorder criterion)
    uint64_t uid = 0;                  // unique per created matrix (key of cached CUDA graphs)
};

struct PeerHalo;
struct PeerReduce;

namespace ngsb {

// Product + neighbour exchange in ONE kernel (distributed CG, peer-memory data path): a warp that has finished a slice
// holding interface rows stores those rows' results straight into the neighbours' receive areas over NVLink, while the
// other warps keep multiplying; the kernel's last block publishes the sequence flags and this rank's partial of <s, A s>.
// Replaces the separate halo_push_kernel of ParallelBaseVector::Cumulate (parallel/parallelvvector.cpp:247-272: ISend of
// the interface values after the local MultAdd).
struct SellPush {
    const PeerHalo *H;
    const PeerReduce *R;
    const uint32_t *slice_if;     // per scheduled slice position: index of its interface record, 0xffffffff = no interface row
    const int32_t *lane_if;       // [records * 32]: interface dof index of the lane's row, -1 = interior row
    const uint32_t *if_first;     // per interface dof: its copies' positions if_pos[if_first[k] .. if_first[k+1]) in the packed
    const uint32_t *if_pos;       //   neighbour-major exchange list (the tables the unpack kernel adds by)
    int es;
};

// epilogue selector for the fused dot
enum SpmvEpi { EPI_NONE = 0, EPI_DOT_OUT = 1, EPI_CG_KSS = 2 };

struct SpmvArgs {
    const ngsb_csr *A;
    const double *x;
    double *y;
    double sr, si;          // scale
    bool accumulate;        // y += s*A*x  vs  y = s*A*x
    int epi;                // SpmvEpi
    const double *dotvec;   // vector dotted with the result rows (EPI_DOT_OUT / EPI_CG_KSS)
    int dot_conj;           // complex: conjugate the RESULT (argument) in the dot
    double *dot_out;        // EPI_DOT_OUT: device (re,im)
    CgState *state;         // EPI_CG_KSS (also: skip when state->done)
    // row subset for the distributed path (interior/boundary split); nblocks==0 -> all
    uint32_t block_begin, block_end;
    bool use_range;
    // SELL only: restrict the product to the scheduled slices slice_list[0 .. nlist) (ascending positions of the schedule).
    // The distributed CG multiplies the slices that hold interface rows first, pushes them to the neighbours and
    // multiplies the interior slices while the values travel (dist.cu).  dot_add: (re,im) added to the fused dot of this
    // launch (the partial of the other half); skip_overflow: the overflow sums were already computed by the first half.
    const uint32_t *slice_list;
    uint32_t nlist;
    const double *dot_add;
    bool skip_overflow;
    // (inner matrices) y and dotvec are in the caller's numbering: index them through d_row_user instead of d_row_of
    bool user_rows;
    // fused neighbour exchange (see SellPush; DEVICE pointer); needs epi != EPI_NONE (the last-block finish publishes the flags)
    const SellPush *push;
    const uint32_t *push_slice_src;   // the matrix' slice schedule with bit 31 set on the slices that hold interface rows (device)
};

int spmv_launch(const SpmvArgs &a);
int sell_build(ngsb_csr *A, const uint64_t *h_rowptr);
void sell_free(ngsb_csr *A);
int sell_launch(const SpmvArgs &a);
// y_k += alpha_k * A * x_k for four real vectors in one sweep over the matrix (MultiVector MultAdd)
// csrview.cu: the CSR column / value arrays on demand (option csr_keep)
void csr_release(ngsb_csr *A);
bool csr_release_wanted(const ngsb_csr *A);
int csr_ensure(const ngsb_csr *A);
int sell_extract_diag(const ngsb_csr *A, const uint8_t *d_bits, double *d_diag, int *d_status);
// reorder.cu
int csr_maybe_reorder(ngsb_csr *A, bool *made);
int launch_perm_gather(ngsb_ctx *ctx, const double *in, const uint32_t *perm, size_t n, int es, double *out);
// the same for entries of `es` doubles that lie `stride` doubles apart (in and out)
int launch_perm_gather_strided(ngsb_ctx *ctx, const double *in, const uint32_t *perm, size_t n, int es, int stride, double *out);
int launch_perm_bits(ngsb_ctx *ctx, const uint8_t *in, const uint32_t *perm, size_t n, uint8_t *out);
uint64_t next_uid();
int sell_launch_multi4(const ngsb_csr *A, const double *const x[4], double *const y[4], const double alpha[4]);

} // namespace ngsb
