// sell.cu -- sliced-ELLPACK (SELL-32) SpMV: the default kernel behind SparseMatrix<TM>::MultAdd.
//
// Why not vector-CSR: ncu on the sub-warp-per-row CSR kernel (profiles/r1_ncu_spmv_subwarp.txt)
// shows it bound by the L1TEX data pipe (l1tex__data_pipe_lsu_wavefronts 86 % of peak, DRAM only
// 64 %): with W lanes per row every warp-wide load touches G = 32/W different rows, so the value
// and column loads split into G wavefronts each and the x gather into one wavefront per distinct
// 128-byte line (about 13 for FE rows) -- 0.66 wavefronts per entry, 1 wavefront/clk/SM = 5.1 TB/s
// ceiling.  Staging the matrix stream through TMA does not change that count.
//
// SELL-32 removes the split: a slice is 32 consecutive rows, lane l owns row 32 s + l and walks it
// sequentially; the slice is stored entry-major ([j][lane]), so one warp-wide load of values
// (16 B per lane: packets of two entries) or columns is fully coalesced, and because
// neighbouring dofs of a finite-element numbering couple to neighbouring dofs, the 32 gathers of
// one step fall into few lines.  Each lane accumulates its row in storage order with one
// accumulator -- the reference's own summation order (RowTimesVector, linalg/sparsematrix.hpp:
// 625-632).  Rows are padded to the slice width with (val 0, col = last column of the row); rows
// longer than `cap` keep their first cap entries in the slice and the rest in a small overflow CSR
// that a second kernel reduces beforehand.
//
// Grid-stride kernel (about 96 CTAs per SM, 8 slices per CTA step, interleaved so that the resident CTAs sweep each
// stripe of the schedule as one moving window and x lines are shared through L1/L2); it fuses
// y = s A x (+ y), the dot <dotvec, result> and the CG scalar step (deterministic last-block finish).
#include "krylov.cuh"
#include "peer.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

namespace ngsb {

int device_scan_u64(ngsb_ctx *ctx, uint64_t *d_a, uint64_t n);   // vec.cu

struct SellParams {
    const uint64_t *slice_off;
    const uint32_t *slice_src;    // slice t of the schedule is slice slice_src[t] of the sigma-sorted row order
    const uint32_t *row_of;       // slot (= 32*slice + lane) -> matrix row, 0xffffffff for the padding of the last slice
    const int32_t *scol;
    const double *sval;
    const uint16_t *scol16;       // 16-bit column offsets (same indexing as scol) of the compressed slices, or NULL
    const int32_t *sbase;         // per (slice, entry step j): smallest column of the 32 lanes, index slice_off/32 + j
    const uint8_t *slice_c16;     // per scheduled slice: 1 = columns are sbase + scol16
    int pf_steps;                 // compressed slices: L2 prefetch distance inside a slice in steps of 4 packets (0 = off)
    int pf_next;                  // compressed slices: packets of the warp's NEXT slice prefetched into L2 at slice start (0 = off)
    const int32_t *slice_ovf;     // per slice: first overflow entry of this slice or -1 (NULL: none)
    const uint32_t *ovf_slot;     // slots of the rows with an overflow part, ascending
    const double *ovf_sum;        // raw overflow sums, es doubles per overflow row
    uint32_t novf;
    uint32_t nslices;
    uint64_t nrows;
    const double *x;
    double *y;
    double sr, si;
    int accumulate, epi, dot_conj;
    const double *dotvec;
    double *dot_out;
    CgState *state;
    double *partials;
    unsigned int *counter;
    const uint32_t *slice_list;   // LIST kernels: scheduled positions to process (ascending), nlist of them
    uint32_t nlist;
    const double *dot_add;        // (re,im) added to the fused dot by the finishing thread, or NULL
    const SellPush *push;         // PUSH kernels: interface rows go straight to the neighbours (descriptor in device memory)
};

__device__ __forceinline__ double2 ldg_stream_d2(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ldg_stream_i2(const int2 *p)
{
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_stream_d(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_stream_i(const int *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ unsigned short ldg_stream_u16(const unsigned short *p)
{
    unsigned short v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ unsigned int ldg_stream_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// L2 eviction policies: the matrix stream is touched once (evict first), the x entries are re-used by neighbouring slices
// (evict last) -- without hints the 57 GB stream pushes the few MB of live x lines out of the 126 MB L2
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <bool POL> __device__ __forceinline__ double2 ldp_d2(const double2 *p, uint64_t pol)
{
    double2 v;
    if (POL) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
template <bool POL> __device__ __forceinline__ unsigned int ldp_u32(const unsigned int *p, uint64_t pol)
{
    unsigned int v;
    if (POL) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <bool POL> __device__ __forceinline__ double ldp_x(const double *p, uint64_t pol)
{
    double v;
    if (POL) asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    else v = __ldg(p);
    return v;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ double warp_sum_s(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// interface row of a PUSH product: its result (already stored at y) goes to every sharer's receive area now (P2P stores over
// NVLink), at the place that rank's exchange table expects it -- what halo_push_kernel did in a launch of its own after the
// product.  Out of line: about 3 % of the slices come here, the streaming loops keep their 64 registers.
__device__ __noinline__ void sell_push_row(const SellPush *Pp, uint32_t s, int lane, const double *y, uint32_t row)
{
    const SellPush &P = *Pp;
    const uint32_t rec = P.slice_if[s];
    if (rec == 0xffffffffu) return;
    const int k = P.lane_if[(size_t)rec * 32 + lane];
    if (k < 0) return;
    const PeerHalo *H = P.H;
    const int es = P.es;
    const double *yy = y + (size_t)es * row;
    // sequence number / parity of the exchange this product feeds (the previous one was completed by its unpack kernel)
    const unsigned long long xseq = *(volatile const unsigned long long *)H->seq + 1;
    for (uint32_t q = P.if_first[k]; q < P.if_first[k + 1]; q++) {
        const uint32_t kk = P.if_pos[q];
        int pq = 0;
        while (kk >= H->peer_off[pq + 1]) pq++;
        double *dst = H->peer_recv[pq] + (xseq & 1) * H->peer_stride[pq] + (H->peer_my_off[pq] + (kk - H->peer_off[pq])) * (size_t)es;
        for (int c = 0; c < es; c++) dst[c] = yy[c];
    }
}

// VAR (tuning variants of the real-valued inner loop): 0 = 4 packets per step; 1 = 2 packets per step;
// 2 = 4 packets per step with the next step's values/columns prefetched before the gathers of this one
// LIST: walk p.slice_list instead of the whole schedule (interface-first split of the distributed product)
template <int KIND, int VAR, int MINB, bool POL = false, bool LIST = false, bool PUSH = false>
__global__ void __launch_bounds__(256, MINB) sell_spmv_kernel(const SellParams p)
{
    __shared__ double red[64];
    __shared__ int s_last;
    if (p.state != nullptr && p.state->done) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool pushed = false;

    double dr = 0.0, di = 0.0;
    const uint64_t nwalk = LIST ? (uint64_t)p.nlist : (uint64_t)p.nslices;
    for (uint64_t it = (uint64_t)blockIdx.x * 8 + wid; it < nwalk; it += (uint64_t)gridDim.x * 8) {
        const uint64_t s = LIST ? (uint64_t)p.slice_list[it] : it;
        const uint64_t off = p.slice_off[s];
        const uint32_t width = (uint32_t)((p.slice_off[s + 1] - off) >> 5);     // entries per lane
        // PUSH launches get a copy of the schedule whose top bit marks the slices holding interface rows: no extra load, no
        // extra register in the streaming loops
        const uint32_t src_raw = p.slice_src[s];
        const bool has_if = PUSH && (src_raw >> 31);
        const uint64_t src = PUSH ? (src_raw & 0x7fffffffu) : src_raw;
        const uint32_t slot = (uint32_t)(src * 32 + lane);
        // no table when the rows sit in slot order (natural slices; internally reordered matrices in the solvers' numbering)
        const uint32_t row = p.row_of != nullptr ? p.row_of[slot] : ((uint64_t)slot < p.nrows ? slot : 0xffffffffu);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (KIND == NGSB_REAL && p.slice_c16 != nullptr && p.slice_c16[s]) {
            // compressed slice: per packet 16 B of values and 4 B of column offsets per lane, plus one (uniform) pair of
            // 32-bit bases per warp -- 10.125 instead of 12 bytes per entry.  Same summation order as below.
            const double2 *v2 = reinterpret_cast<const double2 *>(p.sval + off) + lane;
            const unsigned int *h2 = reinterpret_cast<const unsigned int *>(p.scol16 + off) + lane;
            const int2 *b2 = reinterpret_cast<const int2 *>(p.sbase + (off >> 5));
            const uint32_t np = width >> 1;
            uint32_t q = 0;
            const uint64_t polf = POL ? l2_policy_evict_first() : 0, poll = POL ? l2_policy_evict_last() : 0;
            if (!LIST && p.pf_next > 0) {
                // the compressed stream carries fewer bytes per dependent step, so the loads of one step no longer cover the DRAM
                // latency: pull the head of this warp's next slice into L2 now (lane l fetches its own 128-byte line)
                const uint64_t sn = s + (uint64_t)gridDim.x * 8;
                if (sn < p.nslices) {
                    const uint64_t offn = p.slice_off[sn];
                    const uint32_t npn = min((uint32_t)((p.slice_off[sn + 1] - offn) >> 6), (uint32_t)p.pf_next);
                    const char *vn = reinterpret_cast<const char *>(p.sval + offn);
                    const char *hn = reinterpret_cast<const char *>(p.scol16 + offn);
                    for (uint32_t l = lane; l < npn * 4; l += 32) prefetch_l2(vn + (size_t)l * 128);      // 512 B of values per packet
                    if ((uint32_t)lane < npn) prefetch_l2(hn + (size_t)lane * 128);                        // 128 B of offsets per packet
                    if (lane == 0) prefetch_l2(p.sbase + (offn >> 5));
                }
            }
            const uint32_t pfd = (uint32_t)p.pf_steps * 4;
#define NGSB_C16(h, b, c0, c1) { c0 = (b).x + (int)((h) & 0xffffu); c1 = (b).y + (int)((h) >> 16); }
            if (VAR == 4 && np >= 4) {
                // "late values": only the index stream (4 + 8 registers) is prefetched one step ahead; the values of a step are
                // requested together with its x gathers.  Same bytes in flight per warp and step as the value-prefetching loop
                // below, 16 registers fewer -> 5 resident CTAs instead of 4.
                int ca0, ca1, cb0, cb1, cc0, cc1, cd0, cd1;
                {
                    const unsigned int ha = ldp_u32<POL>(h2, polf), hb = ldp_u32<POL>(h2 + 32, polf), hc = ldp_u32<POL>(h2 + 64, polf), hd = ldp_u32<POL>(h2 + 96, polf);
                    const int2 ba = __ldg(b2), bb = __ldg(b2 + 1), bc = __ldg(b2 + 2), bd = __ldg(b2 + 3);
                    NGSB_C16(ha, ba, ca0, ca1) NGSB_C16(hb, bb, cb0, cb1) NGSB_C16(hc, bc, cc0, cc1) NGSB_C16(hd, bd, cd0, cd1)
                }
                for (q = 0; q + 4 <= np; q += 4) {
                    const double x0 = ldp_x<POL>(p.x + ca0, poll), x1 = ldp_x<POL>(p.x + ca1, poll), x2 = ldp_x<POL>(p.x + cb0, poll), x3 = ldp_x<POL>(p.x + cb1, poll);
                    const double x4 = ldp_x<POL>(p.x + cc0, poll), x5 = ldp_x<POL>(p.x + cc1, poll), x6 = ldp_x<POL>(p.x + cd0, poll), x7 = ldp_x<POL>(p.x + cd1, poll);
                    const double2 va = ldp_d2<POL>(v2 + (q + 0) * 32, polf), vb = ldp_d2<POL>(v2 + (q + 1) * 32, polf);
                    const double2 vc = ldp_d2<POL>(v2 + (q + 2) * 32, polf), vd = ldp_d2<POL>(v2 + (q + 3) * 32, polf);
                    const bool more = q + 8 <= np;
                    unsigned int ha = 0, hb = 0, hc = 0, hd = 0;
                    int2 ba = make_int2(0, 0), bb = ba, bc = ba, bd = ba;
                    if (more) {
                        ha = ldp_u32<POL>(h2 + (q + 4) * 32, polf); hb = ldp_u32<POL>(h2 + (q + 5) * 32, polf);
                        hc = ldp_u32<POL>(h2 + (q + 6) * 32, polf); hd = ldp_u32<POL>(h2 + (q + 7) * 32, polf);
                        ba = __ldg(b2 + q + 4); bb = __ldg(b2 + q + 5); bc = __ldg(b2 + q + 6); bd = __ldg(b2 + q + 7);
                    }
                    s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
                    s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
                    if (more) { NGSB_C16(ha, ba, ca0, ca1) NGSB_C16(hb, bb, cb0, cb1) NGSB_C16(hc, bc, cc0, cc1) NGSB_C16(hd, bd, cd0, cd1) }
                }
            } else if (np >= 4) {
                double2 va = ldp_d2<POL>(v2, polf), vb = ldp_d2<POL>(v2 + 32, polf), vc = ldp_d2<POL>(v2 + 64, polf), vd = ldp_d2<POL>(v2 + 96, polf);
                int ca0, ca1, cb0, cb1, cc0, cc1, cd0, cd1;
                {
                    const unsigned int ha = ldp_u32<POL>(h2, polf), hb = ldp_u32<POL>(h2 + 32, polf), hc = ldp_u32<POL>(h2 + 64, polf), hd = ldp_u32<POL>(h2 + 96, polf);
                    const int2 ba = __ldg(b2), bb = __ldg(b2 + 1), bc = __ldg(b2 + 2), bd = __ldg(b2 + 3);
                    NGSB_C16(ha, ba, ca0, ca1) NGSB_C16(hb, bb, cb0, cb1) NGSB_C16(hc, bc, cc0, cc1) NGSB_C16(hd, bd, cd0, cd1)
                }
                for (q = 4; q + 4 <= np; q += 4) {
                    double x0 = ldp_x<POL>(p.x + ca0, poll), x1 = ldp_x<POL>(p.x + ca1, poll), x2 = ldp_x<POL>(p.x + cb0, poll), x3 = ldp_x<POL>(p.x + cb1, poll);
                    double x4 = ldp_x<POL>(p.x + cc0, poll), x5 = ldp_x<POL>(p.x + cc1, poll), x6 = ldp_x<POL>(p.x + cd0, poll), x7 = ldp_x<POL>(p.x + cd1, poll);
                    if (pfd && q + pfd + 4 <= np) {
                        // 2 KB of values = 16 lines (lanes 0-15), 512 B of offsets = 4 lines (lanes 16-19) of the step `pf_steps` ahead
                        if (lane < 16) prefetch_l2(reinterpret_cast<const char *>(p.sval + off) + (size_t)(q + pfd) * 512 + lane * 128);
                        else if (lane < 20) prefetch_l2(reinterpret_cast<const char *>(p.scol16 + off) + (size_t)(q + pfd) * 128 + (lane - 16) * 128);
                    }
                    double2 na = ldp_d2<POL>(v2 + (q + 0) * 32, polf), nb = ldp_d2<POL>(v2 + (q + 1) * 32, polf);
                    double2 nc = ldp_d2<POL>(v2 + (q + 2) * 32, polf), nd = ldp_d2<POL>(v2 + (q + 3) * 32, polf);
                    const unsigned int ha = ldp_u32<POL>(h2 + (q + 0) * 32, polf), hb = ldp_u32<POL>(h2 + (q + 1) * 32, polf);
                    const unsigned int hc = ldp_u32<POL>(h2 + (q + 2) * 32, polf), hd = ldp_u32<POL>(h2 + (q + 3) * 32, polf);
                    const int2 ba = __ldg(b2 + q), bb = __ldg(b2 + q + 1), bc = __ldg(b2 + q + 2), bd = __ldg(b2 + q + 3);
                    s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
                    s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
                    va = na; vb = nb; vc = nc; vd = nd;
                    NGSB_C16(ha, ba, ca0, ca1) NGSB_C16(hb, bb, cb0, cb1) NGSB_C16(hc, bc, cc0, cc1) NGSB_C16(hd, bd, cd0, cd1)
                }
                double x0 = ldp_x<POL>(p.x + ca0, poll), x1 = ldp_x<POL>(p.x + ca1, poll), x2 = ldp_x<POL>(p.x + cb0, poll), x3 = ldp_x<POL>(p.x + cb1, poll);
                double x4 = ldp_x<POL>(p.x + cc0, poll), x5 = ldp_x<POL>(p.x + cc1, poll), x6 = ldp_x<POL>(p.x + cd0, poll), x7 = ldp_x<POL>(p.x + cd1, poll);
                s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
                s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
            }
            for (; q < np; q++) {
                const double2 va = ldp_d2<POL>(v2 + q * 32, polf);
                const unsigned int ha = ldp_u32<POL>(h2 + q * 32, polf);
                const int2 ba = __ldg(b2 + q);
                int c0, c1;
                NGSB_C16(ha, ba, c0, c1)
                s0 = fma(va.x, ldp_x<POL>(p.x + c0, poll), s0);
                s0 = fma(va.y, ldp_x<POL>(p.x + c1, poll), s0);
            }
#undef NGSB_C16
        } else if (KIND == NGSB_REAL) {
            const double2 *v2 = reinterpret_cast<const double2 *>(p.sval + off) + lane;
            const int2 *c2 = reinterpret_cast<const int2 *>(p.scol + off) + lane;
            const uint32_t np = width >> 1;
            uint32_t q = 0;
            if (VAR == 2 && np >= 4) {
                double2 va = ldg_stream_d2(v2), vb = ldg_stream_d2(v2 + 32), vc = ldg_stream_d2(v2 + 64), vd = ldg_stream_d2(v2 + 96);
                int2 ca = ldg_stream_i2(c2), cb = ldg_stream_i2(c2 + 32), cc = ldg_stream_i2(c2 + 64), cd = ldg_stream_i2(c2 + 96);
                for (q = 4; q + 4 <= np; q += 4) {
                    double x0 = __ldg(p.x + ca.x), x1 = __ldg(p.x + ca.y), x2 = __ldg(p.x + cb.x), x3 = __ldg(p.x + cb.y);
                    double x4 = __ldg(p.x + cc.x), x5 = __ldg(p.x + cc.y), x6 = __ldg(p.x + cd.x), x7 = __ldg(p.x + cd.y);
                    double2 na = ldg_stream_d2(v2 + (q + 0) * 32), nb = ldg_stream_d2(v2 + (q + 1) * 32);
                    double2 nc = ldg_stream_d2(v2 + (q + 2) * 32), nd = ldg_stream_d2(v2 + (q + 3) * 32);
                    int2 ea = ldg_stream_i2(c2 + (q + 0) * 32), eb = ldg_stream_i2(c2 + (q + 1) * 32);
                    int2 ec = ldg_stream_i2(c2 + (q + 2) * 32), ed = ldg_stream_i2(c2 + (q + 3) * 32);
                    s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
                    s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
                    va = na; vb = nb; vc = nc; vd = nd; ca = ea; cb = eb; cc = ec; cd = ed;
                }
                double x0 = __ldg(p.x + ca.x), x1 = __ldg(p.x + ca.y), x2 = __ldg(p.x + cb.x), x3 = __ldg(p.x + cb.y);
                double x4 = __ldg(p.x + cc.x), x5 = __ldg(p.x + cc.y), x6 = __ldg(p.x + cd.x), x7 = __ldg(p.x + cd.y);
                s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
                s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
            }
            if (VAR == 1)
                for (; q + 2 <= np; q += 2) {
                    double2 va = ldg_stream_d2(v2 + (q + 0) * 32), vb = ldg_stream_d2(v2 + (q + 1) * 32);
                    int2 ca = ldg_stream_i2(c2 + (q + 0) * 32), cb = ldg_stream_i2(c2 + (q + 1) * 32);
                    double x0 = __ldg(p.x + ca.x), x1 = __ldg(p.x + ca.y), x2 = __ldg(p.x + cb.x), x3 = __ldg(p.x + cb.y);
                    s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
                }
            if (VAR == 0 || VAR == 4)
            for (; q + 4 <= np; q += 4) {
                double2 va = ldg_stream_d2(v2 + (q + 0) * 32), vb = ldg_stream_d2(v2 + (q + 1) * 32);
                double2 vc = ldg_stream_d2(v2 + (q + 2) * 32), vd = ldg_stream_d2(v2 + (q + 3) * 32);
                int2 ca = ldg_stream_i2(c2 + (q + 0) * 32), cb = ldg_stream_i2(c2 + (q + 1) * 32);
                int2 cc = ldg_stream_i2(c2 + (q + 2) * 32), cd = ldg_stream_i2(c2 + (q + 3) * 32);
                double x0 = __ldg(p.x + ca.x), x1 = __ldg(p.x + ca.y), x2 = __ldg(p.x + cb.x), x3 = __ldg(p.x + cb.y);
                double x4 = __ldg(p.x + cc.x), x5 = __ldg(p.x + cc.y), x6 = __ldg(p.x + cd.x), x7 = __ldg(p.x + cd.y);
                s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
                s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
            }
            for (; q < np; q++) {
                double2 va = ldg_stream_d2(v2 + q * 32);
                int2 ca = ldg_stream_i2(c2 + q * 32);
                s0 = fma(va.x, __ldg(p.x + ca.x), s0);
                s0 = fma(va.y, __ldg(p.x + ca.y), s0);
            }
        } else if (KIND == NGSB_COMPLEX && p.slice_c16 != nullptr && p.slice_c16[s]) {
            // compressed complex slice: 16 B value + 2 B column offset per lane and step, one 32-bit base per warp and step
            // (18.125 instead of 20 bytes per entry); same loop structure and summation order as the 32-bit branch below
            const double2 *v2 = reinterpret_cast<const double2 *>(p.sval) + off + lane;
            const unsigned short *h1 = p.scol16 + off + lane;
            const int *b1 = p.sbase + (off >> 5);
            const double2 *x2 = reinterpret_cast<const double2 *>(p.x);
            uint32_t q = 0;
            if (width >= 4) {
                double2 va = ldg_stream_d2(v2), vb = ldg_stream_d2(v2 + 32), vc = ldg_stream_d2(v2 + 64), vd = ldg_stream_d2(v2 + 96);
                int ca = __ldg(b1) + (int)ldg_stream_u16(h1), cb = __ldg(b1 + 1) + (int)ldg_stream_u16(h1 + 32);
                int cc = __ldg(b1 + 2) + (int)ldg_stream_u16(h1 + 64), cd = __ldg(b1 + 3) + (int)ldg_stream_u16(h1 + 96);
                for (q = 4; q + 4 <= width; q += 4) {
                    double2 xa = __ldg(x2 + ca), xb = __ldg(x2 + cb), xc = __ldg(x2 + cc), xd = __ldg(x2 + cd);
                    double2 na = ldg_stream_d2(v2 + (q + 0) * 32), nb = ldg_stream_d2(v2 + (q + 1) * 32);
                    double2 nc = ldg_stream_d2(v2 + (q + 2) * 32), nd = ldg_stream_d2(v2 + (q + 3) * 32);
                    const unsigned short ha = ldg_stream_u16(h1 + (q + 0) * 32), hb = ldg_stream_u16(h1 + (q + 1) * 32);
                    const unsigned short hc = ldg_stream_u16(h1 + (q + 2) * 32), hd = ldg_stream_u16(h1 + (q + 3) * 32);
                    const int ba = __ldg(b1 + q), bb = __ldg(b1 + q + 1), bc = __ldg(b1 + q + 2), bd = __ldg(b1 + q + 3);
                    s0 += va.x * xa.x - va.y * xa.y; s1 += va.x * xa.y + va.y * xa.x;
                    s0 += vb.x * xb.x - vb.y * xb.y; s1 += vb.x * xb.y + vb.y * xb.x;
                    s0 += vc.x * xc.x - vc.y * xc.y; s1 += vc.x * xc.y + vc.y * xc.x;
                    s0 += vd.x * xd.x - vd.y * xd.y; s1 += vd.x * xd.y + vd.y * xd.x;
                    va = na; vb = nb; vc = nc; vd = nd;
                    ca = ba + (int)ha; cb = bb + (int)hb; cc = bc + (int)hc; cd = bd + (int)hd;
                }
                double2 xa = __ldg(x2 + ca), xb = __ldg(x2 + cb), xc = __ldg(x2 + cc), xd = __ldg(x2 + cd);
                s0 += va.x * xa.x - va.y * xa.y; s1 += va.x * xa.y + va.y * xa.x;
                s0 += vb.x * xb.x - vb.y * xb.y; s1 += vb.x * xb.y + vb.y * xb.x;
                s0 += vc.x * xc.x - vc.y * xc.y; s1 += vc.x * xc.y + vc.y * xc.x;
                s0 += vd.x * xd.x - vd.y * xd.y; s1 += vd.x * xd.y + vd.y * xd.x;
            }
            for (; q < width; q++) {
                double2 va = ldg_stream_d2(v2 + q * 32);
                double2 xa = __ldg(x2 + (__ldg(b1 + q) + (int)ldg_stream_u16(h1 + q * 32)));
                s0 += va.x * xa.x - va.y * xa.y;
                s1 += va.x * xa.y + va.y * xa.x;
            }
        } else if (KIND == NGSB_COMPLEX) {
            const double2 *v2 = reinterpret_cast<const double2 *>(p.sval) + off + lane;
            const int *c1 = p.scol + off + lane;
            const double2 *x2 = reinterpret_cast<const double2 *>(p.x);
            uint32_t q = 0;
            if (VAR == 2 && width >= 4) {
                // software pipeline: the next four entries are in flight while this step's x gathers resolve
                double2 va = ldg_stream_d2(v2), vb = ldg_stream_d2(v2 + 32), vc = ldg_stream_d2(v2 + 64), vd = ldg_stream_d2(v2 + 96);
                int ca = ldg_stream_i(c1), cb = ldg_stream_i(c1 + 32), cc = ldg_stream_i(c1 + 64), cd = ldg_stream_i(c1 + 96);
                for (q = 4; q + 4 <= width; q += 4) {
                    double2 xa = __ldg(x2 + ca), xb = __ldg(x2 + cb), xc = __ldg(x2 + cc), xd = __ldg(x2 + cd);
                    double2 na = ldg_stream_d2(v2 + (q + 0) * 32), nb = ldg_stream_d2(v2 + (q + 1) * 32);
                    double2 nc = ldg_stream_d2(v2 + (q + 2) * 32), nd = ldg_stream_d2(v2 + (q + 3) * 32);
                    int ea = ldg_stream_i(c1 + (q + 0) * 32), eb = ldg_stream_i(c1 + (q + 1) * 32);
                    int ec = ldg_stream_i(c1 + (q + 2) * 32), ed = ldg_stream_i(c1 + (q + 3) * 32);
                    s0 += va.x * xa.x - va.y * xa.y; s1 += va.x * xa.y + va.y * xa.x;
                    s0 += vb.x * xb.x - vb.y * xb.y; s1 += vb.x * xb.y + vb.y * xb.x;
                    s0 += vc.x * xc.x - vc.y * xc.y; s1 += vc.x * xc.y + vc.y * xc.x;
                    s0 += vd.x * xd.x - vd.y * xd.y; s1 += vd.x * xd.y + vd.y * xd.x;
                    va = na; vb = nb; vc = nc; vd = nd; ca = ea; cb = eb; cc = ec; cd = ed;
                }
                double2 xa = __ldg(x2 + ca), xb = __ldg(x2 + cb), xc = __ldg(x2 + cc), xd = __ldg(x2 + cd);
                s0 += va.x * xa.x - va.y * xa.y; s1 += va.x * xa.y + va.y * xa.x;
                s0 += vb.x * xb.x - vb.y * xb.y; s1 += vb.x * xb.y + vb.y * xb.x;
                s0 += vc.x * xc.x - vc.y * xc.y; s1 += vc.x * xc.y + vc.y * xc.x;
                s0 += vd.x * xd.x - vd.y * xd.y; s1 += vd.x * xd.y + vd.y * xd.x;
            }
            if (VAR != 2)
            for (; q + 2 <= width; q += 2) {
                double2 va = ldg_stream_d2(v2 + (q + 0) * 32), vb = ldg_stream_d2(v2 + (q + 1) * 32);
                int ca = ldg_stream_i(c1 + (q + 0) * 32), cb = ldg_stream_i(c1 + (q + 1) * 32);
                double2 xa = __ldg(x2 + ca), xb = __ldg(x2 + cb);
                s0 += va.x * xa.x - va.y * xa.y;
                s1 += va.x * xa.y + va.y * xa.x;
                s0 += vb.x * xb.x - vb.y * xb.y;
                s1 += vb.x * xb.y + vb.y * xb.x;
            }
            for (; q < width; q++) {
                double2 va = ldg_stream_d2(v2 + q * 32);
                double2 xa = __ldg(x2 + ldg_stream_i(c1 + q * 32));
                s0 += va.x * xa.x - va.y * xa.y;
                s1 += va.x * xa.y + va.y * xa.x;
            }
        } else {
            // 3x3 blocks: nine component planes per entry, [j][k][lane]
            const double *v = p.sval + off * 9 + lane;
#define NGSB_B3_ENTRY(col)                                                                                       \
            {                                                                                                    \
                const double *m = v + (size_t)q * 9 * 32;                                                        \
                const double *xv = p.x + 3 * (size_t)(col);                                                      \
                double m0 = ldg_stream_d(m), m1 = ldg_stream_d(m + 32), m2 = ldg_stream_d(m + 64);               \
                double m3 = ldg_stream_d(m + 96), m4 = ldg_stream_d(m + 128), m5 = ldg_stream_d(m + 160);        \
                double m6 = ldg_stream_d(m + 192), m7 = ldg_stream_d(m + 224), m8 = ldg_stream_d(m + 256);       \
                double x0 = __ldg(xv), x1 = __ldg(xv + 1), x2 = __ldg(xv + 2);                                   \
                s0 += m0 * x0 + m1 * x1 + m2 * x2;                                                               \
                s1 += m3 * x0 + m4 * x1 + m5 * x2;                                                               \
                s2 += m6 * x0 + m7 * x1 + m8 * x2;                                                               \
            }
            if (p.slice_c16 != nullptr && p.slice_c16[s]) {
                // compressed slice (option sell_c16_all): 74.1 instead of 76 bytes per entry
                const unsigned short *h1 = p.scol16 + off + lane;
                const int *b1 = p.sbase + (off >> 5);
                for (uint32_t q = 0; q < width; q++) NGSB_B3_ENTRY(__ldg(b1 + q) + (int)ldg_stream_u16(h1 + q * 32))
            } else {
                const int *c1 = p.scol + off + lane;
                for (uint32_t q = 0; q < width; q++) NGSB_B3_ENTRY(ldg_stream_i(c1 + q * 32))
            }
#undef NGSB_B3_ENTRY
        }
        if (p.slice_ovf != nullptr) {
            int k = p.slice_ovf[src];
            if (k >= 0)
                for (; (uint32_t)k < p.novf && (p.ovf_slot[k] >> 5) == src; k++)
                    if (p.ovf_slot[k] == slot) {
                        if (KIND == NGSB_REAL) s0 += p.ovf_sum[k];
                        else if (KIND == NGSB_COMPLEX) { s0 += p.ovf_sum[2 * k]; s1 += p.ovf_sum[2 * k + 1]; }
                        else { s0 += p.ovf_sum[3 * k]; s1 += p.ovf_sum[3 * k + 1]; s2 += p.ovf_sum[3 * k + 2]; }
                    }
        }
        if (row != 0xffffffffu) {
            // y = s*sum (+ y) and this row's share of the fused dot
            if (KIND == NGSB_REAL) {
                double r = p.sr * s0;
                if (p.accumulate) r += p.y[row];
                p.y[row] = r;
                if (p.epi) dr = fma(p.dotvec[row], r, dr);
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
                double *yy = p.y + 3 * (size_t)row;
                if (p.accumulate) { r0 += yy[0]; r1 += yy[1]; r2 += yy[2]; }
                yy[0] = r0; yy[1] = r1; yy[2] = r2;
                if (p.epi) {
                    const double *v = p.dotvec + 3 * (size_t)row;
                    dr += v[0] * r0 + v[1] * r1 + v[2] * r2;
                }
            }
            if (PUSH && has_if) { sell_push_row(p.push, (uint32_t)s, lane, p.y, row); pushed = true; }
        }
    }
    if (!p.epi) return;
    // this thread's peer stores are visible system-wide before the block takes its ticket.  Only the threads that stored pay
    // for the system-scope fence (a fence by every thread of the grid cost 4 % of the product)
    if (PUSH && pushed) __threadfence_system();
    // deterministic grid-wide finish of the fused dot (+ scalar step of CG)
    dr = warp_sum_s(dr);
    di = warp_sum_s(di);
    if (lane == 0) { red[wid] = dr; red[32 + wid] = di; }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        dr = lane < nw ? red[lane] : 0.0;
        di = lane < nw ? red[32 + lane] : 0.0;
        dr = warp_sum_s(dr);
        di = warp_sum_s(di);
        if (lane == 0) {
            p.partials[2 * blockIdx.x] = dr;
            p.partials[2 * blockIdx.x + 1] = di;
            __threadfence();
            s_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last) return;
    // the last block sums the per-CTA partials in a fixed order (thread t: partials t, t+256, ...; then lanes, then warps)
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += blockDim.x) {
        a += __ldcg(&p.partials[2 * k]);
        b += __ldcg(&p.partials[2 * k + 1]);
    }
    a = warp_sum_s(a);
    b = warp_sum_s(b);
    __syncthreads();                       // red[] is free again
    if (lane == 0) { red[wid] = a; red[32 + wid] = b; }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    {
        const int nw = blockDim.x >> 5;
        a = lane < nw ? red[lane] : 0.0;
        b = lane < nw ? red[32 + lane] : 0.0;
        a = warp_sum_s(a);
        b = warp_sum_s(b);
    }
    if (threadIdx.x == 0) {
        *p.counter = 0;
        if (p.dot_add != nullptr) { a += p.dot_add[0]; b += p.dot_add[1]; }
        if (p.epi == EPI_DOT_OUT) { p.dot_out[0] = a; p.dot_out[1] = b; }
        else if (p.epi == EPI_CG_KSS) cg_finalize_kss(p.state, make_double2(a, b));
    }
    if (PUSH) {
        // every block's stores are out (each fenced before its ticket): tell the neighbours, and send this rank's partial of
        // <s, A s> on its way to all ranks (the unpack kernel sums the partials and does al = wd / kss).  One lane per peer.
        const PeerHalo *H = p.push->H;
        const unsigned long long xseq = *(volatile const unsigned long long *)H->seq + 1;
        __threadfence_system();
        if ((int)threadIdx.x < H->npeers) st_release_sys(H->peer_flags[threadIdx.x] + (xseq & 1) * NGSB_MAX_RANKS + H->rank, xseq);
        if (p.push->R) pr_push_warp(*p.push->R, a, b);
    }
}

// ------------------------------------------------------------------------------------------
// The whole Jacobi-PCG loop of a small real system as ONE persistent cooperative kernel.
//
// Below a few million rows an iteration is three short kernels (1 M dofs: about 95 + 12 + 6 us of work) and the three kernel
// boundaries -- drain, launch, ramp-up against an L2 full of the previous kernel's dirty lines -- cost as much as a third
// of it (167 us per iteration measured on BASELINE configs[0]).  Here the grid stays resident (one launch per solve or per
// batch), the phases of CGSolver<double>::Mult (linalg/cg.cpp:593-620) are separated by grid barriers, and the scalars
// kss, al, wdn, be and the loop condition are computed redundantly but identically by every block from per-block partials
// summed in a fixed order -- no block waits for another one's scalar step.
//   phase A  as = A s (SELL slices, same loops and summation order as sell_spmv_kernel), partial <s, as>      | barrier
//   phase B  al = wd / kss;  u += al s;  d -= al as;  w = C d;  partial <d, w>                                 | barrier
//   phase C  be = wdn / wd;  history;  loop condition;  s = be s + w                                           | barrier
// ------------------------------------------------------------------------------------------
struct CgPersistParams {
    SellParams sp;
    CgVecs v;
    double *part_a, *part_b;        // per-block partials of the two dots
    unsigned int *bar_count;
    unsigned int *bar_gen;
    int iters;                      // iterations this launch may run
};

__device__ __forceinline__ void grid_barrier(unsigned int *count, unsigned int *gen, unsigned int nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int g = *(volatile unsigned int *)gen;
        __threadfence();
        if (atomicAdd(count, 1u) == nblocks - 1) {
            *(volatile unsigned int *)count = 0;
            __threadfence();
            atomicAdd(gen, 1u);
        } else {
            while (*(volatile unsigned int *)gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// sum of the per-block partials in a fixed order, by every block for itself (result in all threads)
__device__ __forceinline__ double block_sum_partials(const double *part, unsigned int n, double *red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double a = 0.0;
    for (unsigned int k = threadIdx.x; k < n; k += blockDim.x) a += __ldcg(part + k);
    a = warp_sum_s(a);
    __syncthreads();
    if (lane == 0) red[wid] = a;
    __syncthreads();
    double t = 0.0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; w++) t += red[w];
    return t;
}

// one row of a real SELL slice (compressed or 32-bit columns): the software-pipelined default loops of sell_spmv_kernel
__device__ __forceinline__ double sell_row_real(const SellParams &p, uint64_t s, uint64_t off, uint32_t width, int lane)
{
    double s0 = 0.0;
    const double2 *v2 = reinterpret_cast<const double2 *>(p.sval + off) + lane;
    const uint32_t np = width >> 1;
    uint32_t q = 0;
    if (p.slice_c16 != nullptr && p.slice_c16[s]) {
        const unsigned int *h2 = reinterpret_cast<const unsigned int *>(p.scol16 + off) + lane;
        const int2 *b2 = reinterpret_cast<const int2 *>(p.sbase + (off >> 5));
#define NGSB_C16(h, b, c0, c1) { c0 = (b).x + (int)((h) & 0xffffu); c1 = (b).y + (int)((h) >> 16); }
        if (np >= 4) {
            double2 va = ldg_stream_d2(v2), vb = ldg_stream_d2(v2 + 32), vc = ldg_stream_d2(v2 + 64), vd = ldg_stream_d2(v2 + 96);
            int ca0, ca1, cb0, cb1, cc0, cc1, cd0, cd1;
            {
                const unsigned int ha = ldg_stream_u32(h2), hb = ldg_stream_u32(h2 + 32), hc = ldg_stream_u32(h2 + 64), hd = ldg_stream_u32(h2 + 96);
                const int2 ba = __ldg(b2), bb = __ldg(b2 + 1), bc = __ldg(b2 + 2), bd = __ldg(b2 + 3);
                NGSB_C16(ha, ba, ca0, ca1) NGSB_C16(hb, bb, cb0, cb1) NGSB_C16(hc, bc, cc0, cc1) NGSB_C16(hd, bd, cd0, cd1)
            }
            for (q = 4; q + 4 <= np; q += 4) {
                double x0 = __ldg(p.x + ca0), x1 = __ldg(p.x + ca1), x2 = __ldg(p.x + cb0), x3 = __ldg(p.x + cb1);
                double x4 = __ldg(p.x + cc0), x5 = __ldg(p.x + cc1), x6 = __ldg(p.x + cd0), x7 = __ldg(p.x + cd1);
                double2 na = ldg_stream_d2(v2 + (q + 0) * 32), nb = ldg_stream_d2(v2 + (q + 1) * 32);
                double2 nc = ldg_stream_d2(v2 + (q + 2) * 32), nd = ldg_stream_d2(v2 + (q + 3) * 32);
                const unsigned int ha = ldg_stream_u32(h2 + (q + 0) * 32), hb = ldg_stream_u32(h2 + (q + 1) * 32);
                const unsigned int hc = ldg_stream_u32(h2 + (q + 2) * 32), hd = ldg_stream_u32(h2 + (q + 3) * 32);
                const int2 ba = __ldg(b2 + q), bb = __ldg(b2 + q + 1), bc = __ldg(b2 + q + 2), bd = __ldg(b2 + q + 3);
                s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
                s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
                va = na; vb = nb; vc = nc; vd = nd;
                NGSB_C16(ha, ba, ca0, ca1) NGSB_C16(hb, bb, cb0, cb1) NGSB_C16(hc, bc, cc0, cc1) NGSB_C16(hd, bd, cd0, cd1)
            }
            double x0 = __ldg(p.x + ca0), x1 = __ldg(p.x + ca1), x2 = __ldg(p.x + cb0), x3 = __ldg(p.x + cb1);
            double x4 = __ldg(p.x + cc0), x5 = __ldg(p.x + cc1), x6 = __ldg(p.x + cd0), x7 = __ldg(p.x + cd1);
            s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
            s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
        }
        for (; q < np; q++) {
            const double2 va = ldg_stream_d2(v2 + q * 32);
            const unsigned int ha = ldg_stream_u32(h2 + q * 32);
            const int2 ba = __ldg(b2 + q);
            int c0, c1;
            NGSB_C16(ha, ba, c0, c1)
            s0 = fma(va.x, __ldg(p.x + c0), s0);
            s0 = fma(va.y, __ldg(p.x + c1), s0);
        }
#undef NGSB_C16
        return s0;
    }
    const int2 *c2 = reinterpret_cast<const int2 *>(p.scol + off) + lane;
    if (np >= 4) {
        double2 va = ldg_stream_d2(v2), vb = ldg_stream_d2(v2 + 32), vc = ldg_stream_d2(v2 + 64), vd = ldg_stream_d2(v2 + 96);
        int2 ca = ldg_stream_i2(c2), cb = ldg_stream_i2(c2 + 32), cc = ldg_stream_i2(c2 + 64), cd = ldg_stream_i2(c2 + 96);
        for (q = 4; q + 4 <= np; q += 4) {
            double x0 = __ldg(p.x + ca.x), x1 = __ldg(p.x + ca.y), x2 = __ldg(p.x + cb.x), x3 = __ldg(p.x + cb.y);
            double x4 = __ldg(p.x + cc.x), x5 = __ldg(p.x + cc.y), x6 = __ldg(p.x + cd.x), x7 = __ldg(p.x + cd.y);
            double2 na = ldg_stream_d2(v2 + (q + 0) * 32), nb = ldg_stream_d2(v2 + (q + 1) * 32);
            double2 nc = ldg_stream_d2(v2 + (q + 2) * 32), nd = ldg_stream_d2(v2 + (q + 3) * 32);
            int2 ea = ldg_stream_i2(c2 + (q + 0) * 32), eb = ldg_stream_i2(c2 + (q + 1) * 32);
            int2 ec = ldg_stream_i2(c2 + (q + 2) * 32), ed = ldg_stream_i2(c2 + (q + 3) * 32);
            s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
            s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
            va = na; vb = nb; vc = nc; vd = nd; ca = ea; cb = eb; cc = ec; cd = ed;
        }
        double x0 = __ldg(p.x + ca.x), x1 = __ldg(p.x + ca.y), x2 = __ldg(p.x + cb.x), x3 = __ldg(p.x + cb.y);
        double x4 = __ldg(p.x + cc.x), x5 = __ldg(p.x + cc.y), x6 = __ldg(p.x + cd.x), x7 = __ldg(p.x + cd.y);
        s0 = fma(va.x, x0, s0); s0 = fma(va.y, x1, s0); s0 = fma(vb.x, x2, s0); s0 = fma(vb.y, x3, s0);
        s0 = fma(vc.x, x4, s0); s0 = fma(vc.y, x5, s0); s0 = fma(vd.x, x6, s0); s0 = fma(vd.y, x7, s0);
    }
    for (; q < np; q++) {
        double2 va = ldg_stream_d2(v2 + q * 32);
        int2 ca = ldg_stream_i2(c2 + q * 32);
        s0 = fma(va.x, __ldg(p.x + ca.x), s0);
        s0 = fma(va.y, __ldg(p.x + ca.y), s0);
    }
    return s0;
}

__global__ void __launch_bounds__(256, 4) cg_persistent_kernel(const CgPersistParams P)
{
    __shared__ double red[8];
    const SellParams &p = P.sp;
    const CgVecs &v = P.v;
    CgState *st = v.state;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int nb = gridDim.x;
    // every block carries the scalar state of the loop itself -- in shared memory, so that the streaming loops of phase A
    // keep their registers (all threads compute the same values; thread 0 stores them)
    __shared__ double sh_wdn;
    __shared__ int sh_n, sh_nhist, sh_done;
    if (threadIdx.x == 0) { sh_wdn = st->wdn[0]; sh_n = st->n; sh_nhist = st->nhist; sh_done = st->done; }
    __syncthreads();
    for (int it = 0; it < P.iters && !sh_done; it++) {
        // ---- phase A: as = A s, partial <s, as>
        double dr = 0.0;
        for (uint64_t s = (uint64_t)blockIdx.x * 8 + wid; s < p.nslices; s += (uint64_t)nb * 8) {
            const uint64_t off = p.slice_off[s];
            const uint32_t width = (uint32_t)((p.slice_off[s + 1] - off) >> 5);
            const uint32_t slot = (uint32_t)((uint64_t)p.slice_src[s] * 32 + lane);
            const uint32_t row = p.row_of != nullptr ? p.row_of[slot] : ((uint64_t)slot < p.nrows ? slot : 0xffffffffu);
            const double r = sell_row_real(p, s, off, width, lane);
            if (row != 0xffffffffu) {
                p.y[row] = r;
                dr = fma(p.x[row], r, dr);
            }
        }
        dr = warp_sum_s(dr);
        __syncthreads();
        if (lane == 0) red[wid] = dr;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += red[w];
            P.part_a[blockIdx.x] = t;
        }
        grid_barrier(P.bar_count, P.bar_gen, nb);
        const double kss = block_sum_partials(P.part_a, nb, red);
        const double wd = sh_wdn;
        if (kss == 0.0) {                                         // `if (kss == 0.0) break;` (uniform over the grid)
            __syncthreads();
            if (threadIdx.x == 0) sh_done = 1;
            __syncthreads();
            break;
        }
        const double al = wd / kss;
        const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (uint64_t)nb * blockDim.x;
        const uint64_t n2 = v.n & ~(uint64_t)1;
        // ---- phase B: u += al s, d -= al as, w = C d, partial <d, w>
        double acc0 = 0.0, acc1 = 0.0;
        for (uint64_t i = 2 * first; i < n2; i += 2 * step) {
            const double2 a2 = *reinterpret_cast<const double2 *>(v.as + i);
            const double2 m2 = *reinterpret_cast<const double2 *>(v.invdiag + i);
            const double2 s2 = *reinterpret_cast<const double2 *>(v.s + i);
            double2 d2 = *reinterpret_cast<double2 *>(v.d + i);
            double2 u2 = *reinterpret_cast<double2 *>(v.u + i);
            u2.x += al * s2.x; u2.y += al * s2.y;
            d2.x -= al * a2.x; d2.y -= al * a2.y;
            const unsigned bits = v.bits ? (unsigned)(v.bits[i >> 3] >> (i & 7)) : 3u;
            double2 w2;
            w2.x = (bits & 1u) ? m2.x * d2.x : 0.0;
            w2.y = (bits & 2u) ? m2.y * d2.y : 0.0;
            *reinterpret_cast<double2 *>(v.u + i) = u2;
            *reinterpret_cast<double2 *>(v.d + i) = d2;
            *reinterpret_cast<double2 *>(v.w + i) = w2;
            acc0 = fma(d2.x, w2.x, acc0);
            acc1 = fma(d2.y, w2.y, acc1);
        }
        if (n2 < v.n && first == 0) {
            const uint64_t i = n2;
            v.u[i] += al * v.s[i];
            const double dn = v.d[i] - al * v.as[i];
            const double wn = (v.bits == nullptr || ((v.bits[i >> 3] >> (i & 7)) & 1)) ? v.invdiag[i] * dn : 0.0;
            v.d[i] = dn;
            v.w[i] = wn;
            acc0 = fma(dn, wn, acc0);
        }
        acc0 += acc1;
        acc0 = warp_sum_s(acc0);
        __syncthreads();
        if (lane == 0) red[wid] = acc0;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += red[w];
            P.part_b[blockIdx.x] = t;
        }
        grid_barrier(P.bar_count, P.bar_gen, nb);
        const double wdn = block_sum_partials(P.part_b, nb, red);
        const double be = wdn / wd;
        if (threadIdx.x == 0) {
            if (blockIdx.x == 0 && sh_nhist < st->hist_cap) v.hist[sh_nhist] = fabs(wdn);
            sh_nhist++;
            sh_wdn = wdn;
            // `while (n++ < maxsteps && Abs(wdn) > err)` of the next pass
            sh_done = !((sh_n++ < st->maxsteps) && (fabs(wdn) > st->err));
        }
        // ---- phase C: s = be s + w  (two roundings, like `s *= be; s += w`); also when the loop ends here, like the reference
        for (uint64_t i = 2 * first; i < n2; i += 2 * step) {
            const double2 a = *reinterpret_cast<const double2 *>(v.s + i);
            const double2 b = *reinterpret_cast<const double2 *>(v.w + i);
            *reinterpret_cast<double2 *>(v.s + i) = make_double2(__dadd_rn(__dmul_rn(a.x, be), b.x), __dadd_rn(__dmul_rn(a.y, be), b.y));
        }
        if (n2 < v.n && first == 0) v.s[n2] = __dadd_rn(__dmul_rn(v.s[n2], be), v.w[n2]);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            st->wd[0] = wd; st->kss[0] = kss; st->al[0] = al; st->be[0] = be;
        }
        grid_barrier(P.bar_count, P.bar_gen, nb);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->wdn[0] = sh_wdn; st->wdn[1] = 0.0;
        st->n = sh_n;
        st->nhist = sh_nhist;
        st->done = sh_done;
    }
}

// can this solve run as the persistent kernel?  (real, Jacobi, no overflow rows, everything 16-byte aligned)
bool cg_persistent_applicable(const ngsb_csr *A, const CgVecs &v)
{
    if (A->kind != NGSB_REAL || A->novf != 0 || A->nslices == 0 || v.invdiag == nullptr || v.master != nullptr || v.dot_out != nullptr || v.R != nullptr) return false;
    if (v.fold_u || v.ip_mode != NGSB_IP_REAL) return false;
    const uintptr_t al = reinterpret_cast<uintptr_t>(v.u) | reinterpret_cast<uintptr_t>(v.d) | reinterpret_cast<uintptr_t>(v.w) | reinterpret_cast<uintptr_t>(v.s) |
                         reinterpret_cast<uintptr_t>(v.as) | reinterpret_cast<uintptr_t>(v.invdiag);
    return (al & 15) == 0;
}

// run up to `iters` iterations of the loop whose state lives in v.state (initialised by the init kernel)
int cg_persistent_launch(const ngsb_csr *A, const CgVecs &v, int iters)
{
    ngsb_ctx *ctx = A->ctx;
    static int occ = 0;
    if (occ == 0) {
        NGSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cg_persistent_kernel, 256, 0));
        if (occ < 1) { set_error("persistent CG: the kernel does not fit an SM"); return NGSB_ERR_UNSUPPORTED; }
    }
    uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)occ;
    const uint64_t need = ((uint64_t)A->nslices + 7) / 8;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    CgPersistParams P;
    memset(&P, 0, sizeof(P));
    SellParams &p = P.sp;
    p.slice_off = A->d_slice_off; p.slice_src = A->d_slice_src; p.row_of = A->row_identity ? nullptr : A->d_row_of; p.scol = A->d_scol; p.sval = A->d_sval;
    if (A->sell_c16_entries > 0 && ctx->sell_c16 != 0) { p.scol16 = A->d_scol16; p.sbase = A->d_sbase; p.slice_c16 = A->d_slice_c16; }
    p.nslices = A->nslices; p.nrows = A->h;
    p.x = v.s; p.y = const_cast<double *>(v.as);
    P.v = v;
    // partials and barrier words live behind the reduction workspace of the context (the fused kernels are not running)
    P.part_a = ctx->d_partials;
    P.part_b = ctx->d_partials + 8192;
    P.bar_count = ctx->d_counter + 2;
    P.bar_gen = ctx->d_counter + 3;
    P.iters = iters;
    NGSB_REQUIRE(grid <= 8192, "persistent CG: grid too large");
    void *args[] = {(void *)&P};
    SpanGuard g(ctx, KC_SPMV);
    NGSB_CUDA(cudaLaunchCooperativeKernel((const void *)cg_persistent_kernel, dim3((unsigned)grid), dim3(256), args, 0, ctx->stream));
    return NGSB_OK;
}

// SparseMatrix<double>::MultAdd(FlatVector alpha, MultiVector x, MultiVector y) (linalg/sparsematrix.cpp:2274-2351): four
// right-hand sides per sweep over the matrix; values and columns are read once, every row keeps four accumulators that
// are summed in storage order exactly like the single-vector kernel (bit-identical results per vector).
struct SellMultiParams {
    const uint64_t *slice_off;
    const uint32_t *slice_src;
    const uint32_t *row_of;
    const int32_t *scol;
    const double *sval;
    uint32_t nslices;
    const double *x[4];
    double *y[4];
    double alpha[4];
};

__global__ void __launch_bounds__(256, 4) sell_spmm4_kernel(const SellMultiParams p)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double *__restrict__ xa = p.x[0], *__restrict__ xb = p.x[1], *__restrict__ xc = p.x[2], *__restrict__ xd = p.x[3];
    for (uint64_t s = (uint64_t)blockIdx.x * 8 + wid; s < p.nslices; s += (uint64_t)gridDim.x * 8) {
        const uint64_t off = p.slice_off[s];
        const uint32_t width = (uint32_t)((p.slice_off[s + 1] - off) >> 5);
        const uint32_t row = p.row_of[(uint32_t)((uint64_t)p.slice_src[s] * 32 + lane)];
        const double2 *v2 = reinterpret_cast<const double2 *>(p.sval + off) + lane;
        const int2 *c2 = reinterpret_cast<const int2 *>(p.scol + off) + lane;
        const uint32_t np = width >> 1;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        uint32_t q = 0;
        for (; q + 2 <= np; q += 2) {
            const double2 va = ldg_stream_d2(v2 + (q + 0) * 32), vb = ldg_stream_d2(v2 + (q + 1) * 32);
            const int2 ca = ldg_stream_i2(c2 + (q + 0) * 32), cb = ldg_stream_i2(c2 + (q + 1) * 32);
            const double a0 = __ldg(xa + ca.x), a1 = __ldg(xa + ca.y), a2 = __ldg(xa + cb.x), a3 = __ldg(xa + cb.y);
            const double b0 = __ldg(xb + ca.x), b1 = __ldg(xb + ca.y), b2 = __ldg(xb + cb.x), b3 = __ldg(xb + cb.y);
            const double c0 = __ldg(xc + ca.x), c1 = __ldg(xc + ca.y), c2_ = __ldg(xc + cb.x), c3 = __ldg(xc + cb.y);
            const double d0 = __ldg(xd + ca.x), d1 = __ldg(xd + ca.y), d2 = __ldg(xd + cb.x), d3 = __ldg(xd + cb.y);
            s0 = fma(va.x, a0, s0); s0 = fma(va.y, a1, s0); s0 = fma(vb.x, a2, s0); s0 = fma(vb.y, a3, s0);
            s1 = fma(va.x, b0, s1); s1 = fma(va.y, b1, s1); s1 = fma(vb.x, b2, s1); s1 = fma(vb.y, b3, s1);
            s2 = fma(va.x, c0, s2); s2 = fma(va.y, c1, s2); s2 = fma(vb.x, c2_, s2); s2 = fma(vb.y, c3, s2);
            s3 = fma(va.x, d0, s3); s3 = fma(va.y, d1, s3); s3 = fma(vb.x, d2, s3); s3 = fma(vb.y, d3, s3);
        }
        for (; q < np; q++) {
            const double2 va = ldg_stream_d2(v2 + q * 32);
            const int2 ca = ldg_stream_i2(c2 + q * 32);
            s0 = fma(va.x, __ldg(xa + ca.x), s0); s0 = fma(va.y, __ldg(xa + ca.y), s0);
            s1 = fma(va.x, __ldg(xb + ca.x), s1); s1 = fma(va.y, __ldg(xb + ca.y), s1);
            s2 = fma(va.x, __ldg(xc + ca.x), s2); s2 = fma(va.y, __ldg(xc + ca.y), s2);
            s3 = fma(va.x, __ldg(xd + ca.x), s3); s3 = fma(va.y, __ldg(xd + ca.y), s3);
        }
        if (row != 0xffffffffu) {
            p.y[0][row] = p.alpha[0] * s0 + p.y[0][row];
            p.y[1][row] = p.alpha[1] * s1 + p.y[1][row];
            p.y[2][row] = p.alpha[2] * s2 + p.y[2][row];
            p.y[3][row] = p.alpha[3] * s3 + p.y[3][row];
        }
    }
}

// overflow part of long rows: one CTA per row, raw sum of val*x over the entries behind `cap`
template <int KIND>
__global__ void __launch_bounds__(256) sell_overflow_kernel(const uint64_t *__restrict__ optr, const int32_t *__restrict__ ocol,
                                                           const double *__restrict__ oval, const double *__restrict__ x,
                                                           double *__restrict__ osum, const CgState *state)
{
    __shared__ double red[3][8];
    if (state != nullptr && state->done) return;
    const uint32_t k = blockIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (uint64_t j = optr[k] + threadIdx.x; j < optr[k + 1]; j += blockDim.x) {
        const int c = ocol[j];
        if (KIND == NGSB_REAL) s0 = fma(oval[j], __ldg(x + c), s0);
        else if (KIND == NGSB_COMPLEX) {
            double2 v = reinterpret_cast<const double2 *>(oval)[j];
            double2 xv = __ldg(reinterpret_cast<const double2 *>(x) + c);
            s0 += v.x * xv.x - v.y * xv.y;
            s1 += v.x * xv.y + v.y * xv.x;
        } else {
            const double *m = oval + 9 * j;
            const double *xv = x + 3 * (size_t)c;
            double x0 = __ldg(xv), x1 = __ldg(xv + 1), x2 = __ldg(xv + 2);
            s0 += m[0] * x0 + m[1] * x1 + m[2] * x2;
            s1 += m[3] * x0 + m[4] * x1 + m[5] * x2;
            s2 += m[6] * x0 + m[7] * x1 + m[8] * x2;
        }
    }
    s0 = warp_sum_s(s0); s1 = warp_sum_s(s1); s2 = warp_sum_s(s2);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = s0; red[1][wid] = s1; red[2][wid] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0, t2 = 0;
        for (int w = 0; w < 8; w++) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
        if (KIND == NGSB_REAL) osum[k] = t0;
        else if (KIND == NGSB_COMPLEX) { osum[2 * k] = t0; osum[2 * k + 1] = t1; }
        else { osum[3 * k] = t0; osum[3 * k + 1] = t1; osum[3 * k + 2] = t2; }
    }
}

// ------------------------------------------------------------------------------------------
// conversion CSR -> SELL-32 on the device
// ------------------------------------------------------------------------------------------
// sigma-sort key of a row: (window of sigma consecutive rows, cap - min(len, cap)); a stable sort by it puts
// the longest rows of every window first and keeps the natural order among rows of equal length
__global__ void __launch_bounds__(256) sell_rowkey_kernel(const uint64_t *__restrict__ rowptr, uint64_t nrows, uint32_t cap, uint32_t sigma,
                                                         uint64_t *__restrict__ key, uint32_t *__restrict__ id)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const uint64_t l = rowptr[r + 1] - rowptr[r];
    const uint32_t len = l > cap ? cap : (uint32_t)l;
    key[r] = ((r / sigma) << 32) | (uint64_t)(cap - len);
    id[r] = (uint32_t)r;
}

__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t *a, uint64_t begin, uint64_t end, uint32_t v, int iota)
{
    const uint64_t t = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < end) a[t] = iota ? (uint32_t)t : v;
}

// per slice: padded length (entries) and the schedule key = smallest first column of its rows.  In a
// finite-element numbering (vertices | edges | faces | cells) every row starts with a vertex dof of
// its patch, so the key places slices of all entity blocks on one spatial axis.
__global__ void __launch_bounds__(256) sell_width_kernel(const uint64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                        const uint32_t *__restrict__ row_of, uint32_t nslices, uint32_t cap, int even,
                                                        uint32_t *__restrict__ slice_len, uint32_t *__restrict__ slice_key)
{
    const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const uint32_t row = row_of[s * 32 + lane];
    uint32_t len = 0, key = 0xffffffffu;
    if (row != 0xffffffffu) {
        const uint64_t a = rowptr[row];
        uint64_t l = rowptr[row + 1] - a;
        len = l > cap ? cap : (uint32_t)l;
        key = l > 0 ? (uint32_t)col[a] : row;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
        key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    }
    if (even) len = (len + 1u) & ~1u;
    if (lane == 0) { slice_len[s] = len * 32u; slice_key[s] = key; }
}

// slice_off[t+1] = padded entries of the slice scheduled at position t (then scanned in place)
__global__ void __launch_bounds__(256) sell_gather_len_kernel(const uint32_t *__restrict__ slice_len, const uint32_t *__restrict__ slice_src,
                                                             uint32_t nslices, uint64_t *__restrict__ slice_off)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nslices) slice_off[t + 1] = slice_len[slice_src[t]];
    if (t == 0) slice_off[0] = 0;
}

template <int KIND>
__global__ void __launch_bounds__(256) sell_fill_kernel(const uint64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                       const double *__restrict__ val, uint32_t nslices, uint32_t cap,
                                                       const uint64_t *__restrict__ slice_off, const uint32_t *__restrict__ slice_src,
                                                       const uint32_t *__restrict__ row_of, int32_t *__restrict__ scol,
                                                       double *__restrict__ sval)
{
    const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const uint64_t off = slice_off[s];
    const uint32_t width = (uint32_t)((slice_off[s + 1] - off) >> 5);
    const uint32_t row = row_of[(uint64_t)slice_src[s] * 32 + lane];
    uint64_t a = 0;
    uint32_t len = 0;
    if (row != 0xffffffffu) {
        a = rowptr[row];
        uint64_t l = rowptr[row + 1] - a;
        len = l > cap ? cap : (uint32_t)l;
    }
    const int32_t padcol = len > 0 ? col[a + len - 1] : 0;
    for (uint32_t j = 0; j < width; j++) {
        const bool real = j < len;
        const int32_t c = real ? col[a + j] : padcol;
        if (KIND == NGSB_REAL) {
            const uint64_t pos = off + ((uint64_t)(j >> 1) * 32 + lane) * 2 + (j & 1);
            scol[pos] = c;
            sval[pos] = real ? val[a + j] : 0.0;
        } else if (KIND == NGSB_COMPLEX) {
            const uint64_t pos = off + (uint64_t)j * 32 + lane;
            scol[pos] = c;
            sval[2 * pos] = real ? val[2 * (a + j)] : 0.0;
            sval[2 * pos + 1] = real ? val[2 * (a + j) + 1] : 0.0;
        } else {
            scol[off + (uint64_t)j * 32 + lane] = c;
#pragma unroll
            for (int k = 0; k < 9; k++) sval[off * 9 + ((uint64_t)j * 9 + k) * 32 + lane] = real ? val[9 * (a + j) + k] : 0.0;
        }
    }
}

// slots whose row is longer than cap -> (slot, row) list (order fixed afterwards by a host sort)
__global__ void __launch_bounds__(256) sell_ovf_collect_kernel(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ row_of,
                                                              uint64_t nslots, uint32_t cap, uint32_t *__restrict__ list, uint32_t *__restrict__ count)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nslots) return;
    const uint32_t row = row_of[t];
    if (row == 0xffffffffu) return;
    if (rowptr[row + 1] - rowptr[row] > cap) {
        const uint32_t k = atomicAdd(count, 1u);
        list[2 * k] = (uint32_t)t;
        list[2 * k + 1] = row;
    }
}

__global__ void __launch_bounds__(256) sell_ovf_copy_kernel(const uint64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                           const double *__restrict__ val, const uint32_t *__restrict__ orows,
                                                           const uint64_t *__restrict__ optr, uint32_t cap, int ms,
                                                           int32_t *__restrict__ ocol, double *__restrict__ oval)
{
    const uint32_t k = blockIdx.x;
    const uint64_t src = rowptr[orows[k]] + cap, n = optr[k + 1] - optr[k], dst = optr[k];
    for (uint64_t j = threadIdx.x; j < n; j += blockDim.x) {
        ocol[dst + j] = col[src + j];
        for (int c = 0; c < ms; c++) oval[(dst + j) * ms + c] = val[(src + j) * ms + c];
    }
}

// 16-bit column compression of the real SELL slices: for every entry step j of a slice the 32 lanes' columns are stored as
// (smallest column of the step) + 16-bit offset when all steps of the slice allow it; otherwise the slice keeps its 32-bit
// columns.  One warp per scheduled slice.  Neighbouring rows of a finite-element numbering couple to neighbouring dofs, so
// the spread inside one step is small except where rows of different structure meet.
template <bool PAIRED>     // PAIRED: real layout (columns in packets of two per lane), else [j][lane] (complex, 3x3 blocks)
__global__ void __launch_bounds__(256) sell_compress_kernel(const uint64_t *__restrict__ slice_off, uint32_t nslices, const int32_t *__restrict__ scol,
                                                           uint16_t *__restrict__ scol16, int32_t *__restrict__ sbase, uint8_t *__restrict__ slice_c16)
{
    const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const uint64_t off = slice_off[s];
    const uint32_t width = (uint32_t)((slice_off[s + 1] - off) >> 5);
    bool ok = true;
    for (uint32_t j = 0; j < width; j++) {
        const uint64_t pos = PAIRED ? off + ((uint64_t)(j >> 1) * 32 + lane) * 2 + (j & 1) : off + (uint64_t)j * 32 + lane;
        const int32_t c = scol[pos];
        int32_t mn = c, mx = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (mx - mn > 65535) ok = false;
    }
    if (ok)
        for (uint32_t j = 0; j < width; j++) {
            const uint64_t pos = PAIRED ? off + ((uint64_t)(j >> 1) * 32 + lane) * 2 + (j & 1) : off + (uint64_t)j * 32 + lane;
            const int32_t c = scol[pos];
            int32_t mn = c;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            scol16[pos] = (uint16_t)(c - mn);
            if (lane == 0) sbase[(off >> 5) + j] = mn;
        }
    if (lane == 0) slice_c16[s] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(256) sell_c16_count_kernel(const uint64_t *__restrict__ slice_off, const uint8_t *__restrict__ slice_c16, uint32_t nslices,
                                                            unsigned long long *__restrict__ out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nslices && slice_c16[t]) atomicAdd(out, (unsigned long long)(slice_off[t + 1] - slice_off[t]));
}

int sell_build(ngsb_csr *A, const uint64_t *h_rowptr)
{
    NvtxRange nv("SELL-32 build");
    ngsb_ctx *ctx = A->ctx;
    const size_t ms = kind_matscalars(A->kind);
    const uint32_t ns = (uint32_t)((A->h + 31) / 32);
    const uint64_t nslots = (uint64_t)ns * 32;
    A->nslices = ns;
    // rows longer than cap keep their tail in the overflow CSR
    uint32_t cap0 = (uint32_t)std::max<double>(64.0, 4.0 * A->mean_row + 0.5);
    if (ctx->sell_cap > 0) cap0 = (uint32_t)ctx->sell_cap;
    cap0 = (cap0 + 1u) & ~1u;
    const uint32_t cap_full = (uint32_t)((std::min<size_t>(A->max_row, (1u << 20)) + 1) & ~(size_t)1);
    // padded size of the natural (unsorted) slices when no row is cut: rows of one slice have similar lengths in an FE
    // numbering (entity by entity), so long rows sit in slices of long rows and cost no padding.  Where that holds
    // (structured numberings) the slices stay in natural order and no row needs the overflow path.
    uint64_t full = 0;
    for (size_t r0 = 0; r0 < A->h; r0 += 32) {
        uint64_t w = 0;
        for (size_t r = r0; r < std::min<size_t>(A->h, r0 + 32); r++) w = std::max<uint64_t>(w, h_rowptr[r + 1] - h_rowptr[r]);
        full += 32 * ((w + 1) & ~(uint64_t)1);
    }
    const bool tight = ctx->sell_cap <= 0 && (double)full <= 1.05 * (double)std::max<size_t>(1, A->nnz) && A->max_row < (1u << 20);
    NGSB_CUDA(cudaMalloc(&A->d_slice_off, ((size_t)ns + 1) * sizeof(uint64_t)));
    NGSB_CUDA(cudaMalloc(&A->d_slice_src, std::max<size_t>(1, ns) * sizeof(uint32_t)));
    NGSB_CUDA(cudaMalloc(&A->d_row_of, std::max<uint64_t>(32, nslots) * sizeof(uint32_t)));
    if (ns == 0) {
        NGSB_CUDA(cudaMemsetAsync(A->d_slice_off, 0, sizeof(uint64_t), ctx->stream));
        A->sell_entries = 0;
        A->sell_cap = cap0;
        NGSB_CUDA(cudaMalloc(&A->d_scol, 16));
        NGSB_CUDA(cudaMalloc(&A->d_sval, 16));
        NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
        return NGSB_OK;
    }
    const unsigned grid_rows = (unsigned)((A->h + 255) / 256), grid_slots = (unsigned)((nslots + 255) / 256), grid_sl = (ns + 255) / 256;
    std::vector<void *> temps;
    auto tmalloc = [&](void **p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes ? bytes : 16); if (e == cudaSuccess) temps.push_back(*p); return e; };
    auto cleanup = [&]() { for (void *p : temps) cudaFree(p); temps.clear(); };
    cudaError_t e1 = cudaSuccess;
    int rc = NGSB_OK;
    // Unstructured numberings (netgen, or any numbering after the Cuthill-McKee reordering): neighbouring rows differ in
    // length (vertex rows of an order-3 space are 5x the mean), so rows are sorted by length inside windows of sigma rows
    // (SELL-C-sigma).  Sorted windows group the long rows, so they can usually stay whole as well: first try without a
    // cap and keep that layout when it pads less than 8 %; otherwise cut rows at cap0 and reduce the tails in the overflow
    // kernel (145 k single-row CTAs per product on the 13.6 M-dof netgen system: 15 % of the product time).
    uint32_t cap = tight ? std::max(cap0, cap_full) : cap0;
    const int attempts = (!tight && ctx->sell_cap <= 0 && ctx->sell_sigma != 0 && ctx->sell_sigma != 1 && cap_full > cap0 && A->max_row < (1u << 20)) ? 2 : 1;
    for (int attempt = 0; attempt < attempts; attempt++) {
        if (attempts == 2) cap = attempt == 0 ? cap_full : cap0;
    // ---- 1. row order: sigma-sort by length inside windows (SELL-C-sigma), identity when sigma <= 1
    // window of the length sort: 65536 rows pad 0.4 % on the 13.6 M-dof netgen system (4096: 4.3 %, 16384: 1.2 %) and the
    // product follows the bytes (1.385 / 1.394 / 1.425 ms, profiles/r2_sweep_netgen14M.jsonl)
    const uint32_t sigma = ctx->sell_sigma >= 0 ? (uint32_t)ctx->sell_sigma : (tight ? 0u : 65536u);
    fill_u32_kernel<<<grid_slots, 256, 0, ctx->stream>>>(A->d_row_of, A->h, nslots, 0xffffffffu, 0);
    if (sigma > 1) {
        uint64_t *d_k = nullptr, *d_k2 = nullptr;
        uint32_t *d_id = nullptr;
        void *d_tmp = nullptr;
        e1 = tmalloc((void **)&d_k, A->h * sizeof(uint64_t));
        if (e1 == cudaSuccess) e1 = tmalloc((void **)&d_k2, A->h * sizeof(uint64_t));
        if (e1 == cudaSuccess) e1 = tmalloc((void **)&d_id, A->h * sizeof(uint32_t));
        if (e1 == cudaSuccess) {
            sell_rowkey_kernel<<<grid_rows, 256, 0, ctx->stream>>>(A->d_rowptr, A->h, cap, sigma, d_k, d_id);
            size_t tmp_bytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_k, d_k2, d_id, A->d_row_of, (int)A->h, 0, 64, ctx->stream);
            e1 = tmalloc(&d_tmp, tmp_bytes);
            if (e1 == cudaSuccess) e1 = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_k, d_k2, d_id, A->d_row_of, (int)A->h, 0, 64, ctx->stream);
        }
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
        cleanup();
    } else {
        fill_u32_kernel<<<grid_rows, 256, 0, ctx->stream>>>(A->d_row_of, 0, A->h, 0, 1);
    }
    if (e1 != cudaSuccess) { set_error("SELL build (row order): %s", cudaGetErrorString(e1)); return NGSB_ERR_CUDA; }
    A->row_identity = sigma <= 1;
    // ---- 2. slice widths, schedule, offsets
    {
        uint32_t *d_len = nullptr, *d_key = nullptr, *d_key2 = nullptr, *d_id = nullptr;
        void *d_tmp = nullptr;
        e1 = tmalloc((void **)&d_len, ns * sizeof(uint32_t));
        if (e1 == cudaSuccess) e1 = tmalloc((void **)&d_key, ns * sizeof(uint32_t));
        if (e1 == cudaSuccess) e1 = tmalloc((void **)&d_key2, ns * sizeof(uint32_t));
        if (e1 == cudaSuccess) e1 = tmalloc((void **)&d_id, ns * sizeof(uint32_t));
        if (e1 != cudaSuccess) { cleanup(); set_error("SELL build: %s", cudaGetErrorString(e1)); return NGSB_ERR_NOMEM; }
        sell_width_kernel<<<grid_slots, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_row_of, ns, cap, A->kind == NGSB_REAL ? 1 : 0, d_len, d_key);
        fill_u32_kernel<<<grid_sl, 256, 0, ctx->stream>>>(d_id, 0, ns, 0, 1);
        // 3x3 blocks: measured slower with the schedule on B200 (x is a small share of the traffic), keep natural order
        const bool schedule = ctx->sell_schedule == 1 ? A->kind != NGSB_BLOCK3 : ctx->sell_schedule == 2;
        if (schedule) {
            // schedule: slices ordered by key (stable radix sort keeps the natural order among equal keys)
            size_t tmp_bytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key2, d_id, A->d_slice_src, (int)ns, 0, 32, ctx->stream);
            e1 = tmalloc(&d_tmp, tmp_bytes);
            if (e1 == cudaSuccess) e1 = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key2, d_id, A->d_slice_src, (int)ns, 0, 32, ctx->stream);
        } else {
            e1 = cudaMemcpyAsync(A->d_slice_src, d_id, ns * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
        }
        if (e1 == cudaSuccess) {
            sell_gather_len_kernel<<<grid_sl, 256, 0, ctx->stream>>>(d_len, A->d_slice_src, ns, A->d_slice_off);
            e1 = cudaGetLastError();
        }
        rc = e1 == cudaSuccess ? device_scan_u64(ctx, A->d_slice_off, (uint64_t)ns + 1) : NGSB_ERR_CUDA;
        if (rc == NGSB_OK) {
            e1 = cudaMemcpyAsync(&A->sell_entries, A->d_slice_off + ns, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
        }
        cleanup();
        if (e1 != cudaSuccess) { set_error("SELL build (schedule): %s", cudaGetErrorString(e1)); return NGSB_ERR_CUDA; }
        if (rc != NGSB_OK) return rc;
    }
        if (attempts == 2 && attempt == 0 && (double)A->sell_entries <= 1.08 * (double)A->nnz) break;      // whole rows: accepted
    }
    A->sell_cap = cap;
    uint32_t novf = 0;
    for (size_t r = 0; r < A->h; r++) novf += (h_rowptr[r + 1] - h_rowptr[r]) > cap;
    A->novf = novf;
    // ---- 3. fill
    e1 = cudaMalloc(&A->d_scol, std::max<size_t>(16, A->sell_entries * sizeof(int32_t)));
    if (e1 == cudaSuccess) e1 = cudaMalloc(&A->d_sval, std::max<size_t>(16, A->sell_entries * ms * sizeof(double)));
    if (e1 != cudaSuccess) { set_error("SELL build: cudaMalloc of %llu entries failed: %s", (unsigned long long)A->sell_entries, cudaGetErrorString(e1)); return NGSB_ERR_NOMEM; }
    if (A->kind == NGSB_REAL) sell_fill_kernel<NGSB_REAL><<<grid_slots, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, ns, cap, A->d_slice_off, A->d_slice_src, A->d_row_of, A->d_scol, A->d_sval);
    else if (A->kind == NGSB_COMPLEX) sell_fill_kernel<NGSB_COMPLEX><<<grid_slots, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, ns, cap, A->d_slice_off, A->d_slice_src, A->d_row_of, A->d_scol, A->d_sval);
    else sell_fill_kernel<NGSB_BLOCK3><<<grid_slots, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, ns, cap, A->d_slice_off, A->d_slice_src, A->d_row_of, A->d_scol, A->d_sval);
    NGSB_CUDA(cudaGetLastError());
    // ---- 3b. 16-bit column offsets for the real kernel (option sell_c16, default on)
    A->sell_c16_entries = 0;
    // complex and 3x3-block matrices ([j][lane] columns): option sell_c16_all, default off until measured
    if ((A->kind == NGSB_REAL ? ctx->sell_c16 != 0 : ctx->sell_c16_all != 0) && A->sell_entries > 0) {
        unsigned long long *d_cnt = nullptr;
        NGSB_CUDA(cudaMalloc(&A->d_scol16, (A->sell_entries + 64) * sizeof(uint16_t)));
        NGSB_CUDA(cudaMalloc(&A->d_sbase, (A->sell_entries / 32 + 8) * sizeof(int32_t)));
        NGSB_CUDA(cudaMalloc(&A->d_slice_c16, ns));
        NGSB_CUDA(cudaMalloc(&d_cnt, sizeof(unsigned long long)));
        NGSB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), ctx->stream));
        NGSB_CUDA(cudaMemsetAsync(A->d_scol16, 0, (A->sell_entries + 64) * sizeof(uint16_t), ctx->stream));
        NGSB_CUDA(cudaMemsetAsync(A->d_sbase, 0, (A->sell_entries / 32 + 8) * sizeof(int32_t), ctx->stream));
        if (A->kind == NGSB_REAL) sell_compress_kernel<true><<<grid_slots, 256, 0, ctx->stream>>>(A->d_slice_off, ns, A->d_scol, A->d_scol16, A->d_sbase, A->d_slice_c16);
        else sell_compress_kernel<false><<<grid_slots, 256, 0, ctx->stream>>>(A->d_slice_off, ns, A->d_scol, A->d_scol16, A->d_sbase, A->d_slice_c16);
        sell_c16_count_kernel<<<grid_sl, 256, 0, ctx->stream>>>(A->d_slice_off, A->d_slice_c16, ns, d_cnt);
        NGSB_CUDA(cudaGetLastError());
        unsigned long long cnt = 0;
        NGSB_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
        NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_cnt);
        A->sell_c16_entries = cnt;
        if ((double)cnt < 0.25 * (double)A->sell_entries) {
            // hardly any slice qualifies (unstructured numberings at scale: 3 % on the 13.6 M-dof netgen system): the 2 bytes
            // per entry of offset storage are better spent elsewhere, every slice keeps its 32-bit columns
            cudaFree(A->d_scol16); A->d_scol16 = nullptr;
            cudaFree(A->d_sbase); A->d_sbase = nullptr;
            cudaFree(A->d_slice_c16); A->d_slice_c16 = nullptr;
            A->sell_c16_entries = 0;
        }
    }
    // ---- 4. overflow CSR of the rows longer than cap
    if (novf) {
        uint32_t *d_list = nullptr, *d_count = nullptr;
        NGSB_CUDA(cudaMalloc(&d_list, (size_t)novf * 2 * sizeof(uint32_t)));
        NGSB_CUDA(cudaMalloc(&d_count, sizeof(uint32_t)));
        NGSB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint32_t), ctx->stream));
        sell_ovf_collect_kernel<<<grid_slots, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_row_of, nslots, cap, d_list, d_count);
        std::vector<uint32_t> list((size_t)novf * 2);
        NGSB_CUDA(cudaMemcpyAsync(list.data(), d_list, list.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_list);
        cudaFree(d_count);
        std::vector<std::pair<uint32_t, uint32_t>> pairs(novf);
        for (uint32_t k = 0; k < novf; k++) pairs[k] = std::make_pair(list[2 * k], list[2 * k + 1]);
        std::sort(pairs.begin(), pairs.end());
        std::vector<uint32_t> oslot(novf), orows(novf);
        std::vector<uint64_t> optr(1, 0);
        std::vector<int32_t> slice_ovf(ns, -1);
        for (uint32_t k = 0; k < novf; k++) {
            oslot[k] = pairs[k].first;
            orows[k] = pairs[k].second;
            if (slice_ovf[oslot[k] >> 5] < 0) slice_ovf[oslot[k] >> 5] = (int32_t)k;
            optr.push_back(optr.back() + (h_rowptr[orows[k] + 1] - h_rowptr[orows[k]] - cap));
        }
        const uint64_t on = optr.back();
        NGSB_CUDA(cudaMalloc(&A->d_ovf_slot, novf * sizeof(uint32_t)));
        NGSB_CUDA(cudaMalloc(&A->d_ovf_rows, novf * sizeof(uint32_t)));
        NGSB_CUDA(cudaMalloc(&A->d_ovf_ptr, optr.size() * sizeof(uint64_t)));
        NGSB_CUDA(cudaMalloc(&A->d_slice_ovf, slice_ovf.size() * sizeof(int32_t)));
        NGSB_CUDA(cudaMalloc(&A->d_ovf_col, on * sizeof(int32_t)));
        NGSB_CUDA(cudaMalloc(&A->d_ovf_val, on * ms * sizeof(double)));
        NGSB_CUDA(cudaMalloc(&A->d_ovf_sum, (size_t)novf * 3 * sizeof(double)));
        NGSB_CUDA(cudaMemcpyAsync(A->d_ovf_slot, oslot.data(), novf * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        NGSB_CUDA(cudaMemcpyAsync(A->d_ovf_rows, orows.data(), novf * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        NGSB_CUDA(cudaMemcpyAsync(A->d_ovf_ptr, optr.data(), optr.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        NGSB_CUDA(cudaMemcpyAsync(A->d_slice_ovf, slice_ovf.data(), slice_ovf.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        sell_ovf_copy_kernel<<<novf, 256, 0, ctx->stream>>>(A->d_rowptr, A->d_col, A->d_val, A->d_ovf_rows, A->d_ovf_ptr, cap, (int)ms, A->d_ovf_col, A->d_ovf_val);
        NGSB_CUDA(cudaGetLastError());
        NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    NGSB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NGSB_OK;
}

void sell_free(ngsb_csr *A)
{
    cudaFree(A->d_slice_off); cudaFree(A->d_slice_src); cudaFree(A->d_row_of); cudaFree(A->d_ovf_slot); cudaFree(A->d_scol); cudaFree(A->d_sval);
    cudaFree(A->d_ovf_rows); cudaFree(A->d_ovf_ptr); cudaFree(A->d_slice_ovf);
    cudaFree(A->d_ovf_col); cudaFree(A->d_ovf_val); cudaFree(A->d_ovf_sum);
    cudaFree(A->d_scol16); cudaFree(A->d_sbase); cudaFree(A->d_slice_c16);
}

int sell_launch(const SpmvArgs &a)
{
    const ngsb_csr *A = a.A;
    ngsb_ctx *ctx = A->ctx;
    SellParams p;
    memset(&p, 0, sizeof(p));
    p.slice_off = A->d_slice_off; p.slice_src = A->d_slice_src; p.row_of = a.user_rows ? A->d_row_user : (A->row_identity ? nullptr : A->d_row_of); p.scol = A->d_scol; p.sval = A->d_sval;
    p.slice_ovf = A->novf ? A->d_slice_ovf : nullptr;
    p.ovf_slot = A->d_ovf_slot; p.ovf_sum = A->d_ovf_sum; p.novf = A->novf;
    p.nslices = A->nslices; p.nrows = A->h;
    if (A->sell_c16_entries > 0 && (A->kind == NGSB_REAL ? ctx->sell_c16 != 0 : ctx->sell_c16_all != 0)) { p.scol16 = A->d_scol16; p.sbase = A->d_sbase; p.slice_c16 = A->d_slice_c16; }
    p.pf_steps = (int)ctx->sell_pf_steps;
    p.pf_next = (int)ctx->sell_pf_next;
    p.x = a.x; p.y = a.y; p.sr = a.sr; p.si = A->kind == NGSB_COMPLEX ? a.si : 0.0;
    p.accumulate = a.accumulate ? 1 : 0; p.epi = a.epi; p.dot_conj = a.dot_conj;
    p.dotvec = a.dotvec; p.dot_out = a.dot_out; p.state = a.state;
    p.partials = ctx->d_partials; p.counter = ctx->d_counter;
    p.slice_list = a.slice_list; p.nlist = a.nlist; p.dot_add = a.dot_add;
    const bool list = a.slice_list != nullptr;
    const bool push = a.push != nullptr;
    if (push) {
        NGSB_REQUIRE(a.epi != EPI_NONE && !list && !a.user_rows && a.push_slice_src, "SpMV: the fused neighbour exchange needs the fused dot and the whole schedule");
        p.push = a.push;
        p.slice_src = a.push_slice_src;
    }
    if (A->novf && !a.skip_overflow) {
        SpanGuard g(ctx, KC_SPMV);
        if (A->kind == NGSB_REAL) sell_overflow_kernel<NGSB_REAL><<<A->novf, 256, 0, ctx->stream>>>(A->d_ovf_ptr, A->d_ovf_col, A->d_ovf_val, a.x, A->d_ovf_sum, a.state);
        else if (A->kind == NGSB_COMPLEX) sell_overflow_kernel<NGSB_COMPLEX><<<A->novf, 256, 0, ctx->stream>>>(A->d_ovf_ptr, A->d_ovf_col, A->d_ovf_val, a.x, A->d_ovf_sum, a.state);
        else sell_overflow_kernel<NGSB_BLOCK3><<<A->novf, 256, 0, ctx->stream>>>(A->d_ovf_ptr, A->d_ovf_col, A->d_ovf_val, a.x, A->d_ovf_sum, a.state);
        NGSB_CUDA(cudaGetLastError());
    }
    // grid = resident CTAs (occupancy query) unless overridden: every CTA stays on its SM for the whole sweep
    typedef void (*kern_t)(const SellParams);
    kern_t kern;
    const long var = ctx->sell_variant;
    if (push) {
        if (A->kind == NGSB_REAL) kern = sell_spmv_kernel<NGSB_REAL, 2, 4, false, false, true>;
        else if (A->kind == NGSB_COMPLEX) kern = sell_spmv_kernel<NGSB_COMPLEX, 2, 4, false, false, true>;
        else kern = sell_spmv_kernel<NGSB_BLOCK3, 0, 5, false, false, true>;
    } else if (list) {
        // interface-first split: the default loops only
        if (A->kind == NGSB_REAL) kern = sell_spmv_kernel<NGSB_REAL, 2, 4, false, true>;
        else if (A->kind == NGSB_COMPLEX) kern = sell_spmv_kernel<NGSB_COMPLEX, 2, 4, false, true>;
        else kern = sell_spmv_kernel<NGSB_BLOCK3, 0, 5, false, true>;
    } else if (A->kind == NGSB_REAL) {
        // default (0): software-pipelined loop, 64 registers, 4 resident CTAs, grid of 8 CTAs per SM --
        // best of the variants swept on B200 at 6-101 M rows (profiles/r1_sweep_sell_variants.txt)
        switch (var) {
        case 1: kern = sell_spmv_kernel<NGSB_REAL, 0, 5>; break;
        case 2: kern = sell_spmv_kernel<NGSB_REAL, 1, 8>; break;
        case 3: kern = sell_spmv_kernel<NGSB_REAL, 2, 4, true>; break;      // + L2 eviction policies in the compressed loop
        case 4: kern = sell_spmv_kernel<NGSB_REAL, 4, 5>; break;            // compressed loop with late values: 48 registers, 5 CTAs/SM
        case 5: kern = sell_spmv_kernel<NGSB_REAL, 4, 4>; break;            // the same loop at 4 CTAs/SM (A/B)
        case 6: kern = sell_spmv_kernel<NGSB_REAL, 4, 6>; break;            // ... at 6 CTAs/SM (40 registers)
        default: kern = sell_spmv_kernel<NGSB_REAL, 2, 4>; break;
        }
    } else if (A->kind == NGSB_COMPLEX) kern = var == 1 ? sell_spmv_kernel<NGSB_COMPLEX, 0, 5> : sell_spmv_kernel<NGSB_COMPLEX, 2, 4>;
    else kern = sell_spmv_kernel<NGSB_BLOCK3, 0, 5>;
    // Grid: CTA b walks the slices b*8+w, b*8+w + 8*grid, ... so the CTAs resident at one time form a window of adjacent
    // slices inside every stripe of 8*grid slices.  With a small (persistent) grid the whole schedule is swept once per
    // wave of CTAs and x is fetched again each time; with ~96 CTAs per SM there are only a few dozen wide stripes, every
    // window slides through its stripe once and the x lines are re-used out of L2.  Measured at 111 M dofs: 10.66 ms with
    // 8 CTAs/SM, 9.31 ms with 96 (profiles/r1_sweep_c16_grid_size.txt); 3x3 blocks +18 %, complex +9 %.
    // ... and about eight slices per warp: below ~30 M rows a grid of 96 CTAs per SM leaves each CTA only 3-4 slices and the
    // per-CTA cost (launch, prologue, ticket) shows in the CG loop (13.6 M-dof netgen system: 589.6 it/s at 96, 595.7 at 48;
    // one rank's slab of the 8-GPU run: 1.490 -> 1.456 ms per iteration, 637 -> 653 it/s on eight GPUs)
    const uint64_t nwork = (uint64_t)(list ? a.nlist : A->nslices);
    long cps = ctx->spmv_ctas_per_sm > 0 ? ctx->spmv_ctas_per_sm
                                         : (long)std::min<uint64_t>(96, std::max<uint64_t>(32, nwork / (64ull * (uint64_t)std::max(1, ctx->sm_count))));
    uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)cps;
    const uint64_t need = (nwork + 7) / 8;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (grid > (uint64_t)MAX_PARTIALS - 8) grid = MAX_PARTIALS - 8;
    if (ctx->spmv_ctas_per_sm <= 0 && need > 0) {
        // whole rounds: every CTA walks the same number of 8-slice steps (46 CTAs per SM on the 14 M-row slab meant 8.1
        // rounds, i.e. a ninth, nearly empty one: 636 instead of 653 it/s on eight GPUs)
        const uint64_t rounds = (need + grid - 1) / grid;
        grid = std::max<uint64_t>(1, (need + rounds - 1) / rounds);
    }
    SpanGuard g(ctx, KC_SPMV);
    kern<<<(unsigned)grid, 256, 0, ctx->stream>>>(p);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

// four real right-hand sides in one sweep; the caller guarantees kind == REAL and no overflow rows
int sell_launch_multi4(const ngsb_csr *A, const double *const x[4], double *const y[4], const double alpha[4])
{
    ngsb_ctx *ctx = A->ctx;
    SellMultiParams p;
    memset(&p, 0, sizeof(p));
    p.slice_off = A->d_slice_off; p.slice_src = A->d_slice_src; p.row_of = A->d_row_of; p.scol = A->d_scol; p.sval = A->d_sval;
    p.nslices = A->nslices;
    for (int k = 0; k < 4; k++) { p.x[k] = x[k]; p.y[k] = y[k]; p.alpha[k] = alpha[k]; }
    uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)(ctx->spmv_ctas_per_sm > 0 ? ctx->spmv_ctas_per_sm : 96);
    const uint64_t need = ((uint64_t)A->nslices + 7) / 8;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    SpanGuard g(ctx, KC_SPMV);
    sell_spmm4_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(p);
    NGSB_CUDA(cudaGetLastError());
    return NGSB_OK;
}

} // namespace ngsb
