// krylov.cuh -- internal interface of the fused CG kernels (krylov.cu), shared with dist.cu
#pragma once
#include "jacobi.cuh"

struct PeerReduce;

namespace ngsb {

struct CgVecs {
    double *u, *d, *w, *s;
    const double *as, *f;
    const double *invdiag;     // NULL: no preconditioner (w aliases d, never stored)
    const uint8_t *bits;       // `inner` BitArray of the Jacobi preconditioner or NULL
    const uint8_t *master;     // distributed: 1 byte per dof, dot restricted to master dofs
                               // (masked inner product, parallel/parallelvvector.cpp:305-314); NULL: all
    double *dot_out;           // distributed: the local dot goes here (2 doubles) and the scalar
                               // step runs after the all-reduce; NULL: last block finalises
    const PeerReduce *R;       // distributed, peer-memory mode: the finishing thread all-reduces the dot over the ranks itself
                               // (partials into the peers' mailboxes, rank-ordered sum) and does the scalar step -- no
                               // separate finalize launch; takes precedence over dot_out
    uint64_t n;                // entries
    CgState *state;
    double *hist;
    double *partials;
    unsigned int *counter;
    int ip_mode;
    int stream;                // option cg_stream_hints: streaming loads / evict-first stores for the vectors nobody reads next
    int chunked;               // option cg_chunked: contiguous chunk per CTA instead of the grid-stride split (A/B)
    int fold_u;                // option cg_fold_u: `u += al s` moves from the update kernel into the direction kernel, which
                               // reads s anyway (10 instead of 11 vector passes per iteration, same arithmetic per entry)
};

// mode 0: init (d = f [- as], w = C d, s = w, <w,d>), mode 1: update (u, d, w, <d,w>)
int cg_launch_fused(ngsb_ctx *ctx, int kind, int mode, const CgVecs &v, int sub);
int cg_launch_dir(ngsb_ctx *ctx, int kind, const CgVecs &v);
// scalar steps from a reduced dot (device, 2 doubles): which 0 init, 1 kss, 2 wdn
// R != NULL (distributed, peer-memory mode): `dot` holds this rank's partial and is all-reduced in place first
int cg_launch_finalize(ngsb_ctx *ctx, int which, CgState *st, double *dot, double *hist, const PeerReduce *R = nullptr);

// hooks that turn the one-GPU GMRES into the parallel-vector version (dist.cu)
struct GmresDist {
    const uint8_t *master;                        // device, 1 byte per dof
    const PeerReduce *R;                          // device; NULL: reduce with `allreduce`
    int (*allreduce)(void *arg, double *d_buf);   // enqueue an all-reduce (sum) of 2 doubles; NULL in peer-memory mode
    int (*cumulate)(void *arg, double *v);        // ParallelBaseVector::Cumulate of a raw local vector
    void *arg;
};
int gmres_solve_impl(const ngsb_csr *A, const ngsb_jacobi *C, const ngsb_vec *f, ngsb_vec *x, double prec, int maxsteps,
                     int initialize, int *steps, double *history, int hist_cap, int *nhist, const GmresDist *dist);

// the whole real Jacobi-PCG loop as one persistent cooperative kernel (sell.cu)
bool cg_persistent_applicable(const ngsb_csr *A, const CgVecs &v);
int cg_persistent_launch(const ngsb_csr *A, const CgVecs &v, int iters);

// solver workspace (cached on the context)
int ws_get_buf(ngsb_ctx *ctx, size_t nscal, double **out);
void ws_put_buf(ngsb_ctx *ctx, size_t nscal, double *p);
int ws_state(ngsb_ctx *ctx, size_t hist_cap, CgState **d_state, CgState **h_state, double **d_hist);

} // namespace ngsb
