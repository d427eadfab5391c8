// krylov.cuh -- internal interface of the fused CG kernels (krylov.cu), shared with dist.cu
#pragma once
#include "jacobi.cuh"

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
    uint64_t n;                // entries
    CgState *state;
    double *hist;
    double *partials;
    unsigned int *counter;
    int ip_mode;
};

// mode 0: init (d = f [- as], w = C d, s = w, <w,d>), mode 1: update (u, d, w, <d,w>)
int cg_launch_fused(ngsb_ctx *ctx, int kind, int mode, const CgVecs &v, int sub);
int cg_launch_dir(ngsb_ctx *ctx, int kind, const CgVecs &v);
// scalar steps from a reduced dot (device, 2 doubles): which 0 init, 1 kss, 2 wdn
int cg_launch_finalize(ngsb_ctx *ctx, int which, CgState *st, const double *dot, double *hist);

// solver workspace (cached on the context)
int ws_get_buf(ngsb_ctx *ctx, size_t nscal, double **out);
void ws_put_buf(ngsb_ctx *ctx, size_t nscal, double *p);
int ws_state(ngsb_ctx *ctx, size_t hist_cap, CgState **d_state, CgState **h_state, double **d_hist);

} // namespace ngsb
