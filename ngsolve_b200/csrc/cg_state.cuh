// cg_state.cuh -- device-resident state of one CGSolver<IPTYPE>::Mult run
// (linalg/cg.cpp:503-633).  All scalars the reference keeps in host variables
// (al, be, wd, wdn, kss, err, n) live here, so the loop never returns to the host:
// the last block of each reduction kernel does the scalar step and evaluates the
// loop condition; every kernel of the iteration exits at once when `done` is set.
#pragma once
#include <stdint.h>

struct CgState {
    double wd[2];
    double wdn[2];
    double kss[2];
    double al[2];
    double be[2];
    double err;        // prec^2 * Abs(wdn0)
    double prec2;      // prec^2
    int n;             // the reference's loop counter n (GetSteps() at exit)
    int maxsteps;
    int done;          // 1: loop finished (condition false, or kss == 0 break)
    int nhist;         // history entries written (capped by hist_cap)
    int hist_cap;
    int cplx;          // scalars are complex
    int u_pending;     // option cg_fold_u: the loop ended in this iteration and `u += al s` is still owed by its
                       // direction kernel (set by cg_finalize_wdn, cleared by the next update kernel's early exit)
    int pad;
};

#ifdef __CUDACC__
__device__ __forceinline__ double cg_abs(const double *z, int cplx)
{
    return cplx ? hypot(z[0], z[1]) : fabs(z[0]);
}

__device__ __forceinline__ void cg_div(const double *a, const double *b, double *out, int cplx)
{
    if (!cplx) { out[0] = a[0] / b[0]; out[1] = 0.0; return; }
    double den = b[0] * b[0] + b[1] * b[1];
    double re = (a[0] * b[0] + a[1] * b[1]) / den;
    double im = (a[1] * b[0] - a[0] * b[1]) / den;
    out[0] = re;
    out[1] = im;
}

// `while (n++ < maxsteps && Abs(wdn) > err)` -- linalg/cg.cpp:593
__device__ __forceinline__ void cg_eval_loop_condition(CgState *st)
{
    bool cont = (st->n++ < st->maxsteps) && (cg_abs(st->wdn, st->cplx) > st->err);
    if (!cont) st->done = 1;
}

// after <w,d> of the initial residual: cg.cpp:576-588
__device__ __forceinline__ void cg_finalize_init(CgState *st, double2 total, double *hist)
{
    st->wdn[0] = total.x;
    st->wdn[1] = st->cplx ? total.y : 0.0;
    if (st->hist_cap > 0) hist[0] = cg_abs(st->wdn, st->cplx);
    st->nhist = 1;
    if (st->wdn[0] == 0.0 && st->wdn[1] == 0.0) { st->wdn[0] = 1.0; st->wdn[1] = 0.0; }
    st->err = st->prec2 * cg_abs(st->wdn, st->cplx);
    st->n = 0;
    st->done = 0;
    cg_eval_loop_condition(st);
}

// after kss = <s, A s>: cg.cpp:596-600
__device__ __forceinline__ void cg_finalize_kss(CgState *st, double2 total)
{
    st->wd[0] = st->wdn[0];
    st->wd[1] = st->wdn[1];
    st->kss[0] = total.x;
    st->kss[1] = st->cplx ? total.y : 0.0;
    if (st->kss[0] == 0.0 && st->kss[1] == 0.0) { st->done = 1; return; }   // `if (kss == 0.0) break;`
    cg_div(st->wd, st->kss, st->al, st->cplx);
}

// after wdn = <d, w>: cg.cpp:609-618 (+ the loop condition of the next pass)
__device__ __forceinline__ void cg_finalize_wdn(CgState *st, double2 total, double *hist)
{
    st->wdn[0] = total.x;
    st->wdn[1] = st->cplx ? total.y : 0.0;
    cg_div(st->wdn, st->wd, st->be, st->cplx);
    if (st->nhist < st->hist_cap) hist[st->nhist] = cg_abs(st->wdn, st->cplx);
    st->nhist++;
    cg_eval_loop_condition(st);
    if (st->done) st->u_pending = 1;
}
#endif
