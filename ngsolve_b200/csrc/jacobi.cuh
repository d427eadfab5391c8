// jacobi.cuh -- internal interface of the Jacobi preconditioner (jacobi.cu)
#pragma once
#include "spmv.cuh"

struct ngsb_jacobi {
    ngsb_ctx *ctx = nullptr;
    size_t n = 0;
    int kind = 0;
    double *d_invdiag = nullptr;   // n * (1 | 2 | 9) doubles
    uint8_t *d_bits = nullptr;     // `inner` BitArray bytes or NULL
};

namespace ngsb {
int jacobi_apply(const ngsb_jacobi *J, double sr, double si, const double *x, double *y, bool accumulate);
}
