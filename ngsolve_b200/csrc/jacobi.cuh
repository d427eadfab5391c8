// jacobi.cuh -- internal interface of the Jacobi preconditioner (jacobi.cu)
#pragma once
#include "spmv.cuh"

struct ngsb_jacobi {
    ngsb_ctx *ctx = nullptr;
    size_t n = 0;
    int kind = 0;
    double *d_invdiag = nullptr;   // n * (1 | 2 | 9) doubles
    uint8_t *d_bits = nullptr;     // `inner` BitArray bytes or NULL
    uint64_t uid = 0;              // unique per created preconditioner (key of cached CUDA graphs)
    // the same preconditioner in the numbering of an internally reordered matrix (built on first use by a fused solver)
    mutable ngsb_jacobi *permuted = nullptr;
    mutable uint64_t permuted_for = 0;      // uid of the matrix whose permutation `permuted` follows
};

namespace ngsb {
// J in the numbering of A->inner (A internally reordered): entries and freedofs bits gathered through A->d_perm, cached on J
int jacobi_for_inner(const ngsb_jacobi *J, const ngsb_csr *A, const ngsb_jacobi **out);
int jacobi_apply(const ngsb_jacobi *J, double sr, double si, const double *x, double *y, bool accumulate);
// JacobiPrecond ctor; `cumulate(arg, diag, doubles_per_entry)` (may be NULL) runs between extracting the
// diagonal and inverting it (AllReduceDofData of the distributed case, linalg/jacobi.cpp:60-61)
int jacobi_build(const ngsb_csr *A, const uint8_t *freebits, int (*cumulate)(void *, double *, int), void *cum_arg, ngsb_jacobi **out);
}
