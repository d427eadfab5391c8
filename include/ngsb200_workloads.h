/*
 * ngsb200_workloads.h -- synthetic finite-element systems for benchmarks and large parity checks.
 *
 * NOT part of the drop-in boundary (include/ngsb200.h): the reference gets its matrices from
 * BilinearForm::Assemble (comp/bilinearform.cpp) on netgen meshes, which do not exist on the
 * GPU box.  This generator assembles the same kind of system directly on the device:
 * H1-conforming order-p (p = 1..4) hierarchical elements on a Kuhn-triangulated box of
 * nx*ny*nz cubes (6 tetrahedra per cube), dofs numbered the NGSolve way
 * (vertices | edges | faces | cells, comp/h1hofespace.cpp:833-880), Dirichlet dofs kept in
 * the matrix and flagged in a freedofs BitArray (dirichlet=".*"), right-hand side (1, v).
 *   kind NGSB_REAL    : grad u . grad v + mass_re * u v            (Poisson: mass_re = 0)
 *   kind NGSB_COMPLEX : grad u . grad v + (mass_re + i mass_im) u v  (shifted Laplace / Helmholtz)
 *   kind NGSB_BLOCK3  : isotropic linear elasticity, Lame (lambda, mu), Mat<3,3> entries
 * A box may be a sub-domain of a larger global grid (offset + global size): the local matrix
 * then holds the contributions of the local elements only, exactly the reference's MPI split
 * (comp/bilinearform.cpp:6492-6494), and dof keys / boundary flags give the ParallelDofs
 * exchange tables.
 */
#ifndef NGSB200_WORKLOADS_H
#define NGSB200_WORKLOADS_H

#include "ngsb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngsb_femgen ngsb_femgen;

typedef struct {
    int order;                 /* 1..4 */
    int kind;                  /* NGSB_REAL, NGSB_COMPLEX, NGSB_BLOCK3 */
    int n[3];                  /* cubes of this box */
    int offset[3];             /* position of the box in the global grid (cubes) */
    int global[3];             /* cubes of the global grid */
    double h;                  /* cube edge length */
    double mass_re, mass_im;   /* coefficient of the mass term (scalar kinds) */
    double lame_lambda, lame_mu; /* NGSB_BLOCK3 */
} ngsb_femgen_desc;

int ngsb_femgen_create(const ngsb_femgen_desc *desc, ngsb_femgen **out);
int ngsb_femgen_destroy(ngsb_femgen *g);
int ngsb_femgen_sizes(const ngsb_femgen *g, size_t *ndof, size_t *global_ndof);
/* per local dof: index in the global numbering, 1 if the dof lies on the surface of this box,
 * 1 if it is a free (non-Dirichlet) dof.  Any pointer may be NULL. */
int ngsb_femgen_dof_info(const ngsb_femgen *g, uint64_t *global_index, uint8_t *on_box_surface, uint8_t *is_free);
/* host assembly (plain loops; for tests and small systems).  rowptr: ndof+1.  Call with
 * col == NULL to get the row pointers only. */
int ngsb_femgen_host(const ngsb_femgen *g, uint64_t *rowptr, int32_t *col, void *val, void *rhs);
/* device assembly straight into a library matrix + right-hand-side vector */
int ngsb_femgen_device(ngsb_ctx *ctx, const ngsb_femgen *g, ngsb_csr **A, ngsb_vec **rhs);

#ifdef __cplusplus
}
#endif
#endif
