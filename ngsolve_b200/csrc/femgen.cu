// femgen.cu -- synthetic finite-element systems assembled on the device (include/ngsb200_workloads.h).
//
// Benchmark / test input generator, not part of the drop-in boundary.  It produces what
// BilinearForm::Assemble (comp/bilinearform.cpp) would hand to the solve path for an
// H1 order-p space on a tetrahedral mesh of a box: a CSR matrix with ascending columns per
// row, dofs numbered vertices | edges | faces | cells (comp/h1hofespace.cpp:833-880),
// Dirichlet dofs kept in the matrix, plus the load vector of f = 1.
//
// Mesh: nx*ny*nz cubes, each cut into the 6 Kuhn tetrahedra (one per permutation pi of the
// axes: v0 = corner, v1 = v0 + e_pi0, v2 = v1 + e_pi1, v3 = v2 + e_pi2).  Global vertex
// numbers grow along every path, so "vertices sorted by global number" (NGSolve's
// orientation rule) is the path order, and the whole mesh is translation invariant: every
// row of the matrix is one of a few stencils (row type = entity type x local dof), built
// once on the host from exact polynomial integration; a row at the boundary of the box
// keeps only the contributions of the tetrahedra that exist there.  The kernels evaluate one
// row per thread twice (count, then fill); no atomics, no sorting, columns come out ascending.
//
// Basis (hierarchical, barycentric l_a, vertices a<b<c of the entity in global order):
//   vertex  l_a;  edge  l_a l_b (l_b-l_a)^k, k<=p-2;  face  l_a l_b l_c (l_b-l_a)^i (l_c-l_a)^j,
//   i+j<=p-3;  cell  l_0 l_1 l_2 l_3 (p=4).
#include "spmv.cuh"
#include "../../include/ngsb200_workloads.h"

#include <algorithm>
#include <array>
#include <map>
#include <tuple>

namespace ngsb {

int csr_adopt_device(ngsb_ctx *ctx, size_t h, size_t w, size_t nnz, uint64_t *d_rowptr, int32_t *d_col, double *d_val, int kind,
                     ngsb_csr **out, bool allow_reorder = true);   // spmv.cu

static const int NBLOCKS = 26;   // 1 vertex + 7 edge + 12 face + 6 cell entity types
static const int MAXTETS = 24;

struct GBlock { uint64_t off; int dim[3]; int mult; int D; };
struct GRowType { int block, sub, tet_begin, ntets, slot_begin, nslots; };
struct GTet { int8_t c[3]; int8_t perm; };
struct GSlot { int16_t block; int8_t d[3]; int8_t sub; int32_t cbegin, cend; };
struct GContrib { int32_t tet; int32_t validx; };

struct GenTables {
    GBlock blocks[NBLOCKS + 1];
    int rt_of_block[NBLOCKS];
    int n[3];
    int nv;                       // doubles per matrix value: 1 real, 2 complex, 9 block3
    int es;                       // doubles per vector entry: 1, 2, 3
    const GRowType *rts;
    const GTet *tets;
    const GSlot *slots;
    const GContrib *contribs;
    const double *values;         // nv per validx
    const double *rhsvals;        // per (row type, tet)
    uint64_t ndof;
};

__host__ __device__ inline void gen_decode(const GenTables &T, uint64_t r, int &b, int &sub, int pos[3])
{
    b = 0;
    while (b + 1 < NBLOCKS && r >= T.blocks[b + 1].off) b++;
    const GBlock &B = T.blocks[b];
    uint64_t local = r - B.off;
    sub = (int)(local % (uint64_t)B.mult);
    uint64_t e = local / (uint64_t)B.mult;
    pos[0] = (int)(e % (uint64_t)B.dim[0]);
    e /= (uint64_t)B.dim[0];
    pos[1] = (int)(e % (uint64_t)B.dim[1]);
    pos[2] = (int)(e / (uint64_t)B.dim[1]);
}

// one matrix row: returns its length; FILL writes columns, values and the load vector entry
template <bool FILL>
__host__ __device__ inline uint32_t gen_row(const GenTables &T, uint64_t r, int32_t *col, double *val, double *rhs)
{
    int b, sub, pos[3];
    gen_decode(T, r, b, sub, pos);
    const GRowType rt = T.rts[T.rt_of_block[b] + sub];
    uint32_t mask = 0;
    for (int t = 0; t < rt.ntets; t++) {
        const GTet tt = T.tets[rt.tet_begin + t];
        bool ok = true;
        for (int a = 0; a < 3; a++) {
            int c = pos[a] + tt.c[a];
            ok = ok && c >= 0 && c < T.n[a];
        }
        if (ok) mask |= 1u << t;
    }
    uint32_t cnt = 0;
    for (int s = 0; s < rt.nslots; s++) {
        const GSlot sl = T.slots[rt.slot_begin + s];
        bool any = false;
        double acc[9];
        if (FILL)
            for (int k = 0; k < 9; k++) acc[k] = 0.0;
        for (int c = sl.cbegin; c < sl.cend; c++) {
            const GContrib cc = T.contribs[c];
            if ((mask >> cc.tet) & 1u) {
                any = true;
                if (FILL)
                    for (int k = 0; k < T.nv; k++) acc[k] += T.values[(size_t)cc.validx * T.nv + k];
            }
        }
        if (any) {
            if (FILL) {
                const GBlock &B = T.blocks[sl.block];
                uint64_t e = ((uint64_t)(pos[2] + sl.d[2]) * B.dim[1] + (uint64_t)(pos[1] + sl.d[1])) * B.dim[0] + (uint64_t)(pos[0] + sl.d[0]);
                col[cnt] = (int32_t)(B.off + e * B.mult + sl.sub);
                for (int k = 0; k < T.nv; k++) val[(size_t)cnt * T.nv + k] = acc[k];
            }
            cnt++;
        }
    }
    if (FILL && rhs) {
        double f = 0.0;
        for (int t = 0; t < rt.ntets; t++)
            if ((mask >> t) & 1u) f += T.rhsvals[rt.tet_begin + t];
        if (T.es == 1) rhs[r] = f;
        else if (T.es == 2) { rhs[2 * r] = f; rhs[2 * r + 1] = 0.0; }
        else { rhs[3 * r] = f; rhs[3 * r + 1] = f; rhs[3 * r + 2] = f; }
    }
    return cnt;
}

__global__ void __launch_bounds__(256) femgen_count_kernel(const GenTables T, uint64_t *rowlen)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < T.ndof) rowlen[r + 1] = gen_row<false>(T, r, nullptr, nullptr, nullptr);
    if (r == 0) rowlen[0] = 0;
}

__global__ void __launch_bounds__(256) femgen_fill_kernel(const GenTables T, const uint64_t *rowptr, int32_t *col, double *val, double *rhs)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= T.ndof) return;
    const uint64_t p = rowptr[r];
    gen_row<true>(T, r, col + p, val + p * T.nv, rhs);
}

// ------------------------------------------------------------------------------------------
// host: polynomials in barycentric coordinates, exact integration
// ------------------------------------------------------------------------------------------
typedef std::array<int, 4> Mono;
typedef std::map<Mono, double> Poly;

static Poly pconst(double c) { Poly p; p[Mono{0, 0, 0, 0}] = c; return p; }
static Poly plam(int k) { Poly p; Mono m{0, 0, 0, 0}; m[k] = 1; p[m] = 1.0; return p; }
static Poly padd(const Poly &a, const Poly &b, double sb = 1.0)
{
    Poly r = a;
    for (auto &t : b) r[t.first] += sb * t.second;
    return r;
}
static Poly pmul(const Poly &a, const Poly &b)
{
    Poly r;
    for (auto &x : a)
        for (auto &y : b) {
            Mono m;
            for (int k = 0; k < 4; k++) m[k] = x.first[k] + y.first[k];
            r[m] += x.second * y.second;
        }
    return r;
}
static Poly ppow(const Poly &a, int n) { Poly r = pconst(1.0); for (int i = 0; i < n; i++) r = pmul(r, a); return r; }
static Poly pderiv(const Poly &a, int k)
{
    Poly r;
    for (auto &x : a)
        if (x.first[k] > 0) { Mono m = x.first; m[k]--; r[m] += x.second * x.first[k]; }
    return r;
}
// integral over the unit-volume tetrahedron: 3! prod e_i! / (sum e + 3)!
static double pint(const Poly &a)
{
    auto fact = [](int n) { double f = 1; for (int i = 2; i <= n; i++) f *= i; return f; };
    double s = 0;
    for (auto &x : a) {
        int tot = 0; double num = 6.0;
        for (int k = 0; k < 4; k++) { tot += x.first[k]; num *= fact(x.first[k]); }
        s += x.second * num / fact(tot + 3);
    }
    return s;
}

struct LocalDof { int etype; int l[4]; int sub; };   // etype 0 vertex,1 edge,2 face,3 cell; l = local vertices

static const int PERMS[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
static const int EDGES[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
static const int FACES[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};

static int face_type_index(int d1, int d2)
{
    int idx = 0;
    for (int a = 1; a < 8; a++)
        for (int b = 1; b < 8; b++)
            if ((a & b) == 0) { if (a == d1 && b == d2) return idx; idx++; }
    return -1;
}

static void make_blocks(int order, const int n[3], GBlock *blocks, uint64_t *total)
{
    const int ne = order - 1, nf = (order - 1) * (order - 2) / 2, nc = (order - 1) * (order - 2) * (order - 3) / 6;
    int Ds[NBLOCKS], mult[NBLOCKS];
    Ds[0] = 0; mult[0] = 1;
    for (int d = 1; d < 8; d++) { Ds[d] = d; mult[d] = ne; }
    for (int a = 1; a < 8; a++)
        for (int b = 1; b < 8; b++)
            if ((a & b) == 0) { int i = 8 + face_type_index(a, b); Ds[i] = a | b; mult[i] = nf; }
    for (int p = 0; p < 6; p++) { Ds[20 + p] = 7; mult[20 + p] = nc; }
    uint64_t off = 0;
    for (int b = 0; b < NBLOCKS; b++) {
        blocks[b].off = off;
        blocks[b].D = Ds[b];
        blocks[b].mult = mult[b] > 0 ? mult[b] : 1;
        uint64_t cnt = 1;
        for (int a = 0; a < 3; a++) {
            blocks[b].dim[a] = n[a] + 1 - ((Ds[b] >> a) & 1);
            cnt *= (uint64_t)blocks[b].dim[a];
        }
        off += mult[b] > 0 ? cnt * (uint64_t)mult[b] : 0;
    }
    blocks[NBLOCKS].off = off;
    blocks[NBLOCKS].mult = 1;
    blocks[NBLOCKS].D = 0;
    for (int a = 0; a < 3; a++) blocks[NBLOCKS].dim[a] = 1;
    *total = off;
}

} // namespace ngsb

using namespace ngsb;

struct ngsb_femgen {
    ngsb_femgen_desc desc;
    GenTables T;                    // host pointers
    GBlock gblocks[NBLOCKS + 1];    // numbering of the global grid
    uint64_t global_ndof = 0;
    int mults[NBLOCKS];             // true dofs per entity (0 for absent blocks)
    std::vector<GRowType> rts;
    std::vector<GTet> tets;
    std::vector<GSlot> slots;
    std::vector<GContrib> contribs;
    std::vector<double> values, rhsvals;
};

static int build_stencils(ngsb_femgen *g)
{
    const ngsb_femgen_desc &d = g->desc;
    const int p = d.order;
    const int ne = p - 1, nf = (p - 1) * (p - 2) / 2, nc = (p - 1) * (p - 2) * (p - 3) / 6;
    // local dofs of one tetrahedron
    std::vector<LocalDof> ld;
    std::vector<Poly> phi;
    for (int v = 0; v < 4; v++) { ld.push_back({0, {v, 0, 0, 0}, 0}); phi.push_back(plam(v)); }
    for (int e = 0; e < 6; e++)
        for (int k = 0; k < ne; k++) {
            int a = EDGES[e][0], b = EDGES[e][1];
            ld.push_back({1, {a, b, 0, 0}, k});
            phi.push_back(pmul(pmul(plam(a), plam(b)), ppow(padd(plam(b), plam(a), -1.0), k)));
        }
    for (int f = 0; f < 4; f++) {
        int sub = 0;
        for (int j = 0; j <= p - 3; j++)
            for (int i = 0; i + j <= p - 3; i++) {
                int a = FACES[f][0], b = FACES[f][1], c = FACES[f][2];
                ld.push_back({2, {a, b, c, 0}, sub++});
                Poly q = pmul(pmul(plam(a), plam(b)), plam(c));
                q = pmul(q, ppow(padd(plam(b), plam(a), -1.0), i));
                q = pmul(q, ppow(padd(plam(c), plam(a), -1.0), j));
                phi.push_back(q);
            }
    }
    for (int k = 0; k < nc; k++) { ld.push_back({3, {0, 1, 2, 3}, k}); phi.push_back(pmul(pmul(plam(0), plam(1)), pmul(plam(2), plam(3)))); }
    const int ndl = (int)ld.size();
    NGSB_REQUIRE(ndl == 4 + 6 * ne + 4 * nf + nc, "femgen: local dof count mismatch");
    // reference integrals
    std::vector<std::array<Poly, 4>> dphi(ndl);
    for (int i = 0; i < ndl; i++)
        for (int k = 0; k < 4; k++) dphi[i][k] = pderiv(phi[i], k);
    std::vector<double> I((size_t)ndl * ndl * 16), M((size_t)ndl * ndl), bl(ndl);
    for (int i = 0; i < ndl; i++) {
        bl[i] = pint(phi[i]);
        for (int j = 0; j < ndl; j++) {
            M[(size_t)i * ndl + j] = pint(pmul(phi[i], phi[j]));
            for (int k = 0; k < 4; k++)
                for (int l = 0; l < 4; l++) I[(((size_t)i * ndl + j) * 4 + k) * 4 + l] = pint(pmul(dphi[i][k], dphi[j][l]));
        }
    }
    const double h = d.h, vol = h * h * h / 6.0;
    const int nv = d.kind == NGSB_REAL ? 1 : (d.kind == NGSB_COMPLEX ? 2 : 9);
    g->T.nv = nv;
    g->T.es = d.kind == NGSB_REAL ? 1 : (d.kind == NGSB_COMPLEX ? 2 : 3);
    g->values.assign((size_t)6 * ndl * ndl * nv, 0.0);
    std::vector<double> bvals((size_t)6 * ndl);
    for (int pi = 0; pi < 6; pi++) {
        double gl[4][3] = {{0}};
        gl[0][PERMS[pi][0]] -= 1.0 / h;
        gl[1][PERMS[pi][0]] += 1.0 / h; gl[1][PERMS[pi][1]] -= 1.0 / h;
        gl[2][PERMS[pi][1]] += 1.0 / h; gl[2][PERMS[pi][2]] -= 1.0 / h;
        gl[3][PERMS[pi][2]] += 1.0 / h;
        for (int i = 0; i < ndl; i++) {
            bvals[(size_t)pi * ndl + i] = vol * bl[i];
            for (int j = 0; j < ndl; j++) {
                double D[3][3];
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++) {
                        double s = 0;
                        for (int k = 0; k < 4; k++)
                            for (int l = 0; l < 4; l++) s += gl[k][a] * gl[l][b] * I[(((size_t)i * ndl + j) * 4 + k) * 4 + l];
                        D[a][b] = vol * s;    // int d_a phi_i d_b phi_j
                    }
                const double K = D[0][0] + D[1][1] + D[2][2], Mij = vol * M[(size_t)i * ndl + j];
                double *out = &g->values[(((size_t)pi * ndl + i) * ndl + j) * nv];
                if (d.kind == NGSB_REAL) out[0] = K + d.mass_re * Mij;
                else if (d.kind == NGSB_COMPLEX) { out[0] = K + d.mass_re * Mij; out[1] = d.mass_im * Mij; }
                else
                    for (int a = 0; a < 3; a++)
                        for (int b = 0; b < 3; b++)   // lambda div u div v + 2 mu eps(u):eps(v); entry (a,b) couples v_a (row) with u_b
                            out[3 * a + b] = d.lame_lambda * D[a][b] + d.lame_mu * D[b][a] + (a == b ? d.lame_mu * K : 0.0);
            }
        }
    }
    // entity description of every local dof per tet type
    auto tet_vmask = [](int pi, int l) { int m = 0; for (int q = 0; q < l; q++) m |= 1 << PERMS[pi][q]; return m; };
    auto dof_entity = [&](int pi, const LocalDof &L, int &block, int &basemask) {
        int m0 = tet_vmask(pi, L.l[0]);
        basemask = m0;
        if (L.etype == 0) block = 0;
        else if (L.etype == 1) block = tet_vmask(pi, L.l[1]) ^ m0;
        else if (L.etype == 2) {
            int m1 = tet_vmask(pi, L.l[1]), m2 = tet_vmask(pi, L.l[2]);
            block = 8 + face_type_index(m1 ^ m0, m2 ^ m1);
        } else { block = 20 + pi; basemask = 0; }
    };
    g->mults[0] = 1;
    for (int b = 1; b < 8; b++) g->mults[b] = ne;
    for (int b = 8; b < 20; b++) g->mults[b] = nf;
    for (int b = 20; b < 26; b++) g->mults[b] = nc;
    int nrt = 0;
    for (int b = 0; b < NBLOCKS; b++) { g->T.rt_of_block[b] = nrt; nrt += g->mults[b]; }
    g->rts.resize(nrt);
    for (int b = 0; b < NBLOCKS; b++)
        for (int sub = 0; sub < g->mults[b]; sub++) {
            GRowType rt;
            rt.block = b; rt.sub = sub;
            rt.tet_begin = (int)g->tets.size();
            rt.slot_begin = (int)g->slots.size();
            typedef std::tuple<int, int, int, int, int> Key;       // block, dz, dy, dx, sub
            std::map<Key, std::vector<GContrib>> smap;
            int nt = 0;
            for (int o = 0; o < 8; o++)
                for (int pi = 0; pi < 6; pi++) {
                    // the row dof must be a local dof of tet (cube = base - o, pi) with base mask o
                    int irow = -1;
                    for (int i = 0; i < ndl; i++) {
                        int bb, bm;
                        dof_entity(pi, ld[i], bb, bm);
                        if (bb == b && bm == o && ld[i].sub == sub) { irow = i; break; }
                    }
                    if (irow < 0) continue;
                    NGSB_REQUIRE(nt < MAXTETS, "femgen: too many incident tetrahedra");
                    GTet tt;
                    for (int a = 0; a < 3; a++) tt.c[a] = (int8_t)(-((o >> a) & 1));
                    tt.perm = (int8_t)pi;
                    g->tets.push_back(tt);
                    g->rhsvals.push_back(bvals[(size_t)pi * ndl + irow]);
                    for (int j = 0; j < ndl; j++) {
                        int bb, bm;
                        dof_entity(pi, ld[j], bb, bm);
                        Key key(bb, ((bm >> 2) & 1) - ((o >> 2) & 1), ((bm >> 1) & 1) - ((o >> 1) & 1), (bm & 1) - (o & 1), ld[j].sub);
                        smap[key].push_back(GContrib{nt, (pi * ndl + irow) * ndl + j});
                    }
                    nt++;
                }
            rt.ntets = nt;
            for (auto &kv : smap) {
                GSlot sl;
                sl.block = (int16_t)std::get<0>(kv.first);
                sl.d[2] = (int8_t)std::get<1>(kv.first);
                sl.d[1] = (int8_t)std::get<2>(kv.first);
                sl.d[0] = (int8_t)std::get<3>(kv.first);
                sl.sub = (int8_t)std::get<4>(kv.first);
                sl.cbegin = (int32_t)g->contribs.size();
                for (auto &c : kv.second) g->contribs.push_back(c);
                sl.cend = (int32_t)g->contribs.size();
                g->slots.push_back(sl);
            }
            rt.nslots = (int)g->slots.size() - rt.slot_begin;
            g->rts[g->T.rt_of_block[b] + sub] = rt;
        }
    return NGSB_OK;
}

extern "C" int ngsb_femgen_create(const ngsb_femgen_desc *desc, ngsb_femgen **out)
{
    NGSB_REQUIRE(desc && out, "ngsb_femgen_create: NULL argument");
    NGSB_REQUIRE(desc->order >= 1 && desc->order <= 4, "ngsb_femgen_create: order must be 1..4");
    NGSB_REQUIRE(kind_valid(desc->kind), "ngsb_femgen_create: bad kind");
    for (int a = 0; a < 3; a++) {
        NGSB_REQUIRE(desc->n[a] >= 1 && desc->global[a] >= desc->n[a] && desc->offset[a] >= 0 &&
                     desc->offset[a] + desc->n[a] <= desc->global[a], "ngsb_femgen_create: box does not fit the global grid");
    }
    NGSB_REQUIRE(desc->h > 0, "ngsb_femgen_create: h must be positive");
    ngsb_femgen *g = new ngsb_femgen();
    g->desc = *desc;
    uint64_t total = 0;
    make_blocks(desc->order, desc->n, g->T.blocks, &total);
    g->T.ndof = total;
    for (int a = 0; a < 3; a++) g->T.n[a] = desc->n[a];
    make_blocks(desc->order, desc->global, g->gblocks, &g->global_ndof);
    if (total >= (1ull << 31)) { delete g; set_error("ngsb_femgen_create: %llu dofs exceed 32-bit column indices", (unsigned long long)total); return NGSB_ERR_INVALID; }
    int rc = build_stencils(g);
    if (rc != NGSB_OK) { delete g; return rc; }
    g->T.rts = g->rts.data();
    g->T.tets = g->tets.data();
    g->T.slots = g->slots.data();
    g->T.contribs = g->contribs.data();
    g->T.values = g->values.data();
    g->T.rhsvals = g->rhsvals.data();
    *out = g;
    return NGSB_OK;
}

extern "C" int ngsb_femgen_destroy(ngsb_femgen *g)
{
    delete g;
    return NGSB_OK;
}

extern "C" int ngsb_femgen_sizes(const ngsb_femgen *g, size_t *ndof, size_t *global_ndof)
{
    NGSB_REQUIRE(g, "ngsb_femgen_sizes: NULL argument");
    if (ndof) *ndof = g->T.ndof;
    if (global_ndof) *global_ndof = g->global_ndof;
    return NGSB_OK;
}

extern "C" int ngsb_femgen_dof_info(const ngsb_femgen *g, uint64_t *global_index, uint8_t *on_box_surface, uint8_t *is_free)
{
    NGSB_REQUIRE(g, "ngsb_femgen_dof_info: NULL argument");
    const ngsb_femgen_desc &d = g->desc;
    const int64_t nd = (int64_t)g->T.ndof;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nd; r++) {
        int b, sub, pos[3];
        gen_decode(g->T, (uint64_t)r, b, sub, pos);
        const int D = g->T.blocks[b].D;
        bool surf = false, fixed = false;
        uint64_t e = 0;
        const GBlock &G = g->gblocks[b];
        for (int a = 2; a >= 0; a--) {
            const int gp = pos[a] + d.offset[a];
            if (!((D >> a) & 1)) {
                if (pos[a] == 0 || pos[a] == d.n[a]) surf = true;
                if (gp == 0 || gp == d.global[a]) fixed = true;
            }
            e = e * (uint64_t)G.dim[a] + (uint64_t)gp;
        }
        if (global_index) global_index[r] = G.off + e * (uint64_t)G.mult + (uint64_t)sub;
        if (on_box_surface) on_box_surface[r] = surf ? 1 : 0;
        if (is_free) is_free[r] = fixed ? 0 : 1;
    }
    return NGSB_OK;
}

extern "C" int ngsb_femgen_host(const ngsb_femgen *g, uint64_t *rowptr, int32_t *col, void *val, void *rhs)
{
    NGSB_REQUIRE(g && rowptr, "ngsb_femgen_host: NULL argument");
    const int64_t nd = (int64_t)g->T.ndof;
    if (!col) {
        rowptr[0] = 0;
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < nd; r++) rowptr[r + 1] = gen_row<false>(g->T, (uint64_t)r, nullptr, nullptr, nullptr);
        for (int64_t r = 0; r < nd; r++) rowptr[r + 1] += rowptr[r];
        return NGSB_OK;
    }
    NGSB_REQUIRE(val, "ngsb_femgen_host: val is NULL");
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nd; r++)
        gen_row<true>(g->T, (uint64_t)r, col + rowptr[r], (double *)val + rowptr[r] * g->T.nv, (double *)rhs);
    return NGSB_OK;
}

extern "C" int ngsb_femgen_device(ngsb_ctx *ctx, const ngsb_femgen *g, ngsb_csr **A, ngsb_vec **rhs)
{
    NGSB_REQUIRE(ctx && g && A, "ngsb_femgen_device: NULL argument");
    NGSB_CUDA(cudaSetDevice(ctx->device));
    GenTables T = g->T;
    void *d_rts = nullptr, *d_tets = nullptr, *d_slots = nullptr, *d_contribs = nullptr, *d_values = nullptr, *d_rhsvals = nullptr;
    uint64_t *d_rowptr = nullptr;
    int32_t *d_col = nullptr;
    double *d_val = nullptr;
    ngsb_vec *fv = nullptr;
    int rc = NGSB_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == NGSB_OK) { set_error("ngsb_femgen_device: %s: %s", what, cudaGetErrorString(e)); rc = e == cudaErrorMemoryAllocation ? NGSB_ERR_NOMEM : NGSB_ERR_CUDA; }
        return rc != NGSB_OK;
    };
    auto up = [&](void **dst, const void *src, size_t bytes) {
        if (fail(cudaMalloc(dst, bytes ? bytes : 16), "cudaMalloc(table)")) return;
        fail(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream), "upload table");
    };
    up(&d_rts, g->rts.data(), g->rts.size() * sizeof(GRowType));
    up(&d_tets, g->tets.data(), g->tets.size() * sizeof(GTet));
    up(&d_slots, g->slots.data(), g->slots.size() * sizeof(GSlot));
    up(&d_contribs, g->contribs.data(), g->contribs.size() * sizeof(GContrib));
    up(&d_values, g->values.data(), g->values.size() * sizeof(double));
    up(&d_rhsvals, g->rhsvals.data(), g->rhsvals.size() * sizeof(double));
    T.rts = (const GRowType *)d_rts; T.tets = (const GTet *)d_tets; T.slots = (const GSlot *)d_slots;
    T.contribs = (const GContrib *)d_contribs; T.values = (const double *)d_values; T.rhsvals = (const double *)d_rhsvals;
    const uint64_t nd = T.ndof;
    const unsigned grid = (unsigned)((nd + 255) / 256);
    uint64_t nnz = 0;
    if (rc == NGSB_OK) fail(cudaMalloc(&d_rowptr, (nd + 1) * sizeof(uint64_t)), "cudaMalloc(rowptr)");
    if (rc == NGSB_OK) {
        femgen_count_kernel<<<grid, 256, 0, ctx->stream>>>(T, d_rowptr);
        ctx->launches++;
        fail(cudaGetLastError(), "count launch");
        if (rc == NGSB_OK) rc = device_scan_u64(ctx, d_rowptr, nd + 1);
        fail(cudaMemcpyAsync(&nnz, d_rowptr + nd, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream), "read nnz");
        fail(cudaStreamSynchronize(ctx->stream), "count/scan");
    }
    const size_t slack = 16;
    if (rc == NGSB_OK) fail(cudaMalloc(&d_col, (nnz + slack) * sizeof(int32_t)), "cudaMalloc(col)");
    if (rc == NGSB_OK) fail(cudaMalloc(&d_val, (nnz + slack) * T.nv * sizeof(double)), "cudaMalloc(val)");
    if (rc == NGSB_OK && rhs) rc = ngsb_vec_create(ctx, nd, g->desc.kind, &fv);
    if (rc == NGSB_OK) {
        cudaMemsetAsync(d_col + nnz, 0, slack * sizeof(int32_t), ctx->stream);
        cudaMemsetAsync(d_val + nnz * T.nv, 0, slack * T.nv * sizeof(double), ctx->stream);
        femgen_fill_kernel<<<grid, 256, 0, ctx->stream>>>(T, d_rowptr, d_col, d_val, fv ? fv->d : nullptr);
        ctx->launches++;
        fail(cudaGetLastError(), "fill launch");
        fail(cudaStreamSynchronize(ctx->stream), "fill");
    }
    cudaFree(d_rts); cudaFree(d_tets); cudaFree(d_slots); cudaFree(d_contribs); cudaFree(d_values); cudaFree(d_rhsvals);
    if (rc == NGSB_OK) {
        rc = csr_adopt_device(ctx, nd, nd, nnz, d_rowptr, d_col, d_val, g->desc.kind, A);
        if (rc == NGSB_OK) { d_rowptr = nullptr; d_col = nullptr; d_val = nullptr; }
    }
    if (rc != NGSB_OK) {
        cudaFree(d_rowptr); cudaFree(d_col); cudaFree(d_val);
        if (fv) ngsb_vec_destroy(fv);
        return rc;
    }
    if (rhs) *rhs = fv;
    return NGSB_OK;
}
