// ngsb200_ngla.cpp -- NGSolve-side adapter: binds libngsb200 (include/ngsb200.h) behind ngla's
// BaseVector / BaseMatrix interface and registers it in the device-creator registries, exactly where
// ngscuda plugs in (ngscuda/cuda_linalg.cpp:85-115, ngscuda/python_ngscuda.cpp:20-25).
//
// Build (inside an NGSolve build tree / against an installed NGSolve):
//   g++ -O2 -std=c++20 -shared -fPIC $(python -m pybind11 --includes) -I$NGSOLVE/include \
//       -I<repo>/include ngsb200_ngla.cpp -L$NGSOLVE/lib -lngla -lngstd -lngcore \
//       -L<repo>/ngsolve_b200/lib -lngsb200 -o _ngsb200$(python3-config --extension-suffix)
// After `import ngsolve.ngsb200` (or `import _ngsb200`) unchanged scripts run on the B200 path:
//   fdev = f.vec.CreateDeviceVector(); adev = a.mat.CreateDeviceMatrix(); jdev = jac.CreateDeviceMatrix()
//   inv = CGSolver(adev, jdev, maxsteps=2000); res = (inv * fdev).Evaluate()
//
// NOT compiled in this repository's CI: NGSolve itself is not available on the GPU box.
#include <la.hpp>
#include <pybind11/pybind11.h>

#include "ngsb200.h"

namespace ngla
{
  using namespace ngbla;
  using namespace ngcore;

  static ngsb_ctx * TheCtx ()
  {
    static ngsb_ctx * ctx = [] {
      ngsb_ctx * c = nullptr;
      if (ngsb_ctx_create (-1, &c) != NGSB_OK) throw Exception (ngsb_last_error());   // reads NGS_CUDA_DEVICE_INDEX
      return c;
    }();
    return ctx;
  }
  static void Check (int rc) { if (rc != NGSB_OK) throw Exception (ngsb_last_error()); }

  template <typename SCAL> constexpr int KindOf (int es)
  { return std::is_same_v<SCAL,Complex> ? NGSB_COMPLEX : (es == 3 ? NGSB_BLOCK3 : NGSB_REAL); }

  // BaseScalar kept on the device (linalg/basescalar.hpp:17-29; ngscuda's UnifiedScalar, ngscuda/unifiedvector.hpp:113-136):
  // a solver written with vec.InnerProduct(other, scal) / vec.Add(scal, other) never brings its scalars to the host
  class B200Scalar : public BaseScalar
  {
    ngsb_scalar * h = nullptr;
    bool cplx;
  public:
    B200Scalar (bool acplx) : cplx(acplx) { Check (ngsb_scalar_create (TheCtx(), &h)); }
    ~B200Scalar () { ngsb_scalar_destroy (h); }
    ngsb_scalar * Handle () const { return h; }
    bool IsComplex () const override { return cplx; }
    void Set (double d) override { double z[2] = {d,0}; cplx = false; Check (ngsb_scalar_set (h, z)); }
    void Set (Complex c) override { double z[2] = {c.real(),c.imag()}; cplx = true; Check (ngsb_scalar_set (h, z)); }
    double GetD () const override { double z[2]; Check (ngsb_scalar_get (h, z)); return z[0]; }        // synchronises
    Complex GetC () const override { double z[2]; Check (ngsb_scalar_get (h, z)); return Complex (z[0], z[1]); }
    static B200Scalar & Cast (BaseScalar & s)
    {
      auto p = dynamic_cast<B200Scalar*> (&s);
      if (!p) throw Exception ("B200Vector: scalar is not a device scalar (use vec.CreateScalar())");
      return *p;
    }
  };

  // ---- vector: host mirror + dirty flags like UnifiedVector (ngscuda/unifiedvector.hpp:8-98) ----------
  template <typename SCAL>
  class B200Vector : public S_BaseVector<SCAL>
  {
    ngsb_vec * dev = nullptr;
    mutable Array<SCAL> host;
    mutable bool host_uptodate = false, dev_uptodate = true;
    shared_ptr<BaseVector> parent;   // keeps the storage of a Range() view alive
  public:
    B200Vector (size_t asize, int aes = 1)
    {
      this->size = asize; this->entrysize = aes * (std::is_same_v<SCAL,Complex> ? 2 : 1);
      Check (ngsb_vec_create (TheCtx(), asize, KindOf<SCAL>(aes), &dev));
    }
    B200Vector (const BaseVector & v) : B200Vector (v.Size(), v.EntrySize() / (v.IsComplex() ? 2 : 1)) { *this = v; }
    B200Vector (ngsb_vec * view, size_t asize, int aes, shared_ptr<BaseVector> aparent) : dev(view), parent(aparent)
    { this->size = asize; this->entrysize = aes; }
    ~B200Vector () { ngsb_vec_destroy (dev); }

    ngsb_vec * Dev () const { UpdateDevice(); return dev; }
    ngsb_vec * DevW () { UpdateDevice(); host_uptodate = false; return dev; }
    void UpdateDevice () const
    {
      if (dev_uptodate) return;
      Check (ngsb_vec_h2d (dev, host.Data(), 0, this->size));
      dev_uptodate = true;
    }
    void UpdateHost () const
    {
      if (host.Size() != this->size * size_t(EntryScalars())) host.SetSize (this->size * EntryScalars());
      if (host_uptodate) return;
      UpdateDevice ();
      Check (ngsb_vec_d2h (dev, host.Data(), 0, this->size));
      host_uptodate = true;
    }
    int EntryScalars () const { return this->entrysize / (std::is_same_v<SCAL,Complex> ? 2 : 1); }

    // host access: FVDouble()/FVComplex()/Memory() must hand out coherent host memory
    void * Memory () const throw () override { UpdateHost(); dev_uptodate = false; return host.Data(); }
    FlatVector<double> FVDouble () const override
    { UpdateHost(); dev_uptodate = false; return FlatVector<double> (this->size * this->entrysize, (double*)host.Data()); }
    FlatVector<Complex> FVComplex () const override
    {
      if constexpr (!std::is_same_v<SCAL,Complex>) throw Exception ("FVComplex called for real B200Vector");
      UpdateHost(); dev_uptodate = false; return FlatVector<Complex> (this->size * EntryScalars(), (Complex*)host.Data());
    }

    // a read-only operand: a device vector as it is, a host vector through a temporary upload (what
    // UnifiedVectorWrapper does, ngscuda/unifiedvector.hpp:100-136; `*dev = hostvec` in CreateDeviceVector(copy=True),
    // linalg/python_linalg.cpp:656-662, arrives here as Set(1.0, hostvec))
    struct Operand
    {
      const B200Vector * p;
      std::unique_ptr<B200Vector> tmp;
      ngsb_vec * Dev () const { return p->Dev(); }
    };
    static Operand Cast (const BaseVector & v)
    {
      if (auto p = dynamic_cast<const B200Vector*> (&v)) return Operand { p, nullptr };
      if (v.IsComplex() != std::is_same_v<SCAL,Complex>) throw Exception ("B200Vector: real/complex mismatch of operands");
      auto t = std::make_unique<B200Vector> (v.Size(), v.EntrySize() / (v.IsComplex() ? 2 : 1));
      Check (ngsb_vec_h2d (t->dev, v.Memory(), 0, v.Size()));
      t->host_uptodate = false; t->dev_uptodate = true;
      const B200Vector * q = t.get();
      return Operand { q, std::move(t) };
    }
    // a written operand must live on the device
    static B200Vector & CastW (BaseVector & v)
    {
      auto p = dynamic_cast<B200Vector*> (&v);
      if (!p) throw Exception ("B200Vector: result vector is not a device vector (use CreateDeviceVector / CreateColVector of the device matrix)");
      return *p;
    }

    BaseVector & SetScalar (double s) override { double z[2] = {s,0}; host_uptodate = false; dev_uptodate = true; Check (ngsb_vec_set_scalar (dev, z)); return *this; }
    BaseVector & Scale (double s) override { double z[2] = {s,0}; Check (ngsb_vec_scale (DevW(), z)); return *this; }
    BaseVector & Scale (Complex s) override { double z[2] = {s.real(),s.imag()}; Check (ngsb_vec_scale (DevW(), z)); return *this; }
    BaseVector & Set (double s, const BaseVector & v) override { double z[2] = {s,0}; Check (ngsb_vec_set (DevW(), z, Cast(v).Dev())); return *this; }
    BaseVector & Set (Complex s, const BaseVector & v) override { double z[2] = {s.real(),s.imag()}; Check (ngsb_vec_set (DevW(), z, Cast(v).Dev())); return *this; }
    BaseVector & Add (double s, const BaseVector & v) override { double z[2] = {s,0}; Check (ngsb_vec_axpy (DevW(), z, Cast(v).Dev())); return *this; }
    BaseVector & Add (Complex s, const BaseVector & v) override { double z[2] = {s.real(),s.imag()}; Check (ngsb_vec_axpy (DevW(), z, Cast(v).Dev())); return *this; }
    double InnerProductD (const BaseVector & v2) const override
    { double out[2]; Check (ngsb_vec_dot (Dev(), Cast(v2).Dev(), 0, out)); return out[0]; }
    Complex InnerProductC (const BaseVector & v2, bool conjugate = false) const override
    { double out[2]; Check (ngsb_vec_dot (Dev(), Cast(v2).Dev(), conjugate, out)); return Complex (out[0], out[1]); }
    double L2Norm () const override { double n; Check (ngsb_vec_nrm2 (Dev(), &n)); return n; }
    // the scalar-by-reference variants (linalg/basevector.cpp:259-298): result / factor stay in device memory
    void InnerProduct (const BaseVector & v2, BaseScalar & scal, bool conjugate = false) const override
    { Check (ngsb_vec_dot_dev (Dev(), Cast(v2).Dev(), conjugate, B200Scalar::Cast(scal).Handle())); }
    BaseVector & Scale (BaseScalar & scal) override
    { Check (ngsb_vec_scale_dev (DevW(), B200Scalar::Cast(scal).Handle())); return *this; }
    BaseVector & Add (BaseScalar & scal, const BaseVector & v) override
    { Check (ngsb_vec_axpy_dev (DevW(), B200Scalar::Cast(scal).Handle(), Cast(v).Dev())); return *this; }
    shared_ptr<BaseScalar> CreateScalar () const override { return make_shared<B200Scalar> (std::is_same_v<SCAL,Complex>); }
    AutoVector CreateVector () const override { return make_unique<B200Vector<SCAL>> (this->size, EntryScalars()); }
    AutoVector Range (T_Range<size_t> r) const override
    {
      ngsb_vec * view; Check (ngsb_vec_range (Dev(), r.First(), r.Next(), &view));
      return make_unique<B200Vector<SCAL>> (view, r.Size(), this->entrysize,
                                            const_cast<B200Vector*>(this)->shared_from_this());
    }
    BaseVector & operator= (const BaseVector & v)
    {
      if (auto p = dynamic_cast<const B200Vector*> (&v)) return Set (1.0, *p);
      Check (ngsb_vec_h2d (dev, v.Memory(), 0, this->size));      // host vector -> device (H2D)
      host_uptodate = false; dev_uptodate = true;
      return *this;
    }
  };

  // ---- sparse matrix: replaces DevSparseMatrix for double, Complex and Mat<3,3,double> --------------
  template <typename TM>
  class B200SparseMatrix : public BaseMatrix
  {
    ngsb_csr * A = nullptr;
    size_t h, w;
    static constexpr bool cplx = std::is_same_v<TM,Complex>;
    static constexpr int es = std::is_same_v<TM,Mat<3,3,double>> ? 3 : 1;
    using SCAL = std::conditional_t<cplx,Complex,double>;
  public:
    B200SparseMatrix (const SparseMatrix<TM> & mat) : h(mat.Height()), w(mat.Width())
    {
      // exactly what SparseMatrix::CSR() hands to Python (linalg/python_linalg.cpp:121-138)
      Check (ngsb_csr_create (TheCtx(), h, w, mat.NZE(), (const uint64_t*)mat.GetFirstArray().Data(),
                              mat.GetColIndices().Data(), mat.GetValues().Data(),
                              cplx ? NGSB_COMPLEX : (es == 3 ? NGSB_BLOCK3 : NGSB_REAL), &A));
    }
    ~B200SparseMatrix () { ngsb_csr_destroy (A); }
    ngsb_csr * Handle () const { return A; }
    bool IsComplex () const override { return cplx; }
    int VHeight () const override { return h; }
    int VWidth () const override { return w; }
    AutoVector CreateRowVector () const override { return make_unique<B200Vector<SCAL>> (w, es); }
    AutoVector CreateColVector () const override { return make_unique<B200Vector<SCAL>> (h, es); }
    void Mult (const BaseVector & x, BaseVector & y) const override
    { Check (ngsb_csr_mult (A, B200Vector<SCAL>::Cast(x).Dev(), B200Vector<SCAL>::CastW(y).DevW())); }
    void MultAdd (double s, const BaseVector & x, BaseVector & y) const override
    { double z[2] = {s,0}; Check (ngsb_csr_multadd (A, z, B200Vector<SCAL>::Cast(x).Dev(), B200Vector<SCAL>::CastW(y).DevW())); }
    void MultAdd (Complex s, const BaseVector & x, BaseVector & y) const override
    { double z[2] = {s.real(),s.imag()}; Check (ngsb_csr_multadd (A, z, B200Vector<SCAL>::Cast(x).Dev(), B200Vector<SCAL>::CastW(y).DevW())); }
    void MultTransAdd (double s, const BaseVector & x, BaseVector & y) const override
    { double z[2] = {s,0}; Check (ngsb_csr_multtransadd (A, z, B200Vector<SCAL>::Cast(x).Dev(), B200Vector<SCAL>::CastW(y).DevW())); }
  };

  // ---- Jacobi: replaces DevDiagonalMatrix built from JacobiPrecond<TM> (keeps the freedofs mask) ------
  template <typename TM>
  class B200Jacobi : public BaseMatrix
  {
    ngsb_jacobi * J = nullptr;
    size_t n;
    static constexpr bool cplx = std::is_same_v<TM,Complex>;
    static constexpr int es = std::is_same_v<TM,Mat<3,3,double>> ? 3 : 1;
    using SCAL = std::conditional_t<cplx,Complex,double>;
  public:
    B200Jacobi (const JacobiPrecond<TM> & jac) : n(jac.Height())
    {
      auto inv = jac.GetInverse();                              // Array<TM> invdiag (linalg/jacobi.hpp)
      // `inner` (the freedofs BitArray) is a protected member without getter: reach it through a derived-class
      // member pointer, so that the masked apply is exactly JacobiPrecond::MultAdd (linalg/jacobi.cpp:71-109)
      struct Access : JacobiPrecond<TM> { static auto Ptr () { return &Access::inner; } };
      shared_ptr<BitArray> inner = jac.*(Access::Ptr());
      Check (ngsb_jacobi_create (TheCtx(), n, inv.Data(), cplx ? NGSB_COMPLEX : (es == 3 ? NGSB_BLOCK3 : NGSB_REAL),
                                 inner ? (const uint8_t*)inner->Data() : nullptr, &J));
    }
    ~B200Jacobi () { ngsb_jacobi_destroy (J); }
    ngsb_jacobi * Handle () const { return J; }
    bool IsComplex () const override { return cplx; }
    int VHeight () const override { return n; }
    int VWidth () const override { return n; }
    AutoVector CreateRowVector () const override { return make_unique<B200Vector<SCAL>> (n, es); }
    AutoVector CreateColVector () const override { return make_unique<B200Vector<SCAL>> (n, es); }
    void Mult (const BaseVector & x, BaseVector & y) const override
    { Check (ngsb_jacobi_mult (J, B200Vector<SCAL>::Cast(x).Dev(), B200Vector<SCAL>::CastW(y).DevW())); }
    void MultAdd (double s, const BaseVector & x, BaseVector & y) const override
    { double z[2] = {s,0}; Check (ngsb_jacobi_multadd (J, z, B200Vector<SCAL>::Cast(x).Dev(), B200Vector<SCAL>::CastW(y).DevW())); }
  };

  // ---- block Jacobi: replaces DevBlockJacobiMatrix (ngscuda/dev_blockjacobi.cpp:21-140) --------------------
  class B200BlockJacobi : public BaseMatrix
  {
    ngsb_blockjacobi * J = nullptr;
    size_t n;
  public:
    B200BlockJacobi (const BlockJacobiPrecond<double> & bj) : n(bj.Height())
    {
      auto table = bj.GetBlockTable();                               // shared_ptr<Table<int>>
      const Array<FlatMatrix<double>> & inverses = bj.GetInverses();  // row-major blocks, already inverted on the host
      Array<uint64_t> first(table->Size()+1);
      first[0] = 0;
      size_t entries = 0;
      for (size_t b = 0; b < table->Size(); b++)
        { first[b+1] = first[b] + (*table)[b].Size(); entries += sqr (size_t((*table)[b].Size())); }
      Array<double> flat(entries);
      size_t off = 0;
      for (size_t b = 0; b < table->Size(); b++)
        {
          size_t bs = (*table)[b].Size();
          for (size_t r = 0; r < bs; r++)
            for (size_t c = 0; c < bs; c++)
              flat[off + r*bs + c] = inverses[b](r,c);
          off += bs*bs;
        }
      Check (ngsb_blockjacobi_create_from_inverses (TheCtx(), n, table->Size(), first.Data(), table->AsArray().Data(),
                                                    flat.Data(), &J));
    }
    ~B200BlockJacobi () { ngsb_blockjacobi_destroy (J); }
    int VHeight () const override { return n; }
    int VWidth () const override { return n; }
    AutoVector CreateRowVector () const override { return make_unique<B200Vector<double>> (n, 1); }
    AutoVector CreateColVector () const override { return make_unique<B200Vector<double>> (n, 1); }
    void MultAdd (double s, const BaseVector & x, BaseVector & y) const override
    { Check (ngsb_blockjacobi_multadd (J, s, B200Vector<double>::Cast(x).Dev(), B200Vector<double>::CastW(y).DevW(), 0)); }
    void MultTransAdd (double s, const BaseVector & x, BaseVector & y) const override
    { Check (ngsb_blockjacobi_multadd (J, s, B200Vector<double>::Cast(x).Dev(), B200Vector<double>::CastW(y).DevW(), 1)); }
    void Mult (const BaseVector & x, BaseVector & y) const override
    { Check (ngsb_blockjacobi_mult (J, B200Vector<double>::Cast(x).Dev(), B200Vector<double>::CastW(y).DevW(), 0)); }
  };

  // ---- fused device CG: same interface as DevCGSolver (ngscuda/cuda_linalg.hpp:289-310) ----------------
  class B200CGSolver : public BaseMatrix
  {
    shared_ptr<BaseMatrix> a, c;
    int maxsteps; double prec; mutable int steps = 0;
  public:
    B200CGSolver (shared_ptr<BaseMatrix> aa, shared_ptr<BaseMatrix> ac, int amaxsteps, double aprec)
      : a(aa), c(ac), maxsteps(amaxsteps), prec(aprec) { }
    int VHeight () const override { return a->VWidth(); }
    int VWidth () const override { return a->VHeight(); }
    int GetSteps () const { return steps; }
    AutoVector CreateRowVector () const override { return a->CreateColVector(); }
    AutoVector CreateColVector () const override { return a->CreateRowVector(); }
    void Mult (const BaseVector & f, BaseVector & u) const override
    {
      auto A = dynamic_cast<const B200SparseMatrix<double>*> (a.get());
      auto C = dynamic_cast<const B200Jacobi<double>*> (c.get());
      if (!A) throw Exception ("B200CGSolver: matrix is not a B200SparseMatrix<double>");
      Check (ngsb_cg_solve (A->Handle(), C ? C->Handle() : nullptr, B200Vector<double>::Cast(f).Dev(),
                            B200Vector<double>::CastW(u).DevW(), prec, maxsteps,
                            NGSB_IP_REAL, 1, &steps, nullptr, 0, nullptr));
    }
  };

  template <typename TM> static void RegisterFor ()
  {
    BaseMatrix::RegisterDeviceMatrixCreator (typeid(SparseMatrix<TM>), [] (const BaseMatrix & m) -> shared_ptr<BaseMatrix>
      { return make_shared<B200SparseMatrix<TM>> (dynamic_cast<const SparseMatrix<TM>&> (m)); });
    BaseMatrix::RegisterDeviceMatrixCreator (typeid(JacobiPrecond<TM>), [] (const BaseMatrix & m) -> shared_ptr<BaseMatrix>
      { return make_shared<B200Jacobi<TM>> (dynamic_cast<const JacobiPrecond<TM>&> (m)); });
  }

  void InitNgsB200 ()
  {
    TheCtx ();
    auto dvec = [] (const BaseVector & v, bool) -> shared_ptr<BaseVector>
      {
        if (v.IsComplex()) return make_shared<B200Vector<Complex>> (v);
        return make_shared<B200Vector<double>> (v);
      };
    // the exact dynamic types that reach the registry
    BaseVector::RegisterDeviceVectorCreator (typeid(VVector<double>), dvec);
    BaseVector::RegisterDeviceVectorCreator (typeid(VVector<Complex>), dvec);
    BaseVector::RegisterDeviceVectorCreator (typeid(VVector<Vec<3,double>>), dvec);
    BaseVector::RegisterDeviceVectorCreator (typeid(S_BaseVectorPtr<double>), dvec);
    BaseVector::RegisterDeviceVectorCreator (typeid(S_BaseVectorPtr<Complex>), dvec);
    BaseMatrix::RegisterDeviceMatrixCreator (typeid(BlockJacobiPrecond<double>), [] (const BaseMatrix & m) -> shared_ptr<BaseMatrix>
      { return make_shared<B200BlockJacobi> (dynamic_cast<const BlockJacobiPrecond<double>&> (m)); });
    RegisterFor<double> ();
    RegisterFor<Complex> ();
    RegisterFor<Mat<3,3,double>> ();
  }
}

PYBIND11_MODULE(_ngsb200, m)
{
  namespace py = pybind11;
  ngla::InitNgsB200 ();        // registration at import, like _ngscuda
  py::class_<ngla::B200CGSolver, std::shared_ptr<ngla::B200CGSolver>, ngla::BaseMatrix> (m, "DevCGSolver")
    .def (py::init<std::shared_ptr<ngla::BaseMatrix>, std::shared_ptr<ngla::BaseMatrix>, int, double> (),
          py::arg("mat"), py::arg("pre"), py::arg("maxsteps") = 200, py::arg("precision") = 1e-8)
    .def ("GetSteps", &ngla::B200CGSolver::GetSteps);
}
