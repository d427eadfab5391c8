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
// Compiled by integration/build_adapter.sh against the reference build (oracle/_ref/ngs, which travels to the GPU box) and
// exercised by tests/test_gpu_dropin.py: an unchanged NGSolve script, the reference's own python/krylovspace.py solvers and the
// C++ CGSolver / GMRESSolver factories all run on this layer.
#include <la.hpp>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

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
    // Range() views alias the parent's DEVICE storage but have a host mirror of their own (the reference's
    // UnifiedVectorWrapper does the same and keeps the two coherent by UpdateDevice + InvalidateHost of the wrapped vector,
    // ngscuda/unifiedvector.cpp:376-377).  Here a parent knows its live views and a view its parent, and every access
    // keeps the aliases coherent in both directions:
    //   before an object touches the device copy, pending host writes of its aliases are flushed (in program order);
    //   a device write through one alias invalidates the host mirrors of the others.
    mutable std::vector<const B200Vector*> views;
    const B200Vector * vparent = nullptr;

    void FlushOwn () const
    {
      if (dev_uptodate) return;
      Check (ngsb_vec_h2d (dev, host.Data(), 0, this->size));
      dev_uptodate = true;
    }
    template <typename F> void Ancestors (F f) const
    {
      std::vector<const B200Vector*> chain;
      for (auto p = vparent; p; p = p->vparent) chain.push_back (p);
      for (auto it = chain.rbegin(); it != chain.rend(); ++it) f (*it);          // root first
    }
    template <typename F> void Descendants (F f) const { for (auto v : views) { f (v); v->Descendants (f); } }
    // pending host writes reach the device in program order: ancestors (older), this object, its views
    void FlushAll (bool own) const
    {
      Ancestors ([] (auto p) { p->FlushOwn(); });
      if (own) FlushOwn ();
      Descendants ([] (auto v) { v->FlushOwn(); });
    }
    void InvalidateOthersHost () const
    {
      Ancestors ([] (auto p) { p->host_uptodate = false; });
      Descendants ([] (auto v) { v->host_uptodate = false; });
    }
    // coherent host memory that the caller may write (FVDouble / FVComplex / Memory)
    void HostAccess () const
    {
      if (host.Size() != this->size * size_t(EntryScalars())) host.SetSize (this->size * EntryScalars());
      FlushAll (false);
      if (!host_uptodate)
        {
          Check (ngsb_vec_d2h (dev, host.Data(), 0, this->size));
          host_uptodate = true;
        }
      InvalidateOthersHost ();
      dev_uptodate = false;
    }
  public:
    B200Vector (size_t asize, int aes = 1)
    {
      this->size = asize; this->entrysize = aes * (std::is_same_v<SCAL,Complex> ? 2 : 1);
      Check (ngsb_vec_create (TheCtx(), asize, KindOf<SCAL>(aes), &dev));
    }
    B200Vector (const BaseVector & v) : B200Vector (v.Size(), v.EntrySize() / (v.IsComplex() ? 2 : 1)) { *this = v; }
    B200Vector (ngsb_vec * view, size_t asize, int aes, shared_ptr<BaseVector> aparent, const B200Vector * avparent)
      : dev(view), parent(aparent), vparent(avparent)
    { this->size = asize; this->entrysize = aes; vparent->views.push_back (this); }
    ~B200Vector ()
    {
      if (vparent)
        {
          try { FlushAll (true); } catch (...) { }       // a dying view must not lose its host writes
          auto & l = vparent->views;
          l.erase (std::remove (l.begin(), l.end(), this), l.end());
        }
      ngsb_vec_destroy (dev);
    }

    ngsb_vec * Dev () const { FlushAll (true); return dev; }
    ngsb_vec * DevW () { FlushAll (true); host_uptodate = false; InvalidateOthersHost(); return dev; }
    void UpdateDevice () const { FlushAll (true); }
    int EntryScalars () const { return this->entrysize / (std::is_same_v<SCAL,Complex> ? 2 : 1); }

    // host access: FVDouble()/FVComplex()/Memory() must hand out coherent host memory
    void * Memory () const throw () override { HostAccess(); return host.Data(); }
    FlatVector<double> FVDouble () const override
    { HostAccess(); return FlatVector<double> (this->size * this->entrysize, (double*)host.Data()); }
    FlatVector<Complex> FVComplex () const override
    {
      if constexpr (!std::is_same_v<SCAL,Complex>) throw Exception ("FVComplex called for real B200Vector");
      HostAccess(); return FlatVector<Complex> (this->size * EntryScalars(), (Complex*)host.Data());
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

    BaseVector & SetScalar (double s) override { double z[2] = {s,0}; dev_uptodate = true; Check (ngsb_vec_set_scalar (DevW(), z)); return *this; }
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
      ngsb_vec * view; Check (ngsb_vec_range (Dev(), r.First(), r.Next(), &view));     // Dev(): the device copy is current
      return make_unique<B200Vector<SCAL>> (view, r.Size(), this->entrysize,
                                            const_cast<B200Vector*>(this)->shared_from_this(), this);
    }
    BaseVector & operator= (const BaseVector & v)
    {
      if (auto p = dynamic_cast<const B200Vector*> (&v)) return Set (1.0, *p);
      dev_uptodate = true;                                        // whatever the host mirror held is overwritten as a whole
      Check (ngsb_vec_h2d (DevW(), v.Memory(), 0, this->size));   // host vector -> device (H2D)
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

  // ---- fused Krylov solvers behind the UNCHANGED factories ----------------------------------------------
  // `CGSolver(mat, pre, ...)` / `GMRESSolver(mat, pre, ...)` (linalg/python_linalg.cpp:1773-1830) construct the C++ classes of
  // linalg/cg.hpp, whose Mult drives the operands op by op through virtual calls -- on device operands that is 10+ kernel
  // launches and two host round trips per iteration.  The classes below ARE those solvers (same base class, same accessors,
  // same stopping rule) with Mult replaced by the device-resident loop of the library whenever both operands are objects of
  // this layer the fused loop understands: SparseMatrix<double | Complex | Mat<3,3>> with a JacobiPrecond of the same entry
  // type or no preconditioner.  Anything else (block Jacobi, composite operators, absolute tolerance) runs the inherited
  // reference loop on the device vectors -- never a silently different preconditioner.
  struct Fused { ngsb_csr * A = nullptr; ngsb_jacobi * C = nullptr; bool ok = false; };
  template <typename TM> static bool TryFuse (const BaseMatrix * a, const BaseMatrix * c, Fused & f)
  {
    auto A = dynamic_cast<const B200SparseMatrix<TM>*> (a);
    if (!A) return false;
    if (c)
      {
        auto C = dynamic_cast<const B200Jacobi<TM>*> (c);
        if (!C) return false;
        f.C = C->Handle();
      }
    f.A = A->Handle(); f.ok = true;
    return true;
  }
  static Fused Fuse (const BaseMatrix * a, const BaseMatrix * c)
  {
    Fused f;
    if (!a) return f;
    if (!TryFuse<double> (a, c, f) && !TryFuse<Complex> (a, c, f)) TryFuse<Mat<3,3,double>> (a, c, f);
    return f;
  }
  static bool IsB200Operator (const BaseMatrix * a)
  {
    return dynamic_cast<const B200SparseMatrix<double>*> (a) || dynamic_cast<const B200SparseMatrix<Complex>*> (a)
      || dynamic_cast<const B200SparseMatrix<Mat<3,3,double>>*> (a);
  }

  template <class IPTYPE>
  class B200CG : public CGSolver<IPTYPE>
  {
    using SCAL = typename SCAL_TRAIT<IPTYPE>::SCAL;
    mutable bool last_fused = false;
  public:
    using CGSolver<IPTYPE>::CGSolver;
    bool LastFused () const { return last_fused; }
    void Mult (const BaseVector & f, BaseVector & u) const override
    {
      Fused fu = Fuse (this->a.get(), this->c.get());
      last_fused = fu.ok && !this->stop_absolute && !this->useseed;
      if (!last_fused) { CGSolver<IPTYPE>::Mult (f, u); return; }
      constexpr int ip = std::is_same_v<IPTYPE,double> ? NGSB_IP_REAL : (std::is_same_v<IPTYPE,ComplexConjugate> ? NGSB_IP_COMPLEX_CONJ : NGSB_IP_COMPLEX);
      std::vector<double> hist (this->printrates ? this->maxsteps + 2 : 0);
      int nh = 0;
      Check (ngsb_cg_solve (fu.A, fu.C, B200Vector<SCAL>::Cast(f).Dev(), B200Vector<SCAL>::CastW(u).DevW(), this->prec, this->maxsteps,
                            ip, this->initialize, &this->steps, hist.data(), int(hist.size()), &nh));
      if (this->printrates)      // the reference prints every step as it goes (linalg/cg.cpp:619-620); here after the fact
        for (int k = 0; k < std::min<int> (nh, hist.size()); k++)
          cout << IM(1) << k << " " << sqrt (hist[k]) << endl;
    }
  };

  template <class IPTYPE>
  class B200GMRES : public GMRESSolver<IPTYPE>
  {
    using SCAL = typename SCAL_TRAIT<IPTYPE>::SCAL;
    mutable bool last_fused = false;
  public:
    using GMRESSolver<IPTYPE>::GMRESSolver;
    bool LastFused () const { return last_fused; }
    void Mult (const BaseVector & f, BaseVector & u) const override
    {
      Fused fu = Fuse (this->a.get(), this->c.get());
      last_fused = fu.ok && !this->stop_absolute;
      if (!last_fused) { GMRESSolver<IPTYPE>::Mult (f, u); return; }
      Check (ngsb_gmres_solve (fu.A, fu.C, B200Vector<SCAL>::Cast(f).Dev(), B200Vector<SCAL>::CastW(u).DevW(), this->prec, this->maxsteps,
                               this->initialize, &this->steps, nullptr, 0, nullptr));
    }
  };

  static shared_ptr<KrylovSpaceSolver> MakeCG (shared_ptr<BaseMatrix> mat, shared_ptr<BaseMatrix> pre, bool iscomplex, bool conjugate)
  {
    if (mat->IsComplex()) iscomplex = true;
    if (!iscomplex) return make_shared<B200CG<double>> (mat, pre);
    if (conjugate) return make_shared<B200CG<ComplexConjugate>> (mat, pre);
    return make_shared<B200CG<Complex>> (mat, pre);
  }
  static shared_ptr<KrylovSpaceSolver> MakeGMRES (shared_ptr<BaseMatrix> mat, shared_ptr<BaseMatrix> pre)
  {
    if (!mat->IsComplex()) return make_shared<B200GMRES<double>> (mat, pre);
    return make_shared<B200GMRES<Complex>> (mat, pre);
  }
  static bool WasFused (const KrylovSpaceSolver & s)
  {
    if (auto p = dynamic_cast<const B200CG<double>*> (&s)) return p->LastFused();
    if (auto p = dynamic_cast<const B200CG<Complex>*> (&s)) return p->LastFused();
    if (auto p = dynamic_cast<const B200CG<ComplexConjugate>*> (&s)) return p->LastFused();
    if (auto p = dynamic_cast<const B200GMRES<double>*> (&s)) return p->LastFused();
    if (auto p = dynamic_cast<const B200GMRES<Complex>*> (&s)) return p->LastFused();
    return false;
  }

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
  using namespace ngla;
  py::module_ la = py::module_::import ("ngsolve.la");    // BaseMatrix / KrylovSpaceSolver are registered there
  InitNgsB200 ();                                         // registration at import, like _ngscuda (ngscuda/python_ngscuda.cpp:20-25)

  // The factories scripts already call.  An overload PREPENDED to the existing function object: every name bound to it
  // (`ngsolve.CGSolver`, `ngsolve.la.CGSolver`, a script's `from ngsolve import *`) dispatches here first; for operands that
  // are not device matrices of this layer the overload steps aside (reference_cast_error = "try the next overload") and
  // the reference's own factory runs, unchanged.
  la.def ("CGSolver", [] (shared_ptr<BaseMatrix> mat, shared_ptr<BaseMatrix> pre, bool iscomplex, bool printrates, double precision,
                          int maxsteps, bool conjugate, std::optional<int> maxiter) -> shared_ptr<KrylovSpaceSolver>
          {
            if (!mat || !IsB200Operator (mat.get())) throw py::reference_cast_error ();
            if (maxiter) maxsteps = *maxiter;
            auto solver = MakeCG (mat, pre, iscomplex, conjugate);
            solver->SetPrecision (precision);
            solver->SetMaxSteps (maxsteps);
            solver->SetPrintRates (printrates);
            return solver;
          },
          py::arg("mat"), py::arg("pre"), py::arg("complex") = false, py::arg("printrates") = true, py::arg("precision") = 1e-8,
          py::arg("maxsteps") = 200, py::arg("conjugate") = false, py::arg("maxiter") = py::none(), py::prepend ());
  la.def ("GMRESSolver", [] (shared_ptr<BaseMatrix> mat, shared_ptr<BaseMatrix> pre, bool printrates, double precision, int maxsteps)
          -> shared_ptr<KrylovSpaceSolver>
          {
            if (!mat || !IsB200Operator (mat.get())) throw py::reference_cast_error ();
            auto solver = MakeGMRES (mat, pre);
            solver->SetPrecision (precision);
            solver->SetMaxSteps (maxsteps);
            solver->SetPrintRates (printrates);
            return solver;
          },
          py::arg("mat"), py::arg("pre"), py::arg("printrates") = true, py::arg("precision") = 1e-8, py::arg("maxsteps") = 200,
          py::prepend ());
  // same call as ngscuda.DevCGSolver(mat, pre, ..., precision, maxsteps) (ngscuda/python_ngscuda.cpp:243-266)
  m.def ("DevCGSolver", [] (shared_ptr<BaseMatrix> mat, shared_ptr<BaseMatrix> pre, int maxsteps, double precision, bool printrates)
         {
           auto solver = MakeCG (mat, pre, false, false);
           solver->SetPrecision (precision);
           solver->SetMaxSteps (maxsteps);
           solver->SetPrintRates (printrates);
           return solver;
         }, py::arg("mat"), py::arg("pre"), py::arg("maxsteps") = 200, py::arg("precision") = 1e-8, py::arg("printrates") = false);
  // did the last Mult of this solver run the fused device loop (diagnostics for tests)
  m.def ("WasFused", [] (shared_ptr<KrylovSpaceSolver> s) { return WasFused (*s); });
  m.def ("SetOption", [] (std::string name, long value) { Check (ngsb_ctx_set_option (TheCtx(), name.c_str(), value)); });
  m.def ("LaunchCount", [] () { uint64_t n = 0; Check (ngsb_ctx_launch_count (TheCtx(), &n)); return n; });
  m.def ("ReorderInfo", [] (shared_ptr<BaseMatrix> mat)
         {
           int on = 0; double share = -1;
           ngsb_csr * h = nullptr;
           if (auto p = dynamic_cast<B200SparseMatrix<double>*> (mat.get())) h = p->Handle();
           else if (auto p = dynamic_cast<B200SparseMatrix<Complex>*> (mat.get())) h = p->Handle();
           else if (auto p = dynamic_cast<B200SparseMatrix<Mat<3,3,double>>*> (mat.get())) h = p->Handle();
           if (!h) throw Exception ("ReorderInfo: not a device sparse matrix of this layer");
           Check (ngsb_csr_reorder_info (h, &on, nullptr, &share));
           return py::make_tuple (bool(on), share);
         });
}
