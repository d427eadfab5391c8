// ref_helpers.cpp -- golden-fixture helper, compiled against the REFERENCE build (oracle/_ref/ngs) by
// make_golden_reorder.py with `ngscxx`.  It exposes reference C++ entry points that have no Python binding:
//   reorder(mat, perm)  -> SparseMatrix<TM>::Reorder(perm)   linalg/sparsematrix_impl.hpp:762-783
// Test infrastructure only; never part of the product.
#include <la.hpp>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

namespace py = pybind11;

PYBIND11_MODULE(ref_helpers, m)
{
  m.def("reorder", [] (std::shared_ptr<ngla::BaseMatrix> mat, std::vector<size_t> perm) -> std::shared_ptr<ngla::BaseMatrix>
        {
          auto sp = std::dynamic_pointer_cast<ngla::BaseSparseMatrix> (mat);
          if (!sp) throw ngcore::Exception ("reorder: not a sparse matrix");
          ngcore::Array<size_t> p(perm.size());
          for (size_t i = 0; i < perm.size(); i++) p[i] = perm[i];
          return sp->Reorder (p);
        });
}
