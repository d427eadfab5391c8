// ref_helpers.cpp -- golden-fixture helper, compiled against the REFERENCE build (oracle/_ref/ngs) by
// make_golden_reorder.py with `ngscxx`.  It exposes reference C++ entry points that have no Python binding:
//   reorder(mat, perm)  -> SparseMatrix<TM>::Reorder(perm)   linalg/sparsematrix_impl.hpp:762-783
//   archive(mat, file)  -> SparseMatrix<TM>::DoArchive into a BinaryOutArchive   linalg/sparsematrix_impl.hpp:443-452
// Test infrastructure only; never part of the product.
#include <la.hpp>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

namespace py = pybind11;

PYBIND11_MODULE(ref_helpers, m)
{
  // SparseMatrix<TM>::DoArchive (linalg/sparsematrix_impl.hpp:443-452) through ngcore's BinaryOutArchive: the wire format
  m.def("archive", [] (std::shared_ptr<ngla::BaseMatrix> mat, std::string filename)
        {
          ngcore::BinaryOutArchive ar (filename);
          mat->DoArchive (ar);
        });
  m.def("reorder", [] (std::shared_ptr<ngla::BaseMatrix> mat, std::vector<size_t> perm) -> std::shared_ptr<ngla::BaseMatrix>
        {
          auto sp = std::dynamic_pointer_cast<ngla::BaseSparseMatrix> (mat);
          if (!sp) throw ngcore::Exception ("reorder: not a sparse matrix");
          ngcore::Array<size_t> p(perm.size());
          for (size_t i = 0; i < perm.size(); i++) p[i] = perm[i];
          return sp->Reorder (p);
        });
}
