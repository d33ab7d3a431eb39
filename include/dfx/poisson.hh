// poisson.hh — the element-loop callers of the reference's examples/poisson-pq2.cc on top of the
// C ABI: getOccupationPattern (:137-174), assembleLaplaceMatrix (:176-255) and the symmetric
// Dirichlet treatment (:356-386), with the minimal dune-istl stand-ins they need (BCRSMatrix as
// plain CSR arrays, BlockVector<double> = std::vector<double>).  The set-up work (pattern, element
// matrices, load vector) runs on the device; the matrix lives in host memory afterwards because
// the example hands it to a dune-istl solver, which is outside this path.
#pragma once

#include <cstdint>
#include <vector>

#include "dune_functions.hh"

namespace Dune {

template <class T>
using BlockVector = std::vector<T>;

// compressed-row matrix with ascending column indices per row (what MatrixIndexSet::exportIdx
// produces for a BCRSMatrix<double>)
template <class T>
class BCRSMatrix {
 public:
  std::size_t N() const { return rowPtr.empty() ? 0 : rowPtr.size() - 1; }
  std::size_t M() const { return N(); }
  std::size_t nonzeroes() const { return colIdx.size(); }
  BCRSMatrix& operator=(const T& v) { values.assign(values.size(), v); return *this; }
  // y = A x
  template <class X, class Y> void mv(const X& x, Y& y) const {
    for (std::size_t i = 0; i < N(); ++i) {
      T s = 0;
      for (uint64_t p = rowPtr[i]; p < rowPtr[i + 1]; ++p) s += values[p] * x[colIdx[p]];
      y[i] = s;
    }
  }
  // y -= A x
  template <class X, class Y> void mmv(const X& x, Y& y) const {
    for (std::size_t i = 0; i < N(); ++i)
      for (uint64_t p = rowPtr[i]; p < rowPtr[i + 1]; ++p) y[i] -= values[p] * x[colIdx[p]];
  }
  T entry(std::size_t i, std::size_t j) const {
    for (uint64_t p = rowPtr[i]; p < rowPtr[i + 1]; ++p) if (colIdx[p] == j) return values[p];
    return T(0);
  }
  std::vector<uint64_t> rowPtr;
  std::vector<uint32_t> colIdx;
  std::vector<T> values;
};

namespace Functions {
namespace Poisson {

// getOccupationPattern(feBasis, nb); nb.exportIdx(matrix)
template <class FEBasis>
void getOccupationPattern(const FEBasis& feBasis, BCRSMatrix<double>& matrix) {
  const std::size_t n = feBasis.dimension();
  matrix.rowPtr.assign(n + 1, 0);
  uint64_t nnz = 0;
  Impl::check(dfx_matrix_pattern_host(feBasis.handle(), matrix.rowPtr.data(), nullptr, &nnz));
  matrix.colIdx.assign(nnz, 0);
  Impl::check(dfx_matrix_pattern_host(feBasis.handle(), matrix.rowPtr.data(), matrix.colIdx.data(), &nnz));
  matrix.values.assign(nnz, 0.0);
}

// assembleLaplaceMatrix(feBasis, matrix, rhs, volumeTerm) with a registry function as volume term
template <class FEBasis>
void assembleLaplaceMatrix(const FEBasis& feBasis, BCRSMatrix<double>& matrix, BlockVector<double>& rhs,
                           const Registry::Function& volumeTerm) {
  getOccupationPattern(feBasis, matrix);
  rhs.assign(feBasis.dimension(), 0.0);
  Impl::check(dfx_assemble_laplace_host(feBasis.handle(), matrix.rowPtr.data(), matrix.colIdx.data(), matrix.values.data(),
                                        &volumeTerm.f, rhs.data()));
}

// poisson-pq2.cc:356-386 — x holds the Dirichlet values (zero elsewhere)
template <class BitVector>
void incorporateDirichletValues(BCRSMatrix<double>& stiffnessMatrix, const BitVector& dirichletNodes,
                                const BlockVector<double>& x, BlockVector<double>& rhs) {
  stiffnessMatrix.mmv(x, rhs);
  for (std::size_t i = 0; i < stiffnessMatrix.N(); ++i)
    for (uint64_t p = stiffnessMatrix.rowPtr[i]; p < stiffnessMatrix.rowPtr[i + 1]; ++p) {
      const std::size_t idx = stiffnessMatrix.colIdx[p];
      if (dirichletNodes[i]) stiffnessMatrix.values[p] = (i == idx) ? 1.0 : 0.0;
      else if (dirichletNodes[idx]) stiffnessMatrix.values[p] = 0.0;
    }
  for (std::size_t i = 0; i < stiffnessMatrix.N(); ++i) if (dirichletNodes[i]) rhs[i] = x[i];
}

}  // namespace Poisson
}  // namespace Functions
}  // namespace Dune
