// stokes.hh — the element-loop callers of the reference's examples/stokes-taylorhood.cc on top of the C
// ABI: assembleStokesMatrix (:252-297 with getOccupationPattern :164-250 and getLocalMatrix :49-160).  The
// example's MultiTypeBlockMatrix of BCRSMatrix blocks is one CSR matrix here whose rows and columns are the
// flat container positions of the multi-indices (localView.flatIndex(i)); pattern and values are computed
// on the device, the matrix lives in host memory afterwards because the example hands it to a dune-istl
// solver, which is outside this path.
#pragma once

#include "poisson.hh"

namespace Dune {
namespace Functions {
namespace Stokes {

// assembleStokesMatrix(basis, matrix): the basis is a Taylor-Hood tree
// composite(power<dim>(lagrange<2>(), ims), lagrange<1>(), ims') or TaylorHoodBasis
template <class Basis>
void assembleStokesMatrix(const Basis& basis, BCRSMatrix<double>& matrix) {
  Poisson::getOccupationPattern(basis, matrix);
  Impl::check(dfx_assemble_stokes_host(basis.handle(), matrix.rowPtr.data(), matrix.colIdx.data(), matrix.values.data()));
}

// matrixEntry(matrix, row, col) of the example (:299 ff. uses it to write identity rows), on flat positions
inline double& matrixEntry(BCRSMatrix<double>& matrix, std::size_t row, std::size_t col) {
  for (uint64_t p = matrix.rowPtr[row]; p < matrix.rowPtr[row + 1]; ++p)
    if (matrix.colIdx[p] == col) return matrix.values[p];
  throw RangeError("matrix entry is not in the occupation pattern");
}

}  // namespace Stokes
}  // namespace Functions
}  // namespace Dune
