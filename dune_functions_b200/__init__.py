"""B200-native element-loop hot path of dune-functions.

Python front door over the C ABI (include/dfx.h).  Importing the package does not
load the CUDA library; the first call that needs it does, and fails loudly if
`lib/libdfx.so` has not been built.
"""
from .basis import (  # noqa: F401
    YaspGrid, makeBasis, GlobalBasis, LocalView, lagrange, lagrangeDG, power, composite, taylorHood,
    flatLexicographic, flatInterleaved, blockedLexicographic, blockedInterleaved,
    subspaceBasis, SubspaceBasis, interpolate, interpolateBatch, interpolationPoints, interpolateFromValues,
    makeDiscreteGlobalBasisFunction, DiscreteGlobalBasisFunction, subEntityDOFs, boundaryDOFs,
    gaussian, identity, constant, coord, quadratic, sinprod, gaussian_vec, make_function,
    tensor_gauss, gauss_points_1d, tree_array,
)
from .operators import CSRMatrix, LaplaceOperator, getOccupationPattern, assembleLaplaceMatrix, assembleStokesMatrix, applyDirichlet  # noqa: F401
from ._capi import DfxError, RangeError, NotImplementedDfx, CudaError  # noqa: F401
