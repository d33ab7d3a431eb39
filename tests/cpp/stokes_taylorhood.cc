// stokes_taylorhood.cc — main() of the reference's examples/stokes-taylorhood.cc (:299-480, the
// BLOCKEDBASIS variant) written against the façade: Taylor-Hood basis composite(power<dim>(lagrange<2>,
// blockedInterleaved), lagrange<1>, blockedLexicographic) on YaspGrid<2> 4x4, nested vector types
// (MultiTypeBlockVector<BlockVector<FieldVector<double,dim>>, BlockVector<double>> -> std::tuple of
// std::vectors) behind istlVectorBackend, assembleStokesMatrix, forEachBoundaryDOF on the velocity
// sub-space, masked interpolation of the Dirichlet data (driven cavity), identity rows for boundary
// DOFs through localView.index, restarted GMRES without preconditioner, velocity and pressure as
// DiscreteGlobalBasisFunctions.  dune-istl's solver is replaced by 40 lines of GMRES (outside the path).
// Exit code 0 = all checks passed.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <tuple>

#include <dfx/stokes.hh>

using namespace Dune;

static int failures = 0;
#define CHECK(cond, what)                                                        \
  do { if (!(cond)) { std::printf("FAILED %s (%s:%d)\n", what, __FILE__, __LINE__); ++failures; } } while (0)

// RestartedGMResSolver(op, Richardson(1.0), reduction, restart, maxit) on flat vectors
static double gmres(const BCRSMatrix<double>& A, std::vector<double>& x, const std::vector<double>& b, double reduction,
                    int restart, int maxit, int& iterations) {
  const std::size_t n = b.size();
  auto dot = [&](const std::vector<double>& u, const std::vector<double>& v) { double s = 0; for (std::size_t i = 0; i < n; ++i) s += u[i] * v[i]; return s; };
  std::vector<double> r(n), w(n);
  A.mv(x, r);
  for (std::size_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
  const double r0 = std::sqrt(dot(r, r));
  double res = r0;
  iterations = 0;
  while (res > reduction * r0 && iterations < maxit) {
    std::vector<std::vector<double>> V(1, r);
    for (auto& v : V[0]) v /= res;
    std::vector<std::vector<double>> H(restart + 1, std::vector<double>(restart, 0.0));
    std::vector<double> cs(restart), sn(restart), g(restart + 1, 0.0);
    g[0] = res;
    int k = 0;
    for (; k < restart && iterations < maxit; ++k, ++iterations) {
      A.mv(V[k], w);
      for (int j = 0; j <= k; ++j) { H[j][k] = dot(w, V[j]); for (std::size_t i = 0; i < n; ++i) w[i] -= H[j][k] * V[j][i]; }
      H[k + 1][k] = std::sqrt(dot(w, w));
      V.push_back(w);
      if (H[k + 1][k] > 0) for (auto& v : V[k + 1]) v /= H[k + 1][k];
      for (int j = 0; j < k; ++j) { const double t = cs[j] * H[j][k] + sn[j] * H[j + 1][k]; H[j + 1][k] = -sn[j] * H[j][k] + cs[j] * H[j + 1][k]; H[j][k] = t; }
      const double d = std::hypot(H[k][k], H[k + 1][k]);
      cs[k] = H[k][k] / d; sn[k] = H[k + 1][k] / d;
      H[k][k] = d; H[k + 1][k] = 0;
      g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k];
      res = std::abs(g[k + 1]);
      if (res <= reduction * r0) { ++k; ++iterations; break; }
    }
    std::vector<double> y(k);
    for (int i = k - 1; i >= 0; --i) { double s = g[i]; for (int j = i + 1; j < k; ++j) s -= H[i][j] * y[j]; y[i] = s / H[i][i]; }
    for (int j = 0; j < k; ++j) for (std::size_t i = 0; i < n; ++i) x[i] += y[j] * V[j][i];
    A.mv(x, r);
    for (std::size_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
    res = std::sqrt(dot(r, r));
  }
  return res / r0;
}

int main() {
  try {
    // grid_setup
    const int dim = 2;
    using GridType = YaspGrid<dim>;
    FieldVector<double, dim> upperRight = {1, 1};
    std::array<int, dim> elements = {{4, 4}};
    GridType grid(upperRight, elements);
    using GridView = GridType::LeafGridView;
    GridView gridView = grid.leafGridView();

    // function_space_basis
    using namespace Functions::BasisFactory;
    constexpr std::size_t p = 1;
    auto taylorHoodBasis = makeBasis(gridView, composite(power<dim>(lagrange<p + 1>(), blockedInterleaved()), lagrange<p>(), blockedLexicographic()));

    // linear_algebra_setup
    using VelocityVector = BlockVector<FieldVector<double, dim>>;
    using PressureVector = BlockVector<double>;
    using VectorType = std::tuple<VelocityVector, PressureVector>;
    using VelocityBitVector = BlockVector<FieldVector<char, dim>>;
    using PressureBitVector = BlockVector<char>;
    using BitVectorType = std::tuple<VelocityBitVector, PressureBitVector>;
    using MatrixType = BCRSMatrix<double>;

    // rhs_assembly
    VectorType rhs;
    auto rhsBackend = Functions::istlVectorBackend(rhs);
    rhsBackend.resize(taylorHoodBasis);
    rhsBackend = 0;
    CHECK(std::get<0>(rhs).size() == 81 && std::get<1>(rhs).size() == 25, "rhs sized like the basis");

    // matrix_assembly
    MatrixType stiffnessMatrix;
    Functions::Stokes::assembleStokesMatrix(taylorHoodBasis, stiffnessMatrix);
    const std::size_t n = taylorHoodBasis.dimension();
    CHECK(stiffnessMatrix.N() == n && n == 2 * 81 + 25, "matrix size");
    double asym = 0, ppBlock = 0;
    for (std::size_t i = 0; i < n; ++i)
      for (uint64_t q = stiffnessMatrix.rowPtr[i]; q < stiffnessMatrix.rowPtr[i + 1]; ++q) {
        const std::size_t j = stiffnessMatrix.colIdx[q];
        asym = std::max(asym, std::abs(stiffnessMatrix.values[q] - stiffnessMatrix.entry(j, i)));
        if (i >= 162 && j >= 162) ppBlock = std::max(ppBlock, std::abs(stiffnessMatrix.values[q]));
      }
    CHECK(asym < 1e-13, "Stokes matrix is symmetric");
    CHECK(ppBlock == 0.0, "pressure-pressure block is zero");

    // initialize_boundary_dofs_vector + determine_boundary_dofs
    BitVectorType isBoundary;
    auto isBoundaryBackend = Functions::istlVectorBackend(isBoundary);
    isBoundaryBackend.resize(taylorHoodBasis);
    isBoundaryBackend = false;
    using namespace Indices;
    Functions::forEachBoundaryDOF(Functions::subspaceBasis(taylorHoodBasis, _0), [&](auto&& index) { isBoundaryBackend[index] = true; });
    std::size_t nBoundary = 0;
    for (auto& v : std::get<0>(isBoundary)) nBoundary += (v[0] != 0) + (v[1] != 0);
    for (auto c : std::get<1>(isBoundary)) nBoundary += c != 0;
    CHECK(nBoundary == 2 * 32, "both velocity components of the 32 boundary nodes of the 9x9 lattice, no pressure DOF");

    // interpolate_dirichlet_values
    using Coordinate = FieldVector<double, dim>;
    using VelocityRange = FieldVector<double, dim>;
    auto velocityDirichletData = [](Coordinate x) { return VelocityRange{0.0, double(x[0] < 1e-8)}; };
    Functions::interpolate(Functions::subspaceBasis(taylorHoodBasis, _0), rhs, velocityDirichletData, isBoundary);

    // set_dirichlet_matrix
    auto localView = taylorHoodBasis.localView();
    for (const auto& element : Dune::elements(taylorHoodBasis.gridView())) {
      localView.bind(element);
      for (std::size_t i = 0; i < localView.size(); ++i) {
        auto row = localView.index(i);
        if (isBoundaryBackend[row])
          for (std::size_t j = 0; j < localView.size(); ++j)
            Functions::Stokes::matrixEntry(stiffnessMatrix, localView.flatIndex(i), localView.flatIndex(j)) = (i == j) ? 1 : 0;
      }
    }

    // stokes_solve: start from the rhs vector, GMRES to a reduction of 1e-10
    std::vector<double> b, x;
    Functions::Impl::flattenNested(rhs, b);
    x = b;
    int it = 0;
    const double red = gmres(stiffnessMatrix, x, b, 1e-10, 500, 500, it);     // reduction, restart, maxit as in the example (:482-484)
    std::printf("GMRES: %d iterations, reduction %.2e\n", it, red);
    CHECK(red <= 1e-10, "GMRES converged");
    VectorType sol = rhs;
    const double* px = x.data();
    Functions::Impl::unflattenNested(sol, px);

    // make_result_functions: velocity and pressure as discrete functions of the nested solution vector
    auto velocityFunction = Functions::makeDiscreteGlobalBasisFunction<VelocityRange>(Functions::subspaceBasis(taylorHoodBasis, _0), sol);
    auto pressureFunction = Functions::makeDiscreteGlobalBasisFunction<double>(Functions::subspaceBasis(taylorHoodBasis, _1), sol);
    auto vLeft = velocityFunction(Coordinate{0.0, 0.5}), vRight = velocityFunction(Coordinate{1.0, 0.5}), vMid = velocityFunction(Coordinate{0.5, 0.5});
    CHECK(std::abs(vLeft[0]) < 1e-12 && std::abs(vLeft[1] - 1.0) < 1e-12, "driven wall: velocity (0, 1) on x = 0");
    CHECK(std::abs(vRight[0]) < 1e-12 && std::abs(vRight[1]) < 1e-12, "no slip on x = 1");
    CHECK(vMid[1] < 0.0, "the cavity flow turns: downward velocity in the middle");
    // discrete incompressibility: the pressure rows of the assembled system hold B u = 0
    double div = 0;
    {
      MatrixType B;
      Functions::Stokes::assembleStokesMatrix(taylorHoodBasis, B);
      std::vector<double> y(n);
      B.mv(x, y);
      for (std::size_t i = 162; i < n; ++i) div = std::max(div, std::abs(y[i]));
    }
    CHECK(div < 1e-8, "discrete divergence of the solution vanishes");
    std::printf("u(0.5,0.5) = (%.6f, %.6f), p(0.5,0.5) = %.6f, max |B u| = %.2e\n", vMid[0], vMid[1], pressureFunction(Coordinate{0.5, 0.5})[0], div);
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 2;
  }
  std::printf("%s\n", failures ? "stokes_taylorhood FAILED" : "stokes_taylorhood passed");
  return failures ? 1 : 0;
}
