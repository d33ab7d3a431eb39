// poisson_pq2.cc — the reference's examples/poisson-pq2.cc written against the façade
// (include/dfx/poisson.hh): assemble the Laplace problem with a Q2 basis, mark the boundary DOFs
// with forEachBoundaryDOF, interpolate the Dirichlet values into them, incorporate them
// symmetrically, solve with CG and look at the result through a DiscreteGlobalBasisFunction.
// Differences from the example: a structured 32x32 YaspGrid instead of the refined mixed UGGrid,
// an unpreconditioned CG written out here instead of dune-istl's CGSolver + SeqILDL, checks
// instead of VTK output, and a second run with a known solution.
#include <cmath>
#include <cstdio>

#include <dfx/poisson.hh>

using namespace Dune;

static int failures = 0;
#define CHECK(cond, what)                                                        \
  do { if (!(cond)) { std::printf("FAILED %s (%s:%d)\n", what, __FILE__, __LINE__); ++failures; } } while (0)

template <class FEBasis>
void boundaryTreatment(const FEBasis& feBasis, std::vector<char>& dirichletNodes) {
  dirichletNodes.clear();
  dirichletNodes.resize(feBasis.dimension(), false);
  Functions::forEachBoundaryDOF(feBasis, [&](auto&& index) { dirichletNodes[index] = true; });
}

static double dot(const std::vector<double>& a, const std::vector<double>& b) {
  double s = 0;
  for (std::size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}

// conjugate gradients, reduction of the residual norm by `reduction`
static int cg(const BCRSMatrix<double>& A, std::vector<double>& x, const std::vector<double>& b, double reduction, int maxIt) {
  const std::size_t n = x.size();
  std::vector<double> r(b), p, Ap(n);
  A.mmv(x, r);
  p = r;
  double rr = dot(r, r);
  const double stop = reduction * reduction * rr;
  for (int it = 0; it < maxIt; ++it) {
    if (rr <= stop) return it;
    A.mv(p, Ap);
    const double alpha = rr / dot(p, Ap);
    for (std::size_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
    const double rn = dot(r, r);
    for (std::size_t i = 0; i < n; ++i) p[i] = r[i] + (rn / rr) * p[i];
    rr = rn;
  }
  return maxIt;
}

int main() {
  try {
    const int dim = 2;
    FieldVector<double, dim> L(1.0);
    std::array<int, dim> N = {{32, 32}};
    YaspGrid<dim> grid(L, N);
    auto gridView = grid.leafGridView();
    using GridView = decltype(gridView);

    typedef Functions::LagrangeBasis<GridView, 2> FEBasis;
    FEBasis feBasis(gridView);
    using VectorType = BlockVector<double>;
    using MatrixType = BCRSMatrix<double>;

    // ---- the example's problem: -Laplace u = 10, u = sin(2 pi x_0) on the boundary ----
    {
      VectorType rhs;
      MatrixType stiffnessMatrix;
      Functions::Poisson::assembleLaplaceMatrix(feBasis, stiffnessMatrix, rhs, Functions::Registry::constant(10.0));
      CHECK(stiffnessMatrix.N() == 65 * 65 && rhs.size() == 65 * 65, "Q2 on 32x32: 65^2 DOFs");
      double rowSum = 0, asym = 0, mass = 0;
      for (std::size_t i = 0; i < stiffnessMatrix.N(); ++i) {
        double s = 0;
        for (uint64_t p = stiffnessMatrix.rowPtr[i]; p < stiffnessMatrix.rowPtr[i + 1]; ++p) {
          s += stiffnessMatrix.values[p];
          asym = std::max(asym, std::abs(stiffnessMatrix.values[p] - stiffnessMatrix.entry(stiffnessMatrix.colIdx[p], i)));
        }
        rowSum = std::max(rowSum, std::abs(s));
        mass += rhs[i];
      }
      CHECK(rowSum < 1e-11 && asym < 1e-12, "stiffness matrix: symmetric, constants in the kernel");
      CHECK(std::abs(mass - 10.0) < 1e-12, "load vector sums to the integral of the volume term");

      VectorType x(feBasis.dimension(), 0.0);
      std::vector<char> dirichletNodes;
      boundaryTreatment(feBasis, dirichletNodes);
      std::size_t nb = 0;
      for (char c : dirichletNodes) nb += c;
      CHECK(nb == 4 * 64, "256 boundary DOFs");

      const double pi = std::acos(-1.0);
      auto dirichletValueFunction = [pi](const auto& xx) { return std::sin(2 * pi * xx[0]); };
      interpolate(feBasis, x, dirichletValueFunction, dirichletNodes);

      MatrixType original = stiffnessMatrix;
      VectorType originalRhs = rhs;
      Functions::Poisson::incorporateDirichletValues(stiffnessMatrix, dirichletNodes, x, rhs);

      const int its = cg(stiffnessMatrix, x, rhs, 1e-10, 5000);
      CHECK(its < 5000, "CG converged");
      // the solution satisfies the original equations in every interior row
      VectorType res(originalRhs);
      original.mmv(x, res);
      double worst = 0;
      for (std::size_t i = 0; i < res.size(); ++i) if (!dirichletNodes[i]) worst = std::max(worst, std::abs(res[i]));
      CHECK(worst < 1e-8, "interior residual of the original system");
      std::printf("poisson-pq2: %zu DOFs, %zu non-zeros, CG iterations %d, interior residual %.2e\n",
                  (std::size_t)feBasis.dimension(), stiffnessMatrix.nonzeroes(), its, worst);
    }

    // ---- known solution: -Laplace u = 2 pi^2 sin(pi x) sin(pi y), u = 0 on the boundary ----
    {
      const double pi = std::acos(-1.0);
      VectorType rhs;
      MatrixType stiffnessMatrix;
      Functions::Poisson::assembleLaplaceMatrix(feBasis, stiffnessMatrix, rhs, Functions::Registry::sinprod(pi));
      for (auto& v : rhs) v *= 2 * pi * pi;
      VectorType x(feBasis.dimension(), 0.0);
      std::vector<char> dirichletNodes;
      boundaryTreatment(feBasis, dirichletNodes);
      Functions::Poisson::incorporateDirichletValues(stiffnessMatrix, dirichletNodes, x, rhs);
      cg(stiffnessMatrix, x, rhs, 1e-12, 5000);
      auto xFunction = Functions::makeDiscreteGlobalBasisFunction<double>(feBasis, x);
      std::vector<FieldVector<double, 2>> centre = {FieldVector<double, 2>(0.5)};
      std::vector<double> v;
      xFunction.evaluateAll(centre, v);
      double err = 0;
      for (std::size_t e = 0; e < 32 * 32; ++e) {
        const double X = ((e % 32) + 0.5) / 32, Y = ((e / 32) + 0.5) / 32;
        err = std::max(err, std::abs(v[e] - std::sin(pi * X) * std::sin(pi * Y)));
      }
      CHECK(err < 5e-6, "Q2 solution at the element centres");
      std::printf("known solution: max error at element centres %.2e\n", err);
    }
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 2;
  }
  std::printf("%s\n", failures ? "poisson_pq2 FAILED" : "poisson_pq2 passed");
  return failures ? 1 : 0;
}
