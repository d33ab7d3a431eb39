// interpolation_example.cc — the reference's examples/interpolation.cc (BASELINE config 1)
// written against the façade: scalar P2 interpolation of exp(-|x|^2), a Taylor-Hood
// FL(FL) basis whose velocity / pressure subspaces are interpolated separately, masked
// interpolation of boundary values, and evaluation through DiscreteGlobalBasisFunction.
// Differences from the example: the grid is 64x64 (the config BASELINE.json names) instead
// of 1x1 and the results are checked instead of written to VTK.
#include <cstdio>

#include <dfx/dune_functions.hh>

using namespace Dune;

static int failures = 0;
#define CHECK(cond, what)                                                        \
  do { if (!(cond)) { std::printf("FAILED %s (%s:%d)\n", what, __FILE__, __LINE__); ++failures; } } while (0)

int main() {
  try {
    const int dim = 2;
    FieldVector<double, dim> L(1.0);
    std::array<int, dim> N = {{64, 64}};
    YaspGrid<2> grid(L, N);
    using GridView = YaspGrid<2>::LeafGridView;
    GridView gridView = grid.leafGridView();

    auto f1 = [](const FieldVector<double, 2>& x) { return std::exp(-1.0 * x.two_norm2()); };

    Functions::LagrangeBasis<GridView, 1> p1basis(gridView);
    std::vector<double> x0;
    interpolate(p1basis, x0, f1);                                    // host callable at device-computed nodes
    CHECK(x0.size() == 65 * 65, "P1 dimension 65^2");

    Functions::LagrangeBasis<GridView, 2> p2basis(gridView);
    std::vector<double> x1, x1dev;
    interpolate(p2basis, x1, f1);
    interpolate(p2basis, x1dev, Functions::Registry::gaussian(1.0));  // same f from the device registry
    double diff = 0;
    for (std::size_t i = 0; i < x1.size(); ++i) diff = std::max(diff, std::abs(x1[i] - x1dev[i]));
    CHECK(x1.size() == 129 * 129 && diff < 1e-14, "host callable and registry function agree");

    using namespace Functions::BasisFactory;
    auto taylorHoodBasis = makeBasis(gridView, composite(power<dim>(lagrange<2>(), flatLexicographic()),
                                                         lagrange<1>(), flatLexicographic()));
    std::vector<double> x2(taylorHoodBasis.dimension(), 0.0);
    auto f2 = [](const FieldVector<double, 2>& x) { return x; };      // vector valued
    using namespace Indices;
    interpolate(Functions::subspaceBasis(taylorHoodBasis, _0), x2, f2);
    interpolate(Functions::subspaceBasis(taylorHoodBasis, _1), x2, f1);

    // evaluate velocity at element centres: must reproduce x
    auto velocity = Functions::makeDiscreteGlobalBasisFunction<FieldVector<double, 2>>(Functions::subspaceBasis(taylorHoodBasis, _0), x2);
    std::vector<FieldVector<double, 2>> centre = {FieldVector<double, 2>(0.5)};
    std::vector<double> v;
    velocity.evaluateAll(centre, v);
    double err = 0;
    for (std::size_t e = 0; e < 64 * 64; ++e) {
      err = std::max(err, std::abs(v[2 * e] - ((e % 64) + 0.5) / 64));
      err = std::max(err, std::abs(v[2 * e + 1] - ((e / 64) + 0.5) / 64));
    }
    CHECK(err < 1e-13, "velocity interpolant of f(x) = x is exact");

    // boundary mask + masked interpolation, examples/interpolation.cc:99-122: mark the velocity
    // boundary DOFs of the Taylor-Hood basis, then interpolate f2 into the marked entries only
    std::vector<char> isBoundary(taylorHoodBasis.dimension(), 0);
    std::size_t visits = 0;
    Functions::forEachBoundaryDOF(Functions::subspaceBasis(taylorHoodBasis, _0), [&](auto&& index) {
      isBoundary[index] = true;
      ++visits;
    });
    std::vector<char> isBoundaryDev(taylorHoodBasis.dimension(), 0);
    Functions::markBoundaryDOFs(Functions::subspaceBasis(taylorHoodBasis, _0), isBoundaryDev);
    CHECK(isBoundary == isBoundaryDev, "forEachBoundaryDOF loop and the device mask kernel agree");
    const std::size_t n2 = 129 * 129;
    std::size_t nbv = 0, nbp = 0;
    for (std::size_t i = 0; i < isBoundary.size(); ++i) (i < 2 * n2 ? nbv : nbp) += isBoundary[i];
    CHECK(nbv == 2 * 4 * 128 && nbp == 0, "2 x 512 velocity boundary DOFs, no pressure DOF");
    CHECK(visits == 4 * 64 * 3 * 2, "one visit per (boundary facet, facet DOF, component)");
    std::vector<double> x3(taylorHoodBasis.dimension(), -1.0);
    interpolate(Functions::subspaceBasis(taylorHoodBasis, _0), x3, f2, isBoundary);
    bool ok = true;
    for (std::size_t i = 0; i < x3.size(); ++i) ok = ok && (isBoundary[i] ? x3[i] == x2[i] : x3[i] == -1.0);
    CHECK(ok, "masked interpolation touches exactly the marked DOFs");

    // the other two callback signatures (boundarydofs.hh:39-96) and SubEntityDOFs::contains
    {
      std::size_t n3 = 0, n2arg = 0; bool inside = true;
      Functions::forEachBoundaryDOF(p2basis, [&](auto&& localIndex, const auto& localView, const auto& intersection) {
        ++n3;
        inside = inside && intersection.boundary() &&
                 Functions::subEntityDOFs(localView, intersection.indexInInside(), 1).contains(localIndex);
      });
      Functions::forEachBoundaryDOF(p2basis, [&](auto&& localIndex, const auto& localView) {
        n2arg += localView.index(localIndex).size();
      });
      CHECK(n3 == 4 * 64 * 3 && n2arg == n3 && inside, "three- and two-argument callbacks");
    }

    // local function idiom: bind + evaluate + derivative on one element
    auto u = Functions::makeDiscreteGlobalBasisFunction<double>(p2basis, x1);
    auto lf = localFunction(u);
    auto dlf = derivative(lf);
    auto e = grid.element(64 * 10 + 5);
    lf.bind(e); dlf.bind(e);
    FieldVector<double, 2> xl = {0.25, 0.75};
    const double X = (5 + 0.25) / 64, Y = (10 + 0.75) / 64, exact = std::exp(-(X * X + Y * Y));
    CHECK(std::abs(lf(xl)[0] - exact) < 1e-6, "local function value");
    CHECK(std::abs(dlf(xl)[0] - (-2 * X * exact)) < 1e-3 && std::abs(dlf(xl)[1] - (-2 * Y * exact)) < 1e-3, "local derivative");
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 2;
  }
  std::printf("%s\n", failures ? "interpolation_example FAILED" : "interpolation_example passed");
  return failures ? 1 : 0;
}
