// host_only_test.cc — the parts of the façade that are host bookkeeping (no device call), so this program
// runs in the CPU suite (tests/test_facade_host_cpu.py):
//   subspacebasistest.cc            subspaceBasis from tree-path indices, subspaceBasis of a SubspaceBasis
//                                   joins the paths and refers to the root's root
//   containerdescriptortest.cc      checkSize (:87-103) for the basis list of its main() (:236-300)
//   istlvectorbackend.hh:153-233    resize of nested containers, RangeError for static levels of the wrong size
//   basis.update / dimension / size(prefix), RangeError for an unsupported order (lagrangebasis.hh:795-796)
#include <cstdio>
#include <tuple>

#include <dfx/dune_functions.hh>

using namespace Dune;
using namespace Dune::Functions;
using namespace Dune::Functions::BasisFactory;
using namespace Dune::Indices;

static int failures = 0;
#define CHECK(cond, what)                                                        \
  do { if (!(cond)) { std::printf("FAILED %s (%s:%d)\n", what, __FILE__, __LINE__); ++failures; } } while (0)

template <class CD, class Basis>
void checkSize(const CD& cd, const Basis& basis, std::vector<std::size_t> prefix, std::size_t maxLen) {
  CHECK(basis.size(prefix) == cd.size(), "size(prefix) == containerDescriptor size at prefix");
  if (prefix.size() < maxLen && !cd.isValue())
    for (std::size_t i = 0; i < cd.size(); ++i) {
      auto p = prefix; p.push_back(i);
      checkSize(cd[i], basis, p, maxLen);
    }
}
template <class Factory>
void checkBasis(const YaspGrid<2>::LeafGridView& gv, const Factory& f, const char* name, const char* type = nullptr) {
  auto basis = makeBasis(gv, f);
  auto cd = basis.containerDescriptor();
  CHECK(!cd.isUnknown(), "descriptor known");
  checkSize(cd, basis, {}, 4);
  if (type) CHECK(cd.type() == type, "descriptor type");
  std::printf("containerDescriptor %-30s %s\n", name, cd.type().c_str());
}

// ---- backends/test/istlvectorbackendtest.cc: a mock size provider with non-uniform multi-index lengths
// (:33-80) and checkISTLVectorBackend (:83-138) on the standard-library stand-ins of its vector types
template <std::size_t dim>
struct GlobalBasisMoc {
  std::size_t size(const std::vector<std::size_t>& prefix) const {
    if (prefix.size() == 0) return 2;
    if (prefix.size() == 1) return prefix[0] == 0 ? 23 : (prefix[0] == 1 ? 42 : 0);
    if (prefix.size() == 2) return prefix[0] == 0 ? dim : 0;
    return 0;
  }
};
template <class Vector, std::size_t dim>
void checkISTLVectorBackend(const char* name) {
  GlobalBasisMoc<dim> basis;
  Vector x_raw;
  auto x = istlVectorBackend(x_raw);
  x.resize(basis);
  CHECK(std::tuple_size_v<Vector> == basis.size({}), "resize check: root");
  CHECK(std::get<0>(x_raw).size() == basis.size({0}), "resize check: x_raw[_0]");
  auto sizeOf = [](const auto& c) {
    if constexpr (Dune::Functions::Impl::IsTuple<std::decay_t<decltype(c)>>::value) return std::tuple_size_v<std::decay_t<decltype(c)>>;
    else return c.size();
  };
  for (std::size_t i = 0; i < basis.size({0}); ++i) CHECK(sizeOf(std::get<0>(x_raw)[i]) == basis.size({0, i}), "resize check: x_raw[_0][i]");
  CHECK(std::get<1>(x_raw).size() == basis.size({1}), "resize check: x_raw[_1]");
  auto mi3 = [](std::size_t a, std::size_t b, std::size_t c) { MultiIndex m; m.n_ = 3; m.d_[0] = a; m.d_[1] = b; m.d_[2] = c; return m; };
  auto mi2 = [](std::size_t a, std::size_t b) { MultiIndex m; m.n_ = 2; m.d_[0] = a; m.d_[1] = b; return m; };
  for (std::size_t i = 0; i < 23; ++i) for (std::size_t j = 0; j < dim; ++j) x[mi3(0, i, j)] = 0 + i + j;
  for (std::size_t i = 0; i < 42; ++i) x[mi2(1, i)] = 1 + i;
  const auto& x_const = x;
  bool ok = true;
  for (std::size_t i = 0; i < 23; ++i) for (std::size_t j = 0; j < dim; ++j) ok = ok && x_const[mi3(0, i, j)] == double(0 + i + j);
  for (std::size_t i = 0; i < 42; ++i) ok = ok && x_const[mi2(1, i)] == double(1 + i);
  CHECK(ok, "entries written through multi-indices are read back");
  std::vector<double> flat;
  Dune::Functions::Impl::flattenNested(x_raw, flat);
  CHECK(flat.size() == 23 * dim + 42 && flat[dim + 1] == 2.0 && flat[23 * dim + 5] == 6.0, "lexicographic flat image");
  std::printf("istlVectorBackend %-46s ok\n", name);
}

// ---- gridfunctions/test/gridfunctiontest.hh:23-44 integrateGridViewFunction with the midpoint rule
// (quadOrder = 1), used by analyticgridviewfunctiontest.cc:52-95 with exactIntegral = 0.5 for f(x) = x_0
template <class GridView, class F>
double integrateGridViewFunction(const GridView& gridView, const F& f) {
  double integral = 0;
  auto fLocal = localFunction(f);
  for (const auto& e : elements(gridView)) {
    auto geometry = e.geometry();
    fLocal.bind(e);
    if (!fLocal.bound()) throw Exception("Function is not bound after bind()");
    FieldVector<double, GridView::dimension> quadPos(0.5);
    integral += fLocal(quadPos) * 1.0 * geometry.integrationElement(quadPos);
    fLocal.unbind();
  }
  return integral;
}
static double x_component_A(const FieldVector<double, 2>& x) { return x[0]; }

// HierarchicNodeToRangeMap + flatVectorView on nested range types (hierarchicnodetorangemap.hh:33-48,
// flatvectorview.hh:63-100): pure host templates, checked without a device
static void checkNodeToRangeMap() {
  using namespace Dune::Functions::Impl;
  using TH = std::tuple<FieldVector<double, 3>, double>;
  const TH y{FieldVector<double, 3>{1.0, 2.0, 3.0}, 4.0};
  CHECK(rangeEntry(y, {0, 0}) == 1.0 && rangeEntry(y, {0, 2}) == 3.0 && rangeEntry(y, {1}) == 4.0, "tuple<FieldVector<3>, double>[treePath]");
  CHECK(rangeEntry(5.5, {0, 1}) == 5.5 && rangeEntry(5.5, {}) == 5.5, "a range without index access is used whole");
  const FieldMatrix<double, 2, 3> m{{{1, 2, 3}, {4, 5, 6}}};
  CHECK(rangeEntry(m, {1, 2}) == 6.0 && rangeEntry(m, {0, 1}) == 2.0, "FieldMatrix[i][j]");
  CHECK(rangeEntry(m, {1}) == 4.0 && rangeEntry(m, {}) == 1.0, "a path that ends early takes the flat view's first entry (row-major)");
  const std::array<FieldVector<double, 2>, 2> a{{FieldVector<double, 2>{1, 2}, FieldVector<double, 2>{3, 4}}};
  CHECK(rangeEntry(a, {1, 0}) == 3.0, "array<FieldVector<2>, 2>[i][j]");
  const FieldVector<double, 1> one{7.0};
  CHECK(rangeEntry(one, {}) == 7.0, "FieldVector<double,1> as a scalar entry");
  bool threw = false;
  try { rangeEntry(y, {2}); } catch (const RangeError&) { threw = true; }
  CHECK(threw, "tree path beyond the tuple throws RangeError");
  threw = false;
  try { rangeEntry(m, {2, 0}); } catch (const RangeError&) { threw = true; }
  CHECK(threw, "tree path beyond the matrix rows throws RangeError");
  TH z{};
  assignRangeEntry(z, {0, 1}, 2.5);
  assignRangeEntry(z, {1}, -1.0);
  CHECK(std::get<0>(z)[1] == 2.5 && std::get<0>(z)[0] == 0.0 && std::get<1>(z) == -1.0, "assignment into y[treePath]");
  double s = 0;
  assignRangeEntry(s, {0, 0}, 3.0);
  CHECK(s == 3.0, "assignment into a scalar range");
}

int main() {
  checkNodeToRangeMap();
  using Grid = YaspGrid<2>;
  Grid grid(FieldVector<double, 2>(1.0), {{2, 2}});
  auto gv = grid.leafGridView();
  try {
    // ---- subspacebasistest.cc
    auto sameSubspace = [](const auto& a, const auto& b) { return &a.rootBasis() == &b.rootBasis() && a.prefixPath() == b.prefixPath() && a.rootNode() == b.rootNode(); };
    {
      auto basis = makeBasis(gv, composite(power<3>(lagrange<3>(), blockedInterleaved()), lagrange<1>(), blockedLexicographic()));
      auto sb0 = subspaceBasis(basis, _0);
      auto sb01_a = subspaceBasis(sb0, 1);
      auto sb01_b = subspaceBasis(basis, _0, 1);
      CHECK(sameSubspace(sb01_a, sb01_b), "subspaceBasis(subspaceBasis(b,tp1),tp2) equals subspaceBasis(b,(tp1,tp2))");
      CHECK((sb01_a.prefixPath() == std::vector<std::size_t>{0, 1}), "joined tree path");
      CHECK(!sameSubspace(sb0, sb01_a) && !sameSubspace(sb01_a, subspaceBasis(basis, _0, 2)), "different sub-trees differ");
      CHECK(sb01_a.dimension() == basis.dimension(), "dimension() is the root's (subspacebasis.hh:79-94)");
      bool threw = false;
      try { subspaceBasis(basis, _0, 3); } catch (const RangeError&) { threw = true; }
      CHECK(threw, "no such child");
    }
    // ---- containerdescriptortest.cc main()
    checkBasis(gv, lagrange<1>(), "lagrange<1>", "UniformVector<Value>");
    checkBasis(gv, power<2>(lagrange<1>(), blockedInterleaved()), "power<2>(Q1,BI)", "UniformVector<UniformArray<Value,2>>");
    checkBasis(gv, composite(lagrange<1>(), lagrange<2>(), blockedLexicographic()), "composite(Q1,Q2,BL)", "Array<UniformVector<Value>,2>");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), flatLexicographic()), flatLexicographic()), "FL(FL)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), flatInterleaved()), flatLexicographic()), "FL(FI)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), blockedLexicographic()), flatLexicographic()), "FL(BL)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), blockedInterleaved()), flatLexicographic()), "FL(BI)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), flatLexicographic()), flatInterleaved()), "FI(FL)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), blockedLexicographic()), flatInterleaved()), "FI(BL)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), blockedInterleaved()), flatInterleaved()), "FI(BI)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), flatInterleaved()), blockedLexicographic()), "BL(FI)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), blockedLexicographic()), blockedLexicographic()), "BL(BL)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), blockedInterleaved()), blockedLexicographic()), "BL(BI)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), flatLexicographic()), blockedInterleaved()), "BI(FL)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), blockedLexicographic()), blockedInterleaved()), "BI(BL)");
    checkBasis(gv, power<2>(power<3>(lagrange<2>(), blockedInterleaved()), blockedInterleaved()), "BI(BI)");
    checkBasis(gv, power<2>(composite(power<1>(power<1>(lagrange<1>(), blockedInterleaved()), blockedLexicographic()),
                                      power<2>(lagrange<1>(), blockedInterleaved()), power<3>(lagrange<1>(), blockedLexicographic()),
                                      blockedLexicographic()), blockedLexicographic()), "nested blocked tree");
    checkBasis(gv, composite(composite(power<2>(lagrange<1>(), flatInterleaved()), power<1>(power<2>(lagrange<1>(), flatLexicographic()), flatLexicographic()), flatLexicographic()),
                             composite(power<1>(power<1>(lagrange<1>(), flatLexicographic()), flatInterleaved()), power<2>(lagrange<1>(), flatInterleaved()),
                                       power<3>(lagrange<1>(), flatLexicographic()), flatLexicographic()),
                             flatLexicographic()), "nested flat tree", "UniformVector<Value>");
    // ---- istlVectorBackend(v).resize(basis) on nested containers
    {
      auto th = makeBasis(gv, composite(power<2>(lagrange<2>(), blockedInterleaved()), lagrange<1>(), blockedLexicographic()));
      std::tuple<std::vector<FieldVector<double, 2>>, std::vector<double>> v;
      auto backend = istlVectorBackend(v);
      backend.resize(th);
      backend = 3.0;
      CHECK(std::get<0>(v).size() == 25 && std::get<1>(v).size() == 9, "tuple of block vectors sized like the basis");
      MultiIndex mi; mi.n_ = 3; mi.d_[0] = 0; mi.d_[1] = 7; mi.d_[2] = 1;
      backend[mi] = 5.0;
      CHECK(std::get<0>(v)[7][1] == 5.0 && std::get<0>(v)[7][0] == 3.0, "vector[multiIndex] resolves digit by digit");
      MultiIndex mp; mp.n_ = 2; mp.d_[0] = 1; mp.d_[1] = 4;
      backend[mp] = -1.0;
      CHECK(std::get<1>(v)[4] == -1.0, "pressure entry");
      bool threw = false;
      std::tuple<std::vector<FieldVector<double, 3>>, std::vector<double>> wrongBlock;
      try { istlVectorBackend(wrongBlock).resize(th); } catch (const RangeError&) { threw = true; }
      CHECK(threw, "FieldVector<double,3> blocks for a 2-d velocity: RangeError");
      // the copy pair of the nested interpolate / DiscreteGlobalBasisFunction paths: flat image and back
      std::vector<double> flat;
      Dune::Functions::Impl::flattenNested(v, flat);
      CHECK(flat.size() == th.dimension() && flat[2 * 7 + 1] == 5.0 && flat[50 + 4] == -1.0, "flat image in container order");
      for (std::size_t i = 0; i < flat.size(); ++i) flat[i] = (double)i;
      const double* pf = flat.data();
      Dune::Functions::Impl::unflattenNested(v, pf);
      CHECK(pf == flat.data() + flat.size() && std::get<0>(v)[3][1] == 7.0 && std::get<1>(v)[2] == 52.0, "unflatten restores every entry");
      // resize through the sub-space basis (what interpolate(subspaceBasis(b, _0), x, f) does) and a second time
      std::tuple<std::vector<FieldVector<double, 2>>, std::vector<double>> w;
      Dune::Functions::Impl::resizeFor(w, subspaceBasis(th, _0));
      Dune::Functions::Impl::resizeFor(w, subspaceBasis(th, _1));
      CHECK(std::get<0>(w).size() == 25 && std::get<1>(w).size() == 9, "resize through a SubspaceBasis uses the root's sizes");
      std::array<std::vector<double>, 2> arr;
      TaylorHoodBasis<Grid::LeafGridView> thb(gv);
      istlVectorBackend(arr).resize(thb);
      CHECK(arr[0].size() == 50 && arr[1].size() == 9, "Array<FlatVector,2> of TaylorHoodBasis");
    }
    // ---- istlvectorbackendtest.cc main(): TupleVector / MultiTypeBlockVector -> std::tuple, BlockVector -> std::vector
    {
      using FV1 = FieldVector<double, 1>;
      checkISTLVectorBackend<std::tuple<std::vector<std::vector<double>>, std::vector<double>>, 2>("TV<V<V<double>>, V<double>>");
      checkISTLVectorBackend<std::tuple<std::vector<std::vector<FV1>>, std::vector<FV1>>, 2>("TV<V<BV<FV<double,1>>>, V<FV<double,1>>>");
      checkISTLVectorBackend<std::tuple<std::vector<std::array<FV1, 5>>, std::vector<double>>, 5>("TV<V<A<FV<double,1>,5>>, V<double>>");
      checkISTLVectorBackend<std::tuple<std::vector<FieldVector<double, 5>>, std::vector<FV1>>, 5>("MTBV<BV<FV<double,5>>, BV<FV<double,1>>>");
      checkISTLVectorBackend<std::tuple<std::vector<std::tuple<FV1, double, FV1>>, std::vector<FV1>>, 3>("MTBV<V<MTBV<FV1, double, FV1>>, BV<FV1>>");
      // an empty resizable entry where a coefficient is expected (:158-167)
      bool threw = false;
      std::tuple<std::vector<std::vector<std::vector<double>>>, std::vector<double>> tooDeep;
      try { istlVectorBackend(tooDeep).resize(GlobalBasisMoc<2>{}); } catch (const RangeError&) { threw = true; }
      CHECK(threw, "dynamically sized vector of size 0 in place of a scalar coefficient");
      threw = false;
      std::tuple<std::vector<double>, std::vector<double>> tooFlat;
      try { istlVectorBackend(tooFlat).resize(GlobalBasisMoc<2>{}); } catch (const RangeError&) { threw = true; }
      CHECK(threw, "scalar entry where size(prefix) = dim (:223-231)");
    }
    // ---- analyticgridviewfunctiontest.cc on YaspGrid<2> 10x10
    {
      Grid grid10(FieldVector<double, 2>(1.0), {{10, 10}});
      auto gridView = grid10.leafGridView();
      using Domain = FieldVector<double, 2>;
      auto f = [](const Domain& x) { return x[0]; };
      CHECK(std::abs(integrateGridViewFunction(gridView, AnalyticGridViewFunction<decltype(f), Grid::LeafGridView>(f, gridView)) - 0.5) < 1e-10, "manual construction");
      CHECK(std::abs(integrateGridViewFunction(gridView, makeAnalyticGridViewFunction(f, gridView)) - 0.5) < 1e-10, "makeAnalyticGridViewFunction");
      CHECK(std::abs(integrateGridViewFunction(gridView, makeGridViewFunction(f, gridView)) - 0.5) < 1e-10, "makeGridViewFunction");
      CHECK(std::abs(integrateGridViewFunction(gridView, makeAnalyticGridViewFunction(&x_component_A, gridView)) - 0.5) < 1e-10, "free function");
      auto gf = makeAnalyticGridViewFunction([](Domain) { return 1.0; }, gridView);
      auto gf2 = gf;                                   // copyable
      auto lf = localFunction(gf2);
      auto lf2 = lf;
      CHECK(!lf2.bound(), "not bound before bind()");
      lf2.bind(grid10.element(37));
      CHECK(lf2.bound() && lf2(Domain{0.0, 0.0}) == 1.0 && gf(Domain{0.0, 0.0}) == 1.0 && lf2.localContext().index == 37, "evaluation of function and local function");
      auto g = grid10.element(37).geometry();
      CHECK(std::abs(g.global(Domain{0.5, 0.5})[0] - 0.75) < 1e-15 && std::abs(g.global(Domain{0.5, 0.5})[1] - 0.35) < 1e-15, "element geometry");
      CHECK(std::abs(g.local(g.global(Domain{0.25, 0.75}))[1] - 0.75) < 1e-14 && std::abs(g.volume() - 0.01) < 1e-15, "local / volume");
      bool threw = false;
      lf2.unbind();
      try { lf2(Domain{0.0, 0.0}); } catch (const Exception&) { threw = true; }
      CHECK(threw, "unbound local function");
    }
    // ---- localfunctioncopytest.cc (the host part: a copy of a bound local function is bound) and the
    // construction checks of DiscreteGlobalBasisFunction
    {
      Grid grid2(FieldVector<double, 2>(1.0), {{2, 2}});
      LagrangeBasis<Grid::LeafGridView, 1> basis(grid2.leafGridView());
      std::vector<double> x;
      auto xbe = istlVectorBackend(x);
      xbe.resize(basis);
      xbe = 1.0;
      CHECK(x.size() == basis.dimension() && x[3] == 1.0, "BlockVector<double> sized and filled through the backend");
      auto f = makeDiscreteGlobalBasisFunction<double>(basis, xbe);
      auto localF = localFunction(f);
      for (const auto& element : elements(grid2.leafGridView())) {
        localF.bind(element);
        CHECK(localF.bound(), "Check if LocalFunction was bound.");
        auto localF_copy = localF;
        CHECK(localF_copy.bound(), "Check if copied LocalFunction was bound.");
      }
      bool threw = false;
      try { derivative(derivative(localF)); } catch (const NotImplemented&) { threw = true; }
      CHECK(threw, "second derivative throws NotImplemented (discreteglobalbasisfunction.hh:630-633)");
      threw = false;
      std::vector<double> tooShort(basis.dimension() - 1);
      try { makeDiscreteGlobalBasisFunction<double>(basis, tooShort); } catch (const RangeError&) { threw = true; }
      CHECK(threw, "coefficient vector of the wrong size");
      threw = false;
      auto th = makeBasis(grid2.leafGridView(), composite(power<2>(lagrange<2>(), blockedInterleaved()), lagrange<1>()));
      std::tuple<std::vector<FieldVector<double, 2>>, std::vector<double>> nested;
      try { makeDiscreteGlobalBasisFunction<double>(subspaceBasis(th, _1), nested); } catch (const RangeError&) { threw = true; }
      CHECK(threw, "empty nested coefficient container");
      istlVectorBackend(nested).resize(th);
      auto p = makeDiscreteGlobalBasisFunction<double>(subspaceBasis(th, _1), nested);
      using V2 = FieldVector<double, 2>;
      auto u = makeDiscreteGlobalBasisFunction<V2>(subspaceBasis(th, _0), nested);
      CHECK(p.numComponents() == 1 && u.numComponents() == 2, "range entries of the sub-trees");
    }
    // ---- update, dimension, errors
    {
      auto basis = makeBasis(gv, lagrange(3));
      CHECK(basis.dimension() == 49, "Q3 on 2x2");
      Grid other(FieldVector<double, 2>(1.0), {{5, 3}});
      basis.update(other.leafGridView());
      CHECK(basis.dimension() == 16 * 10 && basis.size() == 160, "after update");
      bool threw = false;
      try { makeBasis(gv, lagrange(18)); } catch (const RangeError&) { threw = true; }
      CHECK(threw, "order > 17 throws RangeError");
    }
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 2;
  }
  std::printf("%s\n", failures ? "host_only_test FAILED" : "host_only_test passed");
  return failures ? 1 : 0;
}
