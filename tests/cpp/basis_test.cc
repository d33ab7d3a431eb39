// basis_test.cc — the reference's generic basis harness, restated against the façade:
// checkBasis (dune/functions/functionspacebases/test/basistest.hh:692-713) with
// checkLocalView (:275-383: size <= maxSize, every local index exactly once),
// checkBasisIndices (:190-227: consecutive index tree from (0,...,0), size(prefix)
// consistent, dimension() == number of indices), and the known answers of
// lagrangebasistest.cc, lagrangedgbasistest.cc:31-66, taylorhoodbasistest.cc:35-146,
// compositebasistest.cc:66-122, gridviewfunctionspacebasistest.cc:115-183.
// Runs on the GPU through libdfx.so; exit code 0 = all checks passed.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <set>
#include <tuple>

#include <dfx/dune_functions.hh>

using namespace Dune;
using namespace Dune::Functions;
using namespace Dune::Functions::BasisFactory;

static int failures = 0;
#define CHECK(cond, what)                                                        \
  do { if (!(cond)) { std::printf("FAILED %s (%s:%d)\n", what, __FILE__, __LINE__); ++failures; } } while (0)

template <class Basis>
void checkBasis(const Basis& basis, const char* name) {
  auto localView = basis.localView();
  std::set<MultiIndex> all;
  for (const auto& e : elements(basis.gridView())) {
    localView.bind(e);
    CHECK(localView.size() <= localView.maxSize(), "localView.size() <= maxSize()");
    std::set<MultiIndex> local;
    for (std::size_t i = 0; i < localView.size(); ++i) local.insert(localView.index(i));
    CHECK(local.size() == localView.size(), "each local index exactly once");
    all.insert(local.begin(), local.end());
  }
  CHECK(all.size() == basis.dimension(), "dimension() == number of multi-indices");
  // consecutive index tree + size(prefix)
  std::map<std::vector<std::size_t>, std::set<std::size_t>> children;
  for (const auto& m : all)
    for (std::size_t p = 0; p < m.size(); ++p) {
      std::vector<std::size_t> prefix(m.d_, m.d_ + p);
      children[prefix].insert(m[p]);
    }
  for (const auto& [prefix, ch] : children) {
    CHECK(*ch.begin() == 0 && *ch.rbegin() + 1 == ch.size(), "digits below a prefix are consecutive from 0");
    CHECK(basis.size(prefix) == ch.size(), "size(prefix) consistent");
  }
  for (const auto& m : all) {
    std::vector<std::size_t> full(m.d_, m.d_ + m.size());
    CHECK(basis.size(full) == 0, "size(full multi-index) == 0");
  }
  std::printf("checkBasis %-34s dimension %8zu  %s\n", name, (std::size_t)basis.dimension(), failures ? "FAILED" : "ok");
}

// containerdescriptortest.cc:87-131: checkSize (the descriptor's size below every prefix equals
// basis.size(prefix)) and checkMultiIndices (every multi-index walks through the descriptor)
template <class CD, class Basis>
void checkDescriptorSize(const CD& cd, const Basis& basis, std::vector<std::size_t> prefix, std::size_t maxLen) {
  CHECK(basis.size(prefix) == cd.size(), "size(prefix) == containerDescriptor size at prefix");
  if (prefix.size() < maxLen)
    for (std::size_t i = 0; i < cd.size(); ++i) {
      auto p = prefix; p.push_back(i);
      checkDescriptorSize(cd[i], basis, p, maxLen);
    }
}
template <class Basis>
void checkContainerDescriptor(const Basis& basis, const char* name, const char* expectedType = nullptr) {
  auto cd = basis.containerDescriptor();
  CHECK(!cd.isUnknown(), "container descriptor known");
  std::size_t maxLen = 0;
  auto localView = basis.localView();
  for (const auto& e : elements(basis.gridView())) {
    localView.bind(e);
    for (std::size_t i = 0; i < localView.size(); ++i) {
      auto mi = localView.index(i);
      maxLen = std::max<std::size_t>(maxLen, mi.size());
      const auto* d = &cd;
      for (std::size_t j = 0; j < mi.size(); ++j) {
        CHECK(mi[j] < d->size(), "mi[j] < cd.size()");
        if (mi[j] >= d->size()) break;
        d = &(*d)[mi[j]];
      }
      CHECK(d->isValue(), "a full multi-index addresses a Value");
    }
  }
  checkDescriptorSize(cd, basis, {}, maxLen);
  if (expectedType) CHECK(cd.type() == expectedType, "descriptor type as in the reference");
  std::printf("containerDescriptor %-26s %s  %s\n", name, cd.type().c_str(), failures ? "FAILED" : "ok");
}

// tensor Gauss rule with 3 points per direction on [0,1]^dim
template <int dim>
void gauss3(std::vector<FieldVector<double, dim>>& x, std::vector<double>& w) {
  const double p[3] = {0.5 - 0.5 * std::sqrt(0.6), 0.5, 0.5 + 0.5 * std::sqrt(0.6)}, ww[3] = {5.0 / 18, 8.0 / 18, 5.0 / 18};
  int n = 1; for (int j = 0; j < dim; ++j) n *= 3;
  for (int q = 0; q < n; ++q) {
    FieldVector<double, dim> pt; double wt = 1; int r = q;
    for (int j = dim - 1; j >= 0; --j) { pt[j] = p[r % 3]; wt *= ww[r % 3]; r /= 3; }   // last coordinate fastest
    x.push_back(pt); w.push_back(wt);
  }
}

// gridviewfunctionspacebasistest.cc:115-183: the interpolant of x_0 integrates to 0.5
template <class Basis>
void checkIntegralOfX0(const Basis& basis, const char* name) {
  constexpr int dim = Basis::GridView::dimension;
  std::vector<double> x;
  interpolate(basis, x, [](const FieldVector<double, dim>& p) { return p[0]; });
  auto u = makeDiscreteGlobalBasisFunction<double>(basis, x);
  std::vector<FieldVector<double, dim>> pts; std::vector<double> w, v;
  gauss3<dim>(pts, w);
  u.evaluateAll(pts, v);
  const std::size_t ne = basis.gridView().grid().numElements();
  double integral = 0;
  for (std::size_t e = 0; e < ne; ++e)
    for (std::size_t q = 0; q < pts.size(); ++q) integral += v[e * pts.size() + q] * w[q] / ne;
  CHECK(std::abs(integral - 0.5) < 1e-10, "integral of interpolated x_0 == 0.5");
  std::printf("integral   %-34s %.15f\n", name, integral);
}

int main() {
  using Grid2 = YaspGrid<2>;
  using Grid3 = YaspGrid<3>;
  Grid2 grid2(FieldVector<double, 2>(1.0), {{10, 10}});
  Grid3 grid3(FieldVector<double, 3>(1.0), {{3, 4, 2}});
  auto gv2 = grid2.leafGridView();
  auto gv3 = grid3.leafGridView();
  try {
    // lagrangebasistest.cc: orders 1..4, compile-time and run-time order
    checkBasis(LagrangeBasis<Grid2::LeafGridView, 1>(gv2), "LagrangeBasis<1> 2d");
    checkBasis(LagrangeBasis<Grid2::LeafGridView, 2>(gv2), "LagrangeBasis<2> 2d");
    checkBasis(LagrangeBasis<Grid3::LeafGridView, 3>(gv3), "LagrangeBasis<3> 3d");
    checkBasis(makeBasis(gv3, lagrange(4)), "lagrange(4) 3d");
    checkBasis(makeBasis(gv2, lagrange<0>()), "lagrange<0> 2d");
    // lagrangedgbasistest.cc:31-66
    checkBasis(LagrangeDGBasis<Grid2::LeafGridView, 2>(gv2), "LagrangeDGBasis<2> 2d");
    checkBasis(makeBasis(gv2, lagrangeDG(2)), "lagrangeDG(2) 2d");
    // taylorhoodbasistest.cc:35-146
    checkBasis(TaylorHoodBasis<Grid2::LeafGridView>(gv2), "TaylorHoodBasis 2d");
    checkBasis(makeBasis(gv3, taylorHood()), "taylorHood() 3d");
    // compositebasistest.cc:66-122 / makebasistest.cc:65-153: every merging strategy
    checkBasis(makeBasis(gv2, composite(power<2>(lagrange<2>(), flatLexicographic()), lagrange<1>(), flatLexicographic())), "FL(FL) Taylor-Hood");
    checkBasis(makeBasis(gv2, composite(power<2>(lagrange<2>(), flatInterleaved()), lagrange<1>(), blockedLexicographic())), "BL(FI) Taylor-Hood");
    checkBasis(makeBasis(gv2, composite(power<2>(lagrange<2>(), blockedLexicographic()), lagrange<1>(), flatLexicographic())), "FL(BL) Taylor-Hood");
    checkBasis(makeBasis(gv3, composite(power<3>(lagrange<2>(), blockedInterleaved()), lagrange<1>())), "BL(BI) Taylor-Hood 3d");
    checkBasis(makeBasis(gv2, power<2>(composite(lagrange<1>(), lagrangeDG<1>(), blockedLexicographic()), blockedInterleaved())), "BI(BL(Q1,DG1))");
    checkBasis(makeBasis(gv2, composite(power<2>(power<2>(lagrange<1>(), flatLexicographic()), blockedInterleaved()), lagrange<2>())), "BL(BI(FL(Q1)),Q2)");

    // containerdescriptortest.cc:236-262 on its 2x2 grid
    {
      Grid2 small(FieldVector<double, 2>(1.0), {{2, 2}});
      auto gvs = small.leafGridView();
      checkContainerDescriptor(makeBasis(gvs, lagrange<1>()), "lagrange<1>", "UniformVector<Value>");
      checkContainerDescriptor(makeBasis(gvs, power<2>(lagrange<1>(), blockedInterleaved())), "power<2>(Q1,BI)",
                               "UniformVector<UniformArray<Value,2>>");
      checkContainerDescriptor(makeBasis(gvs, composite(lagrange<1>(), lagrange<2>(), blockedLexicographic())), "composite(Q1,Q2,BL)",
                               "Array<UniformVector<Value>,2>");
      checkContainerDescriptor(makeBasis(gvs, power<2>(power<3>(lagrange<2>(), blockedLexicographic()), flatLexicographic())), "FL(BL(Q2))");
      checkContainerDescriptor(makeBasis(gvs, power<2>(power<3>(lagrange<2>(), blockedInterleaved()), blockedLexicographic())), "BL(BI(Q2))");
      checkContainerDescriptor(makeBasis(gvs, power<2>(power<3>(lagrange<2>(), flatInterleaved()), blockedInterleaved())), "BI(FI(Q2))");
      checkContainerDescriptor(TaylorHoodBasis<Grid2::LeafGridView>(gvs), "TaylorHoodBasis", "Array<UniformVector<Value>,2>");
      checkContainerDescriptor(makeBasis(gv3, composite(power<3>(lagrange<2>(), blockedInterleaved()), lagrange<1>())), "BL(BI) Taylor-Hood 3d",
                               "Tuple<UniformVector<UniformArray<Value,3>>,UniformVector<Value>>");
    }

    checkIntegralOfX0(makeBasis(gv2, lagrange<0>()), "lagrange<0> 2d");
    checkIntegralOfX0(LagrangeBasis<Grid2::LeafGridView, 3>(gv2), "LagrangeBasis<3> 2d");
    checkIntegralOfX0(LagrangeBasis<Grid3::LeafGridView, 4>(gv3), "LagrangeBasis<4> 3d");
    checkIntegralOfX0(LagrangeDGBasis<Grid3::LeafGridView, 2>(gv3), "LagrangeDGBasis<2> 3d");

    // taylorhoodbasistest.cc:58-64: pressure DOFs addressed by indexSet.index(vertex) carry x_0
    {
      TaylorHoodBasis<Grid2::LeafGridView> th(gv2);
      std::vector<double> x(th.dimension(), 0.0);
      auto pressure = subspaceBasis(th, Indices::_1);
      const std::size_t offset = 2 * 21 * 21;         // flat position of multi-index (1, j): after (0, *)
      for (int vy = 0; vy <= 10; ++vy)
        for (int vx = 0; vx <= 10; ++vx) x[offset + vx + 11 * vy] = vx / 10.0;
      auto p = makeDiscreteGlobalBasisFunction<double>(pressure, x);
      std::vector<FieldVector<double, 2>> pts; std::vector<double> w, v;
      gauss3<2>(pts, w);
      p.evaluateAll(pts, v);
      double integral = 0;
      for (std::size_t e = 0; e < 100; ++e) for (std::size_t q = 0; q < 9; ++q) integral += v[e * 9 + q] * w[q] / 100;
      CHECK(std::abs(integral - 0.5) < 1e-10, "Taylor-Hood pressure integral == 0.5");
    }
    // checkInterpolationConsistency (gridfunctions/test/discreteglobalbasisfunctiontest.cc:49-77):
    // interpolating the discrete function of a coefficient vector returns the vector, both through
    // its local function (grid function) and through a lambda that forces evaluation at the nodes
    {
      auto basis = makeBasis(gv2, power<2>(lagrange<2>(), blockedInterleaved()));
      std::vector<double> x;
      interpolate(basis, x, [](const FieldVector<double, 2>& p) { return FieldVector<double, 2>{p[1], p[0]}; });
      auto f = makeDiscreteGlobalBasisFunction<FieldVector<double, 2>>(basis, x);
      std::vector<double> y;
      interpolate(basis, y, f);
      double diff = 0;
      for (std::size_t i = 0; i < x.size(); ++i) diff = std::max(diff, std::abs(x[i] - y[i]));
      CHECK(y.size() == x.size() && diff < 1e-10, "interpolation of a DiscreteGlobalBasisFunction reproduces its coefficients");
      // "By wrapping into a lambda, we force interpolate to use the global evaluation of f" (:58-59)
      auto fGlobal = [&](const FieldVector<double, 2>& p) { return f(p); };
      std::vector<double> yg;
      interpolate(basis, yg, fGlobal);
      diff = 0;
      for (std::size_t i = 0; i < x.size(); ++i) diff = std::max(diff, std::abs(x[i] - yg[i]));
      CHECK(yg.size() == x.size() && diff < 1e-10, "interpolation through the global evaluation reproduces the coefficients");
      bool outside = false;
      try { f(FieldVector<double, 2>{1.5, 0.5}); } catch (const Exception&) { outside = true; }
      CHECK(outside, "evaluation outside of the grid throws");
      // Q2 function into the Q1 basis of the same grid: the vertex values
      LagrangeBasis<Grid2::LeafGridView, 2> q2(gv2);
      LagrangeBasis<Grid2::LeafGridView, 1> q1(gv2);
      std::vector<double> c2, c1, c1ref;
      interpolate(q2, c2, Registry::gaussian(1.0));
      interpolate(q1, c1ref, Registry::gaussian(1.0));
      interpolate(q1, c1, makeDiscreteGlobalBasisFunction<double>(q2, c2));
      diff = 0;
      for (std::size_t i = 0; i < c1.size(); ++i) diff = std::max(diff, std::abs(c1[i] - c1ref[i]));
      CHECK(diff < 1e-14, "Q2 discrete function interpolated into Q1 = nodal values");
    }
    // lagrangebasistest.cc:134-310 (createNonUniformCubeGrid, orders 1..5): cube grids whose vertex ids are
    // not ordered like the lattice, so that edge and facet DOFs of order >= 3 have to be permuted
    // (Experimental::LagrangeFaceDOFPermutation, lagrangebasis.hh:429-697) for the basis to conform.
    // Conformity is checked through what it implies: a cubic is reproduced exactly at points on both
    // sides of every element interface.
    {
      Grid2 tw2(FieldVector<double, 2>(1.0), {{3, 2}});
      Grid3 tw3(FieldVector<double, 3>(1.0), {{2, 2, 2}});
      std::vector<std::uint64_t> ids2(12), ids3(27);
      for (std::size_t v = 0; v < ids2.size(); ++v) ids2[v] = (v * 7 + 3) % 12;     // 7 is coprime to 12 and 27:
      for (std::size_t v = 0; v < ids3.size(); ++v) ids3[v] = (v * 7 + 3) % 27;     // permutations of the ids
      tw2.setGlobalVertexIds(ids2);
      tw3.setGlobalVertexIds(ids3);
      auto cubic2 = [](const FieldVector<double, 2>& p) { return p[0] * p[0] * p[0] + 2 * p[0] * p[1] * p[1] - p[1]; };
      auto cubic3 = [](const FieldVector<double, 3>& p) { return p[0] * p[0] * p[0] + 2 * p[0] * p[1] * p[1] - p[1] + p[2] * p[2] * p[0]; };
      for (int order = 1; order <= 5; ++order) {
        auto b2 = makeBasis(tw2.leafGridView(), lagrange(order));
        auto b3 = makeBasis(tw3.leafGridView(), lagrange(order));
        checkBasis(b2, "lagrange(order) 2d, permuted vertex ids");
        checkBasis(b3, "lagrange(order) 3d, permuted vertex ids");
        if (order < 3) continue;
        std::vector<double> x2, x3;
        interpolate(b2, x2, cubic2);
        interpolate(b3, x3, cubic3);
        auto f2 = makeDiscreteGlobalBasisFunction<double>(b2, x2);
        auto f3 = makeDiscreteGlobalBasisFunction<double>(b3, x3);
        double err = 0;
        for (int i = 0; i < 40; ++i) {
          FieldVector<double, 2> p{std::fmod(0.137 * i + 0.05, 1.0), std::fmod(0.291 * i + 0.11, 1.0)};
          FieldVector<double, 3> q{p[0], p[1], std::fmod(0.173 * i + 0.07, 1.0)};
          err = std::max(err, std::abs(f2(p)[0] - cubic2(p)));
          err = std::max(err, std::abs(f3(q)[0] - cubic3(q)));
        }
        CHECK(err < 1e-12, "cubic reproduced by lagrange(order >= 3) with permuted face DOFs");
      }
      // without ids the same grid gives another numbering of the same space: coefficients are a permutation
      Grid3 plain(FieldVector<double, 3>(1.0), {{2, 2, 2}});
      auto bp = makeBasis(plain.leafGridView(), lagrange(4));
      auto bt = makeBasis(tw3.leafGridView(), lagrange(4));
      std::vector<double> xp, xt;
      interpolate(bp, xp, cubic3);
      interpolate(bt, xt, cubic3);
      CHECK(xp != xt, "permuted vertex ids change the numbering of order-4 DOFs");
      std::sort(xp.begin(), xp.end()); std::sort(xt.begin(), xt.end());
      CHECK(xp == xt, "same coefficients in another order");
    }
    // nested coefficient containers as in examples/stokes-taylorhood.cc:323-346 (MultiTypeBlockVector of a
    // BlockVector<FieldVector<double,dim>> and a BlockVector<double>; here std::tuple / std::vector) and the
    // Array<FlatVector,2> of TaylorHoodBasis: sized like ISTLVectorBackend::resize (istlvectorbackend.hh:153-233),
    // entries in lexicographic order = the flat container order
    {
      using VelocityVector = std::vector<FieldVector<double, 3>>;
      using PressureVector = std::vector<double>;
      using VectorType = std::tuple<VelocityVector, PressureVector>;
      auto th = makeBasis(gv3, composite(power<3>(lagrange<2>(), blockedInterleaved()), lagrange<1>()));
      auto velocity = [](const FieldVector<double, 3>& p) { return FieldVector<double, 3>{p[1], -p[0], p[2] * p[0]}; };
      auto pressure = [](const FieldVector<double, 3>& p) { return p[0] + 2 * p[1]; };
      VectorType x;
      std::vector<double> flat;
      interpolate(subspaceBasis(th, Indices::_0), x, velocity);
      interpolate(subspaceBasis(th, Indices::_1), x, pressure);
      interpolate(subspaceBasis(th, Indices::_0), flat, velocity);
      interpolate(subspaceBasis(th, Indices::_1), flat, pressure);
      const std::size_t n2 = 7 * 9 * 5, n1 = 4 * 5 * 3;
      CHECK(std::get<0>(x).size() == n2 && std::get<1>(x).size() == n1, "nested container sized from the container descriptor");
      bool same = flat.size() == 3 * n2 + n1;
      for (std::size_t i = 0; same && i < n2; ++i) for (int c = 0; c < 3; ++c) same = std::get<0>(x)[i][c] == flat[3 * i + c];
      for (std::size_t j = 0; same && j < n1; ++j) same = std::get<1>(x)[j] == flat[3 * n2 + j];
      CHECK(same, "nested entries in lexicographic order are the flat container order");
      // a discrete function over the nested container evaluates like the one over the flat vector
      std::vector<FieldVector<double, 3>> pts; std::vector<double> w, vn, vf;
      gauss3<3>(pts, w);
      makeDiscreteGlobalBasisFunction<FieldVector<double, 3>>(subspaceBasis(th, Indices::_0), x).evaluateAll(pts, vn);
      makeDiscreteGlobalBasisFunction<FieldVector<double, 3>>(subspaceBasis(th, Indices::_0), flat).evaluateAll(pts, vf);
      CHECK(vn == vf && !vn.empty(), "DiscreteGlobalBasisFunction over a nested coefficient container");
      // HierarchicNodeToRangeMap on a nested range (hierarchicnodetorangemap.hh:33-48): the WHOLE Taylor-Hood tree
      // with f returning tuple<FieldVector<double,3>, double> — leaf (0, c) takes get<0>(y)[c], leaf (1) takes
      // get<1>(y) (the range type of examples/stokes-taylorhood.cc:419-438) — equals the two sub-space calls
      {
        using Range = std::tuple<FieldVector<double, 3>, double>;
        auto both = [&](const FieldVector<double, 3>& p) { return Range{velocity(p), pressure(p)}; };
        std::vector<double> whole;
        interpolate(th, whole, both);
        CHECK(whole == flat, "interpolate(tree, f -> tuple<FieldVector<3>, double>) resolves y[treePath]");
        VectorType xw;
        interpolate(th, xw, both);
        CHECK(std::get<0>(xw) == std::get<0>(x) && std::get<1>(xw) == std::get<1>(x), "the same into the nested container");
        // and back: the typed Range of the discrete function of the whole tree
        auto uh = makeDiscreteGlobalBasisFunction<Range>(th, flat);
        const FieldVector<double, 3> xg{0.4, 0.55, 0.3};
        const Range y = uh.range(xg);
        const auto flatValues = uh(xg);
        CHECK(flatValues.size() == 4, "four range entries");
        CHECK(std::get<0>(y)[0] == flatValues[0] && std::get<0>(y)[1] == flatValues[1] && std::get<0>(y)[2] == flatValues[2] &&
              std::get<1>(y) == flatValues[3], "DiscreteGlobalBasisFunction<tuple<FieldVector<3>, double>>::range assigns y[treePath]");
        // velocity (p1, -p0, p2 p0) is reproduced exactly by Q2, the pressure by Q1
        CHECK(std::abs(std::get<0>(y)[0] - xg[1]) < 1e-13 && std::abs(std::get<0>(y)[2] - xg[2] * xg[0]) < 1e-13 &&
              std::abs(std::get<1>(y) - (xg[0] + 2 * xg[1])) < 1e-13, "values of the typed range");
      }
      // power<2>(power<2>(lagrange<1>)): tree path (i, j) -> y[i][j], a FieldMatrix range (row-major, flatvectorview.hh:63-78)
      {
        auto pp = makeBasis(gv3, power<2>(power<2>(lagrange<1>(), flatInterleaved()), flatLexicographic()));
        auto m = [](const FieldVector<double, 3>& p) { return FieldMatrix<double, 2, 2>{{{p[0], p[1]}, {p[2], 1.0 + p[0] * p[1]}}}; };
        std::vector<double> xm, xs(pp.dimension(), 0.0);
        interpolate(pp, xm, m);
        interpolate(subspaceBasis(pp, Indices::_0, Indices::_0), xs, [](const FieldVector<double, 3>& p) { return p[0]; });
        interpolate(subspaceBasis(pp, Indices::_0, Indices::_1), xs, [](const FieldVector<double, 3>& p) { return p[1]; });
        interpolate(subspaceBasis(pp, Indices::_1, Indices::_0), xs, [](const FieldVector<double, 3>& p) { return p[2]; });
        interpolate(subspaceBasis(pp, Indices::_1, Indices::_1), xs, [](const FieldVector<double, 3>& p) { return 1.0 + p[0] * p[1]; });
        CHECK(xm == xs && !xm.empty(), "FieldMatrix range: leaf (i, j) takes y[i][j]");
        auto um = makeDiscreteGlobalBasisFunction<FieldMatrix<double, 2, 2>>(pp, xm);
        const auto ym = um.range(FieldVector<double, 3>{0.25, 0.5, 0.75});
        CHECK(std::abs(ym[0][0] - 0.25) < 1e-13 && std::abs(ym[0][1] - 0.5) < 1e-13 && std::abs(ym[1][0] - 0.75) < 1e-13, "FieldMatrix range of the discrete function");
      }
      // masked interpolation with a nested bit vector of the same shape
      std::tuple<std::vector<std::array<char, 3>>, std::vector<char>> bits;
      std::get<0>(bits).assign(n2, std::array<char, 3>{1, 0, 1});
      std::get<1>(bits).assign(n1, 0);
      VectorType y = x;
      interpolate(subspaceBasis(th, Indices::_0), y, [](const FieldVector<double, 3>&) { return FieldVector<double, 3>{7.0, 7.0, 7.0}; }, bits);
      same = true;
      for (std::size_t i = 0; same && i < n2; ++i)
        same = std::get<0>(y)[i][0] == 7.0 && std::get<0>(y)[i][1] == std::get<0>(x)[i][1] && std::get<0>(y)[i][2] == 7.0;
      CHECK(same && std::get<1>(y) == std::get<1>(x), "nested bit vector masks the nested coefficients");
      // TaylorHoodBasis: Array<FlatVector,2>
      TaylorHoodBasis<Grid3::LeafGridView> thb(gv3);
      std::array<std::vector<double>, 2> z;
      interpolate(subspaceBasis(thb, Indices::_1), z, pressure);
      CHECK(z[0].size() == 3 * n2 && z[1].size() == n1, "Array<FlatVector,2> of TaylorHoodBasis");
      same = true;
      for (std::size_t j = 0; same && j < n1; ++j) same = z[1][j] == flat[3 * n2 + j];
      CHECK(same, "pressure coefficients in the second block");
      // a statically sized level of the wrong size: RangeError (istlvectorbackend.hh:170, :208, :232)
      bool threwNested = false;
      std::tuple<std::vector<double>> wrong;
      try { interpolate(th, wrong, Registry::constant(1.0)); } catch (const RangeError&) { threwNested = true; }
      CHECK(threwNested, "tuple of the wrong size throws RangeError");
    }
    // basis.update(gridView) (defaultglobalbasis.hh:137-141; lagrangebasistest.cc:163-175 updates and checks again)
    {
      auto basis = makeBasis(gv3, composite(power<3>(lagrange<2>(), blockedInterleaved()), lagrange<1>()));
      const std::size_t before = basis.dimension();
      Grid3 other(FieldVector<double, 3>(1.0), {{2, 5, 3}});
      basis.update(other.leafGridView());
      CHECK(basis.dimension() != before && basis.dimension() == 3 * 5 * 11 * 7 + 3 * 6 * 4, "dimension after update");
      checkBasis(basis, "BL(BI) Taylor-Hood 3d, updated");
      checkContainerDescriptor(basis, "updated Taylor-Hood");
    }
    // error behaviour: RangeError for an unsupported order (lagrangebasis.hh:795-796)
    bool threw = false;
    try { makeBasis(gv3, lagrange(18)); } catch (const RangeError&) { threw = true; }
    CHECK(threw, "order > 17 throws RangeError");
    // derivative of derivative: NotImplemented (discreteglobalbasisfunction.hh:630-633)
    {
      LagrangeBasis<Grid2::LeafGridView, 2> b(gv2);
      std::vector<double> x; interpolate(b, x, Registry::gaussian(1.0));
      auto u = makeDiscreteGlobalBasisFunction<double>(b, x);
      auto lf = localFunction(u);
      threw = false;
      try { derivative(derivative(lf)); } catch (const NotImplemented&) { threw = true; }
      CHECK(threw, "second derivative throws NotImplemented");
    }
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 2;
  }
  std::printf("%s\n", failures ? "basis_test FAILED" : "basis_test passed");
  return failures ? 1 : 0;
}
