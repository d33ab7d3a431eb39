// dune_functions.hh — C++ façade that keeps the dune-functions names for the element-loop
// path on top of the C ABI (include/dfx.h).  Host code written against the reference
//
//     using namespace Dune::Functions::BasisFactory;
//     auto basis = makeBasis(gridView, composite(power<dim>(lagrange<2>(), flatLexicographic()),
//                                                lagrange<1>(), flatLexicographic()));
//     interpolate(subspaceBasis(basis, _0), x, f);
//     auto u  = makeDiscreteGlobalBasisFunction<double>(basis, x);
//     auto lf = localFunction(u);  lf.bind(element);  lf(xLocal);  derivative(lf)(xLocal);
//     auto lv = basis.localView(); lv.bind(element);  lv.index(i);
//
// compiles against this header (plus the minimal Dune::YaspGrid / FieldVector stand-ins
// below, since dune-grid / dune-common are not part of this repository) and computes on
// the GPU.  Reference interfaces mirrored (dune-functions checkout):
//   BasisFactory::makeBasis            functionspacebases/defaultglobalbasis.hh:205-209
//   lagrange / lagrangeDG / power / composite / taylorHood
//                                      lagrangebasis.hh:1006-1027, lagrangedgbasis.hh:150-170,
//                                      powerbasis.hh:140-171, compositebasis.hh:420-445,
//                                      taylorhoodbasis.hh:330-335
//   index merging tags                 basistags.hh:192-225
//   GlobalBasis / LocalView concepts   concepts.hh:222-273, defaultlocalview.hh:95-159
//   interpolate                        interpolate.hh:204-302
//   makeDiscreteGlobalBasisFunction    gridfunctions/discreteglobalbasisfunction.hh:457-478
//   subspaceBasis                      subspacebasis.hh:147-157
//   SubspaceLocalView                  subspacelocalview.hh:32-153
//   subEntityDOFs / SubEntityDOFs      subentitydofs.hh:45-236
//   forEachBoundaryDOF                 boundarydofs.hh:39-127
//   containerDescriptor                defaultglobalbasis.hh:180-186, containerdescriptors.hh:91-209
//   istlVectorBackend                  backends/istlvectorbackend.hh:104-312 (on nested standard containers)
//   AnalyticGridViewFunction, makeGridViewFunction
//                                      gridfunctions/analyticgridviewfunction.hh:33-245, gridviewfunction.hh:68-101
//   basis.update(gridView)             defaultglobalbasis.hh:137-141
// Errors: non-zero dfx_status becomes Dune::RangeError / Dune::NotImplemented /
// Dune::Exception, like the DUNE_THROW sites of the reference.  No CPU fallback: every
// compute call goes through libdfx.so and fails if no device is present.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../dfx.h"

namespace Dune {

struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct RangeError : Exception { using Exception::Exception; };
struct NotImplemented : Exception { using Exception::Exception; };

// ---- dune-common stand-ins -------------------------------------------------------------
template <class K, int n>
class FieldVector : public std::array<K, n> {
 public:
  FieldVector() { this->fill(K(0)); }
  FieldVector(const K& v) { this->fill(v); }
  FieldVector(std::initializer_list<K> l) { int i = 0; for (auto& v : l) (*this)[i++] = v; }
  K two_norm2() const { K r = 0; for (auto& v : *this) r += v * v; return r; }
  K two_norm() const { return std::sqrt(two_norm2()); }
  static constexpr int dimension = n;
};
template <class K, int r, int c>
using FieldMatrix = std::array<std::array<K, c>, r>;

namespace Indices {
template <std::size_t i> using index_constant = std::integral_constant<std::size_t, i>;
inline constexpr index_constant<0> _0{};
inline constexpr index_constant<1> _1{};
inline constexpr index_constant<2> _2{};
inline constexpr index_constant<3> _3{};
}  // namespace Indices

// ---- dune-grid stand-in: unrefined equidistant YaspGrid ---------------------------------
template <int dim>
class YaspGrid {
 public:
  static constexpr int dimension = dim;
  using ctype = double;
  // AxisAlignedCubeGeometry of an element: lower = origin + i h, upper = origin + (i+1) h,
  // global(x) = lower + x (upper - lower)   [dune-grid YaspGeometry / dune-geometry AxisAlignedCubeGeometry]
  struct Geometry {
    FieldVector<double, dim> lower, upper;
    FieldVector<double, dim> global(const FieldVector<double, dim>& local) const {
      FieldVector<double, dim> x;
      for (int j = 0; j < dim; ++j) x[j] = lower[j] + local[j] * (upper[j] - lower[j]);
      return x;
    }
    FieldVector<double, dim> local(const FieldVector<double, dim>& global) const {
      FieldVector<double, dim> x;
      for (int j = 0; j < dim; ++j) x[j] = (global[j] - lower[j]) / (upper[j] - lower[j]);
      return x;
    }
    double integrationElement(const FieldVector<double, dim>&) const { return volume(); }
    double volume() const { double v = 1; for (int j = 0; j < dim; ++j) v *= upper[j] - lower[j]; return v; }
    FieldVector<double, dim> center() const { return global(FieldVector<double, dim>(0.5)); }
  };
  struct Element {
    std::size_t index;                     // position in elements(gridView), x fastest
    std::array<int, dim> coord;
    const YaspGrid* grid = nullptr;
    Geometry geometry() const {
      Geometry g;
      for (int j = 0; j < dim; ++j) {
        const double h = (grid->L_[j] - grid->lower_[j]) / grid->N_[j];
        g.lower[j] = grid->lower_[j] + coord[j] * h;
        g.upper[j] = grid->lower_[j] + (coord[j] + 1) * h;
      }
      return g;
    }
  };
  class LeafGridView {
   public:
    static constexpr int dimension = dim;
    using Grid = YaspGrid;
    explicit LeafGridView(const YaspGrid* g) : grid_(g) {}
    const YaspGrid& grid() const { return *grid_; }
    std::size_t size(int codim) const {
      std::size_t r = 0;
      for (unsigned s = 0; s < (1u << dim); ++s) {
        if (__builtin_popcount(s) != dim - codim) continue;
        std::size_t c = 1;
        for (int j = 0; j < dim; ++j) c *= grid_->N_[j] + (((s >> j) & 1) ? 0 : 1);
        r += c;
      }
      return r;
    }
    const YaspGrid* grid_;
  };
  YaspGrid(const FieldVector<double, dim>& L, const std::array<int, dim>& N) : L_(L), N_(N) { lower_.fill(0.0); }
  YaspGrid(const FieldVector<double, dim>& lowerLeft, const FieldVector<double, dim>& upperRight,
           const std::array<int, dim>& N) : L_(upperRight), N_(N) { for (int j = 0; j < dim; ++j) lower_[j] = lowerLeft[j]; }
  LeafGridView leafGridView() const { return LeafGridView(this); }
  std::size_t numElements() const { std::size_t n = 1; for (int j = 0; j < dim; ++j) n *= N_[j]; return n; }
  Element element(std::size_t e) const {
    Element el; el.index = e; el.grid = this;
    for (int j = 0; j < dim; ++j) { el.coord[j] = (int)(e % N_[j]); e /= N_[j]; }
    return el;
  }
  // grid.globalIdSet() restricted to vertices: what Experimental::LagrangeFaceDOFPermutation orients edges
  // and facets with (lagrangebasis.hh:787).  YaspGrid's own ids are monotone in lexicographic position
  // (empty table).  setGlobalVertexIds stands for a grid manager that numbers differently (UGGrid /
  // ALUGrid cube meshes): one id per vertex, lexicographic, x fastest.  Bases read it at construction.
  void setGlobalVertexIds(std::vector<std::uint64_t> ids) {
    std::size_t nv = 1; for (int j = 0; j < dim; ++j) nv *= (std::size_t)N_[j] + 1;
    if (!ids.empty() && ids.size() != nv) throw RangeError("one id per vertex expected");
    vertexIds_ = std::move(ids);
  }
  const std::vector<std::uint64_t>& globalVertexIds() const { return vertexIds_; }
  FieldVector<double, dim> L_, lower_;
  std::array<int, dim> N_;
  std::vector<std::uint64_t> vertexIds_;
};

// elements(gridView): the reference's range-based element loop
template <int dim>
class ElementRange {
 public:
  using Grid = YaspGrid<dim>;
  struct iterator {
    const Grid* g; std::size_t e;
    typename Grid::Element operator*() const { return g->element(e); }
    iterator& operator++() { ++e; return *this; }
    bool operator!=(const iterator& o) const { return e != o.e; }
  };
  explicit ElementRange(const Grid* g) : g_(g) {}
  iterator begin() const { return {g_, 0}; }
  iterator end() const { return {g_, g_->numElements()}; }
 private:
  const Grid* g_;
};
template <class GV>
ElementRange<GV::dimension> elements(const GV& gv) { return ElementRange<GV::dimension>(gv.grid_); }

// intersections(gridView, element): the 2*dim facets of a Yasp element in indexInInside order
// (facet 2*dir + side); boundary() = facet lies on the boundary of the grid
template <int dim>
struct YaspIntersection {
  int face; bool onBoundary;
  int indexInInside() const { return face; }
  bool boundary() const { return onBoundary; }
};
template <class GV>
std::vector<YaspIntersection<GV::dimension>> intersections(const GV& gv, const typename GV::Grid::Element& e) {
  constexpr int dim = GV::dimension;
  std::vector<YaspIntersection<dim>> r;
  for (int f = 0; f < 2 * dim; ++f)
    r.push_back({f, (f % 2 == 0) ? e.coord[f / 2] == 0 : e.coord[f / 2] == gv.grid().N_[f / 2] - 1});
  return r;
}
template <class GV>
bool hasBoundaryIntersections(const GV& gv, const typename GV::Grid::Element& e) {
  for (const auto& i : intersections(gv, e)) if (i.boundary()) return true;
  return false;
}

namespace Functions {

namespace Impl {
inline void check(int status) {
  if (status == DFX_OK) return;
  std::string msg = dfx_last_error();
  if (status == DFX_ERR_RANGE) throw RangeError(msg);
  if (status == DFX_ERR_NOT_IMPLEMENTED) throw NotImplemented(msg);
  throw Exception(msg);
}
struct Handle {
  dfx_basis* h = nullptr;
  std::size_t numbering = 0;
  ~Handle() { if (h) dfx_basis_destroy(h); }
};
}  // namespace Impl

// ---- BasisFactory: pre-basis factories lower to the POD tree descriptor -----------------
namespace BasisFactory {

struct FlatLexicographic { static constexpr int value = DFX_IMS_FLAT_LEXICOGRAPHIC; };
struct FlatInterleaved { static constexpr int value = DFX_IMS_FLAT_INTERLEAVED; };
struct BlockedLexicographic { static constexpr int value = DFX_IMS_BLOCKED_LEXICOGRAPHIC; };
struct BlockedInterleaved { static constexpr int value = DFX_IMS_BLOCKED_INTERLEAVED; };
inline FlatLexicographic flatLexicographic() { return {}; }
inline FlatInterleaved flatInterleaved() { return {}; }
inline BlockedLexicographic blockedLexicographic() { return {}; }
inline BlockedInterleaved blockedInterleaved() { return {}; }

struct PreBasisFactory { std::vector<dfx_tree_node> nodes; };

template <int k> PreBasisFactory lagrange() { return {{{DFX_NODE_LAGRANGE, k, 0, DFX_IMS_NONE}}}; }
inline PreBasisFactory lagrange(int k) { return {{{DFX_NODE_LAGRANGE, k, 0, DFX_IMS_NONE}}}; }
template <int k> PreBasisFactory lagrangeDG() { return {{{DFX_NODE_LAGRANGE_DG, k, 0, DFX_IMS_NONE}}}; }
inline PreBasisFactory lagrangeDG(int k) { return {{{DFX_NODE_LAGRANGE_DG, k, 0, DFX_IMS_NONE}}}; }
inline PreBasisFactory taylorHood() { return {{{DFX_NODE_TAYLOR_HOOD, 0, 0, DFX_IMS_NONE}}}; }

template <std::size_t C, class IMS = BlockedInterleaved>
PreBasisFactory power(const PreBasisFactory& child, IMS = {}) {
  PreBasisFactory f; f.nodes.push_back({DFX_NODE_POWER, 0, (int32_t)C, IMS::value});
  f.nodes.insert(f.nodes.end(), child.nodes.begin(), child.nodes.end());
  return f;
}
template <class IMS = BlockedInterleaved>
PreBasisFactory power(const PreBasisFactory& child, std::size_t C, IMS = {}) {   // run-time exponent, dynamicpowerbasis.hh
  PreBasisFactory f; f.nodes.push_back({DFX_NODE_POWER, 0, (int32_t)C, IMS::value});
  f.nodes.insert(f.nodes.end(), child.nodes.begin(), child.nodes.end());
  return f;
}

namespace Impl {
template <class T> struct IsIMS : std::false_type {};
template <> struct IsIMS<FlatLexicographic> : std::true_type {};
template <> struct IsIMS<BlockedLexicographic> : std::true_type {};
inline void collect(PreBasisFactory&, int&, int&) {}
template <class A, class... R>
void collect(PreBasisFactory& f, int& children, int& ims, const A& a, const R&... rest) {
  if constexpr (IsIMS<A>::value) ims = A::value;
  else { f.nodes.insert(f.nodes.end(), a.nodes.begin(), a.nodes.end()); ++children; }
  collect(f, children, ims, rest...);
}
}  // namespace Impl

// composite(child..., [ims]) — default BlockedLexicographic (compositebasis.hh:437-445)
template <class... Args>
PreBasisFactory composite(const Args&... args) {
  PreBasisFactory f; f.nodes.push_back({DFX_NODE_COMPOSITE, 0, 0, DFX_IMS_BLOCKED_LEXICOGRAPHIC});
  int children = 0, ims = DFX_IMS_BLOCKED_LEXICOGRAPHIC;
  Impl::collect(f, children, ims, args...);
  f.nodes[0].n_children = children; f.nodes[0].ims = ims;
  return f;
}

}  // namespace BasisFactory

// ---- multi-index: ReservedVector-like (defaultlocalview.hh:57-68) -----------------------
class MultiIndex {
 public:
  std::size_t size() const { return n_; }
  std::size_t operator[](std::size_t i) const { return d_[i]; }
  operator std::size_t() const { return d_[0]; }     // flat bases: StaticMultiIndex<size_t,1> conversion
  bool operator==(const MultiIndex& o) const { return n_ == o.n_ && std::equal(d_, d_ + n_, o.d_); }
  bool operator<(const MultiIndex& o) const { return std::lexicographical_compare(d_, d_ + n_, o.d_, o.d_ + o.n_); }
  std::size_t d_[DFX_MAX_DIGITS] = {0, 0, 0, 0, 0, 0};
  std::size_t n_ = 0;
};

template <class GV> class DefaultGlobalBasis;

// tree node handle (nodes.hh:130-209): size(), localIndex(i), child(i)
template <class GV>
class BasisNode {
 public:
  BasisNode(const DefaultGlobalBasis<GV>* b, int node) : basis_(b), node_(node) {
    Impl::check(dfx_basis_node_info(b->handle(), node, &offset_, &size_, &firstLeaf_, &nLeaves_));
  }
  std::size_t size() const { return size_; }
  std::size_t localIndex(std::size_t i) const { return offset_ + i; }
  std::size_t treeIndex() const { return node_; }
  BasisNode child(std::size_t i) const {
    int c = dfx_basis_node_child(basis_->handle(), node_, (int)i);
    if (c < 0) throw RangeError("no such child");
    return BasisNode(basis_, c);
  }
  template <class... I> BasisNode child(std::size_t i, I... rest) const { return child(i).child(rest...); }
  int id() const { return node_; }
  int numLeaves() const { return nLeaves_; }
 private:
  const DefaultGlobalBasis<GV>* basis_;
  int node_, offset_ = 0, size_ = 0, firstLeaf_ = 0, nLeaves_ = 0;
};

// ---- DefaultLocalView: bind + index served from the batched device table ----------------
template <class GV>
class DefaultLocalView {
 public:
  using GlobalBasis = DefaultGlobalBasis<GV>;
  using GridView = GV;
  using Element = typename GV::Grid::Element;
  using size_type = std::size_t;
  using Tree = BasisNode<GV>;
  static constexpr std::size_t window = 1 << 14;     // elements fetched per device call

  explicit DefaultLocalView(const GlobalBasis& b) : basis_(&b), tree_(&b, 0) {
    stride_ = dfx_basis_max_multi_index_size(b.handle());
    nloc_ = dfx_basis_max_node_size(b.handle());
    lens_.resize(nloc_);
    Impl::check(dfx_basis_local_index_sizes(b.handle(), lens_.data()));
    numbering_ = b.numbering();
  }
  void bind(const Element& e) {
    element_ = e; bound_ = true;
    if (numbering_ != basis_->numbering()) {     // basis.update(gridView) since the window was fetched
      numbering_ = basis_->numbering();
      w0_ = w1_ = 0; flatW0_ = 1; flatW1_ = 0;
    }
    if (e.index < w0_ || e.index >= w1_) {
      w0_ = e.index - e.index % window;
      w1_ = std::min<std::size_t>(w0_ + window, dfx_basis_num_elements(basis_->handle()));
      cache_.resize((w1_ - w0_) * nloc_ * stride_);
      Impl::check(dfx_indices_multi_host(basis_->handle(), w0_, w1_, cache_.data(), stride_));
    }
  }
  void unbind() { bound_ = false; }
  bool bound() const { return bound_; }
  const Element& element() const { return element_; }
  const Tree& tree() const { return tree_; }
  size_type size() const { return nloc_; }
  size_type maxSize() const { return nloc_; }
  MultiIndex index(size_type i) const {
    MultiIndex m; m.n_ = lens_[i];
    const uint64_t* p = &cache_[((element_.index - w0_) * nloc_ + i) * stride_];
    for (std::size_t d = 0; d < m.n_; ++d) m.d_[d] = (std::size_t)p[d];
    return m;
  }
  // Not in the reference: the flat container position of index(i), i.e. what a contiguously stored nested
  // container resolves the multi-index to (istlvectorbackend.hh:267-277) — the row / column numbering of
  // the CSR stand-in for the reference's (nested) matrix types
  size_type flatIndex(size_type i) const {
    if (flatW0_ != w0_ || flatW1_ != w1_) {
      flat_.resize((w1_ - w0_) * nloc_);
      Impl::check(dfx_indices_flat_host(basis_->handle(), w0_, w1_, flat_.data()));
      flatW0_ = w0_; flatW1_ = w1_;
    }
    return flat_[(element_.index - w0_) * nloc_ + i];
  }
  const GlobalBasis& globalBasis() const { return *basis_; }
  const dfx_basis* handle() const { return basis_->handle(); }
  int subtreeNode() const { return 0; }
 private:
  mutable std::vector<uint32_t> flat_;
  mutable std::size_t flatW0_ = 1, flatW1_ = 0;
  std::size_t numbering_ = 0;
  const GlobalBasis* basis_;
  Tree tree_;
  Element element_{};
  bool bound_ = false;
  int stride_ = 1, nloc_ = 0;
  std::vector<uint8_t> lens_;
  std::vector<uint64_t> cache_;
  std::size_t w0_ = 0, w1_ = 0;
};

// ---- ContainerDescriptors (containerdescriptors.hh:91-209) as one run-time type -----------------
// The reference builds the descriptor from compile-time types (Unknown, Value, Tuple, Array, Vector,
// UniformArray, UniformVector; FlatVector = UniformVector<Value>, FlatArray<n> = UniformArray<Value,n>);
// the tree description of this façade is a run-time value, so the descriptor is one too: kind(), size(),
// operator[](i), and type() — the reference's type spelled out.
namespace ContainerDescriptors {
class Descriptor {
 public:
  Descriptor() = default;
  Descriptor(int kind, std::size_t size, std::vector<Descriptor> children) : kind_(kind), size_(size), ch_(std::move(children)) {}
  int kind() const { return kind_; }
  std::size_t size() const { return size_; }
  bool isUnknown() const { return kind_ == DFX_CD_UNKNOWN; }
  bool isValue() const { return kind_ == DFX_CD_VALUE; }
  const Descriptor& operator[](std::size_t i) const {
    if (kind_ == DFX_CD_VALUE) return *this;                            // Value::operator[] gives a Value (:104-105)
    if (kind_ == DFX_CD_UNKNOWN || i >= size_) throw RangeError("container descriptor: child index out of range");
    return (kind_ == DFX_CD_UNIFORM_ARRAY || kind_ == DFX_CD_UNIFORM_VECTOR) ? ch_[0] : ch_[i];
  }
  std::string type() const {
    static const char* names[] = {"Unknown", "Value", "Tuple", "Array", "Vector", "UniformArray", "UniformVector"};
    if (kind_ == DFX_CD_UNKNOWN || kind_ == DFX_CD_VALUE) return names[kind_];
    std::string r = std::string(names[kind_]) + "<";
    if (kind_ == DFX_CD_TUPLE) { for (std::size_t i = 0; i < ch_.size(); ++i) r += (i ? "," : "") + ch_[i].type(); return r + ">"; }
    r += ch_[0].type();
    if (kind_ == DFX_CD_ARRAY || kind_ == DFX_CD_UNIFORM_ARRAY) r += "," + std::to_string(size_);
    return r + ">";
  }
 private:
  int kind_ = DFX_CD_UNKNOWN;
  std::size_t size_ = 0;
  std::vector<Descriptor> ch_;
};
}  // namespace ContainerDescriptors

namespace Impl {
inline ContainerDescriptors::Descriptor buildDescriptor(const std::vector<dfx_container_node>& nodes, std::size_t& pos) {
  const dfx_container_node nd = nodes[pos++];
  std::vector<ContainerDescriptors::Descriptor> ch;
  for (int i = 0; i < nd.n_stored; ++i) ch.push_back(buildDescriptor(nodes, pos));
  return ContainerDescriptors::Descriptor(nd.kind, (std::size_t)nd.size, std::move(ch));
}
}  // namespace Impl

// ---- DefaultGlobalBasis (defaultglobalbasis.hh:51-191) -----------------------------------
template <class GV>
class DefaultGlobalBasis {
 public:
  using GridView = GV;
  using LocalView = DefaultLocalView<GV>;
  using size_type = std::size_t;
  using SizePrefix = std::vector<std::size_t>;

  DefaultGlobalBasis(const GV& gv, const BasisFactory::PreBasisFactory& f, int device = 0) : gv_(gv), h_(std::make_shared<Impl::Handle>()) {
    dfx_grid_desc g = gridDesc(gv);
    Impl::check(dfx_basis_create(&g, f.nodes.data(), (int32_t)f.nodes.size(), device, &h_->h));
    const auto& ids = gv.grid().globalVertexIds();
    if (!ids.empty()) Impl::check(dfx_basis_set_vertex_ids_host(h_->h, ids.data(), ids.size()));
  }
  const GV& gridView() const { return gv_; }
  // update(gridView), defaultglobalbasis.hh:137-141: must be called if the grid has changed; local
  // views have to be bound again afterwards
  void update(const GV& gv) {
    dfx_grid_desc g = gridDesc(gv);
    Impl::check(dfx_basis_update(h_->h, &g));
    gv_ = gv;
    ++h_->numbering;
    const auto& ids = gv.grid().globalVertexIds();
    if (!ids.empty()) Impl::check(dfx_basis_set_vertex_ids_host(h_->h, ids.data(), ids.size()));
  }
  size_type dimension() const { return dfx_basis_dimension(h_->h); }
  size_type size() const { return dfx_basis_size(h_->h, nullptr, 0); }
  template <class Prefix> size_type size(const Prefix& prefix) const {
    std::vector<uint64_t> p(prefix.size());
    for (std::size_t i = 0; i < p.size(); ++i) p[i] = prefix[i];
    return dfx_basis_size(h_->h, p.data(), (int32_t)p.size());
  }
  LocalView localView() const { return LocalView(*this); }
  // containerDescriptor(), defaultglobalbasis.hh:180-186
  ContainerDescriptors::Descriptor containerDescriptor() const {
    int32_t n = 0;
    Impl::check(dfx_basis_container_descriptor(h_->h, nullptr, 0, &n, nullptr, 0));
    std::vector<dfx_container_node> nodes((std::size_t)n);
    Impl::check(dfx_basis_container_descriptor(h_->h, nodes.data(), n, &n, nullptr, 0));
    std::size_t pos = 0;
    return Impl::buildDescriptor(nodes, pos);
  }
  const dfx_basis* handle() const { return h_->h; }
  std::size_t numbering() const { return h_->numbering; }   // changes whenever update() renumbers the basis
  static dfx_grid_desc gridDesc(const GV& gv) {
    constexpr int dim = GV::dimension;
    dfx_grid_desc g{};
    g.dim = dim;
    for (int j = 0; j < 3; ++j) {
      g.n_global[j] = g.local_n[j] = j < dim ? gv.grid().N_[j] : 1;
      g.local_begin[j] = 0;
      g.lower[j] = j < dim ? gv.grid().lower_[j] : 0.0;
      g.upper[j] = j < dim ? gv.grid().L_[j] : 1.0;
    }
    return g;
  }
  const DefaultGlobalBasis& rootBasis() const { return *this; }
  int rootNode() const { return 0; }
 private:
  GV gv_;
  std::shared_ptr<Impl::Handle> h_;
};

namespace BasisFactory {
template <class GV>
DefaultGlobalBasis<GV> makeBasis(const GV& gv, const PreBasisFactory& f) { return DefaultGlobalBasis<GV>(gv, f); }
}  // namespace BasisFactory

// the reference's convenience typedefs
template <class GV, int k> struct LagrangeBasis : DefaultGlobalBasis<GV> {
  explicit LagrangeBasis(const GV& gv) : DefaultGlobalBasis<GV>(gv, BasisFactory::lagrange<k>()) {}
};
template <class GV, int k> struct LagrangeDGBasis : DefaultGlobalBasis<GV> {
  explicit LagrangeDGBasis(const GV& gv) : DefaultGlobalBasis<GV>(gv, BasisFactory::lagrangeDG<k>()) {}
};
template <class GV> struct TaylorHoodBasis : DefaultGlobalBasis<GV> {
  explicit TaylorHoodBasis(const GV& gv) : DefaultGlobalBasis<GV>(gv, BasisFactory::taylorHood()) {}
};

// ---- SubspaceBasis (subspacebasis.hh:27-157) ----------------------------------------------
// SubspaceLocalView (subspacelocalview.hh:32-153): the root's local view and indices,
// tree() = the child at the prefix path, size()/index() forward to the root
template <class GV>
class SubspaceLocalView {
 public:
  using RootLocalView = DefaultLocalView<GV>;
  using GridView = GV;
  using Element = typename RootLocalView::Element;
  using Tree = BasisNode<GV>;
  SubspaceLocalView(const DefaultGlobalBasis<GV>& root, int node) : rootLocalView_(root), tree_(&root, node), node_(node) {}
  void bind(const Element& e) { rootLocalView_.bind(e); }
  void unbind() { rootLocalView_.unbind(); }
  bool bound() const { return rootLocalView_.bound(); }
  const Element& element() const { return rootLocalView_.element(); }
  const Tree& tree() const { return tree_; }
  std::size_t size() const { return rootLocalView_.size(); }
  std::size_t maxSize() const { return rootLocalView_.maxSize(); }
  MultiIndex index(std::size_t i) const { return rootLocalView_.index(i); }
  const RootLocalView& rootLocalView() const { return rootLocalView_; }
  const dfx_basis* handle() const { return rootLocalView_.handle(); }
  int subtreeNode() const { return node_; }
 private:
  RootLocalView rootLocalView_;
  Tree tree_;
  int node_;
};

template <class GV>
class SubspaceBasis {
 public:
  using GridView = GV;
  using LocalView = SubspaceLocalView<GV>;
  LocalView localView() const { return LocalView(*root_, node_); }
  SubspaceBasis(const DefaultGlobalBasis<GV>& root, int node, std::vector<std::size_t> prefixPath = {})
      : root_(&root), node_(node), prefixPath_(std::move(prefixPath)) {}
  const DefaultGlobalBasis<GV>& rootBasis() const { return *root_; }
  // tree path of the sub-tree in the root's local ansatz tree (subspacebasis.hh:96-99)
  const std::vector<std::size_t>& prefixPath() const { return prefixPath_; }
  int rootNode() const { return node_; }
  std::size_t dimension() const { return root_->dimension(); }     // the root's, :79-94
  std::size_t size() const { return root_->size(); }
  template <class Prefix> std::size_t size(const Prefix& prefix) const { return root_->size(prefix); }
  auto containerDescriptor() const { return root_->containerDescriptor(); }   // subspacebasis.hh:115-118
  const GV& gridView() const { return root_->gridView(); }
  const dfx_basis* handle() const { return root_->handle(); }
 private:
  const DefaultGlobalBasis<GV>* root_;
  int node_;
  std::vector<std::size_t> prefixPath_;
};
template <class GV, class... I>
SubspaceBasis<GV> subspaceBasis(const DefaultGlobalBasis<GV>& root, I... path) {
  int node = 0;
  std::vector<std::size_t> tp;
  for (std::size_t c : std::initializer_list<std::size_t>{(std::size_t)path...}) {
    node = dfx_basis_node_child(root.handle(), node, (int)c);
    if (node < 0) throw RangeError("subspaceBasis: no such child");
    tp.push_back(c);
  }
  return SubspaceBasis<GV>(root, node, std::move(tp));
}
// subspaceBasis of a SubspaceBasis: no nesting, the tree paths are joined and the result refers to the
// root's root (subspacebasis.hh:130-133, :139-157; subspacebasistest.cc)
template <class GV, class... I>
SubspaceBasis<GV> subspaceBasis(const SubspaceBasis<GV>& sub, I... path) {
  int node = sub.rootNode();
  std::vector<std::size_t> tp = sub.prefixPath();
  for (std::size_t c : std::initializer_list<std::size_t>{(std::size_t)path...}) {
    node = dfx_basis_node_child(sub.handle(), node, (int)c);
    if (node < 0) throw RangeError("subspaceBasis: no such child");
    tp.push_back(c);
  }
  return SubspaceBasis<GV>(sub.rootBasis(), node, std::move(tp));
}

// ---- functions: registry entries run on the device, any other callable at the nodes ------
namespace Registry {
struct Function { dfx_function f; };
inline Function gaussian(double a = 1.0) { Function r{}; r.f.id = DFX_FN_GAUSSIAN; r.f.n_comp = 1; r.f.p[0] = a; return r; }
inline Function identity(int n) { Function r{}; r.f.id = DFX_FN_IDENTITY; r.f.n_comp = n; return r; }
inline Function coord(int j) { Function r{}; r.f.id = DFX_FN_COORD; r.f.n_comp = 1; r.f.p[0] = j; return r; }
inline Function sinprod(double w) { Function r{}; r.f.id = DFX_FN_SINPROD; r.f.n_comp = 1; r.f.p[0] = w; return r; }
inline Function constant(double c) { Function r{}; r.f.id = DFX_FN_CONSTANT; r.f.n_comp = 1; r.f.p[0] = c; return r; }
}  // namespace Registry

template <class R, class B, class C> class DiscreteGlobalBasisFunction;

namespace Impl {
template <class T> struct IsDiscreteGlobalBasisFunction : std::false_type {};
template <class R, class B, class C> struct IsDiscreteGlobalBasisFunction<DiscreteGlobalBasisFunction<R, B, C>> : std::true_type {};
template <class T> struct IsStdTuple : std::false_type {};
template <class... T> struct IsStdTuple<std::tuple<T...>> : std::true_type {};
// flatVectorView(entry)[0] (flatvectorview.hh:27-222): the first scalar of a range entry in flat lexicographic
// order — FieldMatrix row-major (:63-78), so [0][0]; a scalar is its own view.  Every leaf in scope is a
// scalar-valued Lagrange element, whose one range component lands there.
template <class Y> decltype(auto) flatFirst(Y& y) {
  if constexpr (std::is_arithmetic_v<std::remove_cv_t<Y>>) return (y);
  else if constexpr (IsStdTuple<std::remove_cv_t<Y>>::value) return flatFirst(std::get<0>(y));
  else return flatFirst(y[0]);
}
// HierarchicNodeToRangeMap (hierarchicnodetorangemap.hh:33-48): the leaf with tree path (i0, ..., in) — relative
// to the root of the interpolated (sub)tree — takes y[i0]...[in] (resolveStaticMultiIndex); a range without
// index access is forwarded whole for all nodes.  std::tuple levels (TupleVector-like ranges such as
// tuple<FieldVector<double,3>, double> of examples/stokes-taylorhood.cc) are indexed by position.  The visitor
// receives the scalar the leaf's single range component maps to (flatVectorView of the entry).
template <class Y, class Visitor>
void visitRangeEntry(Y& y, const int* path, int len, Visitor&& visit) {
  using T = std::remove_cv_t<Y>;
  if constexpr (std::is_arithmetic_v<T>) { (void)path; (void)len; visit(y); }
  else if (len == 0) visit(flatFirst(y));
  else if constexpr (IsStdTuple<T>::value) {
    bool found = false;
    int i = 0;
    std::apply([&](auto&... e) { ((i++ == path[0] ? (void)(found = true, visitRangeEntry(e, path + 1, len - 1, visit)) : (void)0), ...); }, y);
    if (!found) throw RangeError("tree path leaves the range type (tuple index " + std::to_string(path[0]) + ")");
  } else {
    if ((std::size_t)path[0] >= y.size()) throw RangeError("tree path leaves the range type (index " + std::to_string(path[0]) + ")");
    visitRangeEntry(y[path[0]], path + 1, len - 1, visit);
  }
}
template <class Y> double rangeEntry(const Y& y, const std::vector<int>& path) {
  double r = 0;
  visitRangeEntry(y, path.data(), (int)path.size(), [&](const auto& v) { r = (double)v; });
  return r;
}
template <class Y> void assignRangeEntry(Y& y, const std::vector<int>& path, double value) {
  visitRangeEntry(y, path.data(), (int)path.size(), [&](auto& v) { v = (std::remove_reference_t<decltype(v)>)value; });
}
// tree paths of the leaves below `node`, relative to it, in leaf order (TypeTree::forEachLeafNode)
inline void collectLeafPaths(const dfx_basis* h, int node, std::vector<int>& prefix, std::vector<std::vector<int>>& out) {
  int off, size, firstLeaf, nLeaves;
  check(dfx_basis_node_info(h, node, &off, &size, &firstLeaf, &nLeaves));
  if (dfx_basis_node_child(h, node, 0) < 0) { out.push_back(prefix); return; }
  for (int c = 0;; ++c) {
    const int child = dfx_basis_node_child(h, node, c);
    if (child < 0) break;
    prefix.push_back(c);
    collectLeafPaths(h, child, prefix, out);
    prefix.pop_back();
  }
}
inline std::vector<std::vector<int>> leafPaths(const dfx_basis* h, int node) {
  std::vector<std::vector<int>> out;
  std::vector<int> prefix;
  collectLeafPaths(h, node, prefix, out);
  return out;
}
template <class C> double* coeffData(C& c, std::size_t n) {
  using V = typename C::value_type;
  static_assert(std::is_trivially_copyable_v<V>, "coefficient container must store plain values contiguously");
  const std::size_t per = sizeof(V) / sizeof(double);
  if (c.size() * per != n) c.resize((n + per - 1) / per);      // vector.resize(basis), interpolate.hh:235
  return reinterpret_cast<double*>(c.data());
}

// ---- nested coefficient containers (ISTLVectorBackend, backends/istlvectorbackend.hh:104-312) --------
// vector[multiIndex] resolves digit j at nesting level j (:267-277), so the entries of a nested container
// in lexicographic order ARE the flat container order of include/dfx.h.  A container whose storage is
// one contiguous run of doubles (std::vector<double>, std::vector<FieldVector<double,n>>, ...) is used
// in place; anything else — std::tuple<...> for a Tuple descriptor (TupleVector / MultiTypeBlockVector in
// the reference's Stokes example), std::array / std::vector of vectors — is copied entry by entry.
template <class T> struct IsTuple : std::false_type {};
template <class... T> struct IsTuple<std::tuple<T...>> : std::true_type {};
template <class T, class = void> struct IsFlatBits : std::false_type {};
template <class T> struct IsFlatBits<T, std::void_t<typename T::value_type>> : std::bool_constant<std::is_arithmetic_v<typename T::value_type>> {};
template <class T, class = void> struct HasResize : std::false_type {};
template <class T> struct HasResize<T, std::void_t<decltype(std::declval<T&>().resize(std::size_t(0)))>> : std::true_type {};
template <class T, class = void> struct IsContiguousRun : std::false_type {};
template <class T>
struct IsContiguousRun<T, std::void_t<typename T::value_type, decltype(std::declval<T&>().data()), decltype(std::declval<T&>().resize(std::size_t(0)))>>
    : std::bool_constant<std::is_trivially_copyable_v<typename T::value_type> && sizeof(typename T::value_type) % sizeof(double) == 0> {};

// ISTLVectorBackend::resize(sizeProvider), :153-233.  `sizes` is anything with size() and operator[](i) that
// says how many entries follow a prefix: a container descriptor, or PrefixSizes around a size provider.
//  - resizable level (:153-184): resized to size(prefix); size 0 means "this is a coefficient" — a non-empty
//    container is then accepted as it is (operator[] pads multi-indices with zeros), an empty one is an error;
//  - statically sized level (:186-221; std::tuple, std::array, FieldVector): must already have that size;
//  - scalar (:223-231): size(prefix) must be 0.
template <class SizeProvider>
struct PrefixSizes {
  const SizeProvider* provider;
  std::vector<std::size_t> prefix;
  std::size_t size() const { return provider->size(prefix); }
  PrefixSizes operator[](std::size_t i) const { PrefixSizes r{provider, prefix}; r.prefix.push_back(i); return r; }
};
template <class C, class D> void resizeNested(C& c, const D& sizes) {
  const std::size_t size = sizes.size();
  if constexpr (std::is_arithmetic_v<C>) {
    (void)c;
    if (size != 0) throw RangeError("Can't resize scalar vector entry to size " + std::to_string(size));
  } else if constexpr (IsTuple<C>::value) {
    if (size == 0) return;
    if (std::tuple_size_v<C> != size)
      throw RangeError("Can't resize non-resizable entry of size " + std::to_string(std::tuple_size_v<C>) + " to size " + std::to_string(size));
    std::size_t i = 0;
    std::apply([&](auto&... e) { (resizeNested(e, sizes[i++]), ...); }, c);
  } else if constexpr (HasResize<C>::value) {
    if (size == 0) {
      if (c.size() == 0) throw RangeError("The vector entry should refer to a scalar coefficient, but is a dynamically sized vector of size==0");
      return;
    }
    c.resize(size);
    for (std::size_t i = 0; i < size; ++i) resizeNested(c[i], sizes[i]);
  } else {
    if (size == 0) return;
    if (c.size() != size)
      throw RangeError("Can't resize non-resizable entry of size " + std::to_string(c.size()) + " to size " + std::to_string(size));
    for (std::size_t i = 0; i < size; ++i) resizeNested(c[i], sizes[i]);
  }
}
// the sizes of a basis: its container descriptor when it has one (uniform levels share one child, no ABI
// call per entry), size(prefix) otherwise
template <class T, class = void> struct HasContainerDescriptor : std::false_type {};
template <class T> struct HasContainerDescriptor<T, std::void_t<decltype(std::declval<const T&>().containerDescriptor())>> : std::true_type {};
template <class C, class SizeProvider> void resizeFor(C& c, const SizeProvider& sizeProvider) {
  if constexpr (HasContainerDescriptor<SizeProvider>::value) {
    const auto cd = sizeProvider.containerDescriptor();
    if (!cd.isUnknown()) { resizeNested(c, cd); return; }
  }
  resizeNested(c, PrefixSizes<SizeProvider>{&sizeProvider, {}});
}
template <class C> void flattenNested(const C& c, std::vector<double>& out) {
  if constexpr (std::is_arithmetic_v<C>) out.push_back((double)c);
  else if constexpr (IsTuple<C>::value) std::apply([&](const auto&... e) { (flattenNested(e, out), ...); }, c);
  else for (std::size_t i = 0; i < c.size(); ++i) flattenNested(c[i], out);
}
template <class C> void unflattenNested(C& c, const double*& p) {
  if constexpr (std::is_arithmetic_v<C>) c = (C)*p++;
  else if constexpr (IsTuple<C>::value) std::apply([&](auto&... e) { (unflattenNested(e, p), ...); }, c);
  else for (std::size_t i = 0; i < c.size(); ++i) unflattenNested(c[i], p);
}
// scalar type at the leaves of a nested container (the first leaf's)
template <class C, class = void> struct LeafScalar { using type = C; };
template <class C> struct LeafScalar<C, std::enable_if_t<IsTuple<C>::value>> { using type = typename LeafScalar<std::tuple_element_t<0, C>>::type; };
template <class C> struct LeafScalar<C, std::void_t<typename C::value_type>> { using type = typename LeafScalar<typename C::value_type>::type; };
// vector[multiIndex]: digit j indexes nesting level j (resolveDynamicMultiIndex, common/indexaccess.hh:376-403)
// a multi-index that ends before a scalar is reached is padded with zeros (:127-133, :369-370)
template <class S, class C, class MI> S& resolveNested(C& c, const MI& mi, std::size_t level) {
  if constexpr (std::is_arithmetic_v<C>) { (void)mi; (void)level; return c; }
  else {
    const std::size_t digit = level < mi.size() ? (std::size_t)mi[level] : 0;
    if constexpr (IsTuple<C>::value) {
      S* r = nullptr;
      std::size_t i = 0;
      std::apply([&](auto&... e) { ((i++ == digit ? (void)(r = &resolveNested<S>(e, mi, level + 1)) : (void)0), ...); }, c);
      if (!r) throw RangeError("multi-index digit exceeds the tuple size");
      return *r;
    } else return resolveNested<S>(c[digit], mi, level + 1);
  }
}
struct AllTrue {};
}  // namespace Impl

// istlVectorBackend(v) (backends/istlvectorbackend.hh:297-312): a non-owning view with resize(basis) and
// operator[](multiIndex) on nested standard containers
template <class V>
class ISTLVectorBackend {
 public:
  using Scalar = typename Impl::LeafScalar<V>::type;
  explicit ISTLVectorBackend(V& v) : v_(&v) {}
  template <class SizeProvider> void resize(const SizeProvider& sizeProvider) { Impl::resizeFor(*v_, sizeProvider); }   // :259-265
  template <class MI> Scalar& operator[](const MI& mi) const { return Impl::resolveNested<Scalar>(*v_, mi, 0); }   // :267-277
  template <class T> ISTLVectorBackend& operator=(const T& value) { fill(*v_, value); return *this; }
  V& vector() const { return *v_; }
 private:
  template <class C, class T> static void fill(C& c, const T& value) {
    if constexpr (std::is_arithmetic_v<C>) c = (C)value;
    else if constexpr (Impl::IsTuple<C>::value) std::apply([&](auto&... e) { (fill(e, value), ...); }, c);
    else for (std::size_t i = 0; i < c.size(); ++i) fill(c[i], value);
  }
  V* v_;
};
template <class V> ISTLVectorBackend<V> istlVectorBackend(V& v) { return ISTLVectorBackend<V>(v); }

// interpolate(basis, coeff, f, bitVector) — interpolate.hh:204-282
template <class B, class C, class F, class BV>
void interpolate(const B& basis, C&& coeff, const F& f, const BV& bv) {
  const dfx_basis* h = basis.handle();
  const std::size_t n = dfx_basis_dimension(h);
  using CV = std::remove_cv_t<std::remove_reference_t<C>>;
  const uint8_t* mask = nullptr;
  std::vector<uint8_t> m;
  if constexpr (!std::is_same_v<BV, Impl::AllTrue>) {
    if constexpr (Impl::IsFlatBits<BV>::value) {
      if (bv.size() != n) throw RangeError("bit vector size does not match basis");
      m.resize(n);
      for (std::size_t i = 0; i < n; ++i) m[i] = bv[i] ? 1 : 0;
    } else {
      // a nested bit vector of the same shape as the coefficient container (interpolate.hh:226-233)
      std::vector<double> flat;
      Impl::flattenNested(bv, flat);
      if (flat.size() != n) throw RangeError("bit vector size does not match basis");
      m.resize(n);
      for (std::size_t i = 0; i < n; ++i) m[i] = flat[i] != 0.0 ? 1 : 0;
    }
    mask = m.data();
  }
  if constexpr (!Impl::IsContiguousRun<CV>::value) {
    // nested container: size it like ISTLVectorBackend::resize, run on its flat image, copy back
    Impl::resizeFor(coeff, basis);
    std::vector<double> flat;
    flat.reserve(n);
    Impl::flattenNested(coeff, flat);
    if (flat.size() != n) throw RangeError("nested coefficient container does not match the basis");
    if constexpr (std::is_same_v<BV, Impl::AllTrue>) interpolate(basis, flat, f, Impl::AllTrue{});
    else interpolate(basis, flat, f, m);
    const double* p = flat.data();
    Impl::unflattenNested(coeff, p);
    return;
  } else {
  double* x = Impl::coeffData(coeff, n);
  if constexpr (std::is_same_v<F, Registry::Function>) {
    Impl::check(dfx_interpolate_host(h, basis.rootNode(), &f.f, x, mask));
  } else if constexpr (Impl::IsDiscreteGlobalBasisFunction<F>::value) {
    // a grid function: its local function is evaluated at the interpolation nodes, element by
    // element (gridviewfunction.hh:68-75; discreteglobalbasisfunctiontest.cc:49-77)
    Impl::check(dfx_interpolate_dgbf_host(h, basis.rootNode(), f.basis().handle(), f.basis().rootNode(),
                                          f.coefficientData(), x, mask));
  } else {
    // arbitrary host callable: the device produces the node positions and the leaf of every
    // coefficient (form (2) of include/dfx.h); f itself is the caller's CPU code
    constexpr int dim = B::GridView::dimension;
    std::vector<double> xyz(n * dim);
    std::vector<int32_t> leaf(n);
    Impl::check(dfx_interpolation_points_host(h, xyz.data(), leaf.data()));
    int off, size, firstLeaf, nLeaves;
    Impl::check(dfx_basis_node_info(h, basis.rootNode(), &off, &size, &firstLeaf, &nLeaves));
    const auto paths = Impl::leafPaths(h, basis.rootNode());      // HierarchicNodeToRangeMap: leaf -> y[treePath]
    for (std::size_t p = 0; p < n; ++p) {
      const int l = leaf[p] - firstLeaf;
      if (l < 0 || l >= nLeaves || (mask && !mask[p])) continue;
      FieldVector<double, dim> pos;
      for (int j = 0; j < dim; ++j) pos[j] = xyz[p * dim + j];
      x[p] = Impl::rangeEntry(f(pos), paths[l]);
    }
  }
  }
}
template <class B, class C, class F>
void interpolate(const B& basis, C&& coeff, const F& f) { interpolate(basis, coeff, f, Impl::AllTrue{}); }
// coefficient (and bit) vectors wrapped in istlVectorBackend(...), as the reference's examples pass them
template <class B, class V, class F>
void interpolate(const B& basis, ISTLVectorBackend<V> coeff, const F& f) { interpolate(basis, coeff.vector(), f, Impl::AllTrue{}); }
template <class B, class V, class F, class BV>
void interpolate(const B& basis, ISTLVectorBackend<V> coeff, const F& f, const BV& bv) { interpolate(basis, coeff.vector(), f, bv); }
template <class B, class C, class F, class W>
void interpolate(const B& basis, C&& coeff, const F& f, ISTLVectorBackend<W> bv) { interpolate(basis, coeff, f, bv.vector()); }

// ---- AnalyticGridViewFunction (gridfunctions/analyticgridviewfunction.hh:33-245) -----------------
// A callable f(GlobalCoordinate) as a grid function: operator()(x) = f(x); its local function is bound to
// an element and evaluates f(geometry.global(xLocal)) (:77-81, :105-109).  Host code: this is how an
// arbitrary f reaches interpolate() (gridviewfunction.hh:68-101), which evaluates it at the node positions
// the device produces.
template <class F, class GV>
class AnalyticGridViewFunction {
 public:
  using GridView = GV;
  static constexpr int dim = GV::dimension;
  using Domain = FieldVector<double, dim>;
  using Element = typename GV::Grid::Element;
  AnalyticGridViewFunction(const F& f, const GV& gridView) : f_(f), gv_(gridView) {}
  auto operator()(const Domain& x) const { return f_(x); }
  const GV& gridView() const { return gv_; }
  class LocalFunction {
   public:
    explicit LocalFunction(const F& f) : f_(f) {}
    void bind(const Element& element) { element_ = element; geometry_ = element.geometry(); bound_ = true; }
    void unbind() { bound_ = false; }
    bool bound() const { return bound_; }
    auto operator()(const Domain& xLocal) const {
      if (!bound_) throw Exception("local function is not bound");
      return f_(geometry_.global(xLocal));
    }
    const Element& localContext() const { return element_; }
   private:
    F f_;
    Element element_{};
    typename GV::Grid::Geometry geometry_{};
    bool bound_ = false;
  };
  friend LocalFunction localFunction(const AnalyticGridViewFunction& t) { return LocalFunction(t.f_); }
 private:
  F f_;
  GV gv_;
};
template <class F, class GV>
AnalyticGridViewFunction<F, GV> makeAnalyticGridViewFunction(const F& f, const GV& gridView) { return AnalyticGridViewFunction<F, GV>(f, gridView); }
// makeGridViewFunction (gridviewfunction.hh:68-101): a grid function is handed through unchanged, anything
// else callable is wrapped
template <class F, class GV>
auto makeGridViewFunction(const F& f, const GV& gridView) {
  if constexpr (Impl::IsDiscreteGlobalBasisFunction<F>::value) { (void)gridView; return f; }
  else return AnalyticGridViewFunction<F, GV>(f, gridView);
}

// ---- SubEntityDOFs (subentitydofs.hh:45-236) ------------------------------------------------
// Range of the local indices of the DOFs attached to a sub-entity (or its sub-sub-entities)
// of the element a local view is bound to, plus contains(localIndex).
template <class GV>
class SubEntityDOFs {
 public:
  template <class LocalView>
  SubEntityDOFs& bind(const LocalView& localView, std::size_t subEntityIndex, std::size_t subEntityCodim) {
    const int nloc = (int)localView.size();
    std::vector<uint8_t> flags(nloc);
    std::vector<int32_t> list(nloc);
    int32_t n = 0;
    Impl::check(dfx_subentity_dofs(localView.handle(), localView.subtreeNode(), (int32_t)subEntityIndex,
                                   (int32_t)subEntityCodim, flags.data(), list.data(), &n));
    containedDOFs_.assign(list.begin(), list.begin() + n);
    dofIsContained_.assign(flags.begin(), flags.end());
    return *this;
  }
  template <class LocalView, int dim>
  SubEntityDOFs& bind(const LocalView& localView, const YaspIntersection<dim>& intersection) {
    return bind(localView, intersection.indexInInside(), 1);
  }
  auto begin() const { return containedDOFs_.cbegin(); }
  auto end() const { return containedDOFs_.cend(); }
  std::size_t size() const { return containedDOFs_.size(); }
  std::size_t operator[](std::size_t i) const { return containedDOFs_[i]; }
  bool contains(std::size_t localIndex) const { return dofIsContained_[localIndex]; }
 private:
  std::vector<std::size_t> containedDOFs_;
  std::vector<bool> dofIsContained_;
};
template <class T>
auto subEntityDOFs(const T&) { return SubEntityDOFs<typename T::GridView>{}; }
template <class LocalView>
auto subEntityDOFs(const LocalView& localView, std::size_t subEntityIndex, std::size_t subEntityCodim) {
  SubEntityDOFs<typename LocalView::GridView> s;
  s.bind(localView, subEntityIndex, subEntityCodim);
  return s;
}

// ---- forEachBoundaryDOF (boundarydofs.hh:39-127) ---------------------------------------------
// The three callback signatures of the reference: f(localIndex, localView, intersection),
// f(localIndex, localView), f(globalMultiIndex).  Same visiting order as the reference (a DOF
// may be visited several times).  The element loop runs on the host over the boundary
// elements; bind/index are served from the device index table.  For large grids use
// markBoundaryDOFs below, which does the whole loop in one kernel.
template <class Basis, class F>
void forEachBoundaryDOF(const Basis& basis, F&& f) {
  using GV = typename Basis::GridView;
  using LV = typename Basis::LocalView;
  auto localView = basis.localView();
  SubEntityDOFs<GV> seDOFs;
  const auto& gridView = basis.gridView();
  for (const auto& element : elements(gridView)) {
    if (!hasBoundaryIntersections(gridView, element)) continue;
    localView.bind(element);
    for (const auto& intersection : intersections(gridView, element)) {
      if (!intersection.boundary()) continue;
      for (auto localIndex : seDOFs.bind(localView, intersection)) {
        if constexpr (std::is_invocable_v<F, std::size_t, const LV&, decltype(intersection)>) f(localIndex, localView, intersection);
        else if constexpr (std::is_invocable_v<F, std::size_t, const LV&>) f(localIndex, localView);
        else f(localView.index(localIndex));
      }
    }
  }
}

// Not in the reference: the loop above with the callback `bitVector[index] = true`, done on the
// device in closed form (dfx_boundary_dofs).  bitVector is in flat container order like the
// coefficient vector; entries that are not boundary DOFs of the (sub-space) basis are kept.
template <class Basis, class BV>
void markBoundaryDOFs(const Basis& basis, BV& bitVector) {
  const std::size_t n = dfx_basis_dimension(basis.handle());
  if (bitVector.size() != n) bitVector.resize(n);
  std::vector<uint8_t> m(n);
  for (std::size_t i = 0; i < n; ++i) m[i] = bitVector[i] ? 1 : 0;
  Impl::check(dfx_boundary_dofs_host(basis.handle(), basis.rootNode(), m.data()));
  for (std::size_t i = 0; i < n; ++i) if (m[i]) bitVector[i] = true;
}

// ---- DiscreteGlobalBasisFunction ------------------------------------------------------------
template <class R, class B, class C>
class DiscreteGlobalBasisFunction {
 public:
  using GV = typename B::GridView;
  static constexpr int dim = GV::dimension;
  using Domain = FieldVector<double, dim>;
  using Element = typename GV::Grid::Element;

  // the basis is stored by value (rvalues are moved in, discreteglobalbasisfunction.hh:389-395);
  // the coefficient vector is referenced like an lvalue passed to the reference
  DiscreteGlobalBasisFunction(const B& basis, const C& coeff) : basis_(basis), coeff_(&coeff) {
    if constexpr (Impl::IsContiguousRun<C>::value) {
      if (coeff.size() * sizeof(typename C::value_type) != basis.dimension() * sizeof(double))
        throw RangeError("coefficient vector size does not match basis");
    } else if (flatImage() == nullptr) throw RangeError("coefficient vector size does not match basis");
    int off, size, firstLeaf;
    Impl::check(dfx_basis_node_info(basis.handle(), basis.rootNode(), &off, &size, &firstLeaf, &ncomp_));
  }

  // operator()(x) for WORLD coordinates (discreteglobalbasisfunction.hh:402-410): the element is
  // located in closed form on the device; a point outside the grid throws like the reference's
  // HierarchicSearch.  One device call per point — use evaluateGlobal for many points.
  std::vector<double> operator()(const Domain& x) const {
    std::vector<double> v;
    evaluateGlobal(std::vector<Domain>{x}, v);
    return v;
  }
  void evaluateGlobal(const std::vector<Domain>& x, std::vector<double>& values, std::vector<double>* jac = nullptr) const {
    const std::size_t n = x.size();
    std::vector<double> pts(n * dim);
    for (std::size_t p = 0; p < n; ++p) for (int j = 0; j < dim; ++j) pts[p * dim + j] = x[p][j];
    values.resize(n * ncomp_);
    if (jac) jac->resize(n * ncomp_ * dim);
    std::vector<int32_t> el(n);
    Impl::check(dfx_evaluate_global_host(basis_.handle(), basis_.rootNode(), coefficientData(),
                                         pts.data(), n, values.data(), jac ? jac->data() : nullptr, el.data()));
    for (std::size_t p = 0; p < n; ++p) if (el[p] < 0) throw Exception("GridError: coordinates are outside of the grid");
  }

  // Batched form: values at the reference-element points `xhat` on ALL elements,
  // out[e][q][comp]; jac[e][q][comp][dim] if requested.  This is the fast way to use the path.
  void evaluateAll(const std::vector<Domain>& xhat, std::vector<double>& values, std::vector<double>* jac = nullptr) const {
    const std::size_t ne = dfx_basis_num_elements(basis_.handle()), nq = xhat.size();
    std::vector<double> pts(nq * dim);
    for (std::size_t q = 0; q < nq; ++q) for (int j = 0; j < dim; ++j) pts[q * dim + j] = xhat[q][j];
    values.resize(ne * nq * ncomp_);
    if (jac) jac->resize(ne * nq * ncomp_ * dim);
    Impl::check(dfx_evaluate_host(basis_.handle(), basis_.rootNode(), coefficientData(),
                                  pts.data(), (int32_t)nq, 0, ne, values.data(), jac ? jac->data() : nullptr));
  }

  // localFunction(f): bind(e) then operator()(xLocal) — discreteglobalbasisfunction.hh:124-148, :336-371.
  // Each call evaluates one (element, point) pair on the device; use evaluateAll for loops.
  class LocalFunction {
   public:
    explicit LocalFunction(const DiscreteGlobalBasisFunction& f, bool derivative = false) : f_(&f), derivative_(derivative) {}
    void bind(const Element& e) { e_ = e; bound_ = true; }
    void unbind() { bound_ = false; }
    bool bound() const { return bound_; }
    auto operator()(const Domain& x) const {
      if (!bound_) throw Exception("local function is not bound");
      std::vector<double> v(f_->ncomp_), j(f_->ncomp_ * dim);
      Impl::check(dfx_evaluate_host(f_->basis_.handle(), f_->basis_.rootNode(),
                                    f_->coefficientData(), x.data(), 1, e_.index, e_.index + 1,
                                    derivative_ ? nullptr : v.data(), derivative_ ? j.data() : nullptr));
      return derivative_ ? j : v;       // flat range entries (values) or [comp][dim] (Jacobian)
    }
    friend LocalFunction derivative(const LocalFunction& lf) {
      if (lf.derivative_) throw NotImplemented("derivative of derivative is not implemented");   // :630-633
      return LocalFunction(*lf.f_, true);
    }
   private:
    const DiscreteGlobalBasisFunction* f_;
    Element e_{};
    bool bound_ = false, derivative_;
  };
  friend LocalFunction localFunction(const DiscreteGlobalBasisFunction& f) { return LocalFunction(f); }
  int numComponents() const { return ncomp_; }
  const B& basis() const { return basis_; }
  // The value as the Range type R of makeDiscreteGlobalBasisFunction<R> (discreteglobalbasisfunction.hh:358-368):
  // leaf l of the (sub)tree contributes to nodeToRangeEntry(node, treePath, y), i.e. y[i0]...[in] for its tree
  // path — nested ranges such as tuple<FieldVector<double,3>, double> or FieldMatrix<double,2,2> included.
  using Range = R;
  R toRange(const std::vector<double>& flatEntries) const {
    R y{};
    const auto paths = Impl::leafPaths(basis_.handle(), basis_.rootNode());
    for (std::size_t l = 0; l < paths.size() && l < flatEntries.size(); ++l) Impl::assignRangeEntry(y, paths[l], flatEntries[l]);
    return y;
  }
  R range(const Domain& xGlobal) const { return toRange((*this)(xGlobal)); }
  // flat container order of the coefficients: the container itself if it is one contiguous run, else its
  // flat image, rebuilt at every use because the vector is referenced, not copied (:389-395)
  const double* coefficientData() const {
    if constexpr (Impl::IsContiguousRun<C>::value) return reinterpret_cast<const double*>(coeff_->data());
    else {
      const double* p = flatImage();
      if (!p) throw RangeError("coefficient vector size does not match basis");
      return p;
    }
  }
 private:
  const double* flatImage() const {
    flat_.clear();
    Impl::flattenNested(*coeff_, flat_);
    return flat_.size() == basis_.dimension() ? flat_.data() : nullptr;
  }
  B basis_;
  const C* coeff_;
  mutable std::vector<double> flat_;
  int ncomp_ = 1;
};

template <class R, class B, class C>
DiscreteGlobalBasisFunction<R, B, C> makeDiscreteGlobalBasisFunction(const B& basis, const C& coeff) {
  return DiscreteGlobalBasisFunction<R, B, C>(basis, coeff);
}
// the coefficient vector given as a backend (localfunctioncopytest.cc:58-62 passes istlVectorBackend(x)):
// the wrapped vector is referenced
template <class R, class B, class V>
DiscreteGlobalBasisFunction<R, B, V> makeDiscreteGlobalBasisFunction(const B& basis, const ISTLVectorBackend<V>& backend) {
  return DiscreteGlobalBasisFunction<R, B, V>(basis, backend.vector());
}

}  // namespace Functions
}  // namespace Dune
