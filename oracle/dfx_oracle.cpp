// dfx_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A dependency-free restatement of the dune-functions element loop
// (LocalView::bind/index, interpolate, DiscreteGlobalBasisFunction local
// evaluation) that keeps the reference's loop structure: per element -> bind ->
// PreBasis::indices -> per leaf -> per local DOF / per point.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker / the CPU baseline.  The product
// (dune_functions_b200/csrc) never includes, links or calls anything in here.
//
// PARITY STATUS: the dune-functions logic (index merging, local ordering,
// interpolation loop, evaluation loop) follows the cited reference lines and is
// pinned by the reference's own golden table (doc/manual/dune-functions-bases.tex
// :1101-1292) and known-answer tests (see tests/).  The pieces that live in
// OTHER DUNE modules which are not in this container (dune-grid YaspGrid index
// set + MCMGMapper, dune-localfunctions LagrangeCube, dune-geometry reference
// cube numbering + AxisAlignedCubeGeometry) are restated from their published
// behaviour, one function each, marked [EXT].  Absolute global index values for
// k >= 2 therefore are "parity unpinned" at the dune-grid boundary (no reference
// test fixes them; every reference test is invariant under renumbering).
//
// Build: see oracle/Makefile (g++ -O3 -ffp-contract=off, std::thread only for the
// optional multi-threaded baseline).

#include "../include/dfx.h"

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <thread>

namespace orc {

using size_type = std::size_t;
constexpr int MAXDIM = 3;

// ---------------------------------------------------------------------------
// Multi-index storage: fixed capacity like Dune::ReservedVector
// (defaultlocalview.hh:57-68 picks StaticMultiIndex / ReservedVector).
struct MultiIndex {
  uint8_t n = 0;
  size_type d[DFX_MAX_DIGITS] = {0, 0, 0, 0, 0, 0};
  size_type size() const { return n; }
  size_type& operator[](size_type i) { return d[i]; }
  const size_type& operator[](size_type i) const { return d[i]; }
  void push_back(size_type v) { d[n++] = v; }
  void pop_back() { --n; }
  size_type& back() { return d[n - 1]; }
  void resize(size_type m) { n = (uint8_t)m; }
  bool operator<(const MultiIndex& o) const {
    return std::lexicographical_compare(d, d + n, o.d, o.d + o.n);
  }
};
// dynamicpowerbasis.hh:314-321 / compositebasis.hh:314-321
static void multiIndexPushFront(MultiIndex& M, size_type M0) {
  M.resize(M.size() + 1);
  for (size_type i = M.size() - 1; i > 0; --i) M[i] = M[i - 1];
  M[0] = M0;
}
// dynamicpowerbasis.hh:176-182
static void multiIndexPopFront(MultiIndex& M) {
  for (size_type i = 0; i + 1 < M.size(); ++i) M[i] = M[i + 1];
  M.resize(M.size() - 1);
}

// ---------------------------------------------------------------------------
// [EXT dune-grid] structured grid: element iteration, geometry, index set.
struct Grid {
  int dim;
  int N[MAXDIM];      // local box
  int G[MAXDIM];      // global grid
  int off[MAXDIM];    // local_begin
  double lower[MAXDIM], upper[MAXDIM], h[MAXDIM];
  // grid().globalIdSet() restricted to vertices: id of the vertex with lexicographic position p
  // of the LOCAL box.  Empty = the structured grid's own ids, monotone in the lexicographic
  // position of the GLOBAL grid [EXT YaspGlobalIdSet / persistent index].  A non-empty table
  // stands for a grid manager whose ids are not (UGGrid / ALUGrid cube meshes, refined grids).
  std::vector<size_type> vertexId;

  explicit Grid(const dfx_grid_desc& g) {
    dim = g.dim;
    for (int j = 0; j < MAXDIM; ++j) {
      N[j] = j < dim ? g.local_n[j] : 1;
      G[j] = j < dim ? g.n_global[j] : 1;
      off[j] = j < dim ? g.local_begin[j] : 0;
      lower[j] = j < dim ? g.lower[j] : 0.0;
      upper[j] = j < dim ? g.upper[j] : 1.0;
      // YaspGrid: h = (upperright - lowerleft) / s   [EXT yaspgrid.hh ctor]
      h[j] = (upper[j] - lower[j]) / G[j];
    }
  }
  size_type numElements() const { return (size_type)N[0] * N[1] * N[2]; }
  // elements(gridView) iterate x fastest; index = ix + Nx (iy + Ny iz)  [EXT]
  void elementCoords(size_type e, int c[MAXDIM]) const {
    c[0] = (int)(e % N[0]);
    c[1] = (int)((e / N[0]) % N[1]);
    c[2] = (int)(e / ((size_type)N[0] * N[1]));
  }
  // EquidistantOffsetCoordinates::coordinate(d, i) = origin + i*h   [EXT]
  double coordinate(int j, int iGlobal) const { return lower[j] + iGlobal * h[j]; }
  // globalIdSet().subId(element, v, dim) for the vertex at lattice coordinates c of the local box
  size_type vertexIdAt(const int c[MAXDIM]) const {
    if (vertexId.empty()) {
      size_type lex = 0, stride = 1;
      for (int j = 0; j < dim; ++j) { lex += stride * (size_type)(off[j] + c[j]); stride *= (size_type)(G[j] + 1); }
      return lex;
    }
    size_type lex = 0, stride = 1;
    for (int j = 0; j < dim; ++j) { lex += stride * (size_type)c[j]; stride *= (size_type)(N[j] + 1); }
    return vertexId[lex];
  }

  // ---- index set [EXT YaspIndexSet]: entities of a codim are split into
  // components by shift bitset (bit j set = entity extends in direction j),
  // ordered by ascending bitset value, each numbered lexicographically, x fastest.
  size_type compSize(unsigned s) const {
    size_type r = 1;
    for (int j = 0; j < dim; ++j) r *= (size_type)(N[j] + (((s >> j) & 1) ? 0 : 1));
    return r;
  }
  size_type size(int codim) const {
    size_type r = 0;
    for (unsigned s = 0; s < (1u << dim); ++s)
      if (__builtin_popcount(s) == dim - codim) r += compSize(s);
    return r;
  }
  size_type entityIndex(unsigned s, const int c[MAXDIM]) const {
    size_type offset = 0;
    for (unsigned t = 0; t < s; ++t)
      if (__builtin_popcount(t) == __builtin_popcount(s)) offset += compSize(t);
    size_type lex = 0, stride = 1;
    for (int j = 0; j < dim; ++j) {
      lex += stride * (size_type)c[j];
      stride *= (size_type)(N[j] + (((s >> j) & 1) ? 0 : 1));
    }
    return offset + lex;
  }
};

// [EXT dune-geometry] reference cube sub-entity numbering.  state per direction:
// 0 = at the lower end, 1 = at the upper end, 2 = the sub-entity extends along it.
static void refCubeSubEntity(int dim, int codim, int i, int st[MAXDIM]) {
  st[0] = st[1] = st[2] = 0;
  if (codim == 0) { for (int j = 0; j < dim; ++j) st[j] = 2; return; }
  if (codim == dim) { for (int j = 0; j < dim; ++j) st[j] = (i >> j) & 1; return; }
  if (dim == 2) {           // codim 1: edges 0:x=0 1:x=1 2:y=0 3:y=1
    int dir = i / 2, side = i % 2;
    st[dir] = side; st[1 - dir] = 2; return;
  }
  if (dim == 3 && codim == 1) {   // faces 0:x=0 1:x=1 2:y=0 3:y=1 4:z=0 5:z=1
    int dir = i / 2, side = i % 2;
    for (int j = 0; j < 3; ++j) st[j] = 2;
    st[dir] = side; return;
  }
  if (dim == 3 && codim == 2) {
    // edges 0-3: z-directed at (x,y) = (0,0),(1,0),(0,1),(1,1)
    //       4,5: y-directed at z=0, x=0,1;  6,7: x-directed at z=0, y=0,1
    //       8,9: y-directed at z=1, x=0,1; 10,11: x-directed at z=1, y=0,1
    static const int T[12][3] = {{0,0,2},{1,0,2},{0,1,2},{1,1,2},{0,2,0},{1,2,0},
                                 {2,0,0},{2,1,0},{0,2,1},{1,2,1},{2,0,1},{2,1,1}};
    for (int j = 0; j < 3; ++j) st[j] = T[i][j];
    return;
  }
  assert(false);
}
static int refCubeSubEntityNumber(int dim, int codim, const int st[MAXDIM]) {
  int n = 1;
  if (codim == 0) return 0;
  if (codim == dim) n = 1 << dim;
  else if (dim == 2) n = 4;
  else if (codim == 1) n = 6;
  else n = 12;
  for (int i = 0; i < n; ++i) {
    int t[MAXDIM]; refCubeSubEntity(dim, codim, i, t);
    bool ok = true;
    for (int j = 0; j < dim; ++j) ok = ok && t[j] == st[j];
    if (ok) return i;
  }
  assert(false); return -1;
}

// [EXT dune-geometry] ReferenceElement::subEntities(i, c, cc).contains(ii): is the codim-cc
// sub-entity ii a sub-(sub-)entity of the codim-c sub-entity i?  In state notation: wherever
// the outer entity is pinned to an end, the inner one is pinned to the same end.
static int refCubeSize(int dim, int codim) {
  if (codim == 0) return 1;
  if (codim == dim) return 1 << dim;
  if (dim == 2) return 4;
  return codim == 1 ? 6 : 12;
}
static bool refCubeSubEntityContains(int dim, int i, int c, int ii, int cc) {
  if (cc < c) return false;
  if (i < 0 || i >= refCubeSize(dim, c) || ii < 0 || ii >= refCubeSize(dim, cc)) return false;
  int so[MAXDIM], si[MAXDIM];
  refCubeSubEntity(dim, c, i, so);
  refCubeSubEntity(dim, cc, ii, si);
  for (int j = 0; j < dim; ++j)
    if (so[j] != 2 && si[j] != so[j]) return false;
  return true;
}

// IndexSet::subIndex(e, i, codim)  [EXT YaspEntity::subCompressedIndex]
static size_type yaspSubIndex(const Grid& g, const int ec[MAXDIM], int i, int codim) {
  int st[MAXDIM]; refCubeSubEntity(g.dim, codim, i, st);
  unsigned s = 0; int c[MAXDIM] = {0, 0, 0};
  for (int j = 0; j < g.dim; ++j) {
    if (st[j] == 2) { s |= 1u << j; c[j] = ec[j]; }
    else c[j] = ec[j] + st[j];
  }
  return g.entityIndex(s, c);
}

// [EXT dune-grid] MultipleCodimMultipleGeomTypeMapper for a cube-only grid:
// offsets assigned codim 0..dim; block b>0 -> offset = n; n += size(codim)*b.
struct MCMGMapper {
  const Grid* g; size_type offset[MAXDIM + 1]; size_type block[MAXDIM + 1]; size_type n;
  template <class Layout> MCMGMapper(const Grid& grid, Layout layout) : g(&grid), n(0) {
    for (int c = 0; c <= grid.dim; ++c) {
      block[c] = layout(grid.dim - c); offset[c] = 0;
      if (block[c] > 0) { offset[c] = n; n += grid.size(c) * block[c]; }
    }
  }
  size_type size() const { return n; }
  size_type subIndex(const int ec[MAXDIM], int i, int codim) const {
    return yaspSubIndex(*g, ec, i, codim) * block[codim] + offset[codim];
  }
  size_type index(const int ec[MAXDIM]) const { return subIndex(ec, 0, 0); }
};

// contiguous chunks of [begin,end) on `nthreads` std::threads (CPU baseline only)
template <class F> static void parallelFor(int64_t begin, int64_t end, int nthreads, F fn) {
  if (nthreads <= 1 || end - begin < 2) { fn(begin, end); return; }
  std::vector<std::thread> pool;
  int64_t n = end - begin, chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    int64_t b = begin + t * chunk, e = std::min(end, b + chunk);
    if (b >= e) break;
    pool.emplace_back([=] { fn(b, e); });
  }
  for (auto& th : pool) th.join();
}

static size_type ipow(size_type b, int e) { size_type r = 1; while (e-- > 0) r *= b; return r; }

// ---------------------------------------------------------------------------
// [EXT dune-localfunctions] LagrangeCubeLocalFiniteElement<D,R,dim,k>
struct LocalKey { int subEntity, codim, index; };
struct LagrangeCubeFE {
  int dim, k; int n;
  std::vector<LocalKey> keys;
  LagrangeCubeFE(int dim_, int k_) : dim(dim_), k(k_), n((int)ipow(k_ + 1, dim_)) {
    keys.resize(n);
    if (k == 0) { keys[0] = {0, 0, 0}; return; }
    if (k == 1) { for (int i = 0; i < n; ++i) keys[i] = {i, dim, 0}; return; }
    for (int i = 0; i < n; ++i) {
      int a[MAXDIM]; multiindex(i, a);
      int codim = 0, st[MAXDIM] = {0, 0, 0};
      for (int j = 0; j < dim; ++j) {
        if (a[j] == 0) { st[j] = 0; ++codim; } else if (a[j] == k) { st[j] = 1; ++codim; } else st[j] = 2;
      }
      // 'index' keeps the ordering of i: interpret i in the (k+1)-adic system, drop
      // boundary digits, re-read in the (k-1)-adic system
      int index = 0;
      for (int j = dim - 1; j >= 0; --j)
        if (a[j] > 0 && a[j] < k) index = (k - 1) * index + (a[j] - 1);
      keys[i] = {refCubeSubEntityNumber(dim, codim, st), codim, index};
    }
  }
  int size() const { return n; }
  void multiindex(int i, int a[MAXDIM]) const {
    a[0] = a[1] = a[2] = 0;
    for (int j = 0; j < dim; ++j) { a[j] = i % (k + 1); i /= (k + 1); }
  }
  // 1-d Lagrange polynomial and derivative on nodes m/k
  double p(int i, double x) const {
    double result = 1.0;
    for (int j = 0; j <= k; ++j) if (j != i) result *= (k * x - j) / (i - j);
    return result;
  }
  double dp(int i, double x) const {
    double result = 0.0;
    for (int j = 0; j <= k; ++j) if (j != i) {
      double prod = (k * 1.0) / (i - j);
      for (int l = 0; l <= k; ++l) if (l != i && l != j) prod *= (k * x - l) / (i - l);
      result += prod;
    }
    return result;
  }
  void evaluateFunction(const double* x, double* out) const {
    if (k == 0) { out[0] = 1; return; }
    if (k == 1) {
      for (int i = 0; i < n; ++i) {
        out[i] = 1;
        for (int j = 0; j < dim; ++j) out[i] *= (i & (1 << j)) ? x[j] : 1 - x[j];
      }
      return;
    }
    for (int i = 0; i < n; ++i) {
      int a[MAXDIM]; multiindex(i, a);
      out[i] = 1.0;
      for (int j = 0; j < dim; ++j) out[i] *= p(a[j], x[j]);
    }
  }
  // out[i*dim + j] = d phi_i / d x_j
  void evaluateJacobian(const double* x, double* out) const {
    if (k == 0) { for (int j = 0; j < dim; ++j) out[j] = 0; return; }
    if (k == 1) {
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < dim; ++j) {
          out[i * dim + j] = (i & (1 << j)) ? 1 : -1;
          for (int l = 0; l < dim; ++l)
            if (j != l) out[i * dim + j] *= (i & (1 << l)) ? x[l] : 1 - x[l];
        }
      return;
    }
    for (int i = 0; i < n; ++i) {
      int a[MAXDIM]; multiindex(i, a);
      for (int j = 0; j < dim; ++j) {
        out[i * dim + j] = 1.0;
        for (int l = 0; l < dim; ++l)
          out[i * dim + j] *= (l == j) ? dp(a[l], x[l]) : p(a[l], x[l]);
      }
    }
  }
  // interpolation node of local DOF i
  void node(int i, double* xhat) const {
    if (k == 0) { for (int j = 0; j < dim; ++j) xhat[j] = 0.5; return; }
    int a[MAXDIM]; multiindex(i, a);
    for (int j = 0; j < dim; ++j) xhat[j] = (1.0 * a[j]) / k;
  }
};

// ---------------------------------------------------------------------------
// Experimental::FaceOrientations<dim> restricted to cubes (lagrangebasis.hh:207-423).
// bits 0..11: edge e is flipped against the reference element (:232-235); bits 12+4f..15+4f:
// orientation index 8..15 of quadrilateral facet f (:279-331).
struct FaceOrientations {
  static constexpr int maxEdges = 12, bitsPerFacet = 4;
  uint64_t data = 0;
  static int facetBitOffset(int subEntity) { return maxEdges + bitsPerFacet * subEntity; }   // :218-221
  bool edgeBit(int e) const { return (data >> e) & 1u; }

  // [EXT dune-geometry] re.subEntities(i, codim, dim): local vertex numbers of a sub-entity, ascending
  static int subEntityVertices(int dim, int codim, int i, int v[4]) {
    int st[MAXDIM]; refCubeSubEntity(dim, codim, i, st);
    int n = 0;
    for (int w = 0; w < (1 << dim); ++w) {
      bool in = true;
      for (int j = 0; j < dim; ++j) in = in && (st[j] == 2 || st[j] == ((w >> j) & 1));
      if (in) v[n++] = w;
    }
    return n;
  }
  // [EXT dune-geometry] re.subEntity(facet, 1, i, 2) of the 3-d cube: the facet's i-th edge in the
  // facet's own (quadrilateral) numbering 0:x'=0 1:x'=1 2:y'=0 3:y'=1, x' < y' being the facet's
  // free directions in ascending order
  static int facetEdge(int facet, int i) {
    int st[MAXDIM]; refCubeSubEntity(3, 1, facet, st);
    int fr[2], n = 0;
    for (int j = 0; j < 3; ++j) if (st[j] == 2) fr[n++] = j;
    st[fr[i / 2]] = i % 2;
    return refCubeSubEntityNumber(3, 2, st);
  }

  FaceOrientations() = default;
  // :360-398.  edges: codims <= dim-1 requested; facets: dim == 3 and codim 1 requested
  FaceOrientations(int dim, const size_type* vertexIds, bool edges, bool facets) {
    if (dim == 1 || !edges) return;
    const int edgeCodim = dim - 1;
    for (int e = 0; e < refCubeSize(dim, edgeCodim); ++e) {                 // computeEdgeOrientation :232-235
      int v[4]; subEntityVertices(dim, edgeCodim, e, v);
      if (vertexIds[v[0]] > vertexIds[v[1]]) data |= (uint64_t)1 << e;
    }
    if (dim == 3 && facets)
      for (int f = 0; f < 6; ++f) {                                         // computeQuadrilateralOrientation :279-331
        int v[4]; subEntityVertices(3, 1, f, v);
        const size_type id[4] = {vertexIds[v[0]], vertexIds[v[1]], vertexIds[v[2]], vertexIds[v[3]]};
        const uint32_t edgeOrientationsToMinVertex = 0b11110101110000001000000110001000u;
        unsigned eo = 0;
        for (int i = 0; i < 4; ++i) eo |= (unsigned)edgeBit(facetEdge(f, i)) << i;
        unsigned i_min;
        if (eo == 5) i_min = id[1] < id[2] ? 1 : 2;
        else if (eo == 10) i_min = id[0] < id[3] ? 0 : 3;
        else i_min = (edgeOrientationsToMinVertex >> (2 * eo)) & 3u;
        uint64_t fo = 0;
        if (i_min == 0) fo = 0b1000 | ((uint64_t)(id[2] < id[1]) << 2);
        else if (i_min == 1) fo = 0b1011 | ((uint64_t)(id[0] < id[3]) << 2);
        else if (i_min == 2) fo = 0b1001 | ((uint64_t)(id[3] < id[0]) << 2);
        else fo = 0b1010 | ((uint64_t)(id[1] < id[2]) << 2);
        data |= fo << facetBitOffset(f);
      }
  }
  // :403-410
  unsigned faceOrientationIndex(int dim, int subEntity, int codim) const {
    if (codim == dim - 1) return edgeBit(subEntity);
    if (codim == 1) return (unsigned)((data >> facetBitOffset(subEntity)) & 15u);
    return 0;
  }
};

// Experimental::LagrangeFaceDOFPermutation restricted to cubes (lagrangebasis.hh:429-697)
struct LagrangeFaceDOFPermutation {
  int dim, order;
  std::vector<std::vector<unsigned>> quadTable;    // rows 8..15 of facetFlipTable_ (:596-603)
  // globallyOrientedQuadrilateralDOFTable :518-559
  LagrangeFaceDOFPermutation(int dim_, int order_) : dim(dim_), order(order_), quadTable(8) {
    if (dim != 3 || order < 2) return;
    const unsigned m = order - 1, n = m * m;
    for (unsigned j = 0; j < 8; ++j) {
      quadTable[j].resize(n);
      for (unsigned i = 0; i < n; ++i) {
        unsigned c[2] = {i % m, i / m};
        if ((j & 3u) == 1) { std::swap(c[0], c[1]); c[0] = m - 1 - c[0]; }
        if ((j & 3u) == 2) { c[0] = m - 1 - c[0]; c[1] = m - 1 - c[1]; }
        if ((j & 3u) == 3) { c[0] = m - 1 - c[0]; std::swap(c[0], c[1]); }
        if (j & 4u) std::swap(c[0], c[1]);
        quadTable[j][i] = c[0] + m * c[1];
      }
    }
  }
  // computeFaceOrientations :616-634 (cubes: never the "k == 3 and tetrahedron" branch)
  FaceOrientations computeFaceOrientations(const Grid& g, const int ec[MAXDIM]) const {
    // no codimension requested: the constructor returns before it touches the (lazy) id range, :364-365
    if (dim == 1 || order < 3) return FaceOrientations();
    size_type ids[8];
    for (int v = 0; v < (1 << g.dim); ++v) {
      int c[MAXDIM] = {0, 0, 0};
      for (int j = 0; j < g.dim; ++j) c[j] = ec[j] + ((v >> j) & 1);
      ids[v] = g.vertexIdAt(c);
    }
    return FaceOrientations(dim, ids, true, true);
  }
  // permuteFaceDOF :652-675
  unsigned permuteFaceDOF(int subEntity, int codim, unsigned index, const FaceOrientations& o) const {
    const int k = order;
    if (dim == 1) return index;
    if (dim == 2) {
      if (k > 2 && codim == 1 && o.faceOrientationIndex(dim, subEntity, codim)) return k - 2 - index;
      return index;
    }
    if (k > 2) {
      if (codim == 2 && o.faceOrientationIndex(dim, subEntity, codim)) return k - 2 - index;
      if (codim == 1) return quadTable[o.faceOrientationIndex(dim, subEntity, codim) - 8][index];
    }
    return index;
  }
};

// ---------------------------------------------------------------------------
// functions f (CPU restatement of the registry in include/dfx.h)
static void evalFunction(const dfx_function& f, int dim, const double* x, double* y) {
  switch (f.id) {
    case DFX_FN_CONSTANT: for (int c = 0; c < f.n_comp; ++c) y[c] = f.p[c]; break;
    case DFX_FN_GAUSSIAN: {   // exp(-a * x.two_norm2())
      double s = 0; for (int j = 0; j < dim; ++j) s += x[j] * x[j];
      y[0] = std::exp(-f.p[0] * s); break; }
    case DFX_FN_IDENTITY: for (int c = 0; c < f.n_comp; ++c) y[c] = c < dim ? x[c] : 0.0; break;
    case DFX_FN_COORD: y[0] = x[(int)f.p[0]]; break;
    case DFX_FN_QUADRATIC: {
      double xx[3] = {x[0], dim > 1 ? x[1] : 0.0, dim > 2 ? x[2] : 0.0};
      double r = f.p[0];
      for (int j = 0; j < 3; ++j) r += f.p[1 + j] * xx[j];
      int q = 4;
      for (int i = 0; i < 3; ++i) for (int j = i; j < 3; ++j) r += f.p[q++] * xx[i] * xx[j];
      y[0] = r; break; }
    case DFX_FN_SINPROD: {
      double r = 1; for (int j = 0; j < dim; ++j) r *= std::sin(f.p[0] * x[j]);
      y[0] = r; break; }
    case DFX_FN_GAUSSIAN_VEC: {
      double s = 0; for (int j = 0; j < dim; ++j) s += x[j] * x[j];
      double e = std::exp(-f.p[0] * s);
      for (int c = 0; c < f.n_comp; ++c) y[c] = (c < dim ? x[c] : 1.0) * e;
      break; }
    default: for (int c = 0; c < f.n_comp; ++c) y[c] = 0;
  }
}

// ---------------------------------------------------------------------------
// Element handle as seen by the pre-bases
struct Element { int c[MAXDIM]; size_type index; };

// Leaf record of the local ansatz tree (nodes.hh:130-209)
struct LeafNode { int offset; int size; const LagrangeCubeFE* fe; int treeNode; };

// Pre-basis hierarchy.  `indices` fills the multi-indices of this subtree's local
// DOFs, children in consecutive blocks (nodes.hh:236-245).
struct PreBasis {
  virtual ~PreBasis() {}
  virtual size_type size(MultiIndex prefix) const = 0;
  size_type size() const { return size(MultiIndex()); }
  virtual size_type dimension() const = 0;
  virtual size_type maxNodeSize() const = 0;
  virtual MultiIndex* indices(const Element& e, MultiIndex* it) const = 0;
  virtual int maxMultiIndexSize() const = 0;
  virtual int minMultiIndexSize() const = 0;
  // Vector-backend bookkeeping of the oracle (not a reference method): the number of complete
  // multi-indices of this pre-basis that are lexicographically smaller than every index starting with
  // `prefix` — for a complete index its position in a contiguously stored nested container, which is
  // how vector[multiIndex] resolves (istlvectorbackend.hh:267-277).  The last digit may be one past
  // the end (lower((size())) == dimension()).  Derived from the same merging rules as indices().
  virtual size_type lower(MultiIndex prefix) const = 0;
  // tree bookkeeping: append this subtree's nodes (preorder) and leaves
  virtual void collect(int& nodeCounter, int& offset, std::vector<LeafNode>& leaves,
                       std::vector<std::array<int, 4>>& nodes) const = 0;
};

// lagrangebasis.hh:740-889
struct LagrangePreBasis : PreBasis {
  const Grid& g; int k; LagrangeCubeFE fe; LagrangeFaceDOFPermutation faceDOFPermutation; MCMGMapper mapper;
  LagrangePreBasis(const Grid& grid, int order)
      : g(grid), k(order), fe(grid.dim, order), faceDOFPermutation(grid.dim, order),
        // dofLayout, lagrangebasis.hh:747-762: order 0 -> 1 per element, cube -> (order-1)^dim
        mapper(grid, [order, &grid](int entityDim) -> size_type {
          if (order == 0) return entityDim == grid.dim;
          return ipow(order - 1, entityDim);
        }) {}
  using PreBasis::size;
  size_type size(MultiIndex prefix) const override {      // leafprebasismixin.hh:50-58
    return prefix.size() == 0 ? dimension() : 0;
  }
  size_type dimension() const override { return mapper.size(); }            // :830-833
  size_type lower(MultiIndex prefix) const override { return prefix.size() == 0 ? 0 : prefix[0]; }
  size_type maxNodeSize() const override { return ipow(k + 1, g.dim); }     // :836-841
  int maxMultiIndexSize() const override { return 1; }
  int minMultiIndexSize() const override { return 1; }
  MultiIndex* indices(const Element& e, MultiIndex* it) const override {    // :843-873
    if (k == 1) {
      for (int li = 0; li < fe.size(); ++li) {
        const LocalKey& key = fe.keys[li];
        size_type gi = yaspSubIndex(g, e.c, key.subEntity, key.codim);
        it->n = 1; it->d[0] = gi; ++it;
      }
      return it;
    }
    // Precompute orientations of all faces (:863); with the structured grid's own vertex ids
    // (monotone in lexicographic position) every orientation is the identity
    FaceOrientations faceOrientations = faceDOFPermutation.computeFaceOrientations(g, e.c);
    for (int li = 0; li < fe.size(); ++li) {
      const LocalKey& key = fe.keys[li];
      size_type gi = mapper.subIndex(e.c, key.subEntity, key.codim);
      gi += faceDOFPermutation.permuteFaceDOF(key.subEntity, key.codim, (unsigned)key.index, faceOrientations);   // :868
      it->n = 1; it->d[0] = gi; ++it;
    }
    return it;
  }
  void collect(int& nodeCounter, int& offset, std::vector<LeafNode>& leaves,
               std::vector<std::array<int, 4>>& nodes) const override {
    int id = nodeCounter++;
    nodes.push_back({offset, fe.size(), (int)leaves.size(), 1});
    leaves.push_back({offset, fe.size(), &fe, id});
    offset += fe.size();
  }
};

// lagrangedgbasis.hh:48-134 + leafprebasismappermixin.hh:59-135
struct LagrangeDGPreBasis : PreBasis {
  const Grid& g; int k; LagrangeCubeFE fe; MCMGMapper mapper;
  LagrangeDGPreBasis(const Grid& grid, int order)
      : g(grid), k(order), fe(grid.dim, order),
        // dofLayout lagrangedgbasis.hh:56-79: cube element -> (k+1)^dim, everything else 0
        mapper(grid, [order, &grid](int entityDim) -> size_type {
          return entityDim == grid.dim ? ipow(order + 1, grid.dim) : 0;
        }) {}
  using PreBasis::size;
  size_type size(MultiIndex prefix) const override { return prefix.size() == 0 ? dimension() : 0; }
  size_type dimension() const override { return mapper.size(); }
  size_type lower(MultiIndex prefix) const override { return prefix.size() == 0 ? 0 : prefix[0]; }
  size_type maxNodeSize() const override { return ipow(k + 1, g.dim); }
  int maxMultiIndexSize() const override { return 1; }
  int minMultiIndexSize() const override { return 1; }
  MultiIndex* indices(const Element& e, MultiIndex* it) const override {   // :120-130
    size_type elementOffset = mapper.index(e.c);
    for (int i = 0; i < fe.size(); ++i) { it->n = 1; it->d[0] = elementOffset + i; ++it; }
    return it;
  }
  void collect(int& nodeCounter, int& offset, std::vector<LeafNode>& leaves,
               std::vector<std::array<int, 4>>& nodes) const override {
    int id = nodeCounter++;
    nodes.push_back({offset, fe.size(), (int)leaves.size(), 1});
    leaves.push_back({offset, fe.size(), &fe, id});
    offset += fe.size();
  }
};

// powerbasis.hh:51-123 -> dynamicpowerbasis.hh:48-392
struct PowerPreBasis : PreBasis {
  std::unique_ptr<PreBasis> sub; size_type children; int ims;
  PowerPreBasis(std::unique_ptr<PreBasis> s, size_type c, int ims_) : sub(std::move(s)), children(c), ims(ims_) {}
  using PreBasis::size;
  size_type size(MultiIndex prefix) const override {                        // :138-219
    switch (ims) {
      case DFX_IMS_FLAT_INTERLEAVED:
        if (prefix.size() == 0) return children * sub->size();
        prefix[0] = prefix[0] / children;
        return sub->size(prefix);
      case DFX_IMS_FLAT_LEXICOGRAPHIC:
        if (prefix.size() == 0) return children * sub->size();
        prefix[0] = prefix[0] % sub->size();
        return sub->size(prefix);
      case DFX_IMS_BLOCKED_LEXICOGRAPHIC:
        if (prefix.size() == 0) return children;
        multiIndexPopFront(prefix);
        return sub->size(prefix);
      case DFX_IMS_BLOCKED_INTERLEAVED: {
        if (prefix.size() == 0) return sub->size();
        auto tail = prefix.back();
        prefix.pop_back();
        if (sub->size(prefix) == 0) return 0;
        prefix.push_back(tail);
        auto subSize = sub->size(prefix);
        if (subSize == 0) return children;
        return subSize; }
    }
    return 0;
  }
  size_type dimension() const override { return children * sub->dimension(); }        // :224-227
  size_type lower(MultiIndex prefix) const override {
    if (prefix.size() == 0) return 0;
    const size_type D = sub->dimension();
    switch (ims) {
      case DFX_IMS_FLAT_INTERLEAVED: {            // first digit j*C + c (:264-288)
        const size_type j = prefix[0] / children, c = prefix[0] % children;
        MultiIndex pj; pj.push_back(j);
        MultiIndex pj1; pj1.push_back(j + 1);
        prefix[0] = j;
        const size_type lj = sub->lower(pj);
        return children * lj + c * (sub->lower(pj1) - lj) + (sub->lower(prefix) - lj); }
      case DFX_IMS_FLAT_LEXICOGRAPHIC: {          // first digit c*S + j (:291-312)
        const size_type S = sub->size(), c = prefix[0] / S;
        prefix[0] = prefix[0] % S;
        return c * D + sub->lower(prefix); }
      case DFX_IMS_BLOCKED_LEXICOGRAPHIC: {       // (c, sub index) (:324-347)
        const size_type c = prefix[0];
        multiIndexPopFront(prefix);
        return c * D + sub->lower(prefix); }
      case DFX_IMS_BLOCKED_INTERLEAVED: {         // (sub index, c) (:350-371)
        MultiIndex head = prefix; head.pop_back();
        if (head.size() > 0 && sub->size(head) == 0) return children * sub->lower(head) + prefix.back();
        return children * sub->lower(prefix); }
    }
    return 0;
  }
  size_type maxNodeSize() const override { return children * sub->maxNodeSize(); }    // :230-233
  int maxMultiIndexSize() const override {
    return sub->maxMultiIndexSize() + ((ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC || ims == DFX_IMS_BLOCKED_INTERLEAVED) ? 1 : 0);
  }
  int minMultiIndexSize() const override {
    return sub->minMultiIndexSize() + ((ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC || ims == DFX_IMS_BLOCKED_INTERLEAVED) ? 1 : 0);
  }
  MultiIndex* indices(const Element& e, MultiIndex* multiIndices) const override {    // :263-371
    size_type subTreeSize = sub->maxNodeSize();
    MultiIndex* next = sub->indices(e, multiIndices);
    switch (ims) {
      case DFX_IMS_FLAT_INTERLEAVED:
        for (size_type i = 0; i < subTreeSize; ++i) multiIndices[i][0] *= children;
        for (size_type child = 1; child < children; ++child)
          for (size_type i = 0; i < subTreeSize; ++i) {
            (*next) = multiIndices[i];
            (*next)[0] = multiIndices[i][0] + child;
            ++next;
          }
        break;
      case DFX_IMS_FLAT_LEXICOGRAPHIC: {
        size_type firstIndexEntrySize = sub->size();
        for (size_type child = 1; child < children; ++child)
          for (size_type i = 0; i < subTreeSize; ++i) {
            (*next) = multiIndices[i];
            (*next)[0] += child * firstIndexEntrySize;
            ++next;
          }
        break; }
      case DFX_IMS_BLOCKED_LEXICOGRAPHIC:
        for (size_type i = 0; i < subTreeSize; ++i) multiIndexPushFront(multiIndices[i], 0);
        for (size_type child = 1; child < children; ++child)
          for (size_type i = 0; i < subTreeSize; ++i) {
            (*next) = multiIndices[i];
            (*next)[0] = child;
            ++next;
          }
        break;
      case DFX_IMS_BLOCKED_INTERLEAVED:
        for (size_type i = 0; i < subTreeSize; ++i) multiIndices[i].push_back(0);
        for (size_type child = 1; child < children; ++child)
          for (size_type i = 0; i < subTreeSize; ++i) {
            (*next) = multiIndices[i];
            (*next).back() = child;
            ++next;
          }
        break;
    }
    return next;
  }
  void collect(int& nodeCounter, int& offset, std::vector<LeafNode>& leaves,
               std::vector<std::array<int, 4>>& nodes) const override {
    int id = nodeCounter++;
    nodes.push_back({offset, 0, (int)leaves.size(), 0});
    for (size_type c = 0; c < children; ++c) sub->collect(nodeCounter, offset, leaves, nodes);
    nodes[id][1] = offset - nodes[id][0];
    nodes[id][3] = (int)leaves.size() - nodes[id][2];
  }
};

// compositebasis.hh:55-341
struct CompositePreBasis : PreBasis {
  std::vector<std::unique_ptr<PreBasis>> subs; int ims;
  using PreBasis::size;
  size_type size(MultiIndex prefix) const override {                        // :184-220
    if (ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC) {
      if (prefix.size() == 0) return subs.size();
      auto front = prefix[0];
      multiIndexPopFront(prefix);
      if (front < subs.size()) return subs[front]->size(prefix);
      return 0;
    }
    size_type result = 0;
    if (prefix.size() == 0) { for (auto& s : subs) result += s->size(); }
    else {
      for (auto& s : subs) {
        auto firstDigitSize = s->size();
        if (prefix[0] < firstDigitSize) { result = s->size(prefix); break; }
        prefix[0] -= firstDigitSize;
      }
    }
    return result;
  }
  size_type dimension() const override { size_type r = 0; for (auto& s : subs) r += s->dimension(); return r; }
  size_type lower(MultiIndex prefix) const override {
    if (prefix.size() == 0) return 0;
    size_type before = 0;
    if (ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC) {   // (child, sub index) (:323-338)
      const size_type child = prefix[0];
      multiIndexPopFront(prefix);
      for (size_type i = 0; i < subs.size() && i < child; ++i) before += subs[i]->dimension();
      return child < subs.size() ? before + subs[child]->lower(prefix) : before;
    }
    for (auto& s : subs) {                         // first digit shifted by the earlier children's sizes (:293-312)
      const size_type firstDigitSize = s->size();
      if (prefix[0] < firstDigitSize) return before + s->lower(prefix);
      prefix[0] -= firstDigitSize;
      before += s->dimension();
    }
    return before;
  }
  size_type maxNodeSize() const override { size_type r = 0; for (auto& s : subs) r += s->maxNodeSize(); return r; }
  int maxMultiIndexSize() const override {
    int r = 0; for (auto& s : subs) r = std::max(r, s->maxMultiIndexSize());
    return r + (ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC ? 1 : 0);
  }
  int minMultiIndexSize() const override {
    int r = 99; for (auto& s : subs) r = std::min(r, s->minMultiIndexSize());
    return r + (ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC ? 1 : 0);
  }
  MultiIndex* indices(const Element& e, MultiIndex* multiIndices) const override {    // :293-338
    if (ims == DFX_IMS_FLAT_LEXICOGRAPHIC) {
      size_type firstComponentOffset = 0;
      for (auto& s : subs) {
        size_type subTreeSize = s->maxNodeSize();
        s->indices(e, multiIndices);
        for (size_type i = 0; i < subTreeSize; ++i) multiIndices[i][0] += firstComponentOffset;
        firstComponentOffset += s->size();
        multiIndices += subTreeSize;
      }
      return multiIndices;
    }
    for (size_type child = 0; child < subs.size(); ++child) {
      size_type subTreeSize = subs[child]->maxNodeSize();
      subs[child]->indices(e, multiIndices);
      for (size_type i = 0; i < subTreeSize; ++i) multiIndexPushFront(multiIndices[i], child);
      multiIndices += subTreeSize;
    }
    return multiIndices;
  }
  void collect(int& nodeCounter, int& offset, std::vector<LeafNode>& leaves,
               std::vector<std::array<int, 4>>& nodes) const override {
    int id = nodeCounter++;
    nodes.push_back({offset, 0, (int)leaves.size(), 0});
    for (auto& s : subs) s->collect(nodeCounter, offset, leaves, nodes);
    nodes[id][1] = offset - nodes[id][0];
    nodes[id][3] = (int)leaves.size() - nodes[id][2];
  }
};

// taylorhoodbasis.hh:59-281 (HI = false, the only variant a factory instantiates)
struct TaylorHoodPreBasis : PreBasis {
  const Grid& g; LagrangePreBasis pq1, pq2;
  explicit TaylorHoodPreBasis(const Grid& grid) : g(grid), pq1(grid, 1), pq2(grid, 2) {}
  using PreBasis::size;
  size_type size(MultiIndex prefix) const override {                        // :138-153
    if (prefix.size() == 0) return 2;
    if (prefix.size() == 1) {
      if (prefix[0] == 0) return g.dim * pq2.size();
      if (prefix[0] == 1) return pq1.size();
    }
    return 0;
  }
  size_type dimension() const override { return g.dim * pq2.size() + pq1.size(); }           // :182-185
  size_type lower(MultiIndex prefix) const override {     // (0, i*dim + c) | (1, j)
    if (prefix.size() == 0 || prefix[0] == 0) return prefix.size() > 1 ? prefix[1] : 0;
    if (prefix[0] == 1) return g.dim * pq2.size() + (prefix.size() > 1 ? prefix[1] : 0);
    return dimension();
  }
  size_type maxNodeSize() const override { return g.dim * pq2.maxNodeSize() + pq1.maxNodeSize(); }
  int maxMultiIndexSize() const override { return 2; }
  int minMultiIndexSize() const override { return 2; }
  MultiIndex* indices(const Element& e, MultiIndex* multiIndices) const override {           // :229-251
    for (size_type child = 0; child < (size_type)g.dim; ++child) {
      size_type subTreeSize = pq2.maxNodeSize();
      pq2.indices(e, multiIndices);
      for (size_type i = 0; i < subTreeSize; ++i) {
        multiIndexPushFront(multiIndices[i], 0);
        multiIndices[i][1] = multiIndices[i][1] * g.dim + child;
      }
      multiIndices += subTreeSize;
    }
    size_type subTreeSize = pq1.maxNodeSize();
    pq1.indices(e, multiIndices);
    for (size_type i = 0; i < subTreeSize; ++i) multiIndexPushFront(multiIndices[i], 1);
    multiIndices += subTreeSize;
    return multiIndices;
  }
  void collect(int& nodeCounter, int& offset, std::vector<LeafNode>& leaves,
               std::vector<std::array<int, 4>>& nodes) const override {
    // tree: composite( power<dim>(Q2), Q1 )
    int id = nodeCounter++;
    nodes.push_back({offset, 0, (int)leaves.size(), 0});
    int pid = nodeCounter++;
    nodes.push_back({offset, 0, (int)leaves.size(), 0});
    for (int c = 0; c < g.dim; ++c) pq2.collect(nodeCounter, offset, leaves, nodes);
    nodes[pid][1] = offset - nodes[pid][0];
    nodes[pid][3] = (int)leaves.size() - nodes[pid][2];
    pq1.collect(nodeCounter, offset, leaves, nodes);
    nodes[id][1] = offset - nodes[id][0];
    nodes[id][3] = (int)leaves.size() - nodes[id][2];
  }
};

static std::unique_ptr<PreBasis> buildPreBasis(const Grid& g, const dfx_tree_node* t, int n, int& pos) {
  if (pos >= n) return nullptr;
  const dfx_tree_node& nd = t[pos++];
  switch (nd.kind) {
    case DFX_NODE_LAGRANGE: return std::make_unique<LagrangePreBasis>(g, nd.order);
    case DFX_NODE_LAGRANGE_DG: return std::make_unique<LagrangeDGPreBasis>(g, nd.order);
    case DFX_NODE_TAYLOR_HOOD: return std::make_unique<TaylorHoodPreBasis>(g);
    case DFX_NODE_POWER: {
      auto sub = buildPreBasis(g, t, n, pos);
      if (!sub) return nullptr;
      return std::make_unique<PowerPreBasis>(std::move(sub), (size_type)nd.n_children, nd.ims); }
    case DFX_NODE_COMPOSITE: {
      auto c = std::make_unique<CompositePreBasis>();
      c->ims = nd.ims;
      for (int i = 0; i < nd.n_children; ++i) {
        auto s = buildPreBasis(g, t, n, pos);
        if (!s) return nullptr;
        c->subs.push_back(std::move(s));
      }
      return c; }
  }
  return nullptr;
}

// ---------------------------------------------------------------------------
// DefaultGlobalBasis + DefaultLocalView (defaultglobalbasis.hh:51-191,
// defaultlocalview.hh:39-187)
struct Basis {
  Grid grid;
  std::unique_ptr<PreBasis> pre;
  std::vector<LeafNode> leaves;
  std::vector<std::array<int, 4>> nodes;   // local offset, size, first leaf, #leaves
  int maxDigits, minDigits;
  // vector backend for blocked bases: multi-index -> flat lexicographic position.  Two routes:
  //  (a) `position`: every multi-index of every element collected and sorted (small grids only; kept
  //      as the cross-check of (b), orc_indices_flat_by_enumeration);
  //  (b) flat(): PreBasis::lower — the rank of the multi-index among all multi-indices, built from the
  //      merging rules node by node.
  mutable std::map<MultiIndex, size_type> position;
  mutable bool positionBuilt = false;

  Basis(const dfx_grid_desc& g) : grid(g) {}
  size_type localSize() const { return pre->maxNodeSize(); }

  void bind(size_type e, std::vector<MultiIndex>& indices) const {    // defaultlocalview.hh:95-101
    Element el; grid.elementCoords(e, el.c); el.index = e;
    indices.resize(localSize());
    pre->indices(el, indices.data());
  }
  void buildPositions() const {
    if (positionBuilt) return;
    std::vector<MultiIndex> idx;
    for (size_type e = 0; e < grid.numElements(); ++e) {
      bind(e, idx);
      for (auto& mi : idx) position.emplace(mi, 0);
    }
    size_type p = 0;
    for (auto& kv : position) kv.second = p++;
    positionBuilt = true;
  }
  // vector[multiIndex] of a contiguously stored nested container
  size_type flat(const MultiIndex& mi) const { return maxDigits == 1 ? mi[0] : pre->lower(mi); }
  void warmFlat() const {}     // flat() is stateless (kept: call sites mark where positions are needed)
};

}  // namespace orc

// ===========================================================================
// C interface for ctypes (tests, smoke, bench cpu_baseline)
extern "C" {

void* orc_basis_create(const dfx_grid_desc* g, const dfx_tree_node* t, int32_t n) {
  auto b = std::make_unique<orc::Basis>(*g);
  int pos = 0;
  b->pre = orc::buildPreBasis(b->grid, t, n, pos);
  if (!b->pre || pos != n) return nullptr;
  int nodeCounter = 0, offset = 0;
  b->pre->collect(nodeCounter, offset, b->leaves, b->nodes);
  b->maxDigits = b->pre->maxMultiIndexSize();
  b->minDigits = b->pre->minMultiIndexSize();
  return b.release();
}
void orc_basis_destroy(void* h) { delete (orc::Basis*)h; }
// Vertex ids of the grid's globalIdSet (lagrangebasis.hh:787), one per vertex of the local box in
// lexicographic order; n = 0 restores the structured grid's own (monotone) ids.
int32_t orc_set_vertex_ids(void* h, const uint64_t* ids, uint64_t n) {
  auto* b = (orc::Basis*)h;
  uint64_t nv = 1;
  for (int j = 0; j < b->grid.dim; ++j) nv *= (uint64_t)(b->grid.N[j] + 1);
  if (n != 0 && n != nv) return 1;
  b->grid.vertexId.assign(ids, ids + n);
  b->position.clear(); b->positionBuilt = false;
  return 0;
}
// FaceOrientations of ONE cube with the given vertex ids (2^dim of them), all codimensions
// requested: edge bits (4 in 2-d, 12 in 3-d) and facet orientation indices (6 in 3-d).  Test hook
// for the truth tables of lagrangebasis.hh:132-174.
void orc_face_orientations(int32_t dim, const uint64_t* ids, int32_t* edge_bits, int32_t* facet_index) {
  orc::size_type v[8];
  for (int i = 0; i < (1 << dim); ++i) v[i] = ids[i];
  orc::FaceOrientations o(dim, v, true, true);
  const int ne = dim == 1 ? 0 : orc::refCubeSize(dim, dim - 1);
  for (int e = 0; e < ne; ++e) edge_bits[e] = o.faceOrientationIndex(dim, e, dim - 1);
  if (dim == 3) for (int f = 0; f < 6; ++f) facet_index[f] = o.faceOrientationIndex(dim, f, 1);
}
// T[j][i] of globallyOrientedQuadrilateralDOFTable (lagrangebasis.hh:518-559), j = 0..7
void orc_quad_dof_table(int32_t order, int32_t j, uint32_t* out) {
  orc::LagrangeFaceDOFPermutation p(3, order);
  for (size_t i = 0; i < p.quadTable[j].size(); ++i) out[i] = p.quadTable[j][i];
}
uint64_t orc_dimension(void* h) { return ((orc::Basis*)h)->pre->dimension(); }
uint64_t orc_size_prefix(void* h, const uint64_t* prefix, int32_t len) {
  orc::MultiIndex p; for (int i = 0; i < len; ++i) p.push_back(prefix[i]);
  return ((orc::Basis*)h)->pre->size(p);
}
int32_t orc_max_node_size(void* h) { return (int32_t)((orc::Basis*)h)->pre->maxNodeSize(); }
int32_t orc_max_digits(void* h) { return ((orc::Basis*)h)->maxDigits; }
int32_t orc_min_digits(void* h) { return ((orc::Basis*)h)->minDigits; }
uint64_t orc_num_elements(void* h) { return ((orc::Basis*)h)->grid.numElements(); }
int32_t orc_num_leaves(void* h) { return (int32_t)((orc::Basis*)h)->leaves.size(); }
int32_t orc_num_nodes(void* h) { return (int32_t)((orc::Basis*)h)->nodes.size(); }
void orc_node_info(void* h, int32_t node, int32_t* out4) {
  auto& n = ((orc::Basis*)h)->nodes[node];
  for (int i = 0; i < 4; ++i) out4[i] = n[i];
}

// localView.bind(e); localView.index(i) for e in [e0,e1): digits [ne][nloc][stride], lens [ne][nloc]
void orc_indices(void* h, uint64_t e0, uint64_t e1, uint64_t* digits, uint8_t* lens, int32_t stride) {
  auto* b = (orc::Basis*)h;
  std::vector<orc::MultiIndex> idx;
  size_t nloc = b->localSize();
  for (uint64_t e = e0; e < e1; ++e) {
    b->bind(e, idx);
    for (size_t i = 0; i < nloc; ++i) {
      uint64_t* o = digits + ((e - e0) * nloc + i) * stride;
      for (int p = 0; p < stride; ++p) o[p] = p < (int)idx[i].size() ? idx[i][p] : 0;
      if (lens) lens[(e - e0) * nloc + i] = (uint8_t)idx[i].size();
    }
  }
}
// flat container positions [ne][nloc]
void orc_indices_flat(void* h, uint64_t e0, uint64_t e1, uint64_t* out) {
  auto* b = (orc::Basis*)h;
  b->warmFlat();
  std::vector<orc::MultiIndex> idx;
  size_t nloc = b->localSize();
  for (uint64_t e = e0; e < e1; ++e) {
    b->bind(e, idx);
    for (size_t i = 0; i < nloc; ++i) out[(e - e0) * nloc + i] = b->flat(idx[i]);
  }
}
// the same by brute force: every multi-index of the whole grid collected, sorted lexicographically
// and numbered (small grids only; cross-check of Basis::flat in tests/test_oracle_cpu.py)
void orc_indices_flat_by_enumeration(void* h, uint64_t e0, uint64_t e1, uint64_t* out) {
  auto* b = (orc::Basis*)h;
  b->buildPositions();
  std::vector<orc::MultiIndex> idx;
  size_t nloc = b->localSize();
  for (uint64_t e = e0; e < e1; ++e) {
    b->bind(e, idx);
    for (size_t i = 0; i < nloc; ++i) out[(e - e0) * nloc + i] = b->position.find(idx[i])->second;
  }
}

// interpolate(basis, coeff, f, bitVector) — interpolate.hh:204-260 with
// Imp::interpolateLocal :151-173.  coeff/mask in flat container order.
// nthreads > 1: OpenMP over z-chunks of elements (CPU baseline only; the last
// writer of a shared DOF then depends on thread timing, values agree to 1 ulp of x).
// orc_interpolate_range: the same loop restricted to the elements [e0, e1) of the grid view (parity
// checks on a slice of a grid too large to run whole: the DOFs of those elements get the value the
// loop writes; a DOF shared with an element outside the range may differ from the full loop's last
// writer by the rounding of geometry.global, i.e. 1 ulp of x).
void orc_interpolate_range(void* h, int32_t subtreeNode, const dfx_function* f, double* coeff,
                           const uint8_t* mask, uint64_t e0, uint64_t e1, int32_t nthreads) {
  auto* b = (orc::Basis*)h;
  b->warmFlat();
  const orc::Grid& g = b->grid;
  const int dim = g.dim;
  const auto& nd = b->nodes[subtreeNode];
  const int firstLeaf = nd[2], nLeaves = nd[3];
  orc::parallelFor((int64_t)e0, (int64_t)e1, nthreads, [&](int64_t eBegin, int64_t eEnd) {
    std::vector<orc::MultiIndex> idx;
    for (int64_t e = eBegin; e < eEnd; ++e) {               // for (const auto& e : elements(gridView))
      b->bind((size_t)e, idx);                               // localView.bind(e)
      // localF.bind(e): element geometry (analyticgridviewfunction.hh:77-81)  [EXT YaspGeometry]
      int ec[3]; g.elementCoords((size_t)e, ec);
      double lower[3], upper[3];
      for (int j = 0; j < dim; ++j) {
        lower[j] = g.coordinate(j, g.off[j] + ec[j]);
        upper[j] = g.coordinate(j, g.off[j] + ec[j] + 1);
      }
      for (int l = 0; l < nLeaves; ++l) {                    // forEachLeafNode
        const orc::LeafNode& leaf = b->leaves[firstLeaf + l];
        const orc::LagrangeCubeFE& fe = *leaf.fe;
        std::vector<double> interpolationCoefficients;       // heap allocation per leaf, as in :161
        interpolationCoefficients.resize(fe.size());
        for (int i = 0; i < fe.size(); ++i) {                // LocalInterpolation::interpolate  [EXT]
          double xhat[3], x[3], y[DFX_MAX_LEAVES];
          fe.node(i, xhat);
          // AxisAlignedCubeGeometry::global: lower + local*(upper-lower)   [EXT]
          for (int j = 0; j < dim; ++j) x[j] = lower[j] + xhat[j] * (upper[j] - lower[j]);
          orc::evalFunction(*f, dim, x, y);
          // HierarchicNodeToRangeMap: y[treePath] / scalar y used whole (:33-48)
          interpolationCoefficients[i] = f->n_comp == 1 ? y[0] : y[l];
        }
        for (int i = 0; i < fe.size(); ++i) {
          const orc::MultiIndex& mi = idx[leaf.offset + i];
          size_t p = b->flat(mi);
          if (!mask || mask[p]) coeff[p] = interpolationCoefficients[i];
        }
      }
    }
  });
}
void orc_interpolate(void* h, int32_t subtreeNode, const dfx_function* f, double* coeff,
                     const uint8_t* mask, int32_t nthreads) {
  orc_interpolate_range(h, subtreeNode, f, coeff, mask, 0, ((orc::Basis*)h)->grid.numElements(), nthreads);
}

// interpolate(targetBasis, coeff, f) where f is a DiscreteGlobalBasisFunction over another basis
// on the same grid: makeGridViewFunction hands a grid function through unchanged
// (gridviewfunction.hh:68-75), so interpolateLocal (interpolate.hh:151-173) evaluates
// localFunction(f) at the target's local interpolation nodes — no global search.  This is the
// first variant of checkInterpolationConsistency (discreteglobalbasisfunctiontest.cc:49-77).
int32_t orc_interpolate_dgbf(void* hTarget, int32_t targetNode, void* hSource, int32_t sourceNode,
                             const double* sourceCoeff, double* coeff, const uint8_t* mask) {
  auto* bt = (orc::Basis*)hTarget;
  auto* bs = (orc::Basis*)hSource;
  bt->warmFlat();
  bs->warmFlat();
  const orc::Grid& g = bt->grid;
  const int dim = g.dim;
  const auto& nt = bt->nodes[targetNode];
  const auto& ns = bs->nodes[sourceNode];
  const int nLeavesT = nt[3], nLeavesS = ns[3];
  if (nLeavesS != 1 && nLeavesS != nLeavesT) return 1;
  if (bs->grid.numElements() != g.numElements()) return 2;
  std::vector<orc::MultiIndex> idxT, idxS;
  std::vector<double> localDoFs(bs->localSize()), shapeValues;
  for (size_t e = 0; e < g.numElements(); ++e) {               // for (const auto& e : elements(gridView))
    bt->bind(e, idxT);                                          // localView.bind(e)
    bs->bind(e, idxS);                                          // localF.bind(e): gather of the source's local DOFs
    for (int i = 0; i < ns[1]; ++i) localDoFs[ns[0] + i] = sourceCoeff[bs->flat(idxS[ns[0] + i])];
    for (int l = 0; l < nLeavesT; ++l) {                        // forEachLeafNode of the target tree
      const orc::LeafNode& leaf = bt->leaves[nt[2] + l];
      const orc::LagrangeCubeFE& fe = *leaf.fe;
      std::vector<double> interpolationCoefficients(fe.size());
      const orc::LeafNode& src = bs->leaves[ns[2] + (nLeavesS == 1 ? 0 : l)];   // range entry of this leaf
      for (int i = 0; i < fe.size(); ++i) {
        double xhat[3];
        fe.node(i, xhat);
        shapeValues.resize(src.fe->size());
        src.fe->evaluateFunction(xhat, shapeValues.data());    // LocalFunction::operator()(xhat), :336-371
        double v = 0;
        for (int a = 0; a < src.fe->size(); ++a) v += localDoFs[src.offset + a] * shapeValues[a];
        interpolationCoefficients[i] = v;
      }
      for (int i = 0; i < fe.size(); ++i) {
        size_t p = bt->flat(idxT[leaf.offset + i]);
        if (!mask || mask[p]) coeff[p] = interpolationCoefficients[i];
      }
    }
  }
  (void)dim;
  return 0;
}

// DiscreteGlobalBasisFunction local evaluation at points xhat[nq][dim] on elements
// [e0,e1): values [ne][nq][ncomp], jacobians [ne][nq][ncomp][dim] (either may be NULL)
// discreteglobalbasisfunction.hh:124-148 (bind/gather), :336-371 (values), :583-627 (jacobians)
void orc_evaluate(void* h, int32_t subtreeNode, const double* coeff, const double* xhat, int32_t nq,
                  uint64_t e0, uint64_t e1, double* values, double* jacobians, int32_t nthreads) {
  auto* b = (orc::Basis*)h;
  b->warmFlat();
  const orc::Grid& g = b->grid;
  const int dim = g.dim;
  const auto& nd = b->nodes[subtreeNode];
  const int firstLeaf = nd[2], nLeaves = nd[3];
  orc::parallelFor((int64_t)e0, (int64_t)e1, nthreads, [&](int64_t eBegin, int64_t eEnd) {
    std::vector<orc::MultiIndex> idx;
    std::vector<double> localDoFs(b->localSize());
    std::vector<double> shapeValues, shapeJacobians;
    for (int64_t e = eBegin; e < eEnd; ++e) {
      // LocalFunctionBase::bind: localView.bind + gather of the subtree's DOFs
      b->bind((size_t)e, idx);
      for (int i = 0; i < nd[1]; ++i) {
        int localIndex = nd[0] + i;
        localDoFs[localIndex] = coeff[b->flat(idx[localIndex])];
      }
      int ec[3]; g.elementCoords((size_t)e, ec);
      double jacInv[3];
      for (int j = 0; j < dim; ++j) {
        double lo = g.coordinate(j, g.off[j] + ec[j]), up = g.coordinate(j, g.off[j] + ec[j] + 1);
        jacInv[j] = 1.0 / (up - lo);      // AxisAlignedCubeGeometry::jacobianInverse  [EXT]
      }
      for (int q = 0; q < nq; ++q) {
        const double* x = xhat + (size_t)q * dim;
        for (int l = 0; l < nLeaves; ++l) {
          const orc::LeafNode& leaf = b->leaves[firstLeaf + l];
          const orc::LagrangeCubeFE& fe = *leaf.fe;
          if (values) {
            shapeValues.resize(fe.size());
            fe.evaluateFunction(x, shapeValues.data());
            double v = 0;
            for (int i = 0; i < fe.size(); ++i) v += localDoFs[leaf.offset + i] * shapeValues[i];   // axpy
            values[(((size_t)(e - e0)) * nq + q) * nLeaves + l] = v;
          }
          if (jacobians) {
            shapeJacobians.resize((size_t)fe.size() * dim);
            fe.evaluateJacobian(x, shapeJacobians.data());
            double refJac[3] = {0, 0, 0};
            for (int i = 0; i < fe.size(); ++i)
              for (int j = 0; j < dim; ++j) refJac[j] += localDoFs[leaf.offset + i] * shapeJacobians[(size_t)i * dim + j];
            for (int j = 0; j < dim; ++j)
              jacobians[((((size_t)(e - e0)) * nq + q) * nLeaves + l) * dim + j] = refJac[j] * jacInv[j];
          }
        }
      }
    }
  });
}

// SubEntityDOFs::bind(localView, subEntityIndex, subEntityCodim) — subentitydofs.hh:71-95.
// On a cube-only grid the result is the same for every element.  contained[localSize]
// (dofIsContained_), list[] = containedDOFs_ in push_back order; returns their number.
int32_t orc_subentity_dofs(void* h, int32_t subtreeNode, int32_t subEntityIndex, int32_t subEntityCodim,
                           uint8_t* contained, int32_t* list) {
  auto* b = (orc::Basis*)h;
  const int dim = b->grid.dim;
  const auto& nd = b->nodes[subtreeNode];
  int32_t n = 0;
  for (size_t i = 0; i < b->localSize(); ++i) contained[i] = 0;          // dofIsContained_.assign(size, false)
  for (int l = 0; l < nd[3]; ++l) {                                      // forEachLeafNode(localView.tree())
    const orc::LeafNode& leaf = b->leaves[nd[2] + l];
    const orc::LagrangeCubeFE& fe = *leaf.fe;
    for (int i = 0; i < fe.size(); ++i) {
      const orc::LocalKey& key = fe.keys[i];
      if (orc::refCubeSubEntityContains(dim, subEntityIndex, subEntityCodim, key.subEntity, key.codim)) {
        if (list) list[n] = leaf.offset + i;                             // node.localIndex(i)
        ++n;
        contained[leaf.offset + i] = 1;
      }
    }
  }
  return n;
}

// forEachBoundaryDOF(basis, [&](auto&& index){ mask[index] = 1; }) — boundarydofs.hh:110-128.
// Boundary = boundary of the GLOBAL grid (intersection.boundary(); processor boundaries of a
// slab are not boundary intersections).  Intersections of a Yasp element are visited in
// indexInInside order 0..2*dim-1 (face 2*dir + side)  [EXT].  Returns the number of callback
// invocations (a DOF may be visited several times, as the reference documents).
uint64_t orc_boundary_dofs(void* h, int32_t subtreeNode, uint8_t* mask) {
  auto* b = (orc::Basis*)h;
  b->warmFlat();
  const orc::Grid& g = b->grid;
  const int dim = g.dim;
  std::vector<orc::MultiIndex> idx;
  std::vector<uint8_t> contained(b->localSize());
  std::vector<int32_t> list(b->localSize());
  uint64_t visits = 0;
  for (size_t e = 0; e < g.numElements(); ++e) {
    int ec[3]; g.elementCoords(e, ec);
    bool hasBoundaryIntersections = false;
    for (int j = 0; j < dim; ++j)
      hasBoundaryIntersections = hasBoundaryIntersections || g.off[j] + ec[j] == 0 || g.off[j] + ec[j] == g.G[j] - 1;
    if (!hasBoundaryIntersections) continue;
    b->bind(e, idx);
    for (int face = 0; face < 2 * dim; ++face) {                          // intersections(gridView, element)
      const int dir = face / 2, side = face % 2;
      const bool boundary = side == 0 ? g.off[dir] + ec[dir] == 0 : g.off[dir] + ec[dir] == g.G[dir] - 1;
      if (!boundary) continue;
      const int32_t n = orc_subentity_dofs(h, subtreeNode, face, 1, contained.data(), list.data());
      for (int32_t q = 0; q < n; ++q) {
        mask[b->flat(idx[list[q]])] = 1;                                  // f(localView.index(localIndex))
        ++visits;
      }
    }
  }
  return visits;
}

// ===========================================================================
// SURVEY.md section 8(f) row 2 — the element-loop CALLERS of examples/poisson-pq2.cc, restated:
// getOccupationPattern (:137-174), getLocalMatrix (:35-87), getVolumeTerm (:91-134),
// assembleLaplaceMatrix (:176-255), the symmetric Dirichlet treatment (:356-386).
}  // extern "C"

namespace orc {

// [EXT dune-geometry] QuadratureRules<double,dim>::rule(cube, order): tensor product of the
// Gauss-Legendre rule with m = order/2 + 1 points on [0,1], last coordinate fastest.
static void gaussLegendre01(int m, std::vector<double>& x, std::vector<double>& w) {
  x.resize(m); w.resize(m);
  for (int i = 0; i < m; ++i) {
    long double z = std::cos(3.14159265358979323846264338327950288L * (i + 0.75L) / (m + 0.5L));
    long double pp = 0;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1, p2 = 0;
      for (int j = 1; j <= m; ++j) { long double p3 = p2; p2 = p1; p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j; }
      pp = m * (z * p1 - p2) / (z * z - 1);
      long double dz = p1 / pp;
      z -= dz;
      if (std::fabs((double)dz) < 1e-19) break;
    }
    x[m - 1 - i] = (double)(0.5L * (1 + z));
    w[m - 1 - i] = (double)(1 / ((1 - z * z) * pp * pp));
  }
}
struct QuadratureRule { int n; std::vector<double> pos, weight; };
static QuadratureRule quadratureRule(int dim, int order) {
  const int m = order / 2 + 1;
  std::vector<double> x, w; gaussLegendre01(m, x, w);
  QuadratureRule q; q.n = (int)ipow(m, dim);
  q.pos.resize((size_t)q.n * dim); q.weight.resize(q.n);
  for (int p = 0; p < q.n; ++p) {
    int r = p; double wt = 1;
    for (int j = dim - 1; j >= 0; --j) { int i = r % m; r /= m; q.pos[(size_t)p * dim + j] = x[i]; wt *= w[i]; }
    q.weight[p] = wt;
  }
  return q;
}

// element geometry pieces  [EXT AxisAlignedCubeGeometry]
static void elementGeometry(const Grid& g, const int ec[3], double lower[3], double upper[3], double jacInv[3], double& integrationElement) {
  integrationElement = 1;
  for (int j = 0; j < g.dim; ++j) {
    lower[j] = g.coordinate(j, g.off[j] + ec[j]);
    upper[j] = g.coordinate(j, g.off[j] + ec[j] + 1);
    jacInv[j] = 1.0 / (upper[j] - lower[j]);
    integrationElement *= upper[j] - lower[j];
  }
}

// getLocalMatrix, poisson-pq2.cc:35-87
static void getLocalMatrix(const Basis& b, size_t e, std::vector<double>& elementMatrix) {
  const Grid& g = b.grid; const int dim = g.dim;
  const LagrangeCubeFE& fe = *b.leaves[0].fe;
  const int n = fe.size();
  elementMatrix.assign((size_t)n * n, 0.0);
  int ec[3]; g.elementCoords(e, ec);
  double lower[3], upper[3], jacInv[3], integrationElement;
  elementGeometry(g, ec, lower, upper, jacInv, integrationElement);
  const int order = 2 * (dim * fe.k - 1);                  // localBasis().order() == k
  const QuadratureRule quad = quadratureRule(dim, order < 0 ? 0 : order);
  std::vector<double> referenceJacobians((size_t)n * dim), jacobians((size_t)n * dim);
  for (int pt = 0; pt < quad.n; ++pt) {
    const double* quadPos = &quad.pos[(size_t)pt * dim];
    fe.evaluateJacobian(quadPos, referenceJacobians.data());
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < dim; ++j) jacobians[(size_t)i * dim + j] = referenceJacobians[(size_t)i * dim + j] * jacInv[j];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double dot = 0;
        for (int d = 0; d < dim; ++d) dot += jacobians[(size_t)i * dim + d] * jacobians[(size_t)j * dim + d];
        elementMatrix[(size_t)i * n + j] += dot * quad.weight[pt] * integrationElement;
      }
  }
}

// getVolumeTerm, poisson-pq2.cc:91-134 with localVolumeTerm = f o geometry.global
static void getVolumeTerm(const Basis& b, size_t e, const dfx_function& f, std::vector<double>& localRhs) {
  const Grid& g = b.grid; const int dim = g.dim;
  const LagrangeCubeFE& fe = *b.leaves[0].fe;
  const int n = fe.size();
  localRhs.assign(n, 0.0);
  int ec[3]; g.elementCoords(e, ec);
  double lower[3], upper[3], jacInv[3], integrationElement;
  elementGeometry(g, ec, lower, upper, jacInv, integrationElement);
  const QuadratureRule quad = quadratureRule(dim, dim * fe.k);
  std::vector<double> shapeFunctionValues(n);
  for (int pt = 0; pt < quad.n; ++pt) {
    const double* quadPos = &quad.pos[(size_t)pt * dim];
    double x[3] = {0, 0, 0}, y[DFX_MAX_LEAVES];
    for (int j = 0; j < dim; ++j) x[j] = lower[j] + quadPos[j] * (upper[j] - lower[j]);
    evalFunction(f, dim, x, y);
    const double functionValue = y[0];
    fe.evaluateFunction(quadPos, shapeFunctionValues.data());
    for (int i = 0; i < n; ++i) localRhs[i] += shapeFunctionValues[i] * functionValue * quad.weight[pt] * integrationElement;
  }
}

}  // namespace orc

extern "C" {

// getOccupationPattern (poisson-pq2.cc:137-174) + MatrixIndexSet::exportIdx: CSR pattern with
// ascending column indices; rows and columns are flat container positions.  row_ptr has
// dimension+1 entries; col_idx may be NULL (sizes only).  Returns the number of non-zeros.
uint64_t orc_occupation_pattern(void* h, uint64_t* row_ptr, uint32_t* col_idx) {
  auto* b = (orc::Basis*)h;
  b->warmFlat();
  const size_t n = b->pre->dimension();
  std::vector<std::vector<uint32_t>> nb(n);
  std::vector<orc::MultiIndex> idx;
  for (size_t e = 0; e < b->grid.numElements(); ++e) {
    b->bind(e, idx);
    for (size_t i = 0; i < idx.size(); ++i)
      for (size_t j = 0; j < idx.size(); ++j) nb[b->flat(idx[i])].push_back((uint32_t)b->flat(idx[j]));   // nb.add(iIdx, jIdx)
  }
  uint64_t nnz = 0;
  for (size_t r = 0; r < n; ++r) {
    std::sort(nb[r].begin(), nb[r].end());
    nb[r].erase(std::unique(nb[r].begin(), nb[r].end()), nb[r].end());
    row_ptr[r] = nnz;
    if (col_idx) std::copy(nb[r].begin(), nb[r].end(), col_idx + nnz);
    nnz += nb[r].size();
  }
  row_ptr[n] = nnz;
  return nnz;
}

// assembleLaplaceMatrix (poisson-pq2.cc:176-255) for a single-leaf (scalar) basis: values in
// the order of the given CSR pattern, rhs for the volume term f; either may be NULL.
int32_t orc_assemble_laplace(void* h, const uint64_t* row_ptr, const uint32_t* col_idx, double* values,
                             const dfx_function* f, double* rhs) {
  auto* b = (orc::Basis*)h;
  if (b->leaves.size() != 1) return 1;
  const size_t n = b->pre->dimension();
  if (values) std::fill(values, values + row_ptr[n], 0.0);                // matrix = 0
  if (rhs) std::fill(rhs, rhs + n, 0.0);                                  // rhs = 0
  std::vector<orc::MultiIndex> idx;
  std::vector<double> elementMatrix, localRhs;
  for (size_t e = 0; e < b->grid.numElements(); ++e) {
    b->bind(e, idx);
    const size_t nl = idx.size();
    if (values) {
      orc::getLocalMatrix(*b, e, elementMatrix);
      for (size_t i = 0; i < nl; ++i) {
        const size_t row = idx[i][0];
        for (size_t j = 0; j < nl; ++j) {
          const uint32_t col = (uint32_t)idx[j][0];
          const uint32_t* lo = col_idx + row_ptr[row];
          const uint32_t* hi = col_idx + row_ptr[row + 1];
          const uint32_t* it = std::lower_bound(lo, hi, col);
          values[it - col_idx] += elementMatrix[i * nl + j];              // matrix[row][col] += elementMatrix[i][j]
        }
      }
    }
    if (rhs && f) {
      orc::getVolumeTerm(*b, e, *f, localRhs);
      for (size_t i = 0; i < nl; ++i) rhs[idx[i][0]] += localRhs[i];
    }
  }
  return 0;
}

// assembleStokesMatrix (examples/stokes-taylorhood.cc:252-297) with getLocalMatrix (:49-160) for a
// Taylor-Hood tree composite(power<dim>(Q2), Q1) with any index-merging strategies: values in the
// order of the given CSR pattern over flat container positions (the example's nested
// matrix[row[0]][col[0]][row[1]][col[1]] flattened).  Returns non-zero for other trees.
int32_t orc_assemble_stokes(void* h, const uint64_t* row_ptr, const uint32_t* col_idx, double* values) {
  auto* b = (orc::Basis*)h;
  const orc::Grid& g = b->grid;
  const int dim = g.dim;
  if ((int)b->leaves.size() != dim + 1) return 1;
  for (int k = 0; k < dim; ++k) if (b->leaves[k].fe->k != 2) return 1;
  if (b->leaves[dim].fe->k != 1) return 1;
  b->warmFlat();
  const size_t n = b->pre->dimension();
  std::fill(values, values + row_ptr[n], 0.0);                                        // matrix = 0
  const orc::LagrangeCubeFE& velocityLocalFiniteElement = *b->leaves[0].fe;           // child(_0, 0)
  const orc::LagrangeCubeFE& pressureLocalFiniteElement = *b->leaves[dim].fe;         // child(_1)
  const int nv = velocityLocalFiniteElement.size(), np = pressureLocalFiniteElement.size();
  const size_t nl = b->localSize();
  const int order = 2 * (dim * velocityLocalFiniteElement.k - 1);
  const orc::QuadratureRule quad = orc::quadratureRule(dim, order);
  std::vector<orc::MultiIndex> idx;
  std::vector<double> elementMatrix(nl * nl), referenceJacobians((size_t)nv * dim), jacobians((size_t)nv * dim), pressureValues(np);
  for (size_t e = 0; e < g.numElements(); ++e) {
    b->bind(e, idx);
    std::fill(elementMatrix.begin(), elementMatrix.end(), 0.0);
    int ec[3]; g.elementCoords(e, ec);
    double lower[3], upper[3], jacInv[3], integrationElement;
    orc::elementGeometry(g, ec, lower, upper, jacInv, integrationElement);
    for (int pt = 0; pt < quad.n; ++pt) {
      const double* quadPos = &quad.pos[(size_t)pt * dim];
      velocityLocalFiniteElement.evaluateJacobian(quadPos, referenceJacobians.data());
      for (int i = 0; i < nv; ++i)
        for (int j = 0; j < dim; ++j) jacobians[(size_t)i * dim + j] = referenceJacobians[(size_t)i * dim + j] * jacInv[j];
      // velocity-velocity coupling (:111-121)
      for (int i = 0; i < nv; ++i)
        for (int j = 0; j < nv; ++j) {
          double dot = 0;
          for (int d = 0; d < dim; ++d) dot += jacobians[(size_t)i * dim + d] * jacobians[(size_t)j * dim + d];
          for (int k = 0; k < dim; ++k) {
            const size_t row = b->leaves[k].offset + i, col = b->leaves[k].offset + j;
            elementMatrix[row * nl + col] += dot * quad.weight[pt] * integrationElement;
          }
        }
      // velocity-pressure coupling (:127-152)
      pressureLocalFiniteElement.evaluateFunction(quadPos, pressureValues.data());
      for (int i = 0; i < nv; ++i)
        for (int j = 0; j < np; ++j)
          for (int k = 0; k < dim; ++k) {
            const size_t vIndex = b->leaves[k].offset + i, pIndex = b->leaves[dim].offset + j;
            const double v = jacobians[(size_t)i * dim + k] * pressureValues[j] * quad.weight[pt] * integrationElement;
            elementMatrix[vIndex * nl + pIndex] -= v;
            elementMatrix[pIndex * nl + vIndex] -= v;
          }
    }
    for (size_t i = 0; i < nl; ++i) {
      const size_t row = b->flat(idx[i]);
      for (size_t j = 0; j < nl; ++j) {
        const uint32_t col = (uint32_t)b->flat(idx[j]);
        const uint32_t* lo = col_idx + row_ptr[row];
        const uint32_t* hi = col_idx + row_ptr[row + 1];
        const uint32_t* it = std::lower_bound(lo, hi, col);
        values[it - col_idx] += elementMatrix[i * nl + j];          // matrixEntry(matrix, row, col) += elementMatrix[i][j]
      }
    }
  }
  return 0;
}

// Symmetric incorporation of Dirichlet values (poisson-pq2.cc:356-386): rhs -= A x (x holds the
// Dirichlet values, zero elsewhere), Dirichlet rows/columns -> identity, rhs[i] = x[i] there.
void orc_dirichlet(uint64_t n, const uint64_t* row_ptr, const uint32_t* col_idx, double* values,
                   const uint8_t* dirichlet, const double* x, double* rhs) {
  for (uint64_t i = 0; i < n; ++i) {                                      // stiffnessMatrix.mmv(x, rhs)
    double s = rhs[i];
    for (uint64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) s -= values[p] * x[col_idx[p]];
    rhs[i] = s;
  }
  for (uint64_t i = 0; i < n; ++i) {
    if (dirichlet[i]) {
      rhs[i] = x[i];
      for (uint64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) values[p] = (col_idx[p] == i) ? 1.0 : 0.0;
    } else {
      for (uint64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) if (dirichlet[col_idx[p]]) values[p] = 0.0;
    }
  }
}

// y = A x (BCRSMatrix::mv), row sums in storage order
void orc_csr_mv(uint64_t n, const uint64_t* row_ptr, const uint32_t* col_idx, const double* values,
                const double* x, double* y) {
  for (uint64_t i = 0; i < n; ++i) {
    double s = 0;
    for (uint64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) s += values[p] * x[col_idx[p]];
    y[i] = s;
  }
}

// DiscreteGlobalBasisFunction::operator()(x) with x in world coordinates
// (discreteglobalbasisfunction.hh:402-410): HierarchicSearch::findEntity, local function bound to
// that element, evaluation at geometry.local(x); derivative likewise (:637-650).
// [EXT dune-grid HierarchicSearch on a one-level grid]: the first element in iteration order with
// referenceElement.checkInside(geometry.local(x)); [EXT dune-geometry] checkInside of the cube:
// every coordinate in [-tol, 1 + tol], tol = 64 eps.  [EXT AxisAlignedCubeGeometry::local]
// (x - lower) / (upper - lower).  A point no element contains gets element -1 and NaN results
// (the reference throws GridError).
void orc_evaluate_global(void* h, int32_t subtreeNode, const double* coeff, const double* x, uint64_t nPoints,
                         double* values, double* jacobians, int64_t* elementOut) {
  auto* b = (orc::Basis*)h;
  b->warmFlat();
  const orc::Grid& g = b->grid;
  const int dim = g.dim;
  const auto& nd = b->nodes[subtreeNode];
  const int firstLeaf = nd[2], nLeaves = nd[3];
  const double tol = 64 * 2.220446049250313e-16;
  std::vector<orc::MultiIndex> idx;
  std::vector<double> localDoFs(b->localSize()), shapeValues, shapeJacobians;
  for (uint64_t p = 0; p < nPoints; ++p) {
    const double* xp = x + p * dim;
    int64_t found = -1;
    double local[3] = {0, 0, 0}, jacInv[3] = {1, 1, 1};
    for (size_t e = 0; e < g.numElements() && found < 0; ++e) {       // findEntity: linear search over level 0
      int ec[3]; g.elementCoords(e, ec);
      bool inside = true;
      double l[3] = {0, 0, 0}, ji[3] = {1, 1, 1};
      for (int j = 0; j < dim; ++j) {
        const double lo = g.coordinate(j, g.off[j] + ec[j]), up = g.coordinate(j, g.off[j] + ec[j] + 1);
        l[j] = (xp[j] - lo) / (up - lo);
        ji[j] = 1.0 / (up - lo);
        inside = inside && l[j] >= -tol && l[j] <= 1 + tol;
      }
      if (inside) { found = (int64_t)e; for (int j = 0; j < 3; ++j) { local[j] = l[j]; jacInv[j] = ji[j]; } }
    }
    if (elementOut) elementOut[p] = found;
    if (found < 0) {
      if (values) for (int l = 0; l < nLeaves; ++l) values[p * nLeaves + l] = std::nan("");
      if (jacobians) for (int q = 0; q < nLeaves * dim; ++q) jacobians[p * nLeaves * dim + q] = std::nan("");
      continue;
    }
    b->bind((size_t)found, idx);
    for (int i = 0; i < nd[1]; ++i) localDoFs[nd[0] + i] = coeff[b->flat(idx[nd[0] + i])];
    for (int l = 0; l < nLeaves; ++l) {
      const orc::LeafNode& leaf = b->leaves[firstLeaf + l];
      const orc::LagrangeCubeFE& fe = *leaf.fe;
      if (values) {
        shapeValues.resize(fe.size());
        fe.evaluateFunction(local, shapeValues.data());
        double v = 0;
        for (int i = 0; i < fe.size(); ++i) v += localDoFs[leaf.offset + i] * shapeValues[i];
        values[p * nLeaves + l] = v;
      }
      if (jacobians) {
        shapeJacobians.resize((size_t)fe.size() * dim);
        fe.evaluateJacobian(local, shapeJacobians.data());
        double refJac[3] = {0, 0, 0};
        for (int i = 0; i < fe.size(); ++i)
          for (int j = 0; j < dim; ++j) refJac[j] += localDoFs[leaf.offset + i] * shapeJacobians[(size_t)i * dim + j];
        for (int j = 0; j < dim; ++j) jacobians[(p * nLeaves + l) * dim + j] = refJac[j] * jacInv[j];
      }
    }
  }
}

int32_t orc_max_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int32_t)n : 1;
}

}  // extern "C"
