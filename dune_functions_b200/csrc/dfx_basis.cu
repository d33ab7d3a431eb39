// dfx_basis.cu — host side of a basis handle: pre-basis tree, size(prefix), the
// expanded local ansatz tree and the flattening of the tree into per-leaf affine
// index maps that the kernels consume.
//
// Reference behaviour restated here (paths relative to the dune-functions checkout):
//   size(prefix)        dynamicpowerbasis.hh:138-219, compositebasis.hh:184-220,
//                       leafprebasismixin.hh:50-58, taylorhoodbasis.hh:138-153
//   index merging       dynamicpowerbasis.hh:263-371, compositebasis.hh:293-338,
//                       taylorhoodbasis.hh:229-251
//   local ordering      nodes.hh:236-245 (children occupy consecutive blocks)
//   Lagrange DOF layout lagrangebasis.hh:747-762, lagrangedgbasis.hh:56-79
#include <algorithm>
#include <cstring>

#include "dfx_internal.h"

namespace dfx {

static thread_local std::string g_lastError;
void setError(const std::string& msg) { g_lastError = msg; }
int fail(int code, const std::string& msg) { g_lastError = msg; return code; }
const char* lastErrorCStr() { return g_lastError.c_str(); }

struct Aff { u64 a, b; };            // digit = a*j + b
struct LeafIdx { int kind, order; std::vector<Aff> dig; };

struct PreNode {
  int kind = 0, order = 0, ims = 0, C = 0;
  std::vector<std::shared_ptr<PreNode>> ch;
  u64 dimension = 0;   // number of basis functions
  u64 size0 = 0;       // size() of the index tree root below this node
  int maxNodeSize = 0, minDig = 1, maxDig = 1;
};

static u64 ipow(u64 b, int e) { u64 r = 1; while (e-- > 0) r *= b; return r; }

// ---- DOF lattice of a leaf on the local box ---------------------------------
static void fillLayout(DevLayout& L, const DevGrid& g, int kind, int k) {
  std::memset(&L, 0, sizeof(L));
  L.kind = kind; L.k = k; L.nloc = (int)ipow((u64)k + 1, g.dim);
  const unsigned full = (1u << g.dim) - 1u;
  for (unsigned s = 0; s < 8; ++s)
    for (int j = 0; j < 3; ++j) L.csize[s][j] = 1;
  u64 n = 0; int nc = 0;
  if (kind == DFX_NODE_LAGRANGE_DG) {
    // one block of (k+1)^dim per element, nothing elsewhere (lagrangedgbasis.hh:56-79)
    for (int j = 0; j < g.dim; ++j) L.csize[full][j] = g.N[j];
    L.blk[full] = L.nloc;
    L.compOffset[full] = 0;
    L.compCount[full] = (u64)g.N[0] * g.N[1] * g.N[2] * (u64)L.nloc;
    L.order[nc++] = (int)full;
    n = L.compCount[full];
  } else {
    // MCMGMapper: codim 0 first ... vertices last; within a codim Yasp components by
    // ascending shift value.  Block size per entity of dimension m: (k-1)^m, or for
    // k == 0 one DOF per element only (lagrangebasis.hh:747-762).
    for (int codim = 0; codim <= g.dim; ++codim)
      for (unsigned s = 0; s <= full; ++s) {
        const int m = __builtin_popcount(s);
        if (m != g.dim - codim) continue;
        u64 blk = (k == 0) ? (s == full ? 1 : 0) : ipow((u64)(k - 1), m);
        u64 ents = 1;
        for (int j = 0; j < g.dim; ++j) {
          L.csize[s][j] = g.N[j] + (((s >> j) & 1u) ? 0 : 1);
          ents *= (u64)L.csize[s][j];
        }
        L.blk[s] = (int)blk;
        L.compOffset[s] = n;
        L.compCount[s] = ents * blk;
        if (blk > 0) { L.order[nc++] = (int)s; n += ents * blk; }
      }
  }
  L.ncomp = nc;
  L.dimension = n;
}

// ---- pre-basis tree -----------------------------------------------------------
static std::shared_ptr<PreNode> parseTree(const DevGrid& g, const dfx_tree_node* t, int n, int& pos, int& err,
                                          std::string& msg) {
  if (pos >= n) { err = DFX_ERR_INVALID; msg = "tree descriptor ends early"; return nullptr; }
  const dfx_tree_node& nd = t[pos++];
  auto p = std::make_shared<PreNode>();
  p->kind = nd.kind; p->order = nd.order; p->ims = nd.ims;
  switch (nd.kind) {
    case DFX_NODE_LAGRANGE:
      // RangeError of LagrangePreBasis ctor, lagrangebasis.hh:789-796
      if (nd.order < 0) { err = DFX_ERR_RANGE; msg = "Lagrange order must be >= 0"; return nullptr; }
      // :795 restricts order > 17 in 3-d only ("in 1d and 2d any order is supported"); here the bound holds in every
      // dimension: the kernels' per-direction tables and the packed lattice multi-indices are sized for it
      // (documented in include/dfx.h, dfx_basis_create)
      if (nd.order > kMaxContinuousOrder) {
        err = DFX_ERR_RANGE;
        msg = g.dim == 3 ? "Polynomial order >17 is not supported" : "Polynomial order >17 is not supported by this library (kernel table bound; the reference restricts 3-d only)";
        return nullptr; }
      return p;
    case DFX_NODE_LAGRANGE_DG:
      if (nd.order < 0 || nd.order > kMaxDGOrder) { err = DFX_ERR_RANGE; msg = "LagrangeDG order out of range"; return nullptr; }
      return p;
    case DFX_NODE_TAYLOR_HOOD: {
      // taylorhoodbasis.hh:229-251: (0, i*dim + c) / (1, j)  ==  BL( FI(power<dim>(Q2)), Q1 )
      p->kind = DFX_NODE_COMPOSITE; p->ims = DFX_IMS_BLOCKED_LEXICOGRAPHIC;
      auto v = std::make_shared<PreNode>();
      v->kind = DFX_NODE_POWER; v->ims = DFX_IMS_FLAT_INTERLEAVED; v->C = g.dim;
      auto q2 = std::make_shared<PreNode>(); q2->kind = DFX_NODE_LAGRANGE; q2->order = 2;
      auto q1 = std::make_shared<PreNode>(); q1->kind = DFX_NODE_LAGRANGE; q1->order = 1;
      v->ch.push_back(q2);
      p->ch.push_back(v); p->ch.push_back(q1);
      return p; }
    case DFX_NODE_POWER: {
      if (nd.n_children < 1) { err = DFX_ERR_INVALID; msg = "power node needs exponent >= 1"; return nullptr; }
      if (nd.ims < DFX_IMS_FLAT_LEXICOGRAPHIC || nd.ims > DFX_IMS_BLOCKED_INTERLEAVED) {
        err = DFX_ERR_INVALID; msg = "power node needs an index merging strategy"; return nullptr; }
      p->C = nd.n_children;
      auto c = parseTree(g, t, n, pos, err, msg);
      if (!c) return nullptr;
      p->ch.push_back(c);
      return p; }
    case DFX_NODE_COMPOSITE: {
      if (nd.n_children < 1) { err = DFX_ERR_INVALID; msg = "composite node needs children"; return nullptr; }
      // composite offers the lexicographic strategies only (compositebasis.hh:55-73)
      if (nd.ims != DFX_IMS_FLAT_LEXICOGRAPHIC && nd.ims != DFX_IMS_BLOCKED_LEXICOGRAPHIC) {
        err = DFX_ERR_INVALID; msg = "composite node supports FlatLexicographic / BlockedLexicographic only"; return nullptr; }
      for (int i = 0; i < nd.n_children; ++i) {
        auto c = parseTree(g, t, n, pos, err, msg);
        if (!c) return nullptr;
        p->ch.push_back(c);
      }
      return p; }
  }
  err = DFX_ERR_INVALID; msg = "unknown node kind";
  return nullptr;
}

static bool blocked(int ims) { return ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC || ims == DFX_IMS_BLOCKED_INTERLEAVED; }

// dimension / size() / multi-index length bounds, bottom-up
static void finalize(PreNode& p, const DevGrid& g) {
  if (p.kind == DFX_NODE_LAGRANGE || p.kind == DFX_NODE_LAGRANGE_DG) {
    DevLayout L; fillLayout(L, g, p.kind, p.order);
    p.dimension = L.dimension; p.size0 = L.dimension; p.maxNodeSize = L.nloc;
    p.minDig = p.maxDig = 1;
    return;
  }
  for (auto& c : p.ch) finalize(*c, g);
  if (p.kind == DFX_NODE_POWER) {
    PreNode& s = *p.ch[0];
    p.dimension = (u64)p.C * s.dimension;
    p.maxNodeSize = p.C * s.maxNodeSize;
    p.minDig = s.minDig + (blocked(p.ims) ? 1 : 0);
    p.maxDig = s.maxDig + (blocked(p.ims) ? 1 : 0);
    switch (p.ims) {
      case DFX_IMS_FLAT_INTERLEAVED: case DFX_IMS_FLAT_LEXICOGRAPHIC: p.size0 = (u64)p.C * s.size0; break;
      case DFX_IMS_BLOCKED_LEXICOGRAPHIC: p.size0 = (u64)p.C; break;
      case DFX_IMS_BLOCKED_INTERLEAVED: p.size0 = s.size0; break;
    }
  } else {
    p.dimension = 0; p.maxNodeSize = 0; p.size0 = 0; p.minDig = 99; p.maxDig = 0;
    for (auto& c : p.ch) {
      p.dimension += c->dimension; p.maxNodeSize += c->maxNodeSize;
      p.minDig = std::min(p.minDig, c->minDig); p.maxDig = std::max(p.maxDig, c->maxDig);
      p.size0 += c->size0;
    }
    if (p.ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC) { p.size0 = p.ch.size(); ++p.minDig; ++p.maxDig; }
  }
}

// size(prefix): number of values the next digit can take
static u64 sizeOf(const PreNode& p, std::vector<u64> prefix) {
  if (prefix.empty()) return p.size0;
  if (p.kind == DFX_NODE_LAGRANGE || p.kind == DFX_NODE_LAGRANGE_DG) return 0;
  if (p.kind == DFX_NODE_POWER) {
    const PreNode& s = *p.ch[0];
    switch (p.ims) {
      case DFX_IMS_FLAT_INTERLEAVED: prefix[0] /= (u64)p.C; return sizeOf(s, prefix);
      case DFX_IMS_FLAT_LEXICOGRAPHIC: prefix[0] %= s.size0; return sizeOf(s, prefix);
      case DFX_IMS_BLOCKED_LEXICOGRAPHIC: prefix.erase(prefix.begin()); return sizeOf(s, prefix);
      case DFX_IMS_BLOCKED_INTERLEAVED: {
        u64 tail = prefix.back();
        prefix.pop_back();
        if (sizeOf(s, prefix) == 0) return 0;
        prefix.push_back(tail);
        u64 sub = sizeOf(s, prefix);
        return sub == 0 ? (u64)p.C : sub; }
    }
    return 0;
  }
  if (p.ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC) {
    u64 front = prefix[0];
    prefix.erase(prefix.begin());
    return front < p.ch.size() ? sizeOf(*p.ch[front], prefix) : 0;
  }
  for (auto& c : p.ch) {
    if (prefix[0] < c->size0) return sizeOf(*c, prefix);
    prefix[0] -= c->size0;
  }
  return 0;
}

u64 preSize(const dfx_basis_impl* b, const u64* prefix, int len) {
  return sizeOf(*b->pre, std::vector<u64>(prefix, prefix + len));
}

// ---- containerDescriptor() -------------------------------------------------------------------
// The reference describes the nested container that fits a basis' index tree with compile-time types
// (containerdescriptors.hh:91-209); here the same tree at run time.  `ch` holds the stored children:
// all of them for Tuple / Array / Vector, the single prototype for UniformArray / UniformVector.
// Whether makeDescriptor yields a Tuple or an Array depends on the children's TYPES being equal
// (:115-141), so every node carries the spelled-out type.
struct CD { int kind = DFX_CD_UNKNOWN; u64 size = 0; std::vector<CD> ch; };

static std::string cdType(const CD& d) {
  switch (d.kind) {
    case DFX_CD_VALUE: return "Value";
    case DFX_CD_UNIFORM_VECTOR: return "UniformVector<" + cdType(d.ch[0]) + ">";
    case DFX_CD_UNIFORM_ARRAY: return "UniformArray<" + cdType(d.ch[0]) + "," + std::to_string(d.size) + ">";
    case DFX_CD_ARRAY: return "Array<" + cdType(d.ch[0]) + "," + std::to_string(d.size) + ">";
    case DFX_CD_VECTOR: return "Vector<" + (d.ch.empty() ? std::string("?") : cdType(d.ch[0])) + ">";
    case DFX_CD_TUPLE: {
      std::string r = "Tuple<";
      for (size_t i = 0; i < d.ch.size(); ++i) r += (i ? "," : "") + cdType(d.ch[i]);
      return r + ">"; }
  }
  return "Unknown";
}
static CD cdUnknown() { return CD{}; }
static CD cdValue() { CD d; d.kind = DFX_CD_VALUE; return d; }
static CD cdUniformVector(u64 n, CD child) { CD d; d.kind = DFX_CD_UNIFORM_VECTOR; d.size = n; d.ch.push_back(std::move(child)); return d; }
static CD cdUniformArray(u64 n, CD child) { CD d; d.kind = DFX_CD_UNIFORM_ARRAY; d.size = n; d.ch.push_back(std::move(child)); return d; }
// makeDescriptor(child0, children...): Array if all types agree, Tuple otherwise (:115-141)
static CD cdMake(std::vector<CD> children) {
  bool same = true;
  for (size_t i = 1; i < children.size(); ++i) same = same && cdType(children[i]) == cdType(children[0]);
  CD d; d.kind = same ? DFX_CD_ARRAY : DFX_CD_TUPLE; d.size = children.size(); d.ch = std::move(children);
  return d;
}
// Impl::appendToTree(s, tree) with a static s (:282-289, TreeTransform :209-279)
static CD cdAppend(u64 s, const CD& t) {
  switch (t.kind) {
    case DFX_CD_VALUE: return cdUniformArray(s, t);
    case DFX_CD_TUPLE:
    case DFX_CD_ARRAY: { std::vector<CD> c; for (auto& x : t.ch) c.push_back(cdAppend(s, x)); return cdMake(std::move(c)); }
    case DFX_CD_VECTOR: { CD d = t; for (auto& x : d.ch) x = cdAppend(s, x); return d; }
    case DFX_CD_UNIFORM_ARRAY: return cdUniformArray(t.size, cdAppend(s, t.ch[0]));
    case DFX_CD_UNIFORM_VECTOR: return cdUniformVector(t.size, cdAppend(s, t.ch[0]));
  }
  return cdUnknown();
}
// Impl::flatLexicographicN / flatInterleavedN with a static n (:305-420): the children of the child are
// repeated n times, block-wise (lexicographic) or one after the other (interleaved)
static CD cdFlatN(u64 n, const CD& t, bool interleaved) {
  switch (t.kind) {
    case DFX_CD_UNIFORM_VECTOR: return cdUniformVector(t.size * n, t.ch[0]);
    case DFX_CD_UNIFORM_ARRAY: return cdUniformArray(t.size * n, t.ch[0]);
    case DFX_CD_VECTOR:
    case DFX_CD_ARRAY:
    case DFX_CD_TUPLE: {
      const size_t m = t.ch.size();
      std::vector<CD> c;
      for (size_t i = 0; i < n * m; ++i) c.push_back(t.ch[interleaved ? i / n : i % m]);
      if (t.kind == DFX_CD_TUPLE) return cdMake(std::move(c));
      CD d; d.kind = t.kind; d.size = n * m; d.ch = std::move(c);
      return d; }
  }
  return cdUnknown();
}
static CD cdOf(const PreNode& p) {
  if (p.kind == DFX_NODE_LAGRANGE || p.kind == DFX_NODE_LAGRANGE_DG)
    return cdUniformVector(p.size0, cdValue());                                // leafprebasismixin.hh:66-69
  if (p.kind == DFX_NODE_POWER) {                                              // dynamicpowerbasis.hh:373-387 via powerbasis.hh:119-122
    CD sub = cdOf(*p.ch[0]);
    switch (p.ims) {
      case DFX_IMS_FLAT_INTERLEAVED: return cdFlatN((u64)p.C, sub, true);
      case DFX_IMS_FLAT_LEXICOGRAPHIC: return cdFlatN((u64)p.C, sub, false);
      case DFX_IMS_BLOCKED_LEXICOGRAPHIC: return cdUniformArray((u64)p.C, std::move(sub));
      case DFX_IMS_BLOCKED_INTERLEAVED: return cdAppend((u64)p.C, sub);
    }
    return cdUnknown();
  }
  std::vector<CD> c;                                                           // compositebasis.hh:274-289
  for (auto& q : p.ch) c.push_back(cdOf(*q));
  if (p.ims == DFX_IMS_BLOCKED_LEXICOGRAPHIC) return cdMake(std::move(c));
  // Impl::flatLexicographic (:291-298): only flat children merge
  u64 total = 0;
  for (auto& x : c) {
    if (x.kind != DFX_CD_UNIFORM_VECTOR || x.ch[0].kind != DFX_CD_VALUE) return cdUnknown();
    total += x.size;
  }
  return cdUniformVector(total, cdValue());
}
static void cdFlatten(const CD& d, std::vector<dfx_container_node>& out) {
  dfx_container_node n;
  n.kind = d.kind; n.n_stored = (int32_t)d.ch.size(); n.size = d.size;
  out.push_back(n);
  for (auto& c : d.ch) cdFlatten(c, out);
}
void containerDescriptor(const dfx_basis_impl* b, std::vector<dfx_container_node>& nodes, std::string& type) {
  const CD d = cdOf(*b->pre);
  nodes.clear();
  cdFlatten(d, nodes);
  type = cdType(d);
}

// leaves below p with their multi-index digits as affine maps of the leaf index j
static std::vector<LeafIdx> leavesOf(const PreNode& p) {
  std::vector<LeafIdx> out;
  if (p.kind == DFX_NODE_LAGRANGE || p.kind == DFX_NODE_LAGRANGE_DG) {
    out.push_back({p.kind, p.order, {{1, 0}}});
    return out;
  }
  if (p.kind == DFX_NODE_POWER) {
    const PreNode& s = *p.ch[0];
    auto sub = leavesOf(s);
    for (int c = 0; c < p.C; ++c)
      for (auto l : sub) {
        switch (p.ims) {
          case DFX_IMS_FLAT_INTERLEAVED: l.dig[0].a *= (u64)p.C; l.dig[0].b = l.dig[0].b * (u64)p.C + (u64)c; break;
          case DFX_IMS_FLAT_LEXICOGRAPHIC: l.dig[0].b += (u64)c * s.size0; break;
          case DFX_IMS_BLOCKED_LEXICOGRAPHIC: l.dig.insert(l.dig.begin(), Aff{0, (u64)c}); break;
          case DFX_IMS_BLOCKED_INTERLEAVED: l.dig.push_back(Aff{0, (u64)c}); break;
        }
        out.push_back(l);
      }
    return out;
  }
  u64 firstComponentOffset = 0;
  for (size_t c = 0; c < p.ch.size(); ++c) {
    for (auto l : leavesOf(*p.ch[c])) {
      if (p.ims == DFX_IMS_FLAT_LEXICOGRAPHIC) l.dig[0].b += firstComponentOffset;
      else l.dig.insert(l.dig.begin(), Aff{0, (u64)c});
      out.push_back(l);
    }
    firstComponentOffset += p.ch[c]->size0;
  }
  return out;
}

// expanded local ansatz tree in preorder (power nodes replicate their child)
static int expand(const PreNode& p, std::vector<HostNode>& nodes, int& offset, int& leafCounter) {
  int id = (int)nodes.size();
  nodes.push_back(HostNode{p.kind, p.order, p.ims, offset, 0, leafCounter, 0, {}});
  if (p.kind == DFX_NODE_LAGRANGE || p.kind == DFX_NODE_LAGRANGE_DG) {
    offset += p.maxNodeSize; ++leafCounter;
  } else if (p.kind == DFX_NODE_POWER) {
    for (int c = 0; c < p.C; ++c) { int ch = expand(*p.ch[0], nodes, offset, leafCounter); nodes[id].children.push_back(ch); }
  } else {
    for (auto& c : p.ch) { int ch = expand(*c, nodes, offset, leafCounter); nodes[id].children.push_back(ch); }
  }
  nodes[id].size = offset - nodes[id].localOffset;
  nodes[id].nLeaves = leafCounter - nodes[id].firstLeaf;
  return id;
}

// ---- flat container position = lexicographic rank of the multi-index ----------
struct RankLeaf { u64 n; std::vector<Aff> dig; };

// number of j' in [0, L.n) whose multi-index is lexicographically smaller than D
static u64 countLess(const RankLeaf& L, const std::vector<u64>& D) {
  const size_t len = L.dig.size();
  for (size_t p = 0;; ++p) {
    if (p >= len) return p >= D.size() ? 0 : L.n;
    if (p >= D.size()) return 0;
    const u64 a = L.dig[p].a, b = L.dig[p].b;
    if (a == 0) {
      if (b < D[p]) return L.n;
      if (b > D[p]) return 0;
      continue;
    }
    u64 less = 0;
    if (D[p] > b) { less = (D[p] - b + a - 1) / a; if (less > L.n) less = L.n; }
    if (D[p] >= b && (D[p] - b) % a == 0 && (D[p] - b) / a < L.n) {
      for (size_t q = p + 1;; ++q) {
        if (q >= len) return less + (q >= D.size() ? 0 : 1);
        if (q >= D.size()) return less;
        if (L.dig[q].b < D[q]) return less + 1;
        if (L.dig[q].b > D[q]) return less;
      }
    }
    return less;
  }
}

static u64 rankOf(const std::vector<RankLeaf>& leaves, size_t l, u64 j) {
  std::vector<u64> D;
  for (auto& d : leaves[l].dig) D.push_back(d.a * j + d.b);
  u64 r = 0;
  for (auto& L : leaves) r += countLess(L, D);
  return r;
}

int buildBasis(const dfx_grid_desc* gd, const dfx_tree_node* t, int n, int device, dfx_basis_impl** out) {
  if (!gd || !t || !out || n < 1) return fail(DFX_ERR_INVALID, "null argument");
  if (gd->dim < 1 || gd->dim > 3) return fail(DFX_ERR_INVALID, "grid dimension must be 1, 2 or 3");
  DevGrid g; std::memset(&g, 0, sizeof(g));
  g.dim = gd->dim;
  for (int j = 0; j < 3; ++j) { g.N[j] = 1; g.G[j] = 1; g.off[j] = 0; g.lower[j] = 0; g.h[j] = 1; }
  for (int j = 0; j < gd->dim; ++j) {
    if (gd->n_global[j] < 1 || gd->local_n[j] < 1 || gd->local_begin[j] < 0 ||
        gd->local_begin[j] + gd->local_n[j] > gd->n_global[j])
      return fail(DFX_ERR_INVALID, "local box must be a non-empty part of the grid");
    if (!(gd->upper[j] > gd->lower[j])) return fail(DFX_ERR_INVALID, "grid needs upper > lower");
    g.N[j] = gd->local_n[j]; g.G[j] = gd->n_global[j]; g.off[j] = gd->local_begin[j];
    g.lower[j] = gd->lower[j];
    g.h[j] = (gd->upper[j] - gd->lower[j]) / gd->n_global[j];
  }
  int pos = 0, err = 0; std::string msg;
  auto pre = parseTree(g, t, n, pos, err, msg);
  if (!pre) return fail(err, msg);
  if (pos != n) return fail(DFX_ERR_INVALID, "tree descriptor has trailing nodes");
  finalize(*pre, g);

  auto b = std::make_unique<dfx_basis>();
  b->gridDesc = *gd; b->device = device; b->pre = pre;
  b->treeDesc.assign(t, t + n);
  b->dimension = pre->dimension; b->minDigits = pre->minDig; b->maxDigits = pre->maxDig;
  if (pre->maxDig > DFX_MAX_DIGITS) return fail(DFX_ERR_NOT_IMPLEMENTED, "multi-index longer than DFX_MAX_DIGITS");

  int offset = 0, leafCounter = 0;
  expand(*pre, b->nodes, offset, leafCounter);
  auto leaves = leavesOf(*pre);
  if ((int)leaves.size() > DFX_MAX_LEAVES) return fail(DFX_ERR_NOT_IMPLEMENTED, "more than DFX_MAX_LEAVES leaf nodes");

  DevBasis& D = b->dev; std::memset(&D, 0, sizeof(D));
  D.grid = g; D.nLeaves = (int)leaves.size(); D.nlocTotal = offset; D.maxDigits = pre->maxDig;
  // layouts: one per distinct (kind, order)
  std::vector<RankLeaf> rl;
  int lo = 0;
  for (size_t l = 0; l < leaves.size(); ++l) {
    int li = -1;
    for (int q = 0; q < D.nLayouts; ++q)
      if (D.layouts[q].kind == leaves[l].kind && D.layouts[q].k == leaves[l].order) li = q;
    if (li < 0) {
      if (D.nLayouts == kMaxLayouts) return fail(DFX_ERR_NOT_IMPLEMENTED, "more than 4 distinct leaf types in one tree");
      li = D.nLayouts++;
      fillLayout(D.layouts[li], g, leaves[l].kind, leaves[l].order);
    }
    DevLeaf& dl = D.leaves[l];
    dl.layout = li; dl.localOffset = lo; dl.nloc = D.layouts[li].nloc; dl.ndig = (int)leaves[l].dig.size();
    for (int p = 0; p < dl.ndig; ++p) { dl.digA[p] = leaves[l].dig[p].a; dl.digB[p] = leaves[l].dig[p].b; }
    lo += dl.nloc;
    rl.push_back(RankLeaf{D.layouts[li].dimension, leaves[l].dig});
  }
  // flat positions: rank is affine in j for every strategy combination; verified here
  for (size_t l = 0; l < leaves.size(); ++l) {
    const u64 nl = rl[l].n;
    u64 r0 = rankOf(rl, l, 0);
    u64 a = nl > 1 ? rankOf(rl, l, 1) - r0 : 1;
    for (u64 j : {(u64)2, nl / 2, nl - 1})
      if (j < nl && rankOf(rl, l, j) != r0 + a * j)
        return fail(DFX_ERR_NOT_IMPLEMENTED, "index tree without an affine flat layout");
    D.leaves[l].posA = a; D.leaves[l].posB = r0;
  }
  // per local index: leaf id and lattice multi-index alpha (i = sum alpha_j (k+1)^j)
  b->localTab.resize(offset);
  for (int l = 0; l < D.nLeaves; ++l) {
    const int k = D.layouts[D.leaves[l].layout].k;
    for (int i = 0; i < D.leaves[l].nloc; ++i) {
      int ii = i, a[3];
      for (int j = 0; j < 3; ++j) { a[j] = ii % (k + 1); ii /= (k + 1); }
      b->localTab[D.leaves[l].localOffset + i] = (unsigned)l | (a[0] << 5) | (a[1] << 10) | (a[2] << 15);
    }
  }
  // ... and everything else of the index table that depends on the local index only (IdxLi)
  b->idxTab.resize(offset);
  for (int l = 0; l < D.nLeaves; ++l) {
    const DevLeaf& leaf = D.leaves[l];
    const DevLayout& L = D.layouts[leaf.layout];
    for (int i = 0; i < leaf.nloc; ++i) {
      int ii = i, a[3];
      for (int j = 0; j < 3; ++j) { a[j] = ii % (L.k + 1); ii /= (L.k + 1); }
      const int zero[3] = {0, 0, 0};
      const u64 jbase = layoutIndex(L, D.grid, zero, a, i);
      unsigned s = 0;
      if (L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0) s = (1u << D.grid.dim) - 1u;
      else
        for (int j = 0; j < 3; ++j) if (a[j] > 0 && a[j] < L.k) s |= 1u << j;
      IdxLi& t = b->idxTab[leaf.localOffset + i];
      const u64 blk = (u64)L.blk[s];
      t.P0 = leaf.posA * jbase + leaf.posB;
      t.S = (unsigned)(leaf.posA * blk);
      t.cs0 = (unsigned)L.csize[s][0]; t.cs1 = (unsigned)L.csize[s][1];
      t.ndig = (unsigned)leaf.ndig;
      for (int p = 0; p < DFX_MAX_DIGITS; ++p) {
        t.dA[p] = p < leaf.ndig ? leaf.digA[p] * blk : 0;
        t.dB[p] = p < leaf.ndig ? leaf.digA[p] * jbase + leaf.digB[p] : 0;
      }
    }
  }
  *out = b.release();
  return DFX_OK;
}

// LagrangeCubeLocalBasis 1-d factors: l_i(x) = prod_{j != i} (k x - j)/(i - j) and derivative
double lagrangeP(int k, int i, double x) {
  double r = 1.0;
  for (int j = 0; j <= k; ++j) if (j != i) r *= (k * x - j) / (i - j);
  return r;
}
double lagrangeDP(int k, int i, double x) {
  double r = 0.0;
  for (int j = 0; j <= k; ++j) if (j != i) {
    double prod = (k * 1.0) / (i - j);
    for (int l = 0; l <= k; ++l) if (l != i && l != j) prod *= (k * x - l) / (i - l);
    r += prod;
  }
  return r;
}

}  // namespace dfx
