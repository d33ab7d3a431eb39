// dfx_kernels.cu — generic (any dimension / order / tree) kernels of the path.
// The specialised fast paths live in dfx_eval_fast.cu.
//
//   k_indices          DefaultLocalView::bind -> PreBasis::indices     (defaultlocalview.hh:95-101)
//   k_interpolate      interpolate element loop, owner-writes          (interpolate.hh:254-259, :151-173)
//   k_points           positions of the interpolation nodes            (form (2) of f, include/dfx.h)
//   k_from_values      masked copy of caller-evaluated nodal values
//   k_eval_generic     LocalFunction::bind + operator() (+derivative)  (discreteglobalbasisfunction.hh:124-148,
//                                                                        :336-371, :583-627)
//   k_halo_pack/unpack one lattice plane of shared DOFs <-> contiguous buffer
#include <atomic>

#include "dfx_device.cuh"
#include "dfx_launch.h"

namespace dfx {

std::atomic<unsigned long long> g_launchCount{0};

// ---------------------------------------------------------------------------
__device__ __forceinline__ int findLeaf(const DevBasis& B, int li) {
  int l = 0;
  while (l + 1 < B.nLeaves && li >= B.leaves[l + 1].localOffset) ++l;
  return l;
}

// One thread per (element, local index).  T = uint32_t / uint64_t.
template <class T, bool MULTI>
__global__ void __launch_bounds__(256)
k_indices(const __grid_constant__ DevBasis B, u64 e0, u64 total, T* __restrict__ out, int stride) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const unsigned nloc = (unsigned)B.nlocTotal;
  const u64 el = t / nloc;
  const int li = (int)(t - el * nloc);
  const u64 e = e0 + el;
  int ec[3];
  ec[0] = (int)(e % (u64)B.grid.N[0]);
  const u64 r = e / (u64)B.grid.N[0];
  ec[1] = (int)(r % (u64)B.grid.N[1]);
  ec[2] = (int)(r / (u64)B.grid.N[1]);
  const int l = findLeaf(B, li);
  const DevLeaf& leaf = B.leaves[l];
  const DevLayout& L = B.layouts[leaf.layout];
  const int i = li - leaf.localOffset;
  int a[3];
  localAlpha(L.k, i, a);
  const u64 j = layoutIndexOriented(L, B.grid, B.vid, ec, a, i);
  if (MULTI) {
    for (int p = 0; p < stride; ++p)
      out[t * (u64)stride + p] = p < leaf.ndig ? (T)(leaf.digA[p] * j + leaf.digB[p]) : (T)0;
  } else {
    out[t] = (T)(leaf.posA * j + leaf.posB);
  }
}

template <class T, bool MULTI>
static int launchIndices(const dfx_basis* b, u64 e0, u64 e1, T* out, int stride, cudaStream_t st) {
  const u64 total = (e1 - e0) * (u64)b->dev.nlocTotal;
  if (total == 0) return DFX_OK;
  const u64 blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  k_indices<T, MULTI><<<(unsigned)blocks, 256, 0, st>>>(b->dev, e0, total, out, stride);
  return postLaunch("k_indices");
}

int indicesMulti32(const dfx_basis* b, u64 e0, u64 e1, uint32_t* out, int stride, cudaStream_t st) {
  return launchIndices<uint32_t, true>(b, e0, e1, out, stride, st);
}
int indicesMulti64(const dfx_basis* b, u64 e0, u64 e1, u64* out, int stride, cudaStream_t st) {
  return launchIndices<u64, true>(b, e0, e1, out, stride, st);
}
int indicesFlat32(const dfx_basis* b, u64 e0, u64 e1, uint32_t* out, cudaStream_t st) {
  return launchIndices<uint32_t, false>(b, e0, e1, out, 1, st);
}
int indicesFlat64(const dfx_basis* b, u64 e0, u64 e1, u64* out, cudaStream_t st) {
  return launchIndices<u64, false>(b, e0, e1, out, 1, st);
}

// ---------------------------------------------------------------------------
// Position of the interpolation node of scalar DOF t of layout L, as the LAST
// element (in grid-view iteration order of the global grid) that contains the
// node computes it: that element's write is the one that survives in the
// reference loop (interpolate.hh:254-259 visits elements in order and :166-171
// overwrites).  Returns false if t is out of range.
__device__ __forceinline__ void nodePosition(const DevBasis& B, const DevLayout& L, u64 t, double x[3]) {
  const DevGrid& g = B.grid;
  int ie[3], a[3];
  const int k = L.k;
  dofOwnerNode(B, L, t, ie, a);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j >= g.dim) { x[j] = 0.0; continue; }
    // Lagrange node alpha/k (the element centre for k = 0), LagrangeCubeLocalInterpolation
    const double xhat = k == 0 ? 0.5 : __ddiv_rn(__dmul_rn(1.0, (double)a[j]), (double)k);
    x[j] = elementGlobal(g, j, ie[j], xhat);
  }
}

// One thread per scalar DOF of one layout; f evaluated once per node and scattered to
// every leaf of the subtree that lives on this lattice (a power node's components).
template <class F>
__global__ void __launch_bounds__(256)
k_interpolate(const __grid_constant__ DevBasis B, int layoutId, int firstLeaf, int nLeaves, const F f,
              double* __restrict__ coeff, const uint8_t* __restrict__ mask) {
  const DevLayout& L = B.layouts[layoutId];
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L.dimension) return;
  double x[3];
  nodePosition(B, L, t, x);
  const double cm = f.common(x[0], x[1], x[2]);
  const bool scalar = f.ncomp() == 1;
  for (int l = 0; l < nLeaves; ++l) {
    const DevLeaf& leaf = B.leaves[firstLeaf + l];
    if (leaf.layout != layoutId) continue;
    const u64 p = leaf.posA * t + leaf.posB;
    if (mask == nullptr || mask[p]) coeff[p] = f.comp(scalar ? 0 : l, cm, x[0], x[1], x[2]);
  }
}

int interpolateRegistry(const dfx_basis* b, int firstLeaf, int nLeaves, const dfx_function& fn, double* coeff,
                        const uint8_t* mask, cudaStream_t st) {
  RegistryFn f; f.f = fn; f.dim = b->dev.grid.dim;
  for (int q = 0; q < b->dev.nLayouts; ++q) {
    bool used = false;
    for (int l = 0; l < nLeaves; ++l) used = used || b->dev.leaves[firstLeaf + l].layout == q;
    if (!used) continue;
    const u64 n = b->dev.layouts[q].dimension;
    const u64 blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "basis too large for one launch");
    k_interpolate<RegistryFn><<<(unsigned)blocks, 256, 0, st>>>(b->dev, q, firstLeaf, nLeaves, f, coeff, mask);
    int rc = postLaunch("k_interpolate");
    if (rc) return rc;
  }
  return DFX_OK;
}

__global__ void __launch_bounds__(256)
k_points(const __grid_constant__ DevBasis B, int layoutId, double* __restrict__ xyz, int* __restrict__ leafOut) {
  const DevLayout& L = B.layouts[layoutId];
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L.dimension) return;
  double x[3];
  nodePosition(B, L, t, x);
  for (int l = 0; l < B.nLeaves; ++l) {
    const DevLeaf& leaf = B.leaves[l];
    if (leaf.layout != layoutId) continue;
    const u64 p = leaf.posA * t + leaf.posB;
    for (int j = 0; j < B.grid.dim; ++j) xyz[p * B.grid.dim + j] = x[j];
    leafOut[p] = l;
  }
}

int interpolationPoints(const dfx_basis* b, double* xyz, int* leafOut, cudaStream_t st) {
  for (int q = 0; q < b->dev.nLayouts; ++q) {
    const u64 n = b->dev.layouts[q].dimension;
    const u64 blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "basis too large for one launch");
    k_points<<<(unsigned)blocks, 256, 0, st>>>(b->dev, q, xyz, leafOut);
    int rc = postLaunch("k_points");
    if (rc) return rc;
  }
  return DFX_OK;
}

__global__ void __launch_bounds__(256)
k_from_values(const __grid_constant__ DevBasis B, int leafId, const double* __restrict__ values,
              double* __restrict__ coeff, const uint8_t* __restrict__ mask) {
  const DevLeaf& leaf = B.leaves[leafId];
  const u64 n = B.layouts[leaf.layout].dimension;
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const u64 p = leaf.posA * t + leaf.posB;
  if (mask == nullptr || mask[p]) coeff[p] = values[p];
}

int interpolateFromValues(const dfx_basis* b, int firstLeaf, int nLeaves, const double* values, double* coeff,
                          const uint8_t* mask, cudaStream_t st) {
  for (int l = firstLeaf; l < firstLeaf + nLeaves; ++l) {
    const u64 n = b->dev.layouts[b->dev.leaves[l].layout].dimension;
    const u64 blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "basis too large for one launch");
    k_from_values<<<(unsigned)blocks, 256, 0, st>>>(b->dev, l, values, coeff, mask);
    int rc = postLaunch("k_from_values");
    if (rc) return rc;
  }
  return DFX_OK;
}

// ---------------------------------------------------------------------------
// interpolate(basis, coeff, gridFunction) for a grid function that is evaluated element by
// element (a DiscreteGlobalBasisFunction of another basis): `vals[(e - e0)][i][c]` holds the
// function at local interpolation node i of element e (range entry c).  One thread per
// (element, local node): it writes iff its element is the LAST one of [e0, e1) — in grid-view
// iteration order — that contains the node, i.e. unless the node sits on an upper face whose
// neighbour is also in the range; later bands overwrite earlier ones like the reference's loop.
struct ScatterArgs {
  int layoutId, firstLeaf, nLeaves, ncompVals;
  int tileLog2;                        // a block handles 2^tileLog2 consecutive elements
  u64 e0, e1;
  FastDiv divN0, divN1, divK1;
};

// A block stages the nodal values of a tile of consecutive elements in shared memory (coalesced
// read of `vals`, which is element-major) and then walks (local node, element) with the element
// fastest, so that the lanes of a warp write consecutive entities of one index-set component.
__global__ void __launch_bounds__(256)
k_scatter_nodes(const __grid_constant__ DevBasis B, const __grid_constant__ ScatterArgs A,
                const double* __restrict__ vals, double* __restrict__ coeff, const uint8_t* __restrict__ mask) {
  extern __shared__ __align__(16) double tile[];
  // the layout's per-component tables are indexed by a per-lane shift mask: shared memory, not
  // the (serialising) indexed-constant path
  __shared__ u64 sOff[8];
  __shared__ unsigned sCs0[8], sCs1[8], sBlk[8];
  const DevLayout& L = B.layouts[A.layoutId];
  const DevGrid& g = B.grid;
  if (threadIdx.x < 8) {
    sOff[threadIdx.x] = L.compOffset[threadIdx.x];
    sCs0[threadIdx.x] = (unsigned)L.csize[threadIdx.x][0];
    sCs1[threadIdx.x] = (unsigned)L.csize[threadIdx.x][1];
    sBlk[threadIdx.x] = (unsigned)L.blk[threadIdx.x];
  }
  const unsigned tileEl = 1u << A.tileLog2;
  const u64 elBase = (u64)blockIdx.x * tileEl;                 // relative to e0
  const u64 nEl = A.e1 - A.e0;
  const unsigned cnt = (unsigned)(nEl - elBase < (u64)tileEl ? nEl - elBase : (u64)tileEl);
  const unsigned perEl = (unsigned)L.nloc * (unsigned)A.ncompVals;
  const double* src = vals + elBase * (u64)perEl;
  for (unsigned w = threadIdx.x; w < cnt * perEl; w += blockDim.x) tile[w] = src[w];
  __syncthreads();
  const int k = L.k;
  const unsigned last = (unsigned)A.e1;
  const unsigned work = (unsigned)L.nloc << A.tileLog2;
  for (unsigned w = threadIdx.x; w < work; w += blockDim.x) {
    const unsigned elLoc = w & (tileEl - 1u), i = w >> A.tileLog2;
    if (elLoc >= cnt) continue;
    const unsigned e = (unsigned)(A.e0 + elBase) + elLoc;
    unsigned q, ec0, ec1, ec2;
    A.divN0.divmod(e, q, ec0);
    A.divN1.divmod(q, ec2, ec1);
    u64 jdx;
    if (L.kind == DFX_NODE_LAGRANGE && k > 0) {
      unsigned a0, a1, a2, r;
      A.divK1.divmod(i, r, a0);
      A.divK1.divmod(r, a2, a1);
      // the element one step up in direction j: inside [e0, e1)?  then it writes this node, not us
      if ((int)a0 == k && ec0 + 1 < (unsigned)g.N[0] && e + 1 < last) continue;
      if (g.dim > 1 && (int)a1 == k && ec1 + 1 < (unsigned)g.N[1] && e + (unsigned)g.N[0] < last) continue;
      if (g.dim > 2 && (int)a2 == k && ec2 + 1 < (unsigned)g.N[2] && (u64)e + (u64)g.N[0] * (u64)g.N[1] < (u64)last) continue;
      unsigned s = 0, fr = 0, frs = 1;
      unsigned c0 = ec0, c1 = ec1, c2 = ec2;
      if (a0 > 0 && (int)a0 < k) { s |= 1u; fr += (a0 - 1) * frs; frs *= (unsigned)(k - 1); } else c0 += (int)a0 == k ? 1u : 0u;
      if (a1 > 0 && (int)a1 < k) { s |= 2u; fr += (a1 - 1) * frs; frs *= (unsigned)(k - 1); } else c1 += (int)a1 == k ? 1u : 0u;
      if (a2 > 0 && (int)a2 < k) { s |= 4u; fr += (a2 - 1) * frs; } else c2 += (int)a2 == k ? 1u : 0u;
      if (B.vid != nullptr && k > 2 && g.dim > 1) {
        const int cc[3] = {(int)c0, (int)c1, (int)c2};
        fr = orientFaceDOF<false>(g, B.vid, k, s, cc, fr);
      }
      const u64 ent = (u64)c0 + (u64)sCs0[s] * ((u64)c1 + (u64)sCs1[s] * (u64)c2);
      jdx = sOff[s] + ent * (u64)sBlk[s] + fr;
    } else {
      const u64 ei = (u64)ec0 + (u64)g.N[0] * ((u64)ec1 + (u64)g.N[1] * (u64)ec2);
      jdx = ei * (u64)L.nloc + i;               // element-local DOFs (LagrangeDG, Lagrange<0>)
    }
    const double* v = tile + (elLoc * (unsigned)L.nloc + i) * (unsigned)A.ncompVals;
    for (int l = 0; l < A.nLeaves; ++l) {
      const DevLeaf& leaf = B.leaves[A.firstLeaf + l];
      if (leaf.layout != A.layoutId) continue;
      const u64 p = leaf.posA * jdx + leaf.posB;
      if (mask == nullptr || mask[p]) coeff[p] = v[A.ncompVals == 1 ? 0 : l];
    }
  }
}

int scatterNodes(const dfx_basis* b, int layoutId, int firstLeaf, int nLeaves, int ncompVals, u64 e0, u64 e1,
                 const double* vals, double* coeff, const uint8_t* mask, cudaStream_t st) {
  const DevLayout& L = b->dev.layouts[layoutId];
  if (e1 == e0) return DFX_OK;
  if (e1 > 0xffffffffull) return fail(DFX_ERR_RANGE, "more than 2^32 elements in one box");
  // tile: as many elements (a power of two, <= 64) as fit 40 KB of shared memory
  const size_t perEl = (size_t)L.nloc * ncompVals * sizeof(double);
  int tl = 6;
  while (tl > 0 && (perEl << tl) > 40 * 1024) --tl;
  const size_t smem = perEl << tl;
  if (smem > 200 * 1024) return fail(DFX_ERR_NOT_IMPLEMENTED, "element too large for the scatter tile");
  static std::atomic<unsigned long long> configured{0};
  if (smem > 48 * 1024) { int rc = optInSharedMemory(k_scatter_nodes, b->device, configured); if (rc) return rc; }
  const u64 blocks = ((e1 - e0) + (1ull << tl) - 1) >> tl;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  ScatterArgs A;
  A.layoutId = layoutId; A.firstLeaf = firstLeaf; A.nLeaves = nLeaves; A.ncompVals = ncompVals; A.e0 = e0; A.e1 = e1;
  A.tileLog2 = tl;
  A.divN0 = FastDiv((unsigned)b->dev.grid.N[0]); A.divN1 = FastDiv((unsigned)b->dev.grid.N[1]);
  A.divK1 = FastDiv((unsigned)(L.k + 1));
  k_scatter_nodes<<<(unsigned)blocks, 256, smem, st>>>(b->dev, A, vals, coeff, mask);
  return postLaunch("k_scatter_nodes");
}

// ---------------------------------------------------------------------------
// Generic evaluation: one thread per (element, point); shape tables
// shape[layout][q][i], dshape[layout][q][i][dim] precomputed on the host exactly as
// LagrangeCubeLocalBasis::evaluateFunction / evaluateJacobian produce them.
// Sums run over i ascending like the reference's axpy loop.
template <bool VAL, bool JAC>
__global__ void __launch_bounds__(128)
k_eval_generic(const __grid_constant__ DevBasis B, int firstLeaf, int nLeaves, const double* __restrict__ coeff,
               const double* __restrict__ shape, const double* __restrict__ dshape, ShapeOffsets so, int nq, u64 e0,
               u64 total, double* __restrict__ values, double* __restrict__ jac) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const DevGrid& g = B.grid;
  const u64 el = t / (u64)nq;
  const int q = (int)(t - el * (u64)nq);
  const u64 e = e0 + el;
  int ec[3];
  ec[0] = (int)(e % (u64)g.N[0]);
  const u64 r = e / (u64)g.N[0];
  ec[1] = (int)(r % (u64)g.N[1]);
  ec[2] = (int)(r / (u64)g.N[1]);
  double jinv[3] = {0, 0, 0};
  if (JAC)
    for (int j = 0; j < g.dim; ++j) jinv[j] = elementJacInv(g, j, g.off[j] + ec[j]);
  for (int l = 0; l < nLeaves; ++l) {
    const DevLeaf& leaf = B.leaves[firstLeaf + l];
    const DevLayout& L = B.layouts[leaf.layout];
    const double* sh = shape + so.v[leaf.layout] + (size_t)q * L.nloc;
    const double* dsh = dshape + so.d[leaf.layout] + (size_t)q * L.nloc * g.dim;
    double v = 0, rj[3] = {0, 0, 0};
    for (int i = 0; i < L.nloc; ++i) {
      int a[3];
      localAlpha(L.k, i, a);
      const u64 j = layoutIndexOriented(L, g, B.vid, ec, a, i);
      const double c = coeff[leaf.posA * j + leaf.posB];
      if (VAL) v = __dadd_rn(v, __dmul_rn(c, sh[i]));
      if (JAC)
        for (int d = 0; d < g.dim; ++d) rj[d] = __dadd_rn(rj[d], __dmul_rn(c, dsh[i * g.dim + d]));
    }
    if (VAL) values[t * (u64)nLeaves + l] = v;
    if (JAC)
      for (int d = 0; d < g.dim; ++d) jac[(t * (u64)nLeaves + l) * g.dim + d] = __dmul_rn(rj[d], jinv[d]);
  }
}

int evalGeneric(const dfx_basis* b, int firstLeaf, int nLeaves, const double* coeff, const double* shape,
                const double* dshape, const ShapeOffsets& so, int nq, u64 e0, u64 e1, double* values, double* jac,
                cudaStream_t st) {
  const u64 total = (e1 - e0) * (u64)nq;
  if (total == 0) return DFX_OK;
  const u64 blocks = (total + 127) / 128;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  if (values && jac)
    k_eval_generic<true, true><<<(unsigned)blocks, 128, 0, st>>>(b->dev, firstLeaf, nLeaves, coeff, shape, dshape, so, nq, e0, total, values, jac);
  else if (values)
    k_eval_generic<true, false><<<(unsigned)blocks, 128, 0, st>>>(b->dev, firstLeaf, nLeaves, coeff, shape, dshape, so, nq, e0, total, values, jac);
  else
    k_eval_generic<false, true><<<(unsigned)blocks, 128, 0, st>>>(b->dev, firstLeaf, nLeaves, coeff, shape, dshape, so, nq, e0, total, values, jac);
  return postLaunch("k_eval_generic");
}

// ---------------------------------------------------------------------------
// LocalFunction::bind for a basis with permuted face DOFs (discreteglobalbasisfunction.hh:124-148 with the
// indices of lagrangebasis.hh:863-869): localDoFs[e][i] = dofs[localView.index(i)] for a range of elements,
// written in the coefficient layout of the COMPANION basis (same tree, every leaf discontinuous), whose
// evaluation then runs on the fast element-local kernels.  One thread per (element, local index of the
// subtree); consecutive threads write consecutive entries of an element.
struct GatherArgs {
  int firstLeaf, nLeaves, firstLocal, nLocal;      // subtree: leaves and local index range
  u64 e0, total;
  u64 posA[DFX_MAX_LEAVES], posB[DFX_MAX_LEAVES];  // companion: flat position = posA * (e * nloc_leaf + i) + posB
};

__global__ void __launch_bounds__(256)
k_gather_local(const __grid_constant__ DevBasis B, const __grid_constant__ GatherArgs A, const double* __restrict__ src,
               double* __restrict__ dst) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.total) return;
  const u64 el = t / (unsigned)A.nLocal;
  const int li = A.firstLocal + (int)(t - el * (unsigned)A.nLocal);
  const u64 e = A.e0 + el;
  int ec[3];
  ec[0] = (int)(e % (u64)B.grid.N[0]);
  const u64 r = e / (u64)B.grid.N[0];
  ec[1] = (int)(r % (u64)B.grid.N[1]);
  ec[2] = (int)(r / (u64)B.grid.N[1]);
  const int l = findLeaf(B, li);
  const DevLeaf& leaf = B.leaves[l];
  const DevLayout& L = B.layouts[leaf.layout];
  const int i = li - leaf.localOffset;
  int a[3];
  localAlpha(L.k, i, a);
  const u64 j = layoutIndexOriented(L, B.grid, B.vid, ec, a, i);
  dst[A.posA[l] * (e * (u64)leaf.nloc + (u64)i) + A.posB[l]] = src[leaf.posA * j + leaf.posB];
}

int gatherLocal(const dfx_basis* b, const dfx_basis* companion, int firstLeaf, int nLeaves, const double* src, double* dst,
                u64 e0, u64 e1, cudaStream_t st) {
  GatherArgs A;
  A.firstLeaf = firstLeaf; A.nLeaves = nLeaves;
  A.firstLocal = b->dev.leaves[firstLeaf].localOffset;
  A.nLocal = 0;
  for (int l = firstLeaf; l < firstLeaf + nLeaves; ++l) A.nLocal += b->dev.leaves[l].nloc;
  for (int l = 0; l < DFX_MAX_LEAVES; ++l) { A.posA[l] = companion->dev.leaves[l].posA; A.posB[l] = companion->dev.leaves[l].posB; }
  A.e0 = e0; A.total = (e1 - e0) * (u64)A.nLocal;
  if (A.total == 0) return DFX_OK;
  const u64 blocks = (A.total + 255) / 256;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  k_gather_local<<<(unsigned)blocks, 256, 0, st>>>(b->dev, A, src, dst);
  return postLaunch("k_gather_local");
}

// ---------------------------------------------------------------------------
// Boundary DOFs: forEachBoundaryDOF (boundarydofs.hh:110-128) + SubEntityDOFs::bind
// (subentitydofs.hh:71-95) in closed form.  A DOF is visited by the reference iff it is
// attached to a sub-entity of a boundary facet, i.e. iff in some direction its entity is
// pinned to a lattice plane (not extended) that is the first or last plane of the GLOBAL
// grid.  DG leaves use the same local keys (LagrangeNode, lagrangedgbasis.hh:37-38), so
// the element's own nodes on a boundary facet count; order 0 has one interior key -> none.
// One thread per scalar DOF of one layout; only boundary DOFs write.
__global__ void __launch_bounds__(256)
k_boundary(const __grid_constant__ DevBasis B, int layoutId, int firstLeaf, int nLeaves, uint8_t* __restrict__ mask) {
  const DevLayout& L = B.layouts[layoutId];
  const DevGrid& g = B.grid;
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L.dimension) return;
  const int k = L.k;
  if (k == 0) return;
  bool bnd = false;
  if (L.kind == DFX_NODE_LAGRANGE_DG) {
    const u64 e = t / (u64)L.nloc;
    int a[3];
    localAlpha(k, (int)(t - e * (u64)L.nloc), a);
    int ie[3];
    ie[0] = (int)(e % (u64)g.N[0]) + g.off[0];
    const u64 r = e / (u64)g.N[0];
    ie[1] = (int)(r % (u64)g.N[1]) + g.off[1];
    ie[2] = (int)(r / (u64)g.N[1]) + g.off[2];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j < g.dim) bnd = bnd || (a[j] == 0 && ie[j] == 0) || (a[j] == k && ie[j] == g.G[j] - 1);
  } else {
    int s = L.order[0];
    for (int q = 1; q < L.ncomp; ++q)
      if (t >= L.compOffset[L.order[q]]) s = L.order[q];
    const u64 ent = (t - L.compOffset[s]) / (u64)L.blk[s];
    int c[3];
    c[0] = (int)(ent % (u64)L.csize[s][0]);
    const u64 r2 = ent / (u64)L.csize[s][0];
    c[1] = (int)(r2 % (u64)L.csize[s][1]);
    c[2] = (int)(r2 / (u64)L.csize[s][1]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (j >= g.dim || ((s >> j) & 1)) continue;
      const int gc = g.off[j] + c[j];
      bnd = bnd || gc == 0 || gc == g.G[j];
    }
  }
  if (!bnd) return;
  for (int l = 0; l < nLeaves; ++l) {
    const DevLeaf& leaf = B.leaves[firstLeaf + l];
    if (leaf.layout == layoutId) mask[leaf.posA * t + leaf.posB] = 1;
  }
}

int boundaryDofs(const dfx_basis* b, int firstLeaf, int nLeaves, uint8_t* mask, cudaStream_t st) {
  for (int q = 0; q < b->dev.nLayouts; ++q) {
    bool used = false;
    for (int l = 0; l < nLeaves; ++l) used = used || b->dev.leaves[firstLeaf + l].layout == q;
    if (!used || b->dev.layouts[q].k == 0) continue;
    const u64 n = b->dev.layouts[q].dimension;
    const u64 blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "basis too large for one launch");
    k_boundary<<<(unsigned)blocks, 256, 0, st>>>(b->dev, q, firstLeaf, nLeaves, mask);
    int rc = postLaunch("k_boundary");
    if (rc) return rc;
  }
  return DFX_OK;
}

// ---------------------------------------------------------------------------
// DiscreteGlobalBasisFunction::operator()(x) for world coordinates
// (discreteglobalbasisfunction.hh:402-410) and its derivative (:637-650).  The reference finds
// the element with a HierarchicSearch (first element in iteration order whose reference cube
// contains geometry.local(x) up to 64 eps); on the structured grid that is, per direction, the
// lowest element whose tolerant interval contains x — a closed form around floor((x - lower)/h).
// One thread per point; shape functions are formed on the fly from the 1-d Lagrange values, the
// sum runs over i ascending with phi_i = ((1 * l_{a0}(x0)) * l_{a1}(x1)) * l_{a2}(x2) like
// LagrangeCubeLocalBasis::evaluateFunction.
__device__ __forceinline__ int locateDirection(const DevGrid& g, int j, double x, double& local, double& jinv) {
  const double tol = 64 * 2.220446049250313e-16;
  int i0 = (int)floor((x - g.lower[j]) / g.h[j]);
  const int first = g.off[j], lastEl = g.off[j] + g.N[j] - 1;
  i0 = i0 < first ? first : (i0 > lastEl ? lastEl : i0);
  for (int i = (i0 > first ? i0 - 1 : i0); i <= (i0 < lastEl ? i0 + 1 : i0); ++i) {
    const double lo = gridCoordinate(g, j, i), up = gridCoordinate(g, j, i + 1);
    const double l = __ddiv_rn(__dsub_rn(x, lo), __dsub_rn(up, lo));
    if (l >= -tol && l <= 1 + tol) { local = l; jinv = __ddiv_rn(1.0, __dsub_rn(up, lo)); return i - first; }
  }
  return -1;
}

// l_i(x) and l_i'(x) on the nodes m/k, the reference's product formulas (k >= 2)
__device__ __forceinline__ double lagrange1d(int k, int i, double x) {
  double r = 1.0;
  for (int j = 0; j <= k; ++j) if (j != i) r = __dmul_rn(r, __ddiv_rn(__dsub_rn(__dmul_rn((double)k, x), (double)j), (double)(i - j)));
  return r;
}
__device__ __forceinline__ double dlagrange1d(int k, int i, double x) {
  double result = 0.0;
  for (int j = 0; j <= k; ++j) if (j != i) {
    double prod = __ddiv_rn(__dmul_rn((double)k, 1.0), (double)(i - j));
    for (int l = 0; l <= k; ++l) if (l != i && l != j) prod = __dmul_rn(prod, __ddiv_rn(__dsub_rn(__dmul_rn((double)k, x), (double)l), (double)(i - l)));
    result = __dadd_rn(result, prod);
  }
  return result;
}

template <bool VAL, bool JAC>
__global__ void __launch_bounds__(128)
k_eval_global(const __grid_constant__ DevBasis B, int firstLeaf, int nLeaves, const double* __restrict__ coeff,
              const double* __restrict__ x, u64 nPoints, double* __restrict__ values, double* __restrict__ jac,
              int* __restrict__ elementOut) {
  const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nPoints) return;
  const DevGrid& g = B.grid;
  int ec[3] = {0, 0, 0};
  double xl[3] = {0, 0, 0}, jinv[3] = {1, 1, 1};
  bool found = true;
  for (int j = 0; j < g.dim; ++j) {
    ec[j] = locateDirection(g, j, x[p * g.dim + j], xl[j], jinv[j]);
    found = found && ec[j] >= 0;
  }
  if (elementOut) elementOut[p] = found ? (int)((u64)ec[0] + (u64)g.N[0] * ((u64)ec[1] + (u64)g.N[1] * (u64)ec[2])) : -1;
  if (!found) {
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    for (int l = 0; l < nLeaves; ++l) {
      if (VAL) values[p * (u64)nLeaves + l] = nanv;
      if (JAC) for (int d = 0; d < g.dim; ++d) jac[(p * (u64)nLeaves + l) * g.dim + d] = nanv;
    }
    return;
  }
  double lv[3][kMaxContinuousOrder + 1], ld[3][kMaxContinuousOrder + 1];
  int kPrev = -1;
  for (int l = 0; l < nLeaves; ++l) {
    const DevLeaf& leaf = B.leaves[firstLeaf + l];
    const DevLayout& L = B.layouts[leaf.layout];
    const int k = L.k;
    if (k != kPrev) {
      for (int j = 0; j < g.dim; ++j)
        for (int a = 0; a <= k; ++a) {
          if (k == 0) { lv[j][a] = 1.0; ld[j][a] = 0.0; }
          else if (k == 1) { lv[j][a] = a ? xl[j] : __dsub_rn(1.0, xl[j]); ld[j][a] = a ? 1.0 : -1.0; }
          else { lv[j][a] = lagrange1d(k, a, xl[j]); if (JAC) ld[j][a] = dlagrange1d(k, a, xl[j]); }
        }
      kPrev = k;
    }
    double v = 0.0, rj[3] = {0, 0, 0};
    for (int i = 0; i < L.nloc; ++i) {
      int a[3];
      localAlpha(k, i, a);
      const u64 jdx = layoutIndexOriented(L, g, B.vid, ec, a, i);
      const double c = coeff[leaf.posA * jdx + leaf.posB];
      if (VAL) {
        double phi = 1.0;
        for (int j = 0; j < g.dim; ++j) phi = __dmul_rn(phi, lv[j][a[j]]);
        v = __dadd_rn(v, __dmul_rn(c, phi));
      }
      if (JAC)
        for (int d = 0; d < g.dim; ++d) {
          double dphi = k == 1 ? ld[d][a[d]] : 1.0;
          if (k == 1) { for (int j = 0; j < g.dim; ++j) if (j != d) dphi = __dmul_rn(dphi, lv[j][a[j]]); }
          else for (int j = 0; j < g.dim; ++j) dphi = __dmul_rn(dphi, j == d ? ld[j][a[j]] : lv[j][a[j]]);
          rj[d] = __dadd_rn(rj[d], __dmul_rn(c, dphi));
        }
    }
    if (VAL) values[p * (u64)nLeaves + l] = v;
    if (JAC) for (int d = 0; d < g.dim; ++d) jac[(p * (u64)nLeaves + l) * g.dim + d] = __dmul_rn(rj[d], jinv[d]);
  }
}

int evalGlobal(const dfx_basis* b, int firstLeaf, int nLeaves, const double* coeff, const double* x, u64 nPoints,
               double* values, double* jac, int* elementOut, cudaStream_t st) {
  if (nPoints == 0) return DFX_OK;
  const u64 blocks = (nPoints + 127) / 128;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "too many points for one launch");
  if (values && jac) k_eval_global<true, true><<<(unsigned)blocks, 128, 0, st>>>(b->dev, firstLeaf, nLeaves, coeff, x, nPoints, values, jac, elementOut);
  else if (jac) k_eval_global<false, true><<<(unsigned)blocks, 128, 0, st>>>(b->dev, firstLeaf, nLeaves, coeff, x, nPoints, values, jac, elementOut);
  else k_eval_global<true, false><<<(unsigned)blocks, 128, 0, st>>>(b->dev, firstLeaf, nLeaves, coeff, x, nPoints, values, jac, elementOut);
  return postLaunch("k_eval_global");
}

// ---------------------------------------------------------------------------
// Halo: the lowest (side 0) / highest (side 1) lattice plane in the last direction.
// Buffer order: leaves in tree order, per leaf the index-set components that lie in
// the plane (shift bit of the last direction clear) in layout order.
__device__ __forceinline__ bool haloLocate(const DevBasis& B, int side, u64 t, u64& p) {
  const int dz = B.grid.dim - 1;
  for (int l = 0; l < B.nLeaves; ++l) {
    const DevLeaf& leaf = B.leaves[l];
    const DevLayout& L = B.layouts[leaf.layout];
    if (L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0) continue;
    for (int q = 0; q < L.ncomp; ++q) {
      const int s = L.order[q];
      if ((s >> dz) & 1) continue;
      const u64 plane = L.compCount[s] / (u64)L.csize[s][dz];
      if (t < plane) {
        const u64 j = L.compOffset[s] + plane * (u64)(side ? L.csize[s][dz] - 1 : 0) + t;
        p = leaf.posA * j + leaf.posB;
        return true;
      }
      t -= plane;
    }
  }
  return false;
}

template <bool PACK>
__global__ void __launch_bounds__(256)
k_halo(const __grid_constant__ DevBasis B, int side, u64 total, double* __restrict__ coeff, double* __restrict__ buf) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  u64 p;
  if (!haloLocate(B, side, t, p)) return;
  if (PACK) buf[t] = coeff[p]; else coeff[p] = buf[t];
}

int haloCopy(const dfx_basis* b, int side, bool pack, double* coeff, double* buf, u64 total, cudaStream_t st) {
  if (total == 0) return DFX_OK;
  const u64 blocks = (total + 255) / 256;
  if (pack) k_halo<true><<<(unsigned)blocks, 256, 0, st>>>(b->dev, side, total, coeff, buf);
  else k_halo<false><<<(unsigned)blocks, 256, 0, st>>>(b->dev, side, total, coeff, buf);
  return postLaunch("k_halo");
}

}  // namespace dfx

// ---------------------------------------------------------------------------
// Peer-memory halo link (multi-GPU, one process per GPU): the sender packs its bottom lattice
// plane STRAIGHT INTO the lower neighbour's receive buffer over NVLink (P2P stores from the pack
// kernel) and then raises an arrival counter there; the receiver spins on its own counter,
// unpacks, and acknowledges by writing the consumed counter in the sender's block.  No
// rendezvous, no host involvement, two receive buffers so that a sender may run one step ahead.
// Block layout (doubles): [buf 0 | buf 1 | arrival | consumed | error | timeout cycles | push blocks done |
// pull blocks done] (u64 each).
// ONE kernel per direction: every block first waits (thread 0, bounded spin) for the flag it depends
// on, moves its 256 entries, fences, and counts itself done; the last block to finish raises the
// flag on the other side.  A wait that runs out of time records the step in the sticky error word
// and the kernel moves nothing — and so does every later push / pull until dfx_halo_link_reset_error:
// a stale plane is never unpacked and a buffer the neighbour has not consumed is never overwritten.
namespace dfx {

__device__ __forceinline__ unsigned long long* linkFlag(double* block, u64 n, int which) {
  return reinterpret_cast<unsigned long long*>(block + 2 * n) + which;
}
enum { kLinkArrival = 0, kLinkConsumed = 1, kLinkError = 2, kLinkTimeout = 3, kLinkPushDone = 4, kLinkPullDone = 5, kLinkWords = 6 };

// PUSH: coeff (bottom plane) -> peer buffer of parity step & 1, then peer.arrival = step.
//       Waits for local.consumed >= step - 2 (the acknowledgement of that buffer's previous use lands in MY block).
// PULL: local buffer -> coeff (top plane), then peer.consumed = step.  Waits for local.arrival >= step.
template <bool PUSH>
__global__ void __launch_bounds__(256)
k_halo_link(const __grid_constant__ DevBasis B, u64 total, double* __restrict__ coeff, double* local, double* peer,
            unsigned long long step) {
  __shared__ int go;
  if (threadIdx.x == 0) {
    volatile unsigned long long* err = linkFlag(local, total, kLinkError);
    int ok = *err == 0ull ? 1 : 0;
    const unsigned long long want = PUSH ? (step > 2 ? step - 2 : 0ull) : step;
    if (ok && want > 0) {
      volatile unsigned long long* flag = linkFlag(local, total, PUSH ? kLinkConsumed : kLinkArrival);
      const long long limit = (long long)*linkFlag(local, total, kLinkTimeout);
      const long long t0 = clock64();
      while (*flag < want) {
        __nanosleep(100);
        if (*err != 0ull) { ok = 0; break; }
        if (clock64() - t0 > limit) { atomicCAS(linkFlag(local, total, kLinkError), 0ull, step); ok = 0; break; }
      }
    }
    __threadfence_system();        // acquire: the plane was written before the flag
    go = ok;
  }
  __syncthreads();
  if (!go) return;                 // nothing moves, no flag is raised: the neighbour's wait reports it too
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 p;
  if (t < total && haloLocate(B, PUSH ? 0 : 1, t, p)) {
    if (PUSH) peer[(step & 1ull) * total + t] = coeff[p];
    else coeff[p] = __ldcg(local + (step & 1ull) * total + t);       // written by the peer: not through L1
  }
  __threadfence_system();          // release: this thread's stores before the block counts itself done
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* done = linkFlag(local, total, PUSH ? kLinkPushDone : kLinkPullDone);
    if (atomicAdd(done, 1ull) == (unsigned long long)gridDim.x - 1ull) {
      *done = 0ull;                // the next kernel of this direction starts after this one (same stream)
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(linkFlag(peer, total, PUSH ? kLinkArrival : kLinkConsumed)) = step;
      __threadfence_system();
    }
  }
}

int haloLinkBytes(u64 n, size_t& bytes) { bytes = (2 * n + kLinkWords + 2) * sizeof(double); return DFX_OK; }
size_t haloLinkWordOffset(u64 n, int which) { return 2 * n * sizeof(double) + (size_t)which * sizeof(unsigned long long); }

int haloPush(const dfx_basis* b, dfx_halo_link_impl* l, const double* coeff, u64 step, cudaStream_t st) {
  if (!l->lower) return fail(DFX_ERR_INVALID, "halo link: no lower neighbour attached");
  if (l->n == 0) return DFX_OK;
  const u64 blocks = (l->n + 255) / 256;
  k_halo_link<true><<<(unsigned)blocks, 256, 0, st>>>(b->dev, l->n, const_cast<double*>(coeff), l->local, l->lower, step);
  return postLaunch("k_halo_link");
}

int haloPull(const dfx_basis* b, dfx_halo_link_impl* l, double* coeff, u64 step, cudaStream_t st) {
  if (!l->upper) return fail(DFX_ERR_INVALID, "halo link: no upper neighbour attached");
  if (l->n == 0) return DFX_OK;
  const u64 blocks = (l->n + 255) / 256;
  k_halo_link<false><<<(unsigned)blocks, 256, 0, st>>>(b->dev, l->n, coeff, l->local, l->upper, step);
  return postLaunch("k_halo_link");
}

}  // namespace dfx
