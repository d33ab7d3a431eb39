// dfx_capi.cu — the extern "C" boundary declared in include/dfx.h.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <utility>
#include <vector>

#include "dfx_eval_fast.h"
#include "dfx_launch.h"
#include "dfx_hostbands.h"

using namespace dfx;

namespace {

// testing / profiling knobs of dfx_set_kernel_path (process-wide; atomics so that a test thread
// flipping one cannot tear a read — product code never writes them)
std::atomic<int> g_forcedPath{-1};      // evaluate
std::atomic<int> g_interpPath{-1};      // interpolate
std::atomic<int> g_indicesPath{-1};     // indices
std::atomic<int> g_hostBandKiB{-1};     // result bytes per band of dfx_evaluate_host (-1: 192 MiB)

int ensureDevice(const dfx_basis* b) {
  if (!b) return fail(DFX_ERR_INVALID, "null basis handle");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(DFX_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (b->device < 0 || b->device >= n) return fail(DFX_ERR_CUDA, "device ordinal out of range");
  DFX_CUDA(cudaSetDevice(b->device));
  if (b->vidPending) {
    // ids set from host memory (dfx_basis_set_vertex_ids_host works without a device, like basis creation)
    dfx_basis* mb = const_cast<dfx_basis*>(b);
    if (!mb->dVid) DFX_CUDA(cudaMalloc(&mb->dVid, b->vidHost.size() * sizeof(u64)));
    DFX_CUDA(cudaMemcpy(mb->dVid, b->vidHost.data(), b->vidHost.size() * sizeof(u64), cudaMemcpyHostToDevice));
    mb->dev.vid = (const u64*)mb->dVid;
    mb->dev.orient = nullptr; mb->orientValid = false;
    mb->vidPending = false;
    if (mb->dDev) DFX_CUDA(cudaMemcpy(mb->dDev, &mb->dev, sizeof(DevBasis), cudaMemcpyHostToDevice));
  }
  if (!b->dLocalTab && !b->localTab.empty()) {
    dfx_basis* mb = const_cast<dfx_basis*>(b);
    DFX_CUDA(cudaMalloc(&mb->dLocalTab, b->localTab.size() * sizeof(unsigned)));
    DFX_CUDA(cudaMemcpy(mb->dLocalTab, b->localTab.data(), b->localTab.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    DFX_CUDA(cudaMalloc(&mb->dIdxTab, b->idxTab.size() * sizeof(IdxLi)));
    DFX_CUDA(cudaMemcpy(mb->dIdxTab, b->idxTab.data(), b->idxTab.size() * sizeof(IdxLi), cudaMemcpyHostToDevice));
    DFX_CUDA(cudaMalloc(&mb->dDev, sizeof(DevBasis)));
    DFX_CUDA(cudaMemcpy(mb->dDev, &b->dev, sizeof(DevBasis), cudaMemcpyHostToDevice));
  }
  return DFX_OK;
}

bool fits32(const dfx_basis* b, uint64_t e0, uint64_t e1) {
  const u64 ne = (u64)b->dev.grid.N[0] * b->dev.grid.N[1] * b->dev.grid.N[2];
  return ne <= 0xffffffffull && (e1 - e0) * (u64)b->dev.nlocTotal <= 0xffffffffull;
}

int subtree(const dfx_basis* b, int node, int& firstLeaf, int& nLeaves) {
  if (node < 0 || node >= (int)b->nodes.size()) return fail(DFX_ERR_INVALID, "tree node id out of range");
  firstLeaf = b->nodes[node].firstLeaf;
  nLeaves = b->nodes[node].nLeaves;
  return DFX_OK;
}

int growScratch(dfx_basis* b, int slot, size_t bytes) {
  if (b->dScratchBytes[slot] >= bytes) return DFX_OK;
  if (b->dScratch[slot]) cudaFree(b->dScratch[slot]);
  b->dScratch[slot] = nullptr; b->dScratchBytes[slot] = 0;
  DFX_CUDA(cudaMalloc(&b->dScratch[slot], bytes));
  b->dScratchBytes[slot] = bytes;
  return DFX_OK;
}

int privateStream(dfx_basis* b, cudaStream_t& st) {
  if (!b->stream) {
    cudaStream_t s;
    DFX_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    b->stream = (void*)s;
  }
  st = (cudaStream_t)b->stream;
  return DFX_OK;
}

// Host shape tables for the points: per layout used below the subtree,
//   shape[q][i]          = prod_j l_{alpha_j}(x_j)          (evaluateFunction)
//   dshape[q][i][d]      = prod_j (j==d ? l' : l)(x_j)       (evaluateJacobian)
// with the k = 0 / k = 1 special cases of LagrangeCubeLocalBasis.
int prepareShape(dfx_basis* b, int node, int firstLeaf, int nLeaves, const double* xhat, int nq, ShapeOffsets& so,
                 const double*& dShape, const double*& dDShape, cudaStream_t st) {
  const DevBasis& D = b->dev;
  const int dim = D.grid.dim;
  bool used[kMaxLayouts] = {false, false, false, false};
  for (int l = 0; l < nLeaves; ++l) used[D.leaves[firstLeaf + l].layout] = true;
  size_t tot = 0;
  for (int q = 0; q < D.nLayouts; ++q) {
    so.v[q] = so.d[q] = 0;
    if (!used[q]) continue;
    so.v[q] = (unsigned)tot; tot += (size_t)nq * D.layouts[q].nloc;
  }
  const size_t valTot = tot;
  for (int q = 0; q < D.nLayouts; ++q) {
    if (!used[q]) continue;
    so.d[q] = (unsigned)(tot - valTot); tot += (size_t)nq * D.layouts[q].nloc * dim;
  }
  const bool cached = b->lastShapeNode == node && (int)b->lastXhat.size() == nq * dim &&
                      std::memcmp(b->lastXhat.data(), xhat, sizeof(double) * nq * dim) == 0 && b->dShape;
  if (!cached) {
    std::vector<double> h(tot);
    for (int q = 0; q < D.nLayouts; ++q) {
      if (!used[q]) continue;
      const DevLayout& L = D.layouts[q];
      const int k = L.k;
      for (int p = 0; p < nq; ++p) {
        const double* x = xhat + (size_t)p * dim;
        for (int i = 0; i < L.nloc; ++i) {
          int a[3] = {0, 0, 0};
          { int ii = i; for (int j = 0; j < dim; ++j) { a[j] = ii % (k + 1); ii /= (k + 1); } }
          double* v = &h[so.v[q] + (size_t)p * L.nloc + i];
          double* dv = &h[valTot + so.d[q] + ((size_t)p * L.nloc + i) * dim];
          if (k == 0) { *v = 1; for (int d = 0; d < dim; ++d) dv[d] = 0; continue; }
          if (k == 1) {
            *v = 1;
            for (int j = 0; j < dim; ++j) *v *= (i & (1 << j)) ? x[j] : 1 - x[j];
            for (int d = 0; d < dim; ++d) {
              dv[d] = (i & (1 << d)) ? 1 : -1;
              for (int j = 0; j < dim; ++j) if (j != d) dv[d] *= (i & (1 << j)) ? x[j] : 1 - x[j];
            }
            continue;
          }
          *v = 1.0;
          for (int j = 0; j < dim; ++j) *v *= lagrangeP(k, a[j], x[j]);
          for (int d = 0; d < dim; ++d) {
            dv[d] = 1.0;
            for (int j = 0; j < dim; ++j) dv[d] *= (j == d) ? lagrangeDP(k, a[j], x[j]) : lagrangeP(k, a[j], x[j]);
          }
        }
      }
    }
    if (b->dShapeBytes < tot * sizeof(double)) {
      if (b->dShape) cudaFree(b->dShape);
      b->dShape = nullptr; b->dShapeBytes = 0;
      DFX_CUDA(cudaMalloc(&b->dShape, tot * sizeof(double)));
      b->dShapeBytes = tot * sizeof(double);
    }
    DFX_CUDA(cudaMemcpyAsync(b->dShape, h.data(), tot * sizeof(double), cudaMemcpyHostToDevice, st));
    DFX_CUDA(cudaStreamSynchronize(st));     // h goes out of scope; tables are tiny
    b->lastXhat.assign(xhat, xhat + (size_t)nq * dim);
    b->lastShapeNode = node;
  }
  dShape = (const double*)b->dShape;
  dDShape = (const double*)b->dShape + valTot;
  return DFX_OK;
}

u64 haloCount(const dfx_basis* b) {
  const DevBasis& D = b->dev;
  const int dz = D.grid.dim - 1;
  u64 n = 0;
  for (int l = 0; l < D.nLeaves; ++l) {
    const DevLayout& L = D.layouts[D.leaves[l].layout];
    if (L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0) continue;
    for (int q = 0; q < L.ncomp; ++q) {
      const int s = L.order[q];
      if ((s >> dz) & 1) continue;
      n += L.compCount[s] / (u64)L.csize[s][dz];
    }
  }
  return n;
}

}  // namespace

extern "C" {

int dfx_abi_version(void) { return DFX_ABI_VERSION; }
const char* dfx_last_error(void) { return lastErrorCStr(); }
uint64_t dfx_kernel_launch_count(void) { return g_launchCount.load(); }

int dfx_basis_create(const dfx_grid_desc* grid_host, const dfx_tree_node* tree_host, int32_t n_nodes, int32_t device,
                     dfx_basis** out) {
  dfx_basis_impl* impl = nullptr;
  int rc = buildBasis(grid_host, tree_host, n_nodes, device, &impl);
  if (rc) return rc;
  *out = static_cast<dfx_basis*>(impl);
  return DFX_OK;
}

// everything the handle owns on the device / in pinned memory; the handle can be rebuilt afterwards
static void releaseResources(dfx_basis_impl* b) {
  if (b->companion) { dfx_basis_destroy(static_cast<dfx_basis*>(b->companion)); b->companion = nullptr; }
  b->companionTried = false;
  const bool any = b->dLocalCoeff || b->dDense || b->dJinv || b->dNodal || b->dVid || b->dOrient || b->evStream[0] || b->evStream[1] || b->dLocalTab || b->dDev ||
                   b->dIdxTab || b->dSep || b->dShape || b->pinned || b->dScratch[0] || b->dScratch[1] || b->dScratch[2] || b->stream;
  if (!any) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize((cudaStream_t)b->stream);
  if (b->dLocalCoeff) { cudaFree(b->dLocalCoeff); b->dLocalCoeff = nullptr; b->dLocalCoeffBytes = 0; }
  for (void** p : {&b->dDense, &b->dJinv, &b->dNodal, &b->dVid, &b->dOrient, &b->dShape, &b->dLocalTab, &b->dIdxTab, &b->dDev, &b->dSep,
                   &b->dScratch[0], &b->dScratch[1], &b->dScratch[2]})
    if (*p) { cudaFree(*p); *p = nullptr; }
  if (b->pinned) { cudaFreeHost(b->pinned); b->pinned = nullptr; }
  for (void** p : {&b->evStream[0], &b->evStream[1], &b->stream})
    if (*p) { cudaStreamDestroy((cudaStream_t)*p); *p = nullptr; }
}

void dfx_basis_destroy(dfx_basis* b) {
  if (!b) return;
  releaseResources(b);
  delete b;
}

// DefaultGlobalBasis::update(gridView), defaultglobalbasis.hh:137-141: preBasis_.update(gv) +
// initializeIndices() — the same tree on another grid (box).  The handle stays valid; index tables,
// coefficient vectors and vertex ids that belonged to the old grid do not.
int dfx_basis_update(dfx_basis* b, const dfx_grid_desc* grid_host) {
  if (!b || !grid_host) return fail(DFX_ERR_INVALID, "null argument");
  dfx_basis_impl* fresh = nullptr;
  const std::vector<dfx_tree_node> tree = b->treeDesc;
  int rc = buildBasis(grid_host, tree.data(), (int)tree.size(), b->device, &fresh);
  if (rc) return rc;                                     // the old state is untouched on failure
  releaseResources(b);
  *static_cast<dfx_basis_impl*>(b) = std::move(*fresh);
  delete static_cast<dfx_basis*>(fresh);
  return DFX_OK;
}

// ---- vertex ids for the face-DOF permutation (lagrangebasis.hh:787, :616-634) ---------------------
static u64 numVertices(const dfx_basis* b) {
  u64 n = 1;
  for (int j = 0; j < b->dev.grid.dim; ++j) n *= (u64)(b->dev.grid.N[j] + 1);
  return n;
}
static bool wantsOrientation(const dfx_basis* b) {
  if (b->dev.grid.dim < 2) return false;
  for (int l = 0; l < b->dev.nLeaves; ++l) {
    const DevLayout& L = b->dev.layouts[b->dev.leaves[l].layout];
    if (L.kind == DFX_NODE_LAGRANGE && L.k >= 3) return true;
  }
  return false;
}
static int setVertexIds(dfx_basis* b, const uint64_t* ids, uint64_t n, bool fromHost, cudaStream_t st) {
  if (!b) return fail(DFX_ERR_INVALID, "null basis handle");
  if (ids && n != numVertices(b)) return fail(DFX_ERR_RANGE, "one id per vertex of the local box expected");
  if (!ids) {
    b->dev.vid = nullptr;                  // the table itself stays allocated until destroy: kernels in flight may read it
    b->dev.orient = nullptr; b->orientValid = false;
    b->oriented = false;
    b->vidPending = false;
    b->vidHost.clear();
    if (b->dDev) {
      int rc = ensureDevice(b); if (rc) return rc;
      DFX_CUDA(cudaMemcpyAsync(b->dDev, &b->dev, sizeof(DevBasis), cudaMemcpyHostToDevice, st));
    }
    return DFX_OK;
  }
  if (fromHost) {
    // host bookkeeping only; the upload happens in ensureDevice() at the next device call
    b->vidHost.assign(ids, ids + n);
    b->vidPending = true;
    b->dev.orient = nullptr; b->orientValid = false;
    b->oriented = wantsOrientation(b);
    return DFX_OK;
  }
  int rc = ensureDevice(b); if (rc) return rc;
  if (!b->dVid) DFX_CUDA(cudaMalloc(&b->dVid, n * sizeof(u64)));
  DFX_CUDA(cudaMemcpyAsync(b->dVid, ids, n * sizeof(u64), cudaMemcpyDeviceToDevice, st));
  b->vidPending = false;
  b->vidHost.clear();
  b->dev.vid = (const u64*)b->dVid;
  b->dev.orient = nullptr; b->orientValid = false;
  b->oriented = wantsOrientation(b);
  if (b->dDev) DFX_CUDA(cudaMemcpyAsync(b->dDev, &b->dev, sizeof(DevBasis), cudaMemcpyHostToDevice, st));
  return DFX_OK;
}
int dfx_basis_set_vertex_ids(dfx_basis* b, const uint64_t* ids, uint64_t n, void* stream) {
  return setVertexIds(b, ids, n, false, (cudaStream_t)stream);
}
int dfx_basis_set_vertex_ids_host(dfx_basis* b, const uint64_t* ids_host, uint64_t n) {
  return setVertexIds(b, ids_host, n, true, nullptr);
}
int32_t dfx_basis_has_permuted_face_dofs(const dfx_basis* b) { return b && b->oriented ? 1 : 0; }

uint64_t dfx_basis_dimension(const dfx_basis* b) { return b ? b->dimension : 0; }
uint64_t dfx_basis_size(const dfx_basis* b, const uint64_t* prefix_host, int32_t len) {
  if (!b || len < 0 || (len > 0 && !prefix_host)) return 0;
  return preSize(b, (const u64*)prefix_host, len);
}
int dfx_basis_container_descriptor(const dfx_basis* b, dfx_container_node* out_host, int32_t capacity, int32_t* n_nodes,
                                   char* type_out, int32_t type_capacity) {
  if (!b || !n_nodes || capacity < 0 || (capacity > 0 && !out_host)) return fail(DFX_ERR_INVALID, "null argument");
  std::vector<dfx_container_node> nodes;
  std::string type;
  containerDescriptor(b, nodes, type);
  *n_nodes = (int32_t)nodes.size();
  for (int32_t i = 0; i < capacity && i < (int32_t)nodes.size(); ++i) out_host[i] = nodes[i];
  if (type_out && type_capacity > 0) {
    const size_t n = std::min((size_t)type_capacity - 1, type.size());
    std::memcpy(type_out, type.data(), n);
    type_out[n] = 0;
  }
  return DFX_OK;
}
int32_t dfx_basis_max_node_size(const dfx_basis* b) { return b ? b->dev.nlocTotal : 0; }
int32_t dfx_basis_max_multi_index_size(const dfx_basis* b) { return b ? b->maxDigits : 0; }
int32_t dfx_basis_min_multi_index_size(const dfx_basis* b) { return b ? b->minDigits : 0; }
uint64_t dfx_basis_num_elements(const dfx_basis* b) {
  return b ? (u64)b->dev.grid.N[0] * b->dev.grid.N[1] * b->dev.grid.N[2] : 0;
}
int32_t dfx_basis_num_leaves(const dfx_basis* b) { return b ? b->dev.nLeaves : 0; }
int32_t dfx_basis_num_tree_nodes(const dfx_basis* b) { return b ? (int32_t)b->nodes.size() : 0; }

int dfx_basis_node_info(const dfx_basis* b, int32_t node, int32_t* local_offset, int32_t* size, int32_t* first_leaf,
                        int32_t* n_leaves) {
  if (!b || node < 0 || node >= (int)b->nodes.size()) return fail(DFX_ERR_INVALID, "tree node id out of range");
  const HostNode& n = b->nodes[node];
  if (local_offset) *local_offset = n.localOffset;
  if (size) *size = n.size;
  if (first_leaf) *first_leaf = n.firstLeaf;
  if (n_leaves) *n_leaves = n.nLeaves;
  return DFX_OK;
}

int32_t dfx_basis_node_child(const dfx_basis* b, int32_t node, int32_t child) {
  if (!b || node < 0 || node >= (int)b->nodes.size()) return -1;
  const HostNode& n = b->nodes[node];
  if (child < 0 || child >= (int)n.children.size()) return -1;
  return n.children[child];
}

int dfx_basis_local_index_sizes(const dfx_basis* b, uint8_t* sizes_host) {
  if (!b || !sizes_host) return fail(DFX_ERR_INVALID, "null argument");
  for (int l = 0; l < b->dev.nLeaves; ++l)
    for (int i = 0; i < b->dev.leaves[l].nloc; ++i)
      sizes_host[b->dev.leaves[l].localOffset + i] = (uint8_t)b->dev.leaves[l].ndig;
  return DFX_OK;
}

// ---- indices ------------------------------------------------------------------
static int checkRange(const dfx_basis* b, uint64_t e0, uint64_t e1) {
  if (e0 > e1 || e1 > dfx_basis_num_elements(b)) return fail(DFX_ERR_RANGE, "element range outside the local box");
  return DFX_OK;
}

int dfx_indices_multi(const dfx_basis* b, uint64_t e0, uint64_t e1, uint32_t* out, int32_t stride, void* stream) {
  DFX_RANGE("dfx_indices_multi");
  int rc = ensureDevice(b); if (rc) return rc;
  if ((rc = checkRange(b, e0, e1))) return rc;
  if (stride < b->maxDigits) return fail(DFX_ERR_INVALID, "stride smaller than the longest multi-index");
  if (b->dimension > 0xffffffffull) return fail(DFX_ERR_RANGE, "indices exceed 32 bits, use dfx_indices_multi_u64");
  if (g_indicesPath != 0 && fits32(b, e0, e1)) return indicesFast<uint32_t, true>(b, e0, e1, out, stride, (cudaStream_t)stream);
  return indicesMulti32(b, e0, e1, out, stride, (cudaStream_t)stream);
}
int dfx_indices_multi_u64(const dfx_basis* b, uint64_t e0, uint64_t e1, uint64_t* out, int32_t stride, void* stream) {
  DFX_RANGE("dfx_indices_multi_u64");
  int rc = ensureDevice(b); if (rc) return rc;
  if ((rc = checkRange(b, e0, e1))) return rc;
  if (stride < b->maxDigits) return fail(DFX_ERR_INVALID, "stride smaller than the longest multi-index");
  if (g_indicesPath != 0 && fits32(b, e0, e1)) return indicesFast<u64, true>(b, e0, e1, (u64*)out, stride, (cudaStream_t)stream);
  return indicesMulti64(b, e0, e1, (u64*)out, stride, (cudaStream_t)stream);
}
int dfx_indices_flat(const dfx_basis* b, uint64_t e0, uint64_t e1, uint32_t* out, void* stream) {
  DFX_RANGE("dfx_indices_flat");
  int rc = ensureDevice(b); if (rc) return rc;
  if ((rc = checkRange(b, e0, e1))) return rc;
  if (b->dimension > 0xffffffffull) return fail(DFX_ERR_RANGE, "indices exceed 32 bits, use dfx_indices_flat_u64");
  if (g_indicesPath != 0 && fits32(b, e0, e1)) return indicesFast<uint32_t, false>(b, e0, e1, out, 1, (cudaStream_t)stream);
  return indicesFlat32(b, e0, e1, out, (cudaStream_t)stream);
}
int dfx_indices_flat_u64(const dfx_basis* b, uint64_t e0, uint64_t e1, uint64_t* out, void* stream) {
  DFX_RANGE("dfx_indices_flat_u64");
  int rc = ensureDevice(b); if (rc) return rc;
  if ((rc = checkRange(b, e0, e1))) return rc;
  if (g_indicesPath != 0 && fits32(b, e0, e1)) return indicesFast<u64, false>(b, e0, e1, (u64*)out, 1, (cudaStream_t)stream);
  return indicesFlat64(b, e0, e1, (u64*)out, (cudaStream_t)stream);
}

// ---- interpolate -----------------------------------------------------------------
static int checkFunction(const dfx_function* f, int nLeaves) {
  if (!f) return fail(DFX_ERR_INVALID, "null function");
  if (f->id < DFX_FN_CONSTANT || f->id > DFX_FN_GAUSSIAN_VEC) return fail(DFX_ERR_INVALID, "unknown function id");
  if (f->n_comp < 1 || f->n_comp > DFX_MAX_LEAVES) return fail(DFX_ERR_INVALID, "bad number of range entries");
  // HierarchicNodeToRangeMap: a scalar is used for every leaf, otherwise leaf l takes entry l
  if (f->n_comp != 1 && f->n_comp != nLeaves)
    return fail(DFX_ERR_RANGE, "range of f does not match the leaves of the (sub)tree");
  return DFX_OK;
}

// separable: the caller opted into per-direction factor tables (dfx_interpolate_separable)
static int interpolateImpl(const dfx_basis* b, int32_t node, const dfx_function* f, double* coeff, const uint8_t* mask,
                           bool separable, void* stream) {
  int rc = ensureDevice(b); if (rc) return rc;
  int fl, nl;
  if ((rc = subtree(b, node, fl, nl))) return rc;
  if ((rc = checkFunction(f, nl))) return rc;
  if (!coeff) return fail(DFX_ERR_INVALID, "null coefficient vector");
  const int path = g_interpPath.load();      // -1 automatic, 0 generic, 1 structured per-node, 2 structured with tables
  if (path != 0 && !b->oriented && b->dimension <= 0xffffffffull && dfx_basis_num_elements(b) <= 0xffffffffull) {
    const bool wantSep = path == 2 || (path == -1 && separable);
    return interpolateFast(const_cast<dfx_basis*>(b), fl, nl, *f, coeff, mask, wantSep, (cudaStream_t)stream);
  }
  if (path != 0 && b->oriented && nl == 1 && mask == nullptr) {
    // permuted face DOFs, one continuous leaf of order 3 .. 5: the run sweep with one orientation byte per entity
    rc = interpolateRunsOriented(const_cast<dfx_basis*>(b), fl, *f, coeff, (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  return interpolateRegistry(b, fl, nl, *f, coeff, mask, (cudaStream_t)stream);
}

int dfx_interpolate(const dfx_basis* b, int32_t node, const dfx_function* f, double* coeff, const uint8_t* mask,
                    void* stream) {
  DFX_RANGE("dfx_interpolate");
  return interpolateImpl(b, node, f, coeff, mask, false, stream);
}

// several interpolate calls on one vector; two lean lattice sweeps share one launch
int dfx_interpolate_batch(const dfx_basis* b, int32_t n_parts, const int32_t* nodes_host, const dfx_function* fns_host, double* coeff,
                          const uint8_t* mask, void* stream) {
  DFX_RANGE("dfx_interpolate_batch");
  int rc = ensureDevice(b); if (rc) return rc;
  if (n_parts < 0 || (n_parts > 0 && (!nodes_host || !fns_host))) return fail(DFX_ERR_INVALID, "null argument");
  if (!coeff) return fail(DFX_ERR_INVALID, "null coefficient vector");
  // validate everything before anything is written
  std::vector<std::pair<int, int>> leaves((size_t)n_parts);
  for (int p = 0; p < n_parts; ++p) {
    if ((rc = subtree(b, nodes_host[p], leaves[p].first, leaves[p].second))) return rc;
    if ((rc = checkFunction(&fns_host[p], leaves[p].second))) return rc;
  }
  int p = 0;
  while (p < n_parts) {
    if (p + 1 < n_parts && mask == nullptr && g_interpPath.load() == -1) {
      // the pair must not overlap (then the order of the two calls would matter): disjoint leaf ranges
      const bool disjoint = leaves[p].first + leaves[p].second <= leaves[p + 1].first ||
                            leaves[p + 1].first + leaves[p + 1].second <= leaves[p].first;
      if (disjoint) {
        rc = interpolateBatch2(const_cast<dfx_basis*>(b), leaves[p].first, leaves[p].second, fns_host[p], leaves[p + 1].first,
                               leaves[p + 1].second, fns_host[p + 1], coeff, (cudaStream_t)stream);
        if (rc == DFX_OK) { p += 2; continue; }
        if (rc > 0) return rc;
      }
    }
    if ((rc = interpolateImpl(b, nodes_host[p], &fns_host[p], coeff, mask, false, stream))) return rc;
    ++p;
  }
  return DFX_OK;
}

int dfx_interpolate_separable(const dfx_basis* b, int32_t node, const dfx_function* f, double* coeff, const uint8_t* mask,
                              void* stream) {
  DFX_RANGE("dfx_interpolate_separable");
  return interpolateImpl(b, node, f, coeff, mask, true, stream);
}

int dfx_interpolation_points(const dfx_basis* b, double* out_xyz, int32_t* out_leaf, void* stream) {
  int rc = ensureDevice(b); if (rc) return rc;
  if (!out_xyz || !out_leaf) return fail(DFX_ERR_INVALID, "null output");
  return interpolationPoints(b, out_xyz, out_leaf, (cudaStream_t)stream);
}

int dfx_interpolate_from_values(const dfx_basis* b, int32_t node, const double* values, double* coeff,
                                const uint8_t* mask, void* stream) {
  DFX_RANGE("dfx_interpolate_from_values");
  int rc = ensureDevice(b); if (rc) return rc;
  int fl, nl;
  if ((rc = subtree(b, node, fl, nl))) return rc;
  if (!values || !coeff) return fail(DFX_ERR_INVALID, "null argument");
  return interpolateFromValues(b, fl, nl, values, coeff, mask, (cudaStream_t)stream);
}

// ---- interpolate a DiscreteGlobalBasisFunction into another basis ---------------------------
int dfx_evaluate(const dfx_basis* cb, int32_t node, const double* coeff, const double* xhat_host, int32_t n_q,
                 uint64_t e0, uint64_t e1, double* out_values, double* out_jac, void* stream);

static bool sameGrid(const dfx_basis* a, const dfx_basis* b) {
  const dfx_grid_desc &x = a->gridDesc, &y = b->gridDesc;
  if (x.dim != y.dim) return false;
  for (int j = 0; j < x.dim; ++j)
    if (x.n_global[j] != y.n_global[j] || x.local_begin[j] != y.local_begin[j] || x.local_n[j] != y.local_n[j] ||
        x.lower[j] != y.lower[j] || x.upper[j] != y.upper[j]) return false;
  return true;
}

int dfx_interpolate_dgbf(const dfx_basis* target, int32_t target_node, const dfx_basis* source, int32_t source_node,
                         const double* source_coeff, double* target_coeff, const uint8_t* mask, double* scratch,
                         uint64_t scratch_doubles, void* stream) {
  DFX_RANGE("dfx_interpolate_dgbf");
  int rc = ensureDevice(target); if (rc) return rc;
  if (!source || !source_coeff || !target_coeff) return fail(DFX_ERR_INVALID, "null argument");
  if (source->device != target->device) return fail(DFX_ERR_INVALID, "source and target basis live on different devices");
  if ((rc = ensureDevice(source))) return rc;
  if (!sameGrid(target, source)) return fail(DFX_ERR_RANGE, "source and target basis are not defined on the same grid box");
  int flT, nlT, flS, nlS;
  if ((rc = subtree(target, target_node, flT, nlT))) return rc;
  if ((rc = subtree(source, source_node, flS, nlS))) return rc;
  if (nlS != 1 && nlS != nlT) return fail(DFX_ERR_RANGE, "range of the grid function does not match the leaves of the (sub)tree");
  cudaStream_t st = (cudaStream_t)stream;
  const DevBasis& T = target->dev;
  const int dim = T.grid.dim;
  const u64 ne = dfx_basis_num_elements(target);
  u64 layer = 1;
  for (int j = 0; j + 1 < dim; ++j) layer *= (u64)T.grid.N[j];
  int maxNloc = 1;
  for (int l = 0; l < nlT; ++l) maxNloc = std::max(maxNloc, T.layouts[T.leaves[flT + l].layout].nloc);
  const u64 perElem = (u64)maxNloc * (u64)nlS;
  double* buf = scratch;
  u64 cap = scratch_doubles;
  if (!buf) {
    // no caller scratch: a band buffer owned by the target handle (<= 1 GiB, kept across calls)
    dfx_basis* t = const_cast<dfx_basis*>(target);
    cap = std::min<u64>(ne * perElem, std::max<u64>(layer * perElem, (1ull << 30) / sizeof(double)));
    if (t->dNodalBytes < cap * sizeof(double)) {
      if (t->dNodal) cudaFree(t->dNodal);
      t->dNodal = nullptr; t->dNodalBytes = 0;
      DFX_CUDA(cudaMalloc(&t->dNodal, cap * sizeof(double)));
      t->dNodalBytes = cap * sizeof(double);
    }
    buf = (double*)t->dNodal;
  }
  if (cap < layer * perElem) return fail(DFX_ERR_RANGE, "scratch smaller than one z-layer of nodal values");
  const u64 band = std::min<u64>(ne, (cap / perElem) / layer * layer);
  rc = DFX_OK;
  for (int q = 0; q < T.nLayouts && !rc; ++q) {
    bool used = false;
    for (int l = 0; l < nlT; ++l) used = used || T.leaves[flT + l].layout == q;
    if (!used) continue;
    const DevLayout& L = T.layouts[q];
    // local interpolation nodes of this layout: alpha / k (element centre for k = 0), local index order
    std::vector<double> xhat((size_t)L.nloc * dim);
    for (int i = 0; i < L.nloc; ++i) {
      int ii = i;
      for (int j = 0; j < dim; ++j) {
        const int a = ii % (L.k + 1); ii /= (L.k + 1);
        xhat[(size_t)i * dim + j] = L.k == 0 ? 0.5 : (1.0 * a) / L.k;
      }
    }
    for (u64 e0 = 0; e0 < ne && !rc; e0 += band) {
      const u64 e1 = std::min<u64>(ne, e0 + band);
      rc = dfx_evaluate(source, source_node, source_coeff, xhat.data(), L.nloc, e0, e1, buf, nullptr, st);
      if (!rc) rc = scatterNodes(target, q, flT, nlT, nlS, e0, e1, buf, target_coeff, mask, st);
    }
  }
  return rc;
}

int dfx_interpolate_dgbf_host(const dfx_basis* target, int32_t target_node, const dfx_basis* source, int32_t source_node,
                              const double* source_coeff_host, double* target_coeff_host, const uint8_t* mask_host) {
  dfx_basis* t = const_cast<dfx_basis*>(target);
  int rc = ensureDevice(t); if (rc) return rc;
  if (!source || !source_coeff_host || !target_coeff_host) return fail(DFX_ERR_INVALID, "null argument");
  cudaStream_t st; if ((rc = privateStream(t, st))) return rc;
  const size_t nT = t->dimension, nS = source->dimension;
  if ((rc = growScratch(t, 0, (nT + nS) * sizeof(double)))) return rc;
  double* dT = (double*)t->dScratch[0];
  double* dS = dT + nT;
  uint8_t* dM = nullptr;
  DFX_CUDA(cudaMemcpyAsync(dT, target_coeff_host, nT * sizeof(double), cudaMemcpyHostToDevice, st));   // untouched entries keep their value
  DFX_CUDA(cudaMemcpyAsync(dS, source_coeff_host, nS * sizeof(double), cudaMemcpyHostToDevice, st));
  if (mask_host) {
    if ((rc = growScratch(t, 1, nT))) return rc;
    dM = (uint8_t*)t->dScratch[1];
    DFX_CUDA(cudaMemcpyAsync(dM, mask_host, nT, cudaMemcpyHostToDevice, st));
  }
  if ((rc = dfx_interpolate_dgbf(target, target_node, source, source_node, dS, dT, dM, nullptr, 0, st))) return rc;
  DFX_CUDA(cudaMemcpyAsync(target_coeff_host, dT, nT * sizeof(double), cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

// ---- evaluate ---------------------------------------------------------------------
void dfx_set_kernel_path(int32_t op, int32_t path) {
  if (op == DFX_OP_EVALUATE) g_forcedPath = path;
  else if (op == DFX_OP_INTERPOLATE) g_interpPath = path;
  else if (op == DFX_OP_INDICES) { g_indicesPath = path == 2 ? 1 : path; g_indicesTile = path == 2 ? 0 : 1; }
  else if (op == DFX_OP_HOST_BAND_KIB) g_hostBandKiB = path;
  else if (op == DFX_OP_SMEM_GRADIENT) g_smemGrad = path;
  else if (op == DFX_OP_TUNE_ROWS) g_rowsLean = path > 0 ? path : -1;
  else if (op == DFX_OP_TUNE_EVAL_BLOCK) g_evalBlock = path == 128 ? 128 : (path == 64 ? 64 : 32);
  else if (op == DFX_OP_TUNE_SMEM_TMA) g_smemTma = path == 0 ? 0 : (path == 2 ? 2 : 1);
  else if (op == DFX_OP_TUNE_SMEM_PAIR) g_smemPair = (path == 0 || path == 1 || path == 2 || path == 4) ? path : -1;
  else if (op == DFX_OP_TUNE_SMEM_BLOCKS) g_smemMaxBlocks = path > 0 ? path : 0;
  else if (op == DFX_OP_TUNE_EVAL_Q3) g_evalQ3 = path == 0 ? 0 : 1;
  else if (op == DFX_OP_TUNE_SPLIT_MAP) g_evalSplitByGroup = path == 0 ? 0 : 1;
  else if (op == DFX_OP_TUNE_PAIR_PREFETCH) g_pairL2Ahead = path < 0 ? 0 : path;
  else if (op == DFX_OP_TUNE_STREAM_HINT) g_evalStreamHint = path == 0 ? 0 : 1;
  else if (op == DFX_OP_TUNE_SWIZZLE) g_evalSwizzle = path == 0 ? 0 : (path == 2 ? 2 : 1);
}

// Companion of a basis with permuted face DOFs: the same tree with every Lagrange leaf replaced by the
// LagrangeDG leaf of the same order (TaylorHood written out as composite(power<dim>(Q2, FI), Q1, BL),
// taylorhoodbasis.hh:229-251), plus a coefficient buffer in its layout.  false = not available.
static bool orientedCompanion(dfx_basis* b) {
  if (b->companion && b->dLocalCoeff) return true;
  if (b->companionTried) return false;
  b->companionTried = true;
  std::vector<dfx_tree_node> tree;
  for (const dfx_tree_node& n : b->treeDesc) {
    if (n.kind == DFX_NODE_TAYLOR_HOOD) {
      tree.push_back({DFX_NODE_COMPOSITE, 0, 2, DFX_IMS_BLOCKED_LEXICOGRAPHIC});
      tree.push_back({DFX_NODE_POWER, 0, b->dev.grid.dim, DFX_IMS_FLAT_INTERLEAVED});
      tree.push_back({DFX_NODE_LAGRANGE_DG, 2, 0, DFX_IMS_NONE});
      tree.push_back({DFX_NODE_LAGRANGE_DG, 1, 0, DFX_IMS_NONE});
      continue;
    }
    dfx_tree_node m = n;
    if (m.kind == DFX_NODE_LAGRANGE) {
      if (m.order > kMaxDGOrder) return false;
      m.kind = DFX_NODE_LAGRANGE_DG;
    }
    tree.push_back(m);
  }
  dfx_basis_impl* c = nullptr;
  if (buildBasis(&b->gridDesc, tree.data(), (int)tree.size(), b->device, &c) != DFX_OK) return false;
  // same local ansatz tree, leaf by leaf
  bool same = c->dev.nLeaves == b->dev.nLeaves && c->dev.nlocTotal == b->dev.nlocTotal && c->nodes.size() == b->nodes.size();
  for (int l = 0; same && l < b->dev.nLeaves; ++l)
    same = c->dev.leaves[l].nloc == b->dev.leaves[l].nloc && c->dev.leaves[l].localOffset == b->dev.leaves[l].localOffset;
  void* buf = nullptr;
  if (!same || cudaMalloc(&buf, c->dimension * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    delete static_cast<dfx_basis*>(c);
    return false;
  }
  b->companion = c;
  b->dLocalCoeff = buf; b->dLocalCoeffBytes = c->dimension * sizeof(double);
  return true;
}

int dfx_evaluate_path(const dfx_basis* b, int32_t node, const double* xhat_host, int32_t n_q, int32_t want_jac) {
  if (!b || !xhat_host || n_q < 1) return -1;
  int fl, nl;
  if (subtree(b, node, fl, nl)) return -1;
  FastPlan plan;
  if (b->oriented) return 0;
  if (g_forcedPath == 3) return dmmaPathHas(b, fl, nl) ? 3 : 0;
  const int path = planFastEval(b, fl, nl, xhat_host, n_q, want_jac != 0, g_forcedPath, plan);
  if (path == 0 && g_forcedPath < 0 && dmmaPathHas(b, fl, nl)) return 3;
  return path;
}

int dfx_evaluate(const dfx_basis* cb, int32_t node, const double* coeff, const double* xhat_host, int32_t n_q,
                 uint64_t e0, uint64_t e1, double* out_values, double* out_jac, void* stream) {
  DFX_RANGE("dfx_evaluate");
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  int fl, nl;
  if ((rc = subtree(b, node, fl, nl))) return rc;
  if ((rc = checkRange(b, e0, e1))) return rc;
  if (!coeff || !xhat_host || n_q < 1) return fail(DFX_ERR_INVALID, "null argument");
  if (!out_values && !out_jac) return DFX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (b->oriented) {
    // permuted face DOFs (dfx_basis_set_vertex_ids): the closed-form gathers of the fast paths do not
    // apply.  LocalFunction::bind as its own pass — the element-local coefficients are gathered through
    // layoutIndexOriented into the layout of the companion basis (same tree, discontinuous leaves) —
    // then the companion evaluates on the element-local fast kernels.
    if (g_forcedPath != 0 && nl == 1 && !out_jac) {
      // continuous order 3 at 4^3 points: the register kernel reads one orientation byte per edge / facet of the
      // element and permutes the in-entity index of its gathers — no separate gather pass, no companion
      FastPlan plan;
      if (planFastEval(b, fl, nl, xhat_host, n_q, false, 2, plan) == 2) {
        rc = runQ3Eval(b, plan, fl, coeff, e0, e1, out_values, st);
        if (rc >= 0) return rc;
      }
    }
    if (g_forcedPath != 0 && orientedCompanion(b)) {
      dfx_basis* c = static_cast<dfx_basis*>(b->companion);
      // the gather walks like the fast index table (path 1 of DFX_OP_INDICES) or like k_indices (path 0)
      if (g_indicesPath != 0 && fits32(b, e0, e1)) rc = gatherLocalFast(b, c, fl, nl, coeff, (double*)b->dLocalCoeff, e0, e1, st);
      else rc = gatherLocal(b, c, fl, nl, coeff, (double*)b->dLocalCoeff, e0, e1, st);
      if (rc) return rc;
      return dfx_evaluate(c, node, (const double*)b->dLocalCoeff, xhat_host, n_q, e0, e1, out_values, out_jac, stream);
    }
    // no companion (order beyond the discontinuous kernels, no memory, or path 0 forced): every
    // coefficient is located through layoutIndexOriented at every point
    ShapeOffsets so; const double *dS, *dDS;
    if ((rc = prepareShape(b, node, fl, nl, xhat_host, n_q, so, dS, dDS, st))) return rc;
    return evalGeneric(b, fl, nl, coeff, dS, dDS, so, n_q, e0, e1, out_values, out_jac, st);
  }
  FastPlan plan;
  const int path = g_forcedPath == 3 ? 0 : planFastEval(b, fl, nl, xhat_host, n_q, out_jac != nullptr, g_forcedPath, plan);
  if (path > 0) return runFastEval(b, plan, fl, nl, coeff, e0, e1, out_values, out_jac, st);
  // a subtree of mixed orders (Taylor-Hood: Q2 velocity + Q1 pressure): every run of same-order leaves
  // gets its own plan and writes its range entries into the interleaved output
  if (g_forcedPath < 0 && nl > 1) {
    std::vector<FastPlan> plans;
    std::vector<std::pair<int, int>> runs;
    const DevBasis& D = b->dev;
    bool allFast = true;
    for (int l = 0; l < nl && allFast;) {
      int r = l + 1;
      while (r < nl && D.layouts[D.leaves[fl + r].layout].k == D.layouts[D.leaves[fl + l].layout].k) ++r;
      plans.emplace_back();
      allFast = planFastEval(b, fl + l, r - l, xhat_host, n_q, out_jac != nullptr, -1, plans.back()) > 0;
      runs.push_back({l, r - l});
      l = r;
    }
    if (allFast && runs.size() == 2) {
      // Taylor-Hood shape (order-2 run + order-1 run): one launch that stages whole points
      rc = runSplitEval(b, plans[0], fl + runs[0].first, runs[0].second, &plans[1], fl + runs[1].first, runs[1].second, coeff,
                        e0, e1, out_values, out_jac, st);
      if (rc >= 0) return rc;
    }
    if (allFast && runs.size() > 1) {
      for (size_t q = 0; q < runs.size(); ++q)
        if ((rc = runFastEval(b, plans[q], fl + runs[q].first, runs[q].second, coeff, e0, e1, out_values, out_jac, st,
                              runs[q].first, nl))) return rc;
      return DFX_OK;
    }
  }
  // no tensor-product structure: dense contraction on the FP64 tensor cores when the element is big enough
  if ((g_forcedPath == 3 || g_forcedPath < 0) && dmmaPathHas(b, fl, nl) && dfx_basis_num_elements(b) <= 0xffffffffull)
    return runDmmaEval(b, node, fl, nl, xhat_host, n_q, coeff, e0, e1, out_values, out_jac, st);
  if (g_forcedPath > 0) return fail(DFX_ERR_NOT_IMPLEMENTED, "forced evaluation path does not apply to this configuration");
  ShapeOffsets so; const double *dS, *dDS;
  if ((rc = prepareShape(b, node, fl, nl, xhat_host, n_q, so, dS, dDS, st))) return rc;
  return evalGeneric(b, fl, nl, coeff, dS, dDS, so, n_q, e0, e1, out_values, out_jac, st);
}

int dfx_evaluate_global(const dfx_basis* b, int32_t node, const double* coeff, const double* x, uint64_t n_points,
                        double* out_values, double* out_jac, int32_t* out_element, void* stream) {
  DFX_RANGE("dfx_evaluate_global");
  int rc = ensureDevice(b); if (rc) return rc;
  int fl, nl;
  if ((rc = subtree(b, node, fl, nl))) return rc;
  if (!coeff || (!x && n_points)) return fail(DFX_ERR_INVALID, "null argument");
  if (!out_values && !out_jac) return fail(DFX_ERR_INVALID, "no output requested");
  if (dfx_basis_num_elements(b) > 0x7fffffffull) return fail(DFX_ERR_RANGE, "more than 2^31 elements in one box");
  return evalGlobal(b, fl, nl, coeff, x, n_points, out_values, out_jac, out_element, (cudaStream_t)stream);
}

// ---- host-buffer entry points -------------------------------------------------------
int dfx_interpolate_host(const dfx_basis* cb, int32_t node, const dfx_function* f, double* coeff_host,
                         const uint8_t* mask_host) {
  DFX_RANGE("dfx_interpolate_host");
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!coeff_host) return fail(DFX_ERR_INVALID, "null coefficient vector");
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t n = b->dimension;
  if ((rc = growScratch(b, 0, n * sizeof(double)))) return rc;
  double* dC = (double*)b->dScratch[0];
  uint8_t* dM = nullptr;
  // entries outside the mask / the subtree keep their value: bring them over first
  if (mask_host || node != 0)
    DFX_CUDA(cudaMemcpyAsync(dC, coeff_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (mask_host) {
    if ((rc = growScratch(b, 1, n))) return rc;
    dM = (uint8_t*)b->dScratch[1];
    DFX_CUDA(cudaMemcpyAsync(dM, mask_host, n, cudaMemcpyHostToDevice, st));
  }
  if ((rc = dfx_interpolate(b, node, f, dC, dM, st))) return rc;
  DFX_CUDA(cudaMemcpyAsync(coeff_host, dC, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

// Pipelined: the element range is cut into bands of z-layers; an input stream brings over the
// coefficients each band needs (only what earlier bands have not brought yet) while the work
// stream evaluates the previous band and sends its results back, so both PCIe directions are
// busy at once and the device never holds more than one band of results.
int dfx_evaluate_host(const dfx_basis* cb, int32_t node, const double* coeff_host, const double* xhat_host,
                      int32_t n_q, uint64_t e0, uint64_t e1, double* out_values_host, double* out_jac_host) {
  DFX_RANGE("dfx_evaluate_host");
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  int fl, nl;
  if ((rc = subtree(b, node, fl, nl))) return rc;
  if ((rc = checkRange(b, e0, e1))) return rc;
  if (!coeff_host || !xhat_host || n_q < 1) return fail(DFX_ERR_INVALID, "null argument");
  if (e0 == e1 || (!out_values_host && !out_jac_host)) return DFX_OK;
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  if (!b->evStream[0]) {
    cudaStream_t s;
    DFX_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    b->evStream[0] = (void*)s;
  }
  cudaStream_t sin = (cudaStream_t)b->evStream[0];
  const size_t n = b->dimension;
  const int dim = b->dev.grid.dim;
  if ((rc = growScratch(b, 0, n * sizeof(double)))) return rc;
  double* dC = (double*)b->dScratch[0];
  const size_t perElemV = (size_t)n_q * nl * sizeof(double);
  const size_t perElemJ = perElemV * dim;
  const size_t perElem = (out_values_host ? perElemV : 0) + (out_jac_host ? perElemJ : 0);
  // bands of whole z-layers, about 192 MiB of results each
  u64 layer = 1;
  for (int j = 0; j + 1 < dim; ++j) layer *= (u64)b->dev.grid.N[j];
  const u64 bandBytes = g_hostBandKiB > 0 ? (u64)g_hostBandKiB << 10 : 192ull << 20;
  u64 chunk = std::max<u64>(1, bandBytes / perElem);
  chunk = std::max<u64>(layer, chunk / layer * layer);
  const u64 bandMax = std::min<u64>(chunk, e1 - e0);
  if (out_values_host && (rc = growScratch(b, 1, bandMax * perElemV))) return rc;
  if (out_jac_host && (rc = growScratch(b, 2, bandMax * perElemJ))) return rc;
  double* dV = out_values_host ? (double*)b->dScratch[1] : nullptr;
  double* dJ = out_jac_host ? (double*)b->dScratch[2] : nullptr;
  // with the test knob set, poison the coefficient scratch (NaN) so that a band that gathers
  // something it did not bring over cannot pass unnoticed
  if (g_hostBandKiB > 0) DFX_CUDA(cudaMemsetAsync(dC, 0xFF, n * sizeof(double), sin));
  Intervals have;
  std::vector<cudaEvent_t> events;
  auto cleanup = [&]() { for (auto ev : events) cudaEventDestroy(ev); };
  // band boundaries on multiples of `chunk` so that a band is whole layers except at the ends
  for (u64 c0 = e0; c0 < e1;) {
    const u64 c1 = std::min<u64>(e1, (c0 / chunk + 1) * chunk);
    Intervals need = subtract(layerBandRanges(b, fl, nl, c0 / layer, (c1 - 1) / layer + 1), have);
    for (auto& r : need) {
      cudaError_t e = cudaMemcpyAsync(dC + r.first, coeff_host + r.first, (r.second - r.first) * sizeof(double),
                                      cudaMemcpyHostToDevice, sin);
      if (e != cudaSuccess) { cleanup(); return cudaFail(e, "cudaMemcpyAsync(coefficients)"); }
      have.push_back(r);
    }
    normalise(have);
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cleanup(); return fail(DFX_ERR_CUDA, "cudaEventCreate"); }
    events.push_back(ev);
    cudaEventRecord(ev, sin);
    cudaStreamWaitEvent(st, ev, 0);
    if ((rc = dfx_evaluate(b, node, dC, xhat_host, n_q, c0, c1, dV, dJ, st))) { cudaStreamSynchronize(sin); cleanup(); return rc; }
    cudaError_t e = cudaSuccess;
    if (dV) e = cudaMemcpyAsync((char*)out_values_host + (c0 - e0) * perElemV, dV, (c1 - c0) * perElemV, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && dJ)
      e = cudaMemcpyAsync((char*)out_jac_host + (c0 - e0) * perElemJ, dJ, (c1 - c0) * perElemJ, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) { cudaStreamSynchronize(sin); cleanup(); return cudaFail(e, "cudaMemcpyAsync(results)"); }
    c0 = c1;
  }
  cudaError_t e1s = cudaStreamSynchronize(sin), e2s = cudaStreamSynchronize(st);
  cleanup();
  if (e1s != cudaSuccess) return cudaFail(e1s, "cudaStreamSynchronize");
  if (e2s != cudaSuccess) return cudaFail(e2s, "cudaStreamSynchronize");
  return DFX_OK;
}

int dfx_evaluate_global_host(const dfx_basis* cb, int32_t node, const double* coeff_host, const double* x_host,
                             uint64_t n_points, double* out_values_host, double* out_jac_host, int32_t* out_element_host) {
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  int fl, nl;
  if ((rc = subtree(b, node, fl, nl))) return rc;
  if (!coeff_host || (!x_host && n_points)) return fail(DFX_ERR_INVALID, "null argument");
  if (!out_values_host && !out_jac_host) return fail(DFX_ERR_INVALID, "no output requested");
  if (n_points == 0) return DFX_OK;
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t n = b->dimension, dim = b->dev.grid.dim;
  // the coefficient vector stays on the device between calls with the same host pointer and content
  // version is unknown, so it is sent every time; points and results share one scratch block
  if ((rc = growScratch(b, 0, n * sizeof(double)))) return rc;
  const size_t perPoint = dim + (size_t)nl * (1 + dim);
  if ((rc = growScratch(b, 1, n_points * perPoint * sizeof(double)))) return rc;
  if ((rc = growScratch(b, 2, n_points * sizeof(int32_t)))) return rc;
  double* dC = (double*)b->dScratch[0];
  double* dX = (double*)b->dScratch[1];
  double* dV = dX + n_points * dim;
  double* dJ = dV + n_points * (size_t)nl;
  int32_t* dE = (int32_t*)b->dScratch[2];
  DFX_CUDA(cudaMemcpyAsync(dC, coeff_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
  DFX_CUDA(cudaMemcpyAsync(dX, x_host, n_points * dim * sizeof(double), cudaMemcpyHostToDevice, st));
  if ((rc = dfx_evaluate_global(b, node, dC, dX, n_points, out_values_host ? dV : nullptr, out_jac_host ? dJ : nullptr, dE, st))) return rc;
  if (out_values_host) DFX_CUDA(cudaMemcpyAsync(out_values_host, dV, n_points * (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (out_jac_host) DFX_CUDA(cudaMemcpyAsync(out_jac_host, dJ, n_points * (size_t)nl * dim * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (out_element_host) DFX_CUDA(cudaMemcpyAsync(out_element_host, dE, n_points * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

int dfx_indices_flat_host(const dfx_basis* cb, uint64_t e0, uint64_t e1, uint32_t* out_host) {
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if ((rc = checkRange(b, e0, e1))) return rc;
  if (!out_host) return fail(DFX_ERR_INVALID, "null output");
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t bytes = (e1 - e0) * (size_t)b->dev.nlocTotal * sizeof(uint32_t);
  if (bytes == 0) return DFX_OK;
  if ((rc = growScratch(b, 1, bytes))) return rc;
  if ((rc = dfx_indices_flat(b, e0, e1, (uint32_t*)b->dScratch[1], st))) return rc;
  DFX_CUDA(cudaMemcpyAsync(out_host, b->dScratch[1], bytes, cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

int dfx_indices_multi_host(const dfx_basis* cb, uint64_t e0, uint64_t e1, uint64_t* out_host, int32_t stride) {
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if ((rc = checkRange(b, e0, e1))) return rc;
  if (!out_host) return fail(DFX_ERR_INVALID, "null output");
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t bytes = (e1 - e0) * (size_t)b->dev.nlocTotal * (size_t)stride * sizeof(uint64_t);
  if (bytes == 0) return DFX_OK;
  if ((rc = growScratch(b, 1, bytes))) return rc;
  if ((rc = dfx_indices_multi_u64(b, e0, e1, (uint64_t*)b->dScratch[1], stride, st))) return rc;
  DFX_CUDA(cudaMemcpyAsync(out_host, b->dScratch[1], bytes, cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

int dfx_interpolation_points_host(const dfx_basis* cb, double* out_xyz_host, int32_t* out_leaf_host) {
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!out_xyz_host || !out_leaf_host) return fail(DFX_ERR_INVALID, "null output");
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t n = b->dimension, dim = b->dev.grid.dim;
  if ((rc = growScratch(b, 1, n * dim * sizeof(double)))) return rc;
  if ((rc = growScratch(b, 2, n * sizeof(int32_t)))) return rc;
  if ((rc = dfx_interpolation_points(b, (double*)b->dScratch[1], (int32_t*)b->dScratch[2], st))) return rc;
  DFX_CUDA(cudaMemcpyAsync(out_xyz_host, b->dScratch[1], n * dim * sizeof(double), cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaMemcpyAsync(out_leaf_host, b->dScratch[2], n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

// ---- boundary DOFs ---------------------------------------------------------------------
// state of reference-cube sub-entity (index, codim) per direction: 0 lower end, 1 upper end,
// 2 extended (dune-geometry cube numbering)
static bool subEntityState(int dim, int index, int codim, int st[3]) {
  st[0] = st[1] = st[2] = 2;
  if (codim < 0 || codim > dim || index < 0) return false;
  if (codim == 0) return index == 0;
  if (codim == dim) { if (index >= (1 << dim)) return false; for (int j = 0; j < dim; ++j) st[j] = (index >> j) & 1; return true; }
  if (codim == 1) { if (index >= 2 * dim) return false; st[index / 2] = index % 2; return true; }
  // dim == 3, codim == 2: edges 0-3 z-directed, 4,5 / 8,9 y-directed at z = 0 / 1, 6,7 / 10,11 x-directed
  if (index >= 12) return false;
  if (index < 4) { st[0] = index & 1; st[1] = (index >> 1) & 1; return true; }
  const int r = (index - 4) & 3;
  st[2] = (index - 4) >> 2;
  if (r < 2) st[0] = r; else st[1] = r - 2;
  return true;
}

int dfx_subentity_dofs(const dfx_basis* b, int32_t node, int32_t sub_entity_index, int32_t sub_entity_codim,
                       uint8_t* contained_host, int32_t* out_local_host, int32_t* n) {
  if (!b || !n) return fail(DFX_ERR_INVALID, "null argument");
  int fl, nl, rc;
  if ((rc = subtree(b, node, fl, nl))) return rc;
  const DevBasis& D = b->dev;
  const int dim = D.grid.dim;
  int st[3];
  if (!subEntityState(dim, sub_entity_index, sub_entity_codim, st))
    return fail(DFX_ERR_RANGE, "no such sub-entity of the reference cube");
  if (contained_host) std::memset(contained_host, 0, (size_t)D.nlocTotal);
  int cnt = 0;
  for (int l = fl; l < fl + nl; ++l) {
    const DevLeaf& leaf = D.leaves[l];
    const int k = D.layouts[leaf.layout].k;
    for (int i = 0; i < leaf.nloc; ++i) {
      bool in = true;
      int ii = i;
      for (int j = 0; j < dim; ++j) {
        const int a = ii % (k + 1); ii /= (k + 1);
        const int sd = k == 0 ? 2 : (a == 0 ? 0 : (a == k ? 1 : 2));     // where the DOF's own sub-entity sits
        if (st[j] != 2 && sd != st[j]) in = false;
      }
      if (!in) continue;
      if (contained_host) contained_host[leaf.localOffset + i] = 1;
      if (out_local_host) out_local_host[cnt] = leaf.localOffset + i;
      ++cnt;
    }
  }
  *n = cnt;
  return DFX_OK;
}

int dfx_boundary_dofs(const dfx_basis* b, int32_t node, uint8_t* mask, void* stream) {
  DFX_RANGE("dfx_boundary_dofs");
  int rc = ensureDevice(b); if (rc) return rc;
  int fl, nl;
  if ((rc = subtree(b, node, fl, nl))) return rc;
  if (!mask) return fail(DFX_ERR_INVALID, "null mask");
  return boundaryDofs(b, fl, nl, mask, (cudaStream_t)stream);
}

int dfx_boundary_dofs_host(const dfx_basis* cb, int32_t node, uint8_t* mask_host) {
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!mask_host) return fail(DFX_ERR_INVALID, "null mask");
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t n = b->dimension;
  if ((rc = growScratch(b, 1, n))) return rc;
  uint8_t* dM = (uint8_t*)b->dScratch[1];
  DFX_CUDA(cudaMemcpyAsync(dM, mask_host, n, cudaMemcpyHostToDevice, st));   // other bytes keep their value
  if ((rc = dfx_boundary_dofs(b, node, dM, st))) return rc;
  DFX_CUDA(cudaMemcpyAsync(mask_host, dM, n, cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

// ---- element-loop callers ------------------------------------------------------------------
// typical number of entries of a matrix row of this basis (interior node: every leaf's DOFs of a
// 2^dim element box, about half of it on average), used to pick the SpMV lane grouping
static int meanRowLength(const dfx_basis* b) {
  const DevBasis& D = b->dev;
  double len = 0;
  for (int l = 0; l < D.nLeaves; ++l) {
    const DevLayout& L = D.layouts[D.leaves[l].layout];
    const bool cell = L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0;
    double n = 1;
    for (int j = 0; j < D.grid.dim; ++j) n *= cell ? (L.k + 1) : (1.5 * L.k + 1);
    len += n;
  }
  return (int)len;
}

int dfx_matrix_pattern(const dfx_basis* cb, uint64_t* row_ptr, uint32_t* col_idx, uint64_t* nnz_host, void* stream) {
  DFX_RANGE("dfx_matrix_pattern");
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!row_ptr) return fail(DFX_ERR_INVALID, "null row_ptr");
  return matrixPattern(b, (u64*)row_ptr, col_idx, (u64*)nnz_host, (cudaStream_t)stream);
}

int dfx_assemble_laplace(const dfx_basis* cb, const uint64_t* row_ptr, const uint32_t* col_idx, double* values,
                         const dfx_function* f, double* rhs, void* stream) {
  DFX_RANGE("dfx_assemble_laplace");
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!row_ptr || !col_idx) return fail(DFX_ERR_INVALID, "null pattern");
  if (rhs && (rc = checkFunction(f, 1))) return rc;
  if (b->oriented) return fail(DFX_ERR_NOT_IMPLEMENTED, "Laplace assembly with permuted face DOFs (dfx_basis_set_vertex_ids)");
  return assembleLaplace(b, (const u64*)row_ptr, col_idx, values, f, rhs, (cudaStream_t)stream);
}

int dfx_csr_mv(const dfx_basis* b, const uint64_t* row_ptr, const uint32_t* col_idx, const double* values,
               const double* x, double* y, int32_t subtract, void* stream) {
  DFX_RANGE("dfx_csr_mv");
  int rc = ensureDevice(b); if (rc) return rc;
  if (!row_ptr || !col_idx || !values || !x || !y) return fail(DFX_ERR_INVALID, "null argument");
  return csrMv(b->dimension, (const u64*)row_ptr, col_idx, values, x, y, subtract != 0, (cudaStream_t)stream, meanRowLength(b));
}

int dfx_apply_dirichlet(const dfx_basis* b, const uint64_t* row_ptr, const uint32_t* col_idx, double* values,
                        const uint8_t* dirichlet, const double* x, double* rhs, void* stream) {
  DFX_RANGE("dfx_apply_dirichlet");
  int rc = ensureDevice(b); if (rc) return rc;
  if (!row_ptr || !col_idx || !values || !dirichlet || !x || !rhs) return fail(DFX_ERR_INVALID, "null argument");
  return dirichletRows(b->dimension, (const u64*)row_ptr, col_idx, values, dirichlet, x, rhs, (cudaStream_t)stream, meanRowLength(b));
}

int dfx_assemble_stokes(const dfx_basis* cb, const uint64_t* row_ptr, const uint32_t* col_idx, double* values, void* stream) {
  DFX_RANGE("dfx_assemble_stokes");
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!row_ptr || !col_idx || !values) return fail(DFX_ERR_INVALID, "null argument");
  return assembleStokes(b, (const u64*)row_ptr, col_idx, values, (cudaStream_t)stream);
}

int dfx_apply_laplace(const dfx_basis* cb, const double* x, double* y, const uint8_t* dirichlet, double* scratch,
                      void* stream) {
  DFX_RANGE("dfx_apply_laplace");
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!x || !y || !scratch) return fail(DFX_ERR_INVALID, "null argument");
  if (b->oriented) return fail(DFX_ERR_NOT_IMPLEMENTED, "matrix-free Laplace with permuted face DOFs (dfx_basis_set_vertex_ids)");
  return applyLaplace(b, x, y, dirichlet, scratch, (cudaStream_t)stream);
}

int dfx_matrix_pattern_host(const dfx_basis* cb, uint64_t* row_ptr_host, uint32_t* col_idx_host, uint64_t* nnz_host) {
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!row_ptr_host) return fail(DFX_ERR_INVALID, "null row_ptr");
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t n = b->dimension;
  if ((rc = growScratch(b, 1, (n + 1) * sizeof(u64)))) return rc;
  u64* dRow = (u64*)b->dScratch[1];
  u64 nnz = 0;
  if ((rc = matrixPattern(b, dRow, nullptr, &nnz, st))) return rc;
  DFX_CUDA(cudaMemcpyAsync(row_ptr_host, dRow, (n + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
  if (col_idx_host && nnz) {
    if ((rc = growScratch(b, 2, nnz * sizeof(uint32_t)))) return rc;
    if ((rc = matrixPattern(b, dRow, (uint32_t*)b->dScratch[2], nullptr, st))) return rc;
    DFX_CUDA(cudaMemcpyAsync(col_idx_host, b->dScratch[2], nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  DFX_CUDA(cudaStreamSynchronize(st));
  if (nnz_host) *nnz_host = nnz;
  return DFX_OK;
}

int dfx_assemble_laplace_host(const dfx_basis* cb, const uint64_t* row_ptr_host, const uint32_t* col_idx_host,
                              double* values_host, const dfx_function* f, double* rhs_host) {
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!row_ptr_host || !col_idx_host) return fail(DFX_ERR_INVALID, "null pattern");
  if (rhs_host && (rc = checkFunction(f, 1))) return rc;
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t n = b->dimension;
  const u64 nnz = row_ptr_host[n];
  if ((rc = growScratch(b, 1, (n + 1) * sizeof(u64)))) return rc;
  if ((rc = growScratch(b, 2, std::max<size_t>(1, nnz) * sizeof(uint32_t)))) return rc;
  if ((rc = growScratch(b, 0, (std::max<size_t>(1, nnz) + n) * sizeof(double)))) return rc;
  u64* dRow = (u64*)b->dScratch[1];
  uint32_t* dCol = (uint32_t*)b->dScratch[2];
  double* dVal = (double*)b->dScratch[0];
  double* dRhs = dVal + std::max<size_t>(1, nnz);
  DFX_CUDA(cudaMemcpyAsync(dRow, row_ptr_host, (n + 1) * sizeof(u64), cudaMemcpyHostToDevice, st));
  DFX_CUDA(cudaMemcpyAsync(dCol, col_idx_host, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  if ((rc = assembleLaplace(b, dRow, dCol, values_host ? dVal : nullptr, f, rhs_host ? dRhs : nullptr, st))) return rc;
  if (values_host) DFX_CUDA(cudaMemcpyAsync(values_host, dVal, nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (rhs_host) DFX_CUDA(cudaMemcpyAsync(rhs_host, dRhs, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

int dfx_assemble_stokes_host(const dfx_basis* cb, const uint64_t* row_ptr_host, const uint32_t* col_idx_host,
                             double* values_host) {
  dfx_basis* b = const_cast<dfx_basis*>(cb);
  int rc = ensureDevice(b); if (rc) return rc;
  if (!row_ptr_host || !col_idx_host || !values_host) return fail(DFX_ERR_INVALID, "null argument");
  cudaStream_t st; if ((rc = privateStream(b, st))) return rc;
  const size_t n = b->dimension;
  const u64 nnz = row_ptr_host[n];
  if ((rc = growScratch(b, 1, (n + 1) * sizeof(u64)))) return rc;
  if ((rc = growScratch(b, 2, std::max<size_t>(1, nnz) * sizeof(uint32_t)))) return rc;
  if ((rc = growScratch(b, 0, std::max<size_t>(1, nnz) * sizeof(double)))) return rc;
  u64* dRow = (u64*)b->dScratch[1];
  uint32_t* dCol = (uint32_t*)b->dScratch[2];
  double* dVal = (double*)b->dScratch[0];
  DFX_CUDA(cudaMemcpyAsync(dRow, row_ptr_host, (n + 1) * sizeof(u64), cudaMemcpyHostToDevice, st));
  DFX_CUDA(cudaMemcpyAsync(dCol, col_idx_host, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  if ((rc = assembleStokes(b, dRow, dCol, dVal, st))) return rc;
  DFX_CUDA(cudaMemcpyAsync(values_host, dVal, nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  return DFX_OK;
}

// ---- slab partition helpers ------------------------------------------------------------
int dfx_halo_ranges(const dfx_basis* b, int32_t side, uint64_t* out_begin, uint64_t* out_count, int32_t max_ranges,
                    int32_t* n) {
  if (!b || !out_begin || !out_count || !n) return fail(DFX_ERR_INVALID, "null argument");
  const DevBasis& D = b->dev;
  const int dz = D.grid.dim - 1;
  int cnt = 0;
  for (int l = 0; l < D.nLeaves; ++l) {
    const DevLeaf& leaf = D.leaves[l];
    const DevLayout& L = D.layouts[leaf.layout];
    if (L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0) continue;
    if (leaf.posA != 1) return fail(DFX_ERR_NOT_IMPLEMENTED, "interleaved layout: use dfx_halo_pack/unpack");
    for (int q = 0; q < L.ncomp; ++q) {
      const int s = L.order[q];
      if ((s >> dz) & 1) continue;
      const u64 plane = L.compCount[s] / (u64)L.csize[s][dz];
      if (cnt >= max_ranges) return fail(DFX_ERR_RANGE, "max_ranges too small");
      out_begin[cnt] = leaf.posB + L.compOffset[s] + plane * (u64)(side ? L.csize[s][dz] - 1 : 0);
      out_count[cnt] = plane;
      ++cnt;
    }
  }
  *n = cnt;
  return DFX_OK;
}

uint64_t dfx_halo_size(const dfx_basis* b) { return b ? haloCount(b) : 0; }

int dfx_halo_pack(const dfx_basis* b, int32_t side, const double* coeff, double* buf, void* stream) {
  DFX_RANGE("dfx_halo_pack");
  int rc = ensureDevice(b); if (rc) return rc;
  return haloCopy(b, side, true, const_cast<double*>(coeff), buf, haloCount(b), (cudaStream_t)stream);
}
int dfx_halo_unpack(const dfx_basis* b, int32_t side, const double* buf, double* coeff, void* stream) {
  DFX_RANGE("dfx_halo_unpack");
  int rc = ensureDevice(b); if (rc) return rc;
  return haloCopy(b, side, false, coeff, const_cast<double*>(buf), haloCount(b), (cudaStream_t)stream);
}

int dfx_owned_ranges(const dfx_basis* b, uint64_t* local_begin, uint64_t* count, uint64_t* global_begin,
                     int32_t max_ranges, int32_t* n) {
  if (!b || !local_begin || !count || !global_begin || !n) return fail(DFX_ERR_INVALID, "null argument");
  const DevBasis& D = b->dev;
  const int dim = D.grid.dim, dz = dim - 1;
  for (int j = 0; j < dz; ++j)
    if (D.grid.off[j] != 0 || D.grid.N[j] != D.grid.G[j])
      return fail(DFX_ERR_NOT_IMPLEMENTED, "owned ranges need a slab partition along the last direction");
  // the same tree on the whole grid gives the global numbering
  dfx_grid_desc gg = b->gridDesc;
  for (int j = 0; j < dim; ++j) { gg.local_begin[j] = 0; gg.local_n[j] = gg.n_global[j]; }
  dfx_basis_impl* gimpl = nullptr;
  int rc = buildBasis(&gg, b->treeDesc.data(), (int)b->treeDesc.size(), b->device, &gimpl);
  if (rc) return rc;
  std::unique_ptr<dfx_basis> gb(static_cast<dfx_basis*>(gimpl));
  const DevBasis& G = gb->dev;
  const bool last = D.grid.off[dz] + D.grid.N[dz] == D.grid.G[dz];
  int cnt = 0;
  for (int l = 0; l < D.nLeaves; ++l) {
    const DevLeaf& leaf = D.leaves[l];
    const DevLeaf& gleaf = G.leaves[l];
    if (leaf.posA != 1 || gleaf.posA != 1)
      return fail(DFX_ERR_NOT_IMPLEMENTED, "owned ranges need a non-interleaved flat layout");
    const DevLayout& L = D.layouts[leaf.layout];
    const DevLayout& GL = G.layouts[gleaf.layout];
    for (int q = 0; q < L.ncomp; ++q) {
      const int s = L.order[q];
      const u64 plane = L.compCount[s] / (u64)L.csize[s][dz];    // same in the global layout
      u64 layers = (u64)L.csize[s][dz];
      if (!((s >> dz) & 1) && !last && L.kind != DFX_NODE_LAGRANGE_DG) layers -= 1;   // top plane: upper neighbour's
      if (cnt >= max_ranges) return fail(DFX_ERR_RANGE, "max_ranges too small");
      local_begin[cnt] = leaf.posB + L.compOffset[s];
      count[cnt] = plane * layers;
      global_begin[cnt] = gleaf.posB + GL.compOffset[s] + plane * (u64)D.grid.off[dz];
      ++cnt;
    }
  }
  *n = cnt;
  return DFX_OK;
}

// ---- peer-memory halo link ------------------------------------------------------------------
int dfx_halo_link_create(const dfx_basis* b, dfx_halo_link** out) {
  int rc = ensureDevice(b); if (rc) return rc;
  if (!out) return fail(DFX_ERR_INVALID, "null argument");
  auto* l = new dfx_halo_link();
  l->device = b->device;
  l->n = haloCount(b);
  size_t bytes; haloLinkBytes(l->n, bytes);
  cudaError_t e = cudaMalloc((void**)&l->local, bytes);
  if (e != cudaSuccess) { delete l; return cudaFail(e, "cudaMalloc(halo link)"); }
  e = cudaMemset(l->local, 0, bytes);
  if (e != cudaSuccess) { cudaFree(l->local); delete l; return cudaFail(e, "cudaMemset(halo link)"); }
  *out = l;
  return dfx_halo_link_set_timeout(l, 4.0);
}

// control words of the link block (dfx_kernels.cu): 2 = sticky error, 3 = wait limit in SM clock cycles
static int linkWord(dfx_halo_link* l, int which, unsigned long long* get, const unsigned long long* set) {
  if (!l) return fail(DFX_ERR_INVALID, "null halo link");
  DFX_CUDA(cudaSetDevice(l->device));
  char* p = reinterpret_cast<char*>(l->local) + haloLinkWordOffset(l->n, which);
  if (set) DFX_CUDA(cudaMemcpy(p, set, sizeof(*set), cudaMemcpyHostToDevice));
  if (get) DFX_CUDA(cudaMemcpy(get, p, sizeof(*get), cudaMemcpyDeviceToHost));
  return DFX_OK;
}

int dfx_halo_link_set_timeout(dfx_halo_link* l, double seconds) {
  if (!l || !(seconds > 0)) return fail(DFX_ERR_INVALID, "bad halo link timeout");
  int khz = 0;
  DFX_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, l->device));
  const unsigned long long cycles = (unsigned long long)(seconds * 1e3 * (double)std::max(khz, 1000000));
  return linkWord(l, 3, nullptr, &cycles);
}

int dfx_halo_link_reset_error(dfx_halo_link* l) {
  const unsigned long long zero = 0;
  int rc = linkWord(l, 2, nullptr, &zero);                     // error
  if (!rc) rc = linkWord(l, 4, nullptr, &zero);                // blocks-done counters of both directions
  if (!rc) rc = linkWord(l, 5, nullptr, &zero);
  return rc;
}

void dfx_halo_link_destroy(dfx_halo_link* l) {
  if (!l) return;
  cudaSetDevice(l->device);
  if (l->lower && l->lowerIpc) cudaIpcCloseMemHandle(l->lower);
  if (l->upper && l->upperIpc) cudaIpcCloseMemHandle(l->upper);
  if (l->local) cudaFree(l->local);
  delete l;
}

int dfx_halo_link_export(dfx_halo_link* l, uint8_t ipc_handle_out[64]) {
  if (!l || !ipc_handle_out) return fail(DFX_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  DFX_CUDA(cudaSetDevice(l->device));
  cudaIpcMemHandle_t h;
  DFX_CUDA(cudaIpcGetMemHandle(&h, l->local));
  std::memcpy(ipc_handle_out, &h, 64);
  return DFX_OK;
}

void* dfx_halo_link_block(dfx_halo_link* l) { return l ? l->local : nullptr; }

int dfx_halo_link_attach(dfx_halo_link* l, int32_t which, const uint8_t* ipc_handle, void* raw_block) {
  if (!l || (which != 0 && which != 1) || (!ipc_handle && !raw_block)) return fail(DFX_ERR_INVALID, "bad halo link attachment");
  DFX_CUDA(cudaSetDevice(l->device));
  void* p = raw_block;
  if (ipc_handle) {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, 64);
    DFX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  }
  if (which == 0) { l->lower = (double*)p; l->lowerIpc = ipc_handle != nullptr; }
  else { l->upper = (double*)p; l->upperIpc = ipc_handle != nullptr; }
  return DFX_OK;
}

uint64_t dfx_halo_link_error(dfx_halo_link* l) {
  unsigned long long v = 0;
  if (linkWord(l, 2, &v, nullptr) != DFX_OK) return ~0ull;
  return v;
}

int dfx_halo_push(const dfx_basis* b, dfx_halo_link* l, const double* coeff, uint64_t step, void* stream) {
  DFX_RANGE("dfx_halo_push");
  int rc = ensureDevice(b); if (rc) return rc;
  if (!l || !coeff || step == 0) return fail(DFX_ERR_INVALID, "bad argument");
  if (l->n != haloCount(b)) return fail(DFX_ERR_INVALID, "halo link belongs to another basis");
  return haloPush(b, l, coeff, step, (cudaStream_t)stream);
}

int dfx_halo_pull(const dfx_basis* b, dfx_halo_link* l, double* coeff, uint64_t step, void* stream) {
  DFX_RANGE("dfx_halo_pull");
  int rc = ensureDevice(b); if (rc) return rc;
  if (!l || !coeff || step == 0) return fail(DFX_ERR_INVALID, "bad argument");
  if (l->n != haloCount(b)) return fail(DFX_ERR_INVALID, "halo link belongs to another basis");
  return haloPull(b, l, coeff, step, (cudaStream_t)stream);
}

}  // extern "C"
