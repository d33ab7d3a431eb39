// dfx_launch.h — host-side launcher prototypes (kernels are defined in the .cu files).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>

#include "dfx_internal.h"

namespace dfx {

extern std::atomic<unsigned long long> g_launchCount;

// NVTX range around an ABI entry point (header-only NVTX 3: a no-op unless a profiler is attached;
// `ncu --nvtx --nvtx-include "dfx_evaluate/"` then selects the kernels of one call)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define DFX_RANGE(name) ::dfx::NvtxRange dfx_nvtx_range_(name)
const char* lastErrorCStr();

// offsets (in doubles) of each layout's shape / shape-jacobian table
struct ShapeOffsets { unsigned v[kMaxLayouts]; unsigned d[kMaxLayouts]; };

inline int postLaunch(const char* what) {
  g_launchCount.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(DFX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
  return DFX_OK;
}
inline int cudaFail(cudaError_t e, const char* what) {
  return fail(DFX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define DFX_CUDA(call)                                         \
  do {                                                         \
    cudaError_t e_ = (call);                                   \
    if (e_ != cudaSuccess) return ::dfx::cudaFail(e_, #call);  \
  } while (0)

// Opt a kernel into more than 48 KB of dynamic shared memory.  The attribute belongs to the (function,
// device) pair, so `configured` — a static of the calling launcher, i.e. one per kernel instantiation —
// keeps one bit per device ordinal; the limit is set to the architecture's maximum once, which costs
// nothing at launch (occupancy follows the size actually requested) and cannot be lowered by another
// thread that launches a smaller tile.
template <class Kernel>
inline int optInSharedMemory(Kernel kern, int device, std::atomic<unsigned long long>& configured) {
  const bool tracked = device >= 0 && device < 64;
  if (tracked && ((configured.load(std::memory_order_acquire) >> device) & 1ull)) return DFX_OK;
  DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (tracked) configured.fetch_or(1ull << device, std::memory_order_release);
  return DFX_OK;
}

int indicesMulti32(const dfx_basis* b, u64 e0, u64 e1, uint32_t* out, int stride, cudaStream_t st);
int indicesMulti64(const dfx_basis* b, u64 e0, u64 e1, u64* out, int stride, cudaStream_t st);
int indicesFlat32(const dfx_basis* b, u64 e0, u64 e1, uint32_t* out, cudaStream_t st);
int indicesFlat64(const dfx_basis* b, u64 e0, u64 e1, u64* out, cudaStream_t st);
int interpolateRegistry(const dfx_basis* b, int firstLeaf, int nLeaves, const dfx_function& fn, double* coeff,
                        const uint8_t* mask, cudaStream_t st);
int interpolationPoints(const dfx_basis* b, double* xyz, int* leafOut, cudaStream_t st);
int interpolateFromValues(const dfx_basis* b, int firstLeaf, int nLeaves, const double* values, double* coeff,
                          const uint8_t* mask, cudaStream_t st);
int evalGlobal(const dfx_basis* b, int firstLeaf, int nLeaves, const double* coeff, const double* x, u64 nPoints,
               double* values, double* jac, int* elementOut, cudaStream_t st);
int scatterNodes(const dfx_basis* b, int layoutId, int firstLeaf, int nLeaves, int ncompVals, u64 e0, u64 e1,
                 const double* vals, double* coeff, const uint8_t* mask, cudaStream_t st);
int gatherLocalFast(const dfx_basis* b, const dfx_basis* companion, int firstLeaf, int nLeaves, const double* src, double* dst,
                    u64 e0, u64 e1, cudaStream_t st);
int gatherLocal(const dfx_basis* b, const dfx_basis* companion, int firstLeaf, int nLeaves, const double* src, double* dst,
                u64 e0, u64 e1, cudaStream_t st);
int evalGeneric(const dfx_basis* b, int firstLeaf, int nLeaves, const double* coeff, const double* shape,
                const double* dshape, const ShapeOffsets& so, int nq, u64 e0, u64 e1, double* values, double* jac,
                cudaStream_t st);
int interpolateFast(dfx_basis* b, int firstLeaf, int nLeaves, const dfx_function& fn, double* coeff,
                    const uint8_t* mask, bool wantSep, cudaStream_t st);
int interpolateRunsOriented(dfx_basis* b, int leafId, const dfx_function& fn, double* coeff, cudaStream_t st);
extern int g_rowsLean;
extern int g_indicesTile;
int interpolateBatch2(dfx_basis* b, int flA, int nlA, const dfx_function& fa, int flB, int nlB, const dfx_function& fb,
                      double* coeff, cudaStream_t st);
template <class T, bool MULTI>
int indicesFast(const dfx_basis* b, u64 e0, u64 e1, T* out, int stride, cudaStream_t st);
int boundaryDofs(const dfx_basis* b, int firstLeaf, int nLeaves, uint8_t* mask, cudaStream_t st);
int matrixPattern(dfx_basis* b, u64* rowPtr, uint32_t* col, u64* nnzHost, cudaStream_t st);
int assembleLaplace(dfx_basis* b, const u64* rowPtr, const uint32_t* col, double* values, const dfx_function* f,
                    double* rhs, cudaStream_t st);
int assembleStokes(dfx_basis* b, const u64* rowPtr, const uint32_t* col, double* values, cudaStream_t st);
int applyLaplace(dfx_basis* b, const double* x, double* y, const uint8_t* dirichlet, double* local, cudaStream_t st);
int csrMv(u64 n, const u64* rowPtr, const uint32_t* col, const double* values, const double* x, double* y, bool subtract,
          cudaStream_t st, int meanRow);
int dirichletRows(u64 n, const u64* rowPtr, const uint32_t* col, double* values, const uint8_t* dirichlet, const double* x,
                  double* rhs, cudaStream_t st, int meanRow);
int haloPush(const dfx_basis* b, dfx_halo_link_impl* l, const double* coeff, u64 step, cudaStream_t st);
int haloPull(const dfx_basis* b, dfx_halo_link_impl* l, double* coeff, u64 step, cudaStream_t st);
int haloCopy(const dfx_basis* b, int side, bool pack, double* coeff, double* buf, u64 total, cudaStream_t st);

}  // namespace dfx
