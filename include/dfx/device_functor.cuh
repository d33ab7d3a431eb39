// device_functor.cuh — form (3) of f for interpolate() (see include/dfx.h): C++ callers that
// compile with nvcc pass their own __device__ functor; the kernel is instantiated in the
// caller's translation unit, the node positions and the leaf of every coefficient come from
// the library (dfx_interpolation_points), so the numbering / last-writer rule stay in one place.
//
//   struct MyF { __device__ double operator()(const double* x, int dim, int comp) const { ... } };
//   dfx::interpolate(basis, subtree_node, MyF{}, coeff_dev, mask_dev_or_null, stream);
//
// `comp` is the range entry (leaf number inside the subtree; 0 for a scalar function), i.e.
// what HierarchicNodeToRangeMap (hierarchicnodetorangemap.hh:33-48) selects in the reference.
// Replaces the callable argument of interpolate(), functionspacebases/interpolate.hh:204-211.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../dfx.h"

namespace dfx {

template <class F>
__global__ void __launch_bounds__(256)
k_user_interpolate(F f, int dim, int firstLeaf, int nLeaves, int scalar, unsigned long long n,
                   const double* __restrict__ xyz, const int* __restrict__ leaf, double* __restrict__ coeff,
                   const uint8_t* __restrict__ mask) {
  const unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int l = leaf[p] - firstLeaf;
  if (l < 0 || l >= nLeaves) return;                 // coefficient of a leaf outside the subtree
  if (mask != nullptr && !mask[p]) return;
  double x[3] = {0, 0, 0};
  for (int j = 0; j < dim; ++j) x[j] = xyz[p * dim + j];
  coeff[p] = f(x, dim, scalar ? 0 : l);
}

// Device-pointer, stream-ordered.  `scratch_xyz` ([dimension][dim] doubles) and `scratch_leaf`
// ([dimension] int32) are caller-owned device buffers; pass filled = true to reuse them across
// calls on the same basis.  Returns a dfx_status.
template <class F>
int interpolate(const dfx_basis* basis, int subtree_node, const F& f, bool scalar_function, double* coeff,
                const uint8_t* mask, double* scratch_xyz, int32_t* scratch_leaf, bool filled, int dim,
                cudaStream_t stream) {
  if (!filled) {
    int rc = dfx_interpolation_points(basis, scratch_xyz, scratch_leaf, (void*)stream);
    if (rc != DFX_OK) return rc;
  }
  int off = 0, size = 0, firstLeaf = 0, nLeaves = 0;
  int rc = dfx_basis_node_info(basis, subtree_node, &off, &size, &firstLeaf, &nLeaves);
  if (rc != DFX_OK) return rc;
  const unsigned long long n = dfx_basis_dimension(basis);
  const unsigned long long blocks = (n + 255) / 256;
  k_user_interpolate<F><<<(unsigned)blocks, 256, 0, stream>>>(f, dim, firstLeaf, nLeaves, scalar_function ? 1 : 0, n,
                                                              scratch_xyz, scratch_leaf, coeff, mask);
  return cudaPeekAtLastError() == cudaSuccess ? DFX_OK : DFX_ERR_CUDA;
}

}  // namespace dfx
