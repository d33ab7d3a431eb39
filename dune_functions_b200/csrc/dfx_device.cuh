// dfx_device.cuh — device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>

#include "dfx_internal.h"

namespace dfx {

// Runtime divisor with precomputed magic (Granlund-Montgomery / libdivide "branchfree"
// scheme for 32-bit unsigned): q = (((n - t) >> 1) + t) >> s with t = mulhi(n, m).
struct FastDiv {
  unsigned d, m, s;
  FastDiv() : d(1), m(0), s(0) {}
  explicit FastDiv(unsigned div) : d(div) {
    if (div == 1) { m = 0; s = 0; return; }
    unsigned l = 32 - __builtin_clz(div - 1);           // ceil(log2(d))
    unsigned long long pw = 1ull << l;
    m = (unsigned)(((pw - div) << 32) / div + 1);
    s = l - 1;
  }
  DFX_HD unsigned div(unsigned n) const {
    if (d == 1) return n;
#if defined(__CUDA_ARCH__)
    unsigned t = __umulhi(n, m);
#else
    unsigned t = (unsigned)(((unsigned long long)n * m) >> 32);
#endif
    return (((n - t) >> 1) + t) >> s;
  }
  DFX_HD void divmod(unsigned n, unsigned& q, unsigned& r) const { q = div(n); r = n - q * d; }
};

// coordinate(j, i) = origin + i*h, EquidistantOffsetCoordinates of YaspGrid; written with
// explicit round-to-nearest intrinsics so that no FMA contraction changes the bits.
__device__ __forceinline__ double gridCoordinate(const DevGrid& g, int j, int iGlobal) {
  return __dadd_rn(g.lower[j], __dmul_rn((double)iGlobal, g.h[j]));
}

// AxisAlignedCubeGeometry::global on element iGlobal: lower + xhat*(upper-lower)
__device__ __forceinline__ double elementGlobal(const DevGrid& g, int j, int iGlobal, double xhat) {
  const double lo = gridCoordinate(g, j, iGlobal);
  const double up = gridCoordinate(g, j, iGlobal + 1);
  return __dadd_rn(lo, __dmul_rn(xhat, __dsub_rn(up, lo)));
}

// AxisAlignedCubeGeometry::jacobianInverse diagonal entry: 1/(upper-lower)
__device__ __forceinline__ double elementJacInv(const DevGrid& g, int j, int iGlobal) {
  const double lo = gridCoordinate(g, j, iGlobal);
  const double up = gridCoordinate(g, j, iGlobal + 1);
  return __ddiv_rn(1.0, __dsub_rn(up, lo));
}

// local index -> lattice multi-index alpha, LagrangeCube numbering: i = sum alpha_j (k+1)^j
__device__ __forceinline__ void localAlpha(int k, int i, int a[3]) {
  const int n1 = k + 1;
  a[0] = i % n1; i /= n1;
  a[1] = i % n1; i /= n1;
  a[2] = i;
}

// ---- registry of functions f (include/dfx.h, dfx_fn_id) ----------------------
struct RegistryFn {
  dfx_function f;
  int dim;
  __device__ __forceinline__ int ncomp() const { return f.n_comp; }
  __device__ void operator()(const double* x, double* y) const {
    switch (f.id) {
      case DFX_FN_CONSTANT:
        for (int c = 0; c < f.n_comp; ++c) y[c] = f.p[c];
        break;
      case DFX_FN_GAUSSIAN: {
        double s = 0;
        for (int j = 0; j < dim; ++j) s = __dadd_rn(s, __dmul_rn(x[j], x[j]));   // x.two_norm2()
        y[0] = exp(__dmul_rn(-f.p[0], s));
        break; }
      case DFX_FN_IDENTITY:
        for (int c = 0; c < f.n_comp; ++c) y[c] = c < dim ? x[c] : 0.0;
        break;
      case DFX_FN_COORD: y[0] = x[(int)f.p[0]]; break;
      case DFX_FN_QUADRATIC: {
        double xx[3] = {x[0], dim > 1 ? x[1] : 0.0, dim > 2 ? x[2] : 0.0};
        double r = f.p[0];
        for (int j = 0; j < 3; ++j) r = __dadd_rn(r, __dmul_rn(f.p[1 + j], xx[j]));
        int q = 4;
        for (int i = 0; i < 3; ++i)
          for (int j = i; j < 3; ++j) { r = __dadd_rn(r, __dmul_rn(__dmul_rn(f.p[q], xx[i]), xx[j])); ++q; }
        y[0] = r;
        break; }
      case DFX_FN_SINPROD: {
        double r = 1;
        for (int j = 0; j < dim; ++j) r = __dmul_rn(r, sin(__dmul_rn(f.p[0], x[j])));
        y[0] = r;
        break; }
      case DFX_FN_GAUSSIAN_VEC: {
        double s = 0;
        for (int j = 0; j < dim; ++j) s = __dadd_rn(s, __dmul_rn(x[j], x[j]));
        const double e = exp(__dmul_rn(-f.p[0], s));
        for (int c = 0; c < f.n_comp; ++c) y[c] = __dmul_rn(c < dim ? x[c] : 1.0, e);
        break; }
      default:
        for (int c = 0; c < f.n_comp; ++c) y[c] = 0.0;
    }
  }
};

}  // namespace dfx
