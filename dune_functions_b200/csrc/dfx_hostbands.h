// dfx_hostbands.h — host-only interval bookkeeping of the pipelined dfx_evaluate_host: which
// coefficient ranges a band of z-layers gathers, and what is still missing on the device.
// Plain C++ (no CUDA), shared with tests/harness so that it is checked against the oracle on CPU.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

#include "dfx_internal.h"

namespace dfx {

// Flat-position ranges [begin, end) of the coefficients that the elements of the z-layers
// [zlo, zhi) (z = last grid direction) gather, merged into disjoint sorted intervals.  Every
// index-set component is lexicographic with z slowest, so a band of layers is one index range
// per (leaf, component); interleaved leaves (posA > 1) contribute the hull of their stride.
typedef std::vector<std::pair<u64, u64>> Intervals;
inline void normalise(Intervals& v) {
  std::sort(v.begin(), v.end());
  Intervals out;
  for (auto& r : v) {
    if (r.first >= r.second) continue;
    if (!out.empty() && r.first <= out.back().second) out.back().second = std::max(out.back().second, r.second);
    else out.push_back(r);
  }
  v.swap(out);
}
inline Intervals layerBandRanges(const dfx_basis_impl* b, int fl, int nl, u64 zlo, u64 zhi) {
  const DevBasis& D = b->dev;
  const int dz = D.grid.dim - 1;
  Intervals out;
  for (int l = fl; l < fl + nl; ++l) {
    const DevLeaf& leaf = D.leaves[l];
    const DevLayout& L = D.layouts[leaf.layout];
    for (int q = 0; q < L.ncomp; ++q) {
      const int s = L.order[q];
      const u64 plane = L.compCount[s] / (u64)L.csize[s][dz];
      const u64 top = ((s >> dz) & 1) ? zhi : zhi + 1;               // pinned components own the plane above too
      const u64 j0 = L.compOffset[s] + plane * zlo, j1 = L.compOffset[s] + plane * top;
      if (j1 > j0) out.push_back({leaf.posA * j0 + leaf.posB, leaf.posA * (j1 - 1) + leaf.posB + 1});
    }
  }
  normalise(out);
  return out;
}
// a minus b (both normalised)
inline Intervals subtract(const Intervals& a, const Intervals& b) {
  Intervals out;
  size_t k = 0;
  for (auto r : a) {
    while (k < b.size() && b[k].second <= r.first) ++k;
    size_t m = k;
    u64 cur = r.first;
    while (m < b.size() && b[m].first < r.second) {
      if (b[m].first > cur) out.push_back({cur, b[m].first});
      cur = std::max(cur, b[m].second);
      ++m;
    }
    if (cur < r.second) out.push_back({cur, r.second});
  }
  return out;
}

}  // namespace dfx
