// host_index_harness.cpp — TEST INFRASTRUCTURE.  Compiles the product's host-side
// basis code (dune_functions_b200/csrc/dfx_basis.cu, pure host C++) with g++ and walks
// the per-leaf affine descriptors + layoutIndex() on the CPU exactly like k_indices
// does on the device, so the flattening logic is checked against the oracle without a
// GPU.  Not part of libdfx.so and never used by the product.
#include "../../dune_functions_b200/csrc/dfx_internal.h"

using namespace dfx;

extern "C" int harness_indices(const dfx_grid_desc* g, const dfx_tree_node* t, int n, unsigned long long* digits,
                               int stride, unsigned long long* flat, unsigned long long* dimension) {
  dfx_basis_impl* impl = nullptr;
  int rc = buildBasis(g, t, n, 0, &impl);
  if (rc) return rc;
  const DevBasis& B = impl->dev;
  const u64 ne = (u64)B.grid.N[0] * B.grid.N[1] * B.grid.N[2];
  *dimension = impl->dimension;
  for (u64 e = 0; e < ne; ++e) {
    int ec[3] = {(int)(e % B.grid.N[0]), (int)((e / B.grid.N[0]) % B.grid.N[1]), (int)(e / ((u64)B.grid.N[0] * B.grid.N[1]))};
    for (int li = 0; li < B.nlocTotal; ++li) {
      int l = 0;
      while (l + 1 < B.nLeaves && li >= B.leaves[l + 1].localOffset) ++l;
      const DevLeaf& leaf = B.leaves[l];
      const DevLayout& L = B.layouts[leaf.layout];
      int i = li - leaf.localOffset, a[3], ii = i;
      a[0] = ii % (L.k + 1); ii /= (L.k + 1); a[1] = ii % (L.k + 1); ii /= (L.k + 1); a[2] = ii;
      u64 j = layoutIndex(L, B.grid, ec, a, i);
      u64 o = e * B.nlocTotal + li;
      for (int p = 0; p < stride; ++p) digits[o * stride + p] = p < leaf.ndig ? leaf.digA[p] * j + leaf.digB[p] : 0;
      flat[o] = leaf.posA * j + leaf.posB;
    }
  }
  delete static_cast<dfx_basis*>(impl);
  return 0;
}

// The band bookkeeping of the pipelined dfx_evaluate_host (csrc/dfx_hostbands.h): for the element
// bands [bounds[i], bounds[i+1]) of the whole tree, which flat coefficient positions does band i
// newly bring to the device?  out_band[p] = index of the band whose upload contains position p
// (-1 = never uploaded).  The test checks with the oracle's index table that every band finds what
// its elements gather.
#include "../../dune_functions_b200/csrc/dfx_hostbands.h"

extern "C" int harness_band_uploads(const dfx_grid_desc* g, const dfx_tree_node* t, int n, const unsigned long long* bounds,
                                    int nBands, int* out_band) {
  dfx_basis_impl* impl = nullptr;
  int rc = buildBasis(g, t, n, 0, &impl);
  if (rc) return rc;
  const DevBasis& B = impl->dev;
  u64 layer = 1;
  for (int j = 0; j + 1 < B.grid.dim; ++j) layer *= (u64)B.grid.N[j];
  for (u64 p = 0; p < impl->dimension; ++p) out_band[p] = -1;
  Intervals have;
  for (int i = 0; i < nBands; ++i) {
    const u64 c0 = bounds[i], c1 = bounds[i + 1];
    if (c1 <= c0) continue;
    Intervals need = subtract(layerBandRanges(impl, 0, B.nLeaves, c0 / layer, (c1 - 1) / layer + 1), have);
    for (auto& r : need) {
      for (u64 p = r.first; p < r.second; ++p) {
        if (p >= impl->dimension || out_band[p] != -1) { delete static_cast<dfx_basis*>(impl); return 100; }   // out of range / uploaded twice
        out_band[p] = i;
      }
      have.push_back(r);
    }
    normalise(have);
  }
  delete static_cast<dfx_basis*>(impl);
  return 0;
}
