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

// Face-DOF permutation (dfx_basis_set_vertex_ids): layoutIndexOriented / orientFaceDOF / dofOwnerNode of
// dfx_internal.h are host-device functions, so the exact code the generic kernels run is walked here on
// the CPU with the vertex ids in host memory.  flat[e][li] = what k_indices writes; the return value
// counts (a) DOFs t whose owner node does not map back to t through the oriented index (the node
// k_interpolate / k_points pick for t would belong to another DOF), (b) in-entity indices that the
// inverse permutation does not bring back.
extern "C" long long harness_oriented(const dfx_grid_desc* g, const dfx_tree_node* t, int n, const unsigned long long* vid,
                                      unsigned long long* flat) {
  dfx_basis_impl* impl = nullptr;
  int rc = buildBasis(g, t, n, 0, &impl);
  if (rc) return -rc;
  DevBasis B = impl->dev;
  B.vid = vid;
  const u64 ne = (u64)B.grid.N[0] * B.grid.N[1] * B.grid.N[2];
  for (u64 e = 0; e < ne; ++e) {
    int ec[3] = {(int)(e % B.grid.N[0]), (int)((e / B.grid.N[0]) % B.grid.N[1]), (int)(e / ((u64)B.grid.N[0] * B.grid.N[1]))};
    for (int li = 0; li < B.nlocTotal; ++li) {
      int l = 0;
      while (l + 1 < B.nLeaves && li >= B.leaves[l + 1].localOffset) ++l;
      const DevLeaf& leaf = B.leaves[l];
      const DevLayout& L = B.layouts[leaf.layout];
      int i = li - leaf.localOffset, a[3], ii = i;
      a[0] = ii % (L.k + 1); ii /= (L.k + 1); a[1] = ii % (L.k + 1); ii /= (L.k + 1); a[2] = ii;
      flat[e * B.nlocTotal + li] = leaf.posA * layoutIndexOriented(L, B.grid, B.vid, ec, a, i) + leaf.posB;
    }
  }
  long long bad = 0;
  for (int q = 0; q < B.nLayouts; ++q) {
    const DevLayout& L = B.layouts[q];
    for (u64 d = 0; d < L.dimension; ++d) {
      int ie[3], a[3];
      dofOwnerNode(B, L, d, ie, a);
      int ec[3], li = 0, stride = 1;
      bool inside = true;
      for (int j = 0; j < 3; ++j) {
        ec[j] = ie[j] - B.grid.off[j];
        if (j < B.grid.dim && (ec[j] < 0 || ec[j] >= B.grid.N[j])) inside = false;
        if (j < B.grid.dim) { li += a[j] * stride; stride *= L.k + 1; }
      }
      // an owner in the neighbouring slab sees the node at alpha = 0 of its first layer: the same entity
      // is alpha = k of our last layer
      for (int j = 0; j < B.grid.dim && !inside; ++j)
        if (ec[j] == B.grid.N[j] && a[j] == 0) { ec[j] -= 1; li += L.k * (j == 0 ? 1 : j == 1 ? (L.k + 1) : (L.k + 1) * (L.k + 1)); a[j] = L.k; }
      if (layoutIndexOriented(L, B.grid, B.vid, ec, a, li) != d) ++bad;
    }
    if (B.vid != nullptr && L.kind == DFX_NODE_LAGRANGE && L.k > 2 && B.grid.dim > 1)
      for (unsigned s = 1; s < (1u << B.grid.dim); ++s) {
        int c[3] = {0, 0, 0};
        for (c[2] = 0; c[2] < L.csize[s][2]; ++c[2])
          for (c[1] = 0; c[1] < L.csize[s][1]; ++c[1])
            for (c[0] = 0; c[0] < L.csize[s][0]; ++c[0])
              for (unsigned fr = 0; fr < (unsigned)L.blk[s]; ++fr) {
                const unsigned f = orientFaceDOF<false>(B.grid, B.vid, L.k, s, c, fr);
                if (f >= (unsigned)L.blk[s] || orientFaceDOF<true>(B.grid, B.vid, L.k, s, c, f) != fr) ++bad;
              }
      }
  }
  delete static_cast<dfx_basis*>(impl);
  return bad;
}
