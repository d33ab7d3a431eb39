"""basis.containerDescriptor() (defaultglobalbasis.hh:180-186; containerdescriptors.hh) — host-only logic of
the library, checked on the CPU the way functionspacebases/test/containerdescriptortest.cc does:

  checkSize          (:87-103)  the descriptor's size below every prefix == basis.size(prefix)
  checkMultiIndices  (:105-131) every multi-index of every element walks through the descriptor

with size(prefix) and the multi-indices taken from the ORACLE (another implementation), for the basis list of
that test's main() (:236-300) plus the Taylor-Hood variants, and the descriptor types the reference spells
out (taylorhoodbasis.hh:204-216; the `stokes` example of the test, :133-160)."""
import numpy as np
import pytest

import orc
from dune_functions_b200 import _capi as capi
from dune_functions_b200 import (YaspGrid, GlobalBasis, lagrange, lagrangeDG, power, composite, taylorHood,
                                 subspaceBasis,
                                 flatLexicographic as FL, flatInterleaved as FI,
                                 blockedLexicographic as BL, blockedInterleaved as BI)

Q1, Q2 = lagrange(1), lagrange(2)


def reference_test_bases():
    out = {"lagrange<1>": Q1, "power<2>(Q1,BI)": power(Q1, 2, BI()), "composite(Q1,Q2,BL)": composite(Q1, Q2, BL())}
    names = {FL(): "FL", FI(): "FI", BL(): "BL", BI(): "BI"}
    for outer in (FL(), FI(), BL(), BI()):
        for inner in (FL(), FI(), BL(), BI()):
            out[f"power<2>(power<3>(Q2,{names[inner]}),{names[outer]})"] = power(power(Q2, 3, inner), 2, outer)
    out["complicated_1"] = power(composite(power(power(Q1, 1, BI()), 1, BL()), power(Q1, 2, BI()), power(Q1, 3, BL()), BL()), 2, BL())
    out["complicated_2"] = composite(
        composite(power(Q1, 2, BI()), power(power(Q1, 2, BL()), 1, BL()), BL()),
        composite(power(power(Q1, 1, BL()), 1, BI()), power(Q1, 2, BI()), power(Q1, 3, BL()), BL()), BL())
    out["complicated_flat"] = composite(
        composite(power(Q1, 2, FI()), power(power(Q1, 2, FL()), 1, FL()), FL()),
        composite(power(power(Q1, 1, FL()), 1, FI()), power(Q1, 2, FI()), power(Q1, 3, FL()), FL()), FL())
    return out


def check_size(cd, ob, prefix, max_len):
    """checkSize of the reference test: size(prefix) == size of the descriptor reached through prefix"""
    assert ob.size(prefix) == cd.size(), prefix
    if len(prefix) < max_len:
        # uniform kinds share one child: one representative per distinct child is enough, but walk a few
        idx = range(cd.size()) if cd.size() <= 8 else (0, 1, cd.size() // 2, cd.size() - 1)
        for i in idx:
            check_size(cd[i], ob, prefix + (i,), max_len)


def check_multi_indices(cd, ob):
    dig, lens = ob.indices()
    for row, ln in zip(dig.reshape(-1, dig.shape[-1]), lens.reshape(-1)):
        d = cd
        for j in range(ln):
            assert row[j] < d.size()
            d = d[int(row[j])]
        assert d.kind == capi.CD_VALUE                       # a full multi-index addresses a scalar


@pytest.mark.parametrize("name", list(reference_test_bases()))
def test_descriptor_of_the_reference_test_bases(name):
    tree = reference_test_bases()[name]
    grid = YaspGrid([2, 2])
    b = GlobalBasis(grid, tree)
    ob = orc.OracleBasis(grid, tree)
    cd = b.containerDescriptor()
    assert cd.kind != capi.CD_UNKNOWN, cd
    check_size(cd, ob, (), ob.maxMultiIndexSize())
    check_multi_indices(cd, ob)
    # the library's own size(prefix) agrees too
    assert b.size(()) == cd.size()


def test_known_descriptor_types():
    g3, g2 = YaspGrid([2, 3, 2]), YaspGrid([2, 2])
    n2, n1 = 5 * 7 * 5, 3 * 4 * 3
    # TaylorHoodPreBasis<GV, HI = false>: Array<FlatVector,2>{dim * n2, n1}  (taylorhoodbasis.hh:212-215)
    cd = GlobalBasis(g3, taylorHood()).containerDescriptor()
    assert cd.type() == "Array<UniformVector<Value>,2>"
    assert (cd[0].size(), cd[1].size()) == (3 * n2, n1) and cd[0].isFlatVector()
    # the HI = true form, which is what composite(power<dim>(Q2, BI), Q1, BL) produces (:207-211)
    cd = GlobalBasis(g3, composite(power(Q2, 3, BI()), Q1, BL())).containerDescriptor()
    assert cd.type() == "Tuple<UniformVector<UniformArray<Value,3>>,UniformVector<Value>>"
    assert cd[0].size() == n2 and cd[0][17].size() == 3 and cd[1].size() == n1
    # the `stokes` descriptor of containerdescriptortest.cc:135-137
    cd = GlobalBasis(g3, composite(power(Q2, 3, BL()), Q1, BL())).containerDescriptor()
    assert cd.type() == "Tuple<UniformArray<UniformVector<Value>,3>,UniformVector<Value>>"
    assert [cd[0][c].size() for c in range(3)] == [n2] * 3 and cd[1].size() == n1
    # leaves and flat merges are FlatVectors (leafprebasismixin.hh:66-69; flatLexicographic :291-298)
    for tree, size in ((Q2, 25), (lagrangeDG(2), 36), (power(Q2, 2, FI()), 50), (power(Q2, 2, FL()), 50),
                       (composite(power(Q2, 2, FL()), Q1, FL()), 59)):
        cd = GlobalBasis(g2, tree).containerDescriptor()
        assert cd.isFlatVector() and cd.size() == size and cd.type() == "UniformVector<Value>"
    # two children of the same TYPE merge into an Array, not a Tuple (makeDescriptor, :129-141)
    assert GlobalBasis(g2, composite(Q1, Q2, BL())).containerDescriptor().type() == "Array<UniformVector<Value>,2>"
    # a flat composite over blocked children has no flat container: Unknown (:296-297)
    cd = GlobalBasis(g2, composite(power(Q1, 2, BL()), Q1, FL())).containerDescriptor()
    assert cd.kind == capi.CD_UNKNOWN and cd.type() == "Unknown"
    with pytest.raises(IndexError):
        cd[0]
    # a SubspaceBasis forwards to its root (subspacebasis.hh:115-118)
    root = GlobalBasis(g3, taylorHood())
    assert subspaceBasis(root, 0, 1).containerDescriptor().type() == root.containerDescriptor().type()


def test_descriptor_through_the_c_abi_respects_capacity():
    import ctypes as C
    b = GlobalBasis(YaspGrid([2, 2]), composite(power(Q2, 2, BI()), Q1, BL()))
    L = capi.lib()
    n = C.c_int32()
    buf = C.create_string_buffer(16)
    assert L.dfx_basis_container_descriptor(b._h, None, 0, C.byref(n), buf, 16) == 0
    assert n.value == 6 and buf.value == b"Tuple<UniformVe"           # truncated, terminated
    arr = (capi.ContainerNode * 2)()
    assert L.dfx_basis_container_descriptor(b._h, arr, 2, C.byref(n), None, 0) == 0
    assert n.value == 6 and arr[0].kind == capi.CD_TUPLE and arr[0].n_stored == 2 and arr[1].kind == capi.CD_UNIFORM_VECTOR
    assert L.dfx_basis_container_descriptor(b._h, None, 3, C.byref(n), None, 0) == capi.DFX_ERR_INVALID
