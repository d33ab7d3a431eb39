"""CPU suite: the host-bookkeeping parts of the C++ façade (tests/cpp/host_only_test.cc — subspaceBasis of a
SubspaceBasis, containerDescriptor checkSize for the basis list of containerdescriptortest.cc,
istlVectorBackend(v).resize(basis) on nested containers, basis.update, error behaviour) and the same
sub-space rules through the Python front door.  No device call is made."""
import os
import subprocess

import pytest

import __graft_entry__ as entry
from dune_functions_b200 import (YaspGrid, GlobalBasis, lagrange, power, composite, subspaceBasis, SubspaceBasis,
                                 blockedInterleaved as BI, blockedLexicographic as BL)
from dune_functions_b200 import _capi as capi


def test_host_only_facade_program():
    exes = {os.path.basename(e): e for e in entry.build_facade_tests()}
    out = subprocess.run([exes["host_only_test"]], capture_output=True, text=True, timeout=300)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0
    assert "host_only_test passed" in out.stdout


def test_subspace_of_a_subspace_joins_the_tree_paths():
    """subspacebasistest.cc: subspaceBasis(subspaceBasis(b, tp1), tp2) equals subspaceBasis(b, (tp1, tp2))"""
    basis = GlobalBasis(YaspGrid([2, 2]), composite(power(lagrange(3), 3, BI()), lagrange(1), BL()))
    sb0 = subspaceBasis(basis, 0)
    a, b = subspaceBasis(sb0, 1), subspaceBasis(basis, 0, 1)
    assert a.rootBasis() is basis and b.rootBasis() is basis
    assert a.prefixPath() == b.prefixPath() == (0, 1) and a.node == b.node
    assert SubspaceBasis(SubspaceBasis(basis, 0), 1).node == b.node
    assert subspaceBasis(basis, 0, 2).node != b.node and sb0.node != b.node
    assert a.dimension() == basis.dimension()
    with pytest.raises(capi.RangeError):
        subspaceBasis(sb0, 3)
