"""GPU suite: the CUDA index path (through the C ABI) against the golden fixtures of
tests/golden/ — the manual's Taylor-Hood table, the basistags.hh merging tables and the
Appendix-A numbering vectors — without going through the oracle."""
import numpy as np
import pytest

import dune_functions_b200 as dfx
from dune_functions_b200 import YaspGrid
from test_golden_cpu import COLUMNS, check_th_column, check_strategy, numbering_cases, tree_of

pytestmark = pytest.mark.gpu


def device_indices(grid, tree):
    b = dfx.makeBasis(grid, tree)
    import torch
    dig = b.indices(dtype=torch.int64).cpu().numpy().astype(np.uint64)
    lens = np.broadcast_to(b.localIndexSizes()[None, :], dig.shape[:2])
    return dig, lens


@pytest.mark.parametrize("column", COLUMNS)
def test_device_reproduces_manual_table(column):
    check_th_column(column, device_indices)


@pytest.mark.parametrize("name", ["FlatLexicographic", "FlatInterleaved", "BlockedLexicographic", "BlockedInterleaved"])
def test_device_reproduces_basistags_tables(name):
    check_strategy(name, device_indices)


@pytest.mark.parametrize("name", list(numbering_cases()))
def test_device_reproduces_numbering_examples(name):
    case = numbering_cases()[name]
    b = dfx.makeBasis(YaspGrid(case["grid"]), tree_of(name, case))
    assert b.dimension() == case["dimension"]
    assert b.indices().cpu().numpy()[:, :, 0].tolist() == case["element_indices"]
