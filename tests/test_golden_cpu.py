"""CPU suite: the oracle against the golden fixtures committed under tests/golden/ (generated from
the reference checkout by tests/golden/make_golden.py):

  th_indexing_variants.json  literal copy of the manual's table of the eight Taylor-Hood variants
                             (doc/manual/dune-functions-bases.tex, Table tab:th_indexing_variants)
  merging_strategies.json    literal copy of the doc-comment tables of basistags.hh:52-184
  numbering_examples.json    SURVEY.md Appendix A (documented dune-grid assumptions, not reference output)

The GPU suite (tests/test_gpu_golden.py) checks the CUDA path against the same files."""
import json
import os
import re

import numpy as np
import pytest

import orc
from cases import th_variant
from dune_functions_b200 import (YaspGrid, lagrange, power, composite,
                                 flatLexicographic as FL, flatInterleaved as FI,
                                 blockedLexicographic as BL, blockedInterleaved as BI)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAGS = {"BL": BL, "BI": BI, "FL": FL, "FI": FI}


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def eval_entry(expr, n2):
    """'(0,2n_2+1)' -> (0, 2*n2+1)"""
    e = re.sub(r"(\d)n_2", r"\1*n2", expr).replace("n_2", "n2")
    assert re.fullmatch(r"[\d\s,()+*n]+", e), expr
    v = eval(e, {"__builtins__": {}}, {"n2": n2})
    return tuple(v) if isinstance(v, tuple) else (v,)


def th_expected(column, n2):
    """{(kind, component, index): multi-index} for the rows the manual prints"""
    out = {}
    for key, expr in load("th_indexing_variants.json")["columns"][column].items():
        if key.startswith("v"):
            c, i = key[1:].split(",")
            out[("v", int(c), int(i))] = eval_entry(expr, n2)
        else:
            out[("p", 0, int(key[1:]))] = eval_entry(expr, n2)
    return out


def th_lookup(indices_fn, q2_idx, q1_idx):
    """{(kind, component, scalar index): multi-index} of a Taylor-Hood basis with local layout
    [v0 (27) | v1 (27) | v2 (27) | p (8)] (nodes.hh:236-245)"""
    dig, lens = indices_fn()
    got = {}
    for e in range(dig.shape[0]):
        for c in range(3):
            for i in range(27):
                li = c * 27 + i
                got[("v", c, int(q2_idx[e, i]))] = tuple(int(x) for x in dig[e, li, :lens[e, li]])
        for j in range(8):
            li = 81 + j
            got[("p", 0, int(q1_idx[e, j]))] = tuple(int(x) for x in dig[e, li, :lens[e, li]])
    return got


def check_th_column(column, make_indices):
    """make_indices(grid, tree) -> (digits, lens) of the implementation under test; the Q2 / Q1
    scalar numberings used to look rows up come from the same implementation"""
    outer, inner = column[:2], column[3:5]
    grid = YaspGrid([2, 3, 2])
    q2 = make_indices(grid, lagrange(2))[0][:, :, 0]
    q1 = make_indices(grid, lagrange(1))[0][:, :, 0]
    n2 = int(q2.max()) + 1
    assert n2 == 5 * 7 * 5
    got = th_lookup(lambda: make_indices(grid, th_variant(TAGS[outer], TAGS[inner], 3)), q2, q1)
    exp = th_expected(column, n2)
    assert len(exp) == 15
    for key, mi in exp.items():
        assert got[key] == mi, (column, key)


COLUMNS = ["BL(BL)", "BL(BI)", "BL(FL)", "BL(FI)", "FL(BL)", "FL(BI)", "FL(FL)", "FL(FI)"]


@pytest.mark.parametrize("column", COLUMNS)
def test_oracle_reproduces_manual_table(column):
    check_th_column(column, lambda grid, tree: orc.OracleBasis(grid, tree).indices())


# ---- basistags.hh doc tables -----------------------------------------------------------------
def strategy_cases():
    """per strategy: (tree of {f,g} merged, standalone tree of f, standalone tree of g) whose
    first digits / lengths are exactly those the doc table uses"""
    q1 = lagrange(1)
    return {
        # f first digits {0,1}, g first digits {0,1,2}
        "FlatLexicographic": (composite(power(q1, 2, BL()), power(q1, 3, BL()), FL()), power(q1, 2, BL()), power(q1, 3, BL())),
        # identical children with first digits {0,1,2}
        "FlatInterleaved": (power(power(q1, 3, BL()), 2, FI()), power(q1, 3, BL()), power(q1, 3, BL())),
        "BlockedLexicographic": (composite(q1, lagrange(2), BL()), q1, lagrange(2)),
        "BlockedInterleaved": (power(q1, 2, BI()), q1, q1),
    }


def merged_rule(name):
    """from the literal doc table: child ('f'/'g') -> function(original multi-index) -> merged"""
    t = load("merging_strategies.json")["tables"][name]
    orig = {lab: idx for lab, idx in t["f"] + t["g"]}
    rules = {"f": [], "g": []}
    for lab, merged in t["merged"]:
        rules[lab[0]].append((orig[lab], merged))
    def apply(child, mi):
        for o, m in rules[child]:
            # the symbolic entry (i0, k1, ...) stands for the remaining digits
            sym = [x for x in o if not x.isdigit()]
            assert len(sym) == 1
            lit_o = [int(x) for x in o if x.isdigit()]
            pos_o = o.index(sym[0])
            if pos_o == 0:                       # (i0) style: whole index is symbolic
                if lit_o:
                    continue
                tail = list(mi)
                return tuple(int(x) if x.isdigit() else None for x in m), tail, m.index(sym[0])
            if list(mi[:pos_o]) == lit_o:        # (j, i_j) style: leading literal digits must match
                return tuple(int(x) if x.isdigit() else None for x in m), list(mi[pos_o:]), m.index(sym[0])
        raise KeyError((child, mi))
    def transform(child, mi):
        pattern, tail, at = apply(child, mi)
        out = []
        for p, x in enumerate(pattern):
            if p == at:
                out.extend(tail)
            else:
                out.append(x)
        return tuple(out)
    return transform


def mi_rows(dig, lens):
    return [[tuple(int(x) for x in dig[e, i, :lens[e, i]]) for i in range(dig.shape[1])] for e in range(dig.shape[0])]


def check_strategy(name, make_indices):
    merged_tree, f_tree, g_tree = strategy_cases()[name]
    grid = YaspGrid([2, 2])
    rule = merged_rule(name)
    fm = mi_rows(*make_indices(grid, f_tree))
    gm = mi_rows(*make_indices(grid, g_tree))
    mm = mi_rows(*make_indices(grid, merged_tree))
    for e in range(len(mm)):
        nf = len(fm[e])
        assert len(mm[e]) == nf + len(gm[e])          # children occupy consecutive local blocks
        for i, mi in enumerate(fm[e]):
            assert mm[e][i] == rule("f", mi), (name, "f", mi)
        for i, mi in enumerate(gm[e]):
            assert mm[e][nf + i] == rule("g", mi), (name, "g", mi)


@pytest.mark.parametrize("name", ["FlatLexicographic", "FlatInterleaved", "BlockedLexicographic", "BlockedInterleaved"])
def test_oracle_reproduces_basistags_tables(name):
    check_strategy(name, lambda grid, tree: orc.OracleBasis(grid, tree).indices())


# ---- Appendix A ---------------------------------------------------------------------------------
def numbering_cases():
    d = load("numbering_examples.json")
    return {k: v for k, v in d.items() if k != "source"}


def tree_of(name, case):
    return th_variant(FL, FL, 2) if "th_flfl" in name else lagrange(case["order"])


@pytest.mark.parametrize("name", list(numbering_cases()))
def test_oracle_reproduces_numbering_examples(name):
    case = numbering_cases()[name]
    b = orc.OracleBasis(YaspGrid(case["grid"]), tree_of(name, case))
    assert b.dimension() == case["dimension"]
    assert b.indices()[0][:, :, 0].tolist() == case["element_indices"]


def test_fixtures_are_in_sync_with_the_reference_when_it_is_present():
    """regenerating from /root/reference (build container only) gives the committed files"""
    if not os.path.isdir("/root/reference/doc/manual"):
        pytest.skip("no reference checkout on this machine")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    assert mg.parse_th_table() == load("th_indexing_variants.json")
    assert json.loads(json.dumps(mg.parse_merging_tables())) == load("merging_strategies.json")
    assert json.loads(json.dumps(mg.parse_face_orientation_tables())) == load("face_orientations.json")
