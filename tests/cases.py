"""Shared test cases: grids and basis trees, named like the reference's tests."""
from dune_functions_b200 import (YaspGrid, lagrange, lagrangeDG, power, composite, taylorHood,
                                 flatLexicographic as FL, flatInterleaved as FI,
                                 blockedLexicographic as BL, blockedInterleaved as BI)

GRIDS = {
    "1d_5": YaspGrid([5]),
    "2d_1x1": YaspGrid([1, 1]),
    "2d_3x2": YaspGrid([3, 2]),
    "2d_10x10": YaspGrid([10, 10]),
    "3d_1x1x1": YaspGrid([1, 1, 1]),
    "3d_3x4x5": YaspGrid([3, 4, 5]),
    "3d_2x1x3": YaspGrid([2, 1, 3], upper=[2.0, 0.5, 1.5], lower=[-1.0, 0.25, 0.0]),
}


def th_variant(outer, inner, dim):
    """the eight Taylor-Hood variants of the manual's table (dune-functions-bases.tex:1101-1292)"""
    return composite(power(lagrange(2), dim, inner()), lagrange(1), outer())


def trees(dim):
    t = {
        "lagrange0": lagrange(0),
        "lagrange1": lagrange(1),
        "lagrange2": lagrange(2),
        "lagrange3": lagrange(3),
        "lagrange4": lagrange(4),
        "lagrangeDG0": lagrangeDG(0),
        "lagrangeDG1": lagrangeDG(1),
        "lagrangeDG2": lagrangeDG(2),
        "lagrangeDG4": lagrangeDG(4),
        "taylorHood": taylorHood(),
        "power3_q2_BI": power(lagrange(2), 3, BI()),
        "power2_q1_FI": power(lagrange(1), 2, FI()),
        "power2_q2_BL": power(lagrange(2), 2, BL()),
        "power3_dg1_FL": power(lagrangeDG(1), 3, FL()),
        # makebasistest.cc:65-153 style nestings
        "nested_BL_BI_FL": composite(power(power(lagrange(1), 2, FL()), 2, BI()), lagrange(2), BL()),
        "nested_FL_FI_BL": composite(power(lagrange(2), 2, FI()), power(lagrangeDG(1), 2, BL()), lagrange(1), FL()),
        "nested_pow_of_comp": power(composite(lagrange(2), lagrange(1), FL()), 2, FI()),
        "nested_pow_of_comp_BL": power(composite(lagrange(2), lagrange(1), BL()), 2, BL()),
        "nested_BI_of_comp_BL": power(composite(lagrange(1), lagrange(1), BL()), 3, BI()),
    }
    for on, o in (("BL", BL), ("FL", FL)):
        for inn, i in (("BL", BL), ("BI", BI), ("FL", FL), ("FI", FI)):
            t[f"TH_{on}({inn})"] = th_variant(o, i, dim)
    return t
