"""The C++ façade (include/dfx/dune_functions.hh) exercised by test programs written like the
reference's own tests/examples (tests/cpp/*.cc); they link libdfx.so and need a GPU."""
import os
import subprocess

import pytest

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["basis_test", "interpolation_example", "user_functor", "poisson_pq2", "stokes_taylorhood"])
def test_facade_program(name):
    exes = {os.path.basename(e): e for e in entry.build_facade_tests()}
    out = subprocess.run([exes[name]], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0
    assert f"{name} passed" in out.stdout
