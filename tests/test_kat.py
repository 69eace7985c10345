"""The reference's own unit tests re-stated as known-answer tests over one complete step (tests/kat_cases.py) -- oracle and, when built,
the live reference on the CPU; the CUDA kernel in tests/test_gpu_kat.py."""
import numpy as np
import pytest

import kat_cases as K
from sipnet_b200 import _abi as A

TOL = 1e-9


def check(out_row, want, who):
    for col, val in want.items():
        got = out_row[A.O[col]]
        assert abs(got - val) < TOL, f"{who}: {col} = {got!r}, the reference's unit test expects {val!r}"


@pytest.mark.parametrize("name", sorted(K.CASES))
def test_oracle_reproduces_reference_unit_test(oracle, name):
    flags, params, site, want = K.build(name)
    rc, done, out, _, _ = oracle.run(flags, params, site, want_debug=False)
    assert rc == 0 and done == 1
    check(out[0], want, "oracle")


@pytest.mark.parametrize("name", sorted(K.CASES))
def test_live_reference_agrees_bit_for_bit(oracle, refshim, name):
    flags, params, site, want = K.build(name)
    rc, done, out, dbg = refshim.run(flags, params, site)
    assert rc == 0 and done == 1
    check(out[0], want, "reference")
    _, _, o_out, o_dbg, _ = oracle.run(flags, params, site)
    assert np.array_equal(out, o_out) and np.array_equal(dbg, o_dbg)
