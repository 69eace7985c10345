"""GPU: the reference's own unit tests re-stated as known-answer tests, through the CUDA kernel (tests/kat_cases.py)."""
import numpy as np
import pytest

import kat_cases as K
from sipnet_b200 import _abi as A, api

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("math", [A.MATH_FAST, A.MATH_VALIDATION])
@pytest.mark.parametrize("name", sorted(K.CASES))
def test_kernel_reproduces_reference_unit_test(oracle, name, math):
    flags, params, site, want = K.build(name)
    with api.Ensemble([site], params.reshape(-1, 1), None, flags, outputs=A.OUT_FULL, math=math) as ens:
        ens.run()
        out = ens.output()[:, 0, 0]
    for col, val in want.items():
        assert abs(out[A.O[col]] - val) < TOL, f"{col} = {out[A.O[col]]!r}, the reference's unit test expects {val!r}"
    _, _, o_out, _, _ = oracle.run(flags, params, site, want_debug=False)
    assert np.array_equal(out, o_out[0])                      # and the same bits as the oracle


@pytest.mark.parametrize("math", [A.MATH_FAST, A.MATH_VALIDATION])
def test_tillage_decay_known_answer(math):
    flags, params, site, want = K.tillage_case()
    with api.Ensemble([site], params.reshape(-1, 1), None, flags, outputs=A.OUT_FULL, math=math) as ens:
        for t in range(site.nsteps):                          # d_till_mod after every step (state row)
            ens.run(t, t + 1)
            got = ens.state()[A.S["dTillMod"], 0]
            assert abs(got - want[t]) < 1e-12, (t, got, want[t])
