"""CPU: the plain-C oracle restatement against the committed golden vectors
(generated from the unmodified reference by tests/golden/make_golden.py).
Bar: bit-exact (the oracle performs the same IEEE operations with the same libm
as the reference's gcc -O0 build)."""
import numpy as np
import pytest

from conftest import Golden, golden_names


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden_bit_exact(oracle, name):
    g = Golden(name)
    rc, done, out, dbg, recs = oracle.run(g.flags, g.params, g.site, max_event_records=4096)
    assert rc == g.rc
    assert done == g.nsteps
    assert np.array_equal(out[g.rows], g.out32, equal_nan=True)
    assert np.array_equal(dbg[g.rows], g.dbg, equal_nan=True)
    # whole-series fingerprints (same summation order => bit-exact)
    assert np.array_equal(out[:done].sum(axis=0), g.colsum)
    assert np.array_equal(np.abs(out[:done]).max(axis=0), g.colabs)


def test_golden_covers_hard_branches():
    """The synthetic goldens reach what the smoke goldens miss."""
    from sipnet_b200 import _abi as A
    g = Golden("synth_halfdaily_m0")
    alive = g.dbg[:, A.D["s.isAlive"]]
    assert (alive == 0).any() and (alive == 1).any()          # mortality + re-emergence
    txt = g.events_out.decode()
    for word in ("plant", "harv", "till", "fert", "irrig", "leafon", "plantdeath"):
        assert word in txt
    assert "eventEvap=0.30" in txt                            # canopy irrigation split (events.c:488-493)
