"""CPU: the oracle restatement against the UNMODIFIED reference compiled into
oracle/_ref (skipped when that build is absent).  Bar: bit-exact on every
Envi/Fluxes/Trackers field of every step."""
import numpy as np
import pytest

from conftest import Golden, golden_names
from sipnet_b200 import _abi as A, synth


@pytest.mark.parametrize("name", golden_names())
def test_oracle_vs_reference_full_series(oracle, refshim, name):
    g = Golden(name)
    rc, done, out, dbg = refshim.run(g.flags, g.params, g.site)
    rc2, done2, out2, dbg2, _ = oracle.run(g.flags, g.params, g.site)
    assert (rc, done) == (rc2, done2)
    assert np.array_equal(out[:done], out2[:done], equal_nan=True)
    assert np.array_equal(dbg[:done], dbg2[:done], equal_nan=True)


def test_param_enum_matches_reference_struct(refshim):
    for i, n in enumerate(A.PARAM_NAMES):
        assert refshim.param_offset(n) == 8 * i, n


@pytest.mark.parametrize("variant", ["half-daily", "unequal"])
def test_oracle_vs_reference_wide_prior_events(oracle, refshim, variant):
    site = synth.synth_site(11, 4, variant, with_events=True)
    P = synth.synth_params(12, stream=11)
    deaths = 0
    for m in range(P.shape[1]):
        rc, done, out, dbg = refshim.run(synth.SYNTH_FLAGS, P[:, m], site)
        rc2, done2, out2, dbg2, _ = oracle.run(synth.SYNTH_FLAGS, P[:, m], site)
        assert (rc, done) == (rc2, done2)
        assert np.array_equal(dbg[:done], dbg2[:done], equal_nan=True), f"member {m}"
        deaths += int((np.diff(dbg[:done, A.D["s.isAlive"]]) < 0).sum())
    assert deaths > 0


@pytest.mark.parametrize("flags", [
    dict(gdd=0, soilPhenol=1), dict(gdd=0), dict(flooding=1), dict(litterPool=1, carbonSaturation=1),
    dict(litterPool=1, anaerobic=1), dict(events=0), dict(growthResp=1, leafWater=1),
    dict(litterPool=1, anaerobic=1, nitrogenCycle=1, carbonSaturation=1, flooding=1, growthResp=1, leafWater=1),
])
def test_oracle_vs_reference_flag_matrix(oracle, refshim, flags):
    full = dict(A.DEFAULT_FLAGS)
    full.update(flags)
    site = synth.synth_site(5, 2, "half-daily", with_events=True, gdd_flag=full["gdd"])
    P = synth.synth_params(4, stream=5)
    P[A.P["soilCSaturation"], :] = 2700.0
    P[A.P["waterDrainFrac"], :] = 0.5
    for m in range(P.shape[1]):
        rc, done, out, dbg = refshim.run(full, P[:, m], site)
        rc2, done2, out2, dbg2, _ = oracle.run(full, P[:, m], site)
        assert (rc, done) == (rc2, done2)
        assert np.array_equal(dbg[:done], dbg2[:done], equal_nan=True)


def test_error_codes_match_reference(oracle, refshim):
    base = synth.synth_site(2, 1, "half-daily")
    p = synth.base_param_vector()
    fl = synth.SYNTH_FLAGS
    # event before first climate record -> 5 (frontend.c:217-222)
    s = synth.synth_site(2, 1, "half-daily"); s.events = [(2010, 5, A.EV_TILLAGE, 0, 0.1, 0, 0, 0)]
    assert refshim.run(fl, p, s)[0] == oracle.run(fl, p, s)[0] == 5
    # event on a day with no climate record -> 5 when a later day passes it (events.c:476-481)
    s = synth.synth_site(2, 1, "half-daily"); s.year = s.year.copy(); s.day = s.day.copy()
    keep = s.day != 100
    s2 = type(s)(s.year[keep], s.day[keep], {k: v[keep] for k, v in s.clim.items()},
                 [(2011, 100, A.EV_TILLAGE, 0, 0.1, 0, 0, 0)])
    r1, r2 = refshim.run(fl, p, s2), oracle.run(fl, p, s2)
    assert r1[0] == r2[0] == 5 and r1[1] == r2[1]
    # unknown irrigation method -> 4 (events.c:498-501)
    s = synth.synth_site(2, 1, "half-daily"); s.events = [(2011, 50, A.EV_IRRIGATION, 2, 1.0, 0, 0, 0)]
    r1, r2 = refshim.run(fl, p, s), oracle.run(fl, p, s)
    assert r1[0] == r2[0] == 4 and r1[1] == r2[1]
    # non-positive step length -> 3 (events.c:460-465)
    s = synth.synth_site(2, 1, "half-daily"); s.clim["length"] = s.clim["length"].copy(); s.clim["length"][10] = 0.0
    r1, r2 = refshim.run(fl, p, s), oracle.run(fl, p, s)
    assert r1[0] == r2[0] == 3 and r1[1] == r2[1] == 10
    # allocation parameters must sum below one -> 3 (sipnet.c:1117-1122)
    pb = p.copy(); pb[A.P["leafAllocation"]] = 0.7
    assert refshim.run(fl, pb, base)[0] == oracle.run(fl, pb, base)[0] == 3


def test_oracle_vs_reference_random_flag_combinations(oracle, refshim):
    """30 random valid flag combinations x 6 wide-prior members x 2 years: every debug field of every step, bit for bit."""
    from gpu_util import random_flag_cases
    for trial, f, site, P in random_flag_cases():
        for m in range(P.shape[1]):
            rc, done, out, dbg = refshim.run(f, P[:, m], site)
            rc2, done2, out2, dbg2, _ = oracle.run(f, P[:, m], site)
            assert (rc, done) == (rc2, done2), (trial, m, f)
            assert np.array_equal(dbg[:done], dbg2[:done], equal_nan=True), (trial, m, f)
            assert np.array_equal(out[:done], out2[:done], equal_nan=True), (trial, m, f)
