"""GPU: parity of the CUDA path (through the C ABI) with the oracle.

Bars (north_star): event day/type application and pool-clamping branches EXACT;
all state pools and fluxes within 1e-10 relative (FP64, fmad disabled on the
validation build).  The tolerance is written in gpu_util.RTOL."""
import numpy as np
import pytest

from conftest import Golden, golden_names
from gpu_util import EXACT_DEBUG, RTOL, assert_close, debug_scales, out_scales
from sipnet_b200 import _abi as A, api, synth

pytestmark = pytest.mark.gpu

ALL = A.OUT_FULL | A.OUT_DEBUG | A.OUT_EVENTS


def run_gpu(sites, params, member_site, flags, outputs=ALL, math=A.MATH_VALIDATION, **kw):
    ens = api.Ensemble(sites, params, member_site, flags, outputs=outputs, math=math, max_event_records=4096
                       if outputs & A.OUT_EVENTS else 0, **kw)
    ens.run()
    res = dict(out=ens.output() if outputs & A.OUT_FULL else None,
               dbg=ens.debug() if outputs & A.OUT_DEBUG else None,
               status=ens.status(), state=ens.state(),
               recs=ens.event_records() if outputs & A.OUT_EVENTS else None)
    ens.close()
    return res


BITEXACT = {"checked": 0, "exact": 0}


def check_member(res, m, o_out, o_dbg, o_recs, rtol=RTOL, tag=""):
    T = o_out.shape[0]
    g_out = res["out"][:, :T, m].T
    g_dbg = res["dbg"][:, :T, m].T
    BITEXACT["checked"] += 1
    BITEXACT["exact"] += int(np.array_equal(g_out, o_out, equal_nan=True) and np.array_equal(g_dbg, o_dbg, equal_nan=True))
    for name in EXACT_DEBUG:
        k = A.D[name]
        assert np.array_equal(g_dbg[:, k], o_dbg[:, k]), f"{tag} {name} differs (branch decision)"
    e1 = assert_close(g_out, o_out, out_scales(o_out), A.OUT_NAMES, rtol, tag + " out")
    e2 = assert_close(g_dbg, o_dbg, debug_scales(o_dbg), A.DEBUG_NAMES, rtol, tag + " dbg")
    # zero / non-zero pattern of every pool is a clamp/branch outcome: exact
    pools = slice(0, 13)
    assert np.array_equal(g_dbg[:, pools] == 0, o_dbg[:, pools] == 0), f"{tag} pool clamp pattern differs"
    if o_recs is not None:
        g_recs = res["recs"][m]
        assert [(r.step, r.type, r.variant, r.nval) for r in g_recs] == \
               [(r.step, r.type, r.variant, r.nval) for r in o_recs], f"{tag} event rows differ"
        for gr, orr in zip(g_recs, o_recs):
            for k in range(orr.nval):
                a, b = gr.val[k], orr.val[k]
                assert abs(a - b) <= rtol * max(abs(a), abs(b), 1e-6), (tag, orr.type, k, a, b)
    return max(e1, e2)


@pytest.mark.parametrize("name", golden_names())
def test_golden_cases(oracle, name):
    """Reference fixtures (smoke cases + synthetic event cases) through the CUDA path."""
    g = Golden(name)
    P = np.repeat(g.params[:, None], 3, axis=1)
    res = run_gpu([g.site], P, None, g.flags)
    rc, done, o_out, o_dbg, o_recs = oracle.run(g.flags, g.params, g.site, max_event_records=4096)
    assert rc == 0 and done == g.nsteps
    worst = check_member(res, 0, o_out, o_dbg, o_recs, tag=name)
    # the committed reference rows themselves
    assert_close(res["out"][:, g.rows, 0].T, g.out32, out_scales(o_out), A.OUT_NAMES, RTOL, name + " golden out")
    assert_close(res["dbg"][:, g.rows, 0].T, g.dbg, debug_scales(o_dbg), A.DEBUG_NAMES, RTOL, name + " golden dbg")
    # identical members give identical results
    assert np.array_equal(res["out"][:, :, 0], res["out"][:, :, 2], equal_nan=True)
    assert (res["status"] & (A.ST_BAD_ALLOCATION | A.ST_RING_OVERFLOW | A.ST_NONFINITE)).max() == 0
    print(f"{name}: worst relative error {worst:.3e}")


@pytest.mark.parametrize("variant", ["half-daily", "unequal"])
def test_wide_prior_ensemble_10yr(oracle, variant):
    """C2-shaped: 1 site x 96 wide-prior members x 10 yr, every field of every step."""
    site = synth.synth_site(0, 10, variant)
    P = synth.synth_params(96)
    res = run_gpu([site], P, None, synth.SYNTH_FLAGS)
    worst = 0.0
    for m in range(P.shape[1]):
        rc, done, o_out, o_dbg, o_recs = oracle.run(synth.SYNTH_FLAGS, P[:, m], site, max_event_records=4096)
        assert rc == 0
        worst = max(worst, check_member(res, m, o_out, o_dbg, o_recs, tag=f"{variant} m{m}"))
    print(f"wide prior {variant}: worst relative error {worst:.3e}")


def test_multi_site_events_c3_shape(oracle):
    """C3-shaped: several sites x members with the agronomic schedule; ragged member counts."""
    sites, P, ms, flags = synth.config_c3(nsites=5, members_per_site=7, nyears=4)
    res = run_gpu(sites, P, ms, flags)
    for m in range(P.shape[1]):
        rc, done, o_out, o_dbg, o_recs = oracle.run(flags, P[:, m], sites[ms[m]], max_event_records=4096)
        assert rc == 0
        check_member(res, m, o_out, o_dbg, o_recs, tag=f"c3 m{m}")
    assert (res["status"] & A.ST_DIED).any()


@pytest.mark.parametrize("flags", [
    dict(gdd=0, soilPhenol=1), dict(gdd=0), dict(flooding=1), dict(litterPool=1, carbonSaturation=1),
    dict(litterPool=1, anaerobic=1), dict(events=0), dict(growthResp=1, leafWater=1),
    dict(litterPool=1, anaerobic=1, nitrogenCycle=1, carbonSaturation=1, flooding=1, growthResp=1, leafWater=1),
])
def test_flag_matrix(oracle, flags):
    full = dict(A.DEFAULT_FLAGS)
    full.update(flags)
    site = synth.synth_site(5, 2, "half-daily", with_events=True, gdd_flag=full["gdd"])
    P = synth.synth_params(4, stream=5)
    P[A.P["soilCSaturation"], :] = 2700.0
    P[A.P["waterDrainFrac"], :] = 0.5
    res = run_gpu([site], P, None, full)
    for m in range(P.shape[1]):
        rc, done, o_out, o_dbg, o_recs = oracle.run(full, P[:, m], site, max_event_records=4096)
        check_member(res, m, o_out, o_dbg, o_recs, tag=f"{flags} m{m}")


def test_specialised_kernels_equal_generic(oracle):
    """The compile-time-flag kernels (no DEBUG) produce the same bits as the
    runtime-flag DEBUG kernel on the outputState() columns."""
    for flags in (dict(A.DEFAULT_FLAGS), dict(synth.SYNTH_FLAGS)):
        site = synth.synth_site(7, 3, "half-daily", with_events=True)
        P = synth.synth_params(40, stream=7)
        a = run_gpu([site], P, None, flags, outputs=A.OUT_FULL)
        b = run_gpu([site], P, None, flags, outputs=A.OUT_FULL | A.OUT_DEBUG)
        assert np.array_equal(a["out"], b["out"], equal_nan=True)


@pytest.mark.parametrize("block", [32, 128])
def test_block_size_invariance(block):
    sites, P, ms, flags = synth.config_c3(nsites=3, members_per_site=45, nyears=2)
    a = run_gpu(sites, P, ms, flags, outputs=A.OUT_FULL, block_threads=block)
    b = run_gpu(sites, P, ms, flags, outputs=A.OUT_FULL, block_threads=32)
    assert np.array_equal(a["out"], b["out"], equal_nan=True)


def test_segmented_run_equals_continuous():
    """Chunked sipnet_gpu_run == one run, bit for bit (reference: testRestartMVP.c:253-297)."""
    site = synth.synth_site(9, 3, "unequal", with_events=True)
    P = synth.synth_params(33, stream=9)
    flags = synth.SYNTH_FLAGS
    whole = run_gpu([site], P, None, flags, outputs=A.OUT_FULL | A.OUT_DEBUG)
    ens = api.Ensemble([site], P, None, flags, outputs=A.OUT_FULL | A.OUT_DEBUG, out_steps_capacity=500)
    T = site.nsteps
    outs, dbgs = [], []
    for t0 in range(0, T, 500):
        ens.run(t0, min(T, t0 + 500))
        outs.append(ens.output())
        dbgs.append(ens.debug())
    state = ens.state()
    ens.close()
    assert np.array_equal(np.concatenate(outs, axis=1), whole["out"], equal_nan=True)
    assert np.array_equal(np.concatenate(dbgs, axis=1), whole["dbg"], equal_nan=True)
    assert np.array_equal(state, whole["state"], equal_nan=True)


def test_ragged_site_lengths(oracle):
    s0 = synth.synth_site(20, 2, "half-daily")
    s1 = synth.synth_site(21, 1, "half-daily")
    P = synth.synth_params(10, stream=20)
    ms = np.array([0] * 6 + [1] * 4, np.int32)
    res = run_gpu([s0, s1], P, ms, synth.SYNTH_FLAGS, outputs=A.OUT_FULL)
    for m in range(10):
        site = (s0, s1)[ms[m]]
        rc, done, o_out, _, _ = oracle.run(synth.SYNTH_FLAGS, P[:, m], site, want_debug=False)
        assert_close(res["out"][:, :site.nsteps, m].T, o_out, out_scales(o_out), A.OUT_NAMES, RTOL, f"ragged m{m}")
        assert np.isnan(res["out"][:, site.nsteps:, m]).all()


def test_init_error_codes_match_oracle(oracle):
    p = synth.base_param_vector()[:, None]
    fl = synth.SYNTH_FLAGS

    def gpu_rc(site, params=p):
        try:
            api.Ensemble([site], params, None, fl).close()
            return 0
        except api.SipnetGpuError as e:
            return e.code
    s = synth.synth_site(2, 1, "half-daily"); s.events = [(2010, 5, A.EV_TILLAGE, 0, 0.1, 0, 0, 0)]
    assert gpu_rc(s) == oracle.run(fl, p[:, 0], s)[0] == 5
    s = synth.synth_site(2, 1, "half-daily"); s.events = [(2011, 50, A.EV_IRRIGATION, 2, 1.0, 0, 0, 0)]
    assert gpu_rc(s) == oracle.run(fl, p[:, 0], s)[0] == 4
    s = synth.synth_site(2, 1, "half-daily"); s.clim["length"] = s.clim["length"].copy(); s.clim["length"][10] = 0.0
    assert gpu_rc(s) == oracle.run(fl, p[:, 0], s)[0] == 3
    base = synth.synth_site(2, 1, "half-daily")
    keep = base.day != 100
    s2 = api.SiteData(base.year[keep], base.day[keep], {k: v[keep] for k, v in base.clim.items()},
                      [(2011, 100, A.EV_TILLAGE, 0, 0.1, 0, 0, 0)])
    assert gpu_rc(s2) == oracle.run(fl, p[:, 0], s2)[0] == 5
    # bad allocation: per-member status bit instead of exit(3)
    pb = np.repeat(p, 3, axis=1); pb[A.P["leafAllocation"], 1] = 0.7
    ens = api.Ensemble([base], pb, None, fl)
    ens.run()
    st = ens.status()
    ens.close()
    assert st[1] & A.ST_BAD_ALLOCATION and not (st[0] & A.ST_BAD_ALLOCATION)
    assert oracle.run(fl, pb[:, 1], base)[0] == 3


@pytest.mark.parametrize("variant", ["half-daily", "unequal"])
def test_fast_kernel_is_bit_identical(oracle, variant):
    """The production (optimistic, branch-free) kernel performs the same IEEE operations as the
    general one: bit-identical to the oracle on a 10-year wide-prior ensemble with events, and no
    member needed the exact replay."""
    site = synth.synth_site(0, 10, variant, with_events=True)
    P = synth.synth_params(64)
    res = run_gpu([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL, math=A.MATH_FAST)
    for m in range(P.shape[1]):
        rc, done, o_out, _, _ = oracle.run(synth.SYNTH_FLAGS, P[:, m], site, want_debug=False)
        assert np.array_equal(res["out"][:, :, m].T, o_out, equal_nan=True), f"fast {variant} m{m}"
    assert not (res["status"] & A.ST_REPLAY).any()


@pytest.mark.parametrize("flags", [dict(A.DEFAULT_FLAGS), dict(synth.SYNTH_FLAGS),
                                   dict(A.DEFAULT_FLAGS, growthResp=1, leafWater=1, litterPool=1, waterHResp=0),
                                   dict(A.DEFAULT_FLAGS, gdd=0, soilPhenol=1, flooding=1)])
def test_fast_equals_validation_all_outputs(flags):
    sites, P, ms, _ = synth.config_c3(nsites=4, members_per_site=40, nyears=3)
    kw = dict(outputs=A.OUT_FULL | A.OUT_EVENTS)
    a = run_gpu(sites, P, ms, flags, math=A.MATH_FAST, **kw)
    b = run_gpu(sites, P, ms, flags, math=A.MATH_VALIDATION, **kw)
    assert np.array_equal(a["out"], b["out"], equal_nan=True)
    assert np.array_equal(a["state"], b["state"], equal_nan=True)
    assert np.array_equal(a["status"] & ~np.uint32(A.ST_REPLAY), b["status"])
    for ra, rb in zip(a["recs"], b["recs"]):
        assert [(r.step, r.type, r.variant, r.nval, tuple(r.val)) for r in ra] == \
               [(r.step, r.type, r.variant, r.nval, tuple(r.val)) for r in rb]


def test_fast_kernel_replays_members_outside_its_guards(oracle):
    """exp(x) with |x| >= 512 leaves the optimistic main path (glibc's specialcase branch): a huge
    canopy (lai ~ 2000) makes -attenuation*lai reach that range.  Those members must be flagged,
    re-run by the general kernel and still match the oracle bit for bit; others are untouched."""
    site = synth.synth_site(4, 2, "half-daily", with_events=True)
    P = synth.synth_params(48, stream=4)
    weird = [3, 17, 40, 29]
    for m in weird:
        P[A.P["laiInit"], m] = 2000.0
        P[A.P["leafTurnoverRate"], m] = 0.01
    # member 29 keeps its canopy for the whole run: it leaves the guards in EVERY segment of the segmented run below
    # (a member flagged in one segment must be replayed again in the next ones)
    P[A.P["leafTurnoverRate"], 29] = 0.0
    P[A.P["fracLeafFall"], 29] = 0.0
    res = run_gpu([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL, math=A.MATH_FAST)
    flagged = np.flatnonzero(res["status"] & A.ST_REPLAY)
    assert set(weird) <= set(flagged.tolist())
    for m in list(range(0, 48, 5)) + weird:
        rc, done, o_out, _, _ = oracle.run(synth.SYNTH_FLAGS, P[:, m], site, want_debug=False)
        assert np.array_equal(res["out"][:, :, m].T, o_out, equal_nan=True), f"member {m}"
    # segmented run with replay in the middle
    ens = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL, math=A.MATH_FAST, out_steps_capacity=400)
    outs = []
    for t0 in range(0, site.nsteps, 400):
        ens.run(t0, min(site.nsteps, t0 + 400))
        outs.append(ens.output())
    ens.close()
    seg = np.concatenate(outs, axis=1)
    assert np.array_equal(seg[:, :, 29], res["out"][:, :, 29], equal_nan=True), "member outside the guards in every segment"
    assert np.array_equal(seg, res["out"], equal_nan=True)


def test_zz_validation_build_is_bit_identical():
    """With the glibc-exact exp/pow (sip_libm.cuh) and -fmad=false the validation build performs the
    reference's IEEE operations one for one: every member checked above must have matched the oracle
    (itself bit-identical to the unmodified reference) in EVERY bit of every field of every step."""
    assert BITEXACT["checked"] > 0
    print(f"bit-identical members: {BITEXACT['exact']} of {BITEXACT['checked']}")
    assert BITEXACT["exact"] == BITEXACT["checked"]


def test_run_to_host_pipeline_equals_run_plus_gather():
    """sipnet_gpu_run_to_host (chunked, double-buffered, copy stream) delivers the same bytes."""
    sites, P, ms, flags = synth.config_c3(nsites=2, members_per_site=50, nyears=3)
    ref = run_gpu(sites, P, ms, flags, outputs=A.OUT_FULL, math=A.MATH_FAST)
    ens = api.Ensemble(sites, P, ms, flags, outputs=A.OUT_FULL, math=A.MATH_FAST)
    T = ens.max_steps
    for chunk in (0, 100, 777):
        ens.reset()
        dst = np.full((A.NOUT, T, P.shape[1]), -1.0)
        ens.run_to_host(dst, 0, T, chunk_steps=chunk)
        assert np.array_equal(dst, ref["out"], equal_nan=True), chunk
    # a sub-range after a plain run
    ens.reset()
    ens.run(0, 1000)
    dst = np.empty((A.NOUT, T - 1000, P.shape[1]))
    ens.run_to_host(dst, 1000, T, chunk_steps=300)
    assert np.array_equal(dst, ref["out"][:, 1000:], equal_nan=True)
    assert np.array_equal(ens.state(), ref["state"], equal_nan=True)
    ens.close()


def test_set_state_resumes_a_run_exactly():
    """sipnet_gpu_set_state (restartLoadCheckpoint semantics): state + ring gathered after the first part of a run
    and loaded into a FRESH handle continue bit-identically, for an ensemble with the full event schedule, in both
    ring layouts (compact and the reference's 250 slots), debug rows (yearly/total trackers) included."""
    site = synth.synth_site(4, 3, "unequal", with_events=True)
    P = synth.synth_params(96, stream=9)
    T, cut = site.nsteps, 1111
    for slots, outputs, math in ((0, A.OUT_FULL, A.MATH_FAST), (A.RING_SLOTS_REFERENCE, A.OUT_FULL | A.OUT_DEBUG, A.MATH_VALIDATION)):
        kw = dict(outputs=outputs | A.OUT_EVENTS, math=math, ring_slots=slots, max_event_records=400)
        whole = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, **kw)
        assert whole.ring_slots == (slots or whole.ring_slots) and (slots == 0 or whole.ring_slots == 250)
        whole.run(0, T)
        want, want_state = whole.output(), whole.state()
        want_ring = whole.ring()
        want_dbg = whole.debug() if outputs & A.OUT_DEBUG else None
        first = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, **kw)
        first.run(0, cut)
        state, (rv, rw) = first.state(), first.ring()
        first.close()
        second = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, **kw)
        second.set_state(state, rv, rw, next_step=cut)
        second.run(cut, T)
        assert np.array_equal(second.output(), want[:, cut:], equal_nan=True)
        assert np.array_equal(second.state(), want_state, equal_nan=True)
        got_ring = second.ring()
        assert np.array_equal(got_ring[0], want_ring[0]) and np.array_equal(got_ring[1], want_ring[1])
        if want_dbg is not None:
            assert np.array_equal(second.debug(), want_dbg[:, cut:], equal_nan=True)
        # a cursor outside the ring is rejected like restart.c:976-981
        bad = state.copy()
        bad[A.S["meanLast"], 3] = second.ring_slots
        with pytest.raises(api.SipnetGpuError) as e:
            second.set_state(bad, rv, rw, next_step=cut)
        assert e.value.code == 9
        whole.close(); second.close()


def _run_summary(sites, P, ms, flags, static, **kw):
    import os
    os.environ["SIPNET_GPU_STATIC_SCHED"] = "1" if static else "0"
    try:
        ens = api.Ensemble(sites, P, ms, flags, math=A.MATH_FAST, **kw)
    finally:
        os.environ.pop("SIPNET_GPU_STATIC_SCHED", None)
    T = ens.max_steps
    res = []
    for t0 in range(0, T, 600):                                  # segments that do not divide into work items evenly
        ens.run(t0, min(T, t0 + 600))
        res.append((ens.mean(), ens.variance()))
    out = dict(state=ens.state(), status=ens.status(), ring=ens.ring(), moments=res,
               loglik=ens.loglik() if kw["outputs"] & A.OUT_LOGLIK else None,
               events=ens.event_counts() if kw["outputs"] & A.OUT_EVENTS else None)
    ens.close()
    return out


def _same(a, b):
    assert np.array_equal(a["state"], b["state"], equal_nan=True)
    assert np.array_equal(a["status"], b["status"])
    assert all(np.array_equal(x, y) for x, y in zip(a["ring"], b["ring"]))
    for (m1, v1), (m2, v2) in zip(a["moments"], b["moments"]):
        assert np.array_equal(m1, m2, equal_nan=True) and np.array_equal(v1, v2, equal_nan=True)
    if a["loglik"] is not None:
        assert np.array_equal(a["loglik"], b["loglik"])
    if a["events"] is not None:
        assert np.array_equal(a["events"], b["events"])


def test_dynamic_scheduling_is_bit_identical(oracle):
    """More member blocks than resident CTAs: the persistent grid hands out (block, 256-step sub-range) work items.
    Same bits as one CTA per block over the whole range -- single site with log-likelihood, and many sites with
    events and unequal record counts -- and the oracle on sampled members."""
    # (1) one site, 40 000 members (313 blocks of 128 > 296 resident), NEE likelihood + moments
    site = synth.synth_site(0, 2, "half-daily")
    rc, done, o_out, _, _ = oracle.run(synth.SYNTH_FLAGS, synth.synth_params(1)[:, 0], site, want_debug=False)
    site.nee_obs = synth.synth_obs(o_out[:, A.O["nee"]].copy())
    P = synth.synth_params(40000, stream=31)
    kw = dict(outputs=A.OUT_MOMENTS | A.OUT_LOGLIK, summary_cols=[A.O["nee"], A.O["soilWater"]], out_steps_capacity=600,
              nee_sigma=0.7)
    dyn = _run_summary([site], P, None, synth.SYNTH_FLAGS, False, **kw)
    _same(dyn, _run_summary([site], P, None, synth.SYNTH_FLAGS, True, **kw))
    for m in (0, 127, 128, 20011, 39999):                         # and the oracle for a few of them
        rc, done, out, _, _ = oracle.run(synth.SYNTH_FLAGS, P[:, m], site, want_debug=False)
        assert rc == 0
        assert dyn["state"][A.S["plantWoodC"], m] + dyn["state"][A.S["plantCAccountingDelta"], m] == out[-1, A.O["plantWoodC"]]
        assert dyn["state"][A.S["soilC"], m] == out[-1, A.O["soilC"]] and dyn["state"][A.S["totNee"], m] == out[-1, A.O["cumNEE"]]
    # (2) 400 sites x 100 members with the event schedule; every 7th site one year shorter
    sites, P, ms, flags = synth.config_c3(nsites=400, members_per_site=100, nyears=2)
    for s in range(0, 400, 7):
        short = synth.synth_site(s, 1, "half-daily", with_events=True)
        sites[s] = short
    kw = dict(outputs=A.OUT_MOMENTS | A.OUT_EVENTS, summary_cols=[A.O["nee"]], out_steps_capacity=600, max_event_records=8)
    _same(_run_summary(sites, P, ms, flags, False, **kw), _run_summary(sites, P, ms, flags, True, **kw))


def test_random_flag_combinations_are_bit_identical(oracle):
    """The same 30 random flag combinations as tests/test_oracle_vs_reference.py, optimistic kernel vs oracle: all 32
    output columns of every step and the final state, bit for bit."""
    from gpu_util import random_flag_cases
    for trial, f, site, P in random_flag_cases():
        with api.Ensemble([site], P, None, f, outputs=A.OUT_FULL, math=A.MATH_FAST) as ens:
            ens.run()
            out = ens.output()
            status = ens.status()
        for m in range(P.shape[1]):
            rc, done, o_out, _, _ = oracle.run(f, P[:, m], site, want_debug=False)
            if status[m] & A.ST_BAD_ALLOCATION:
                assert rc == 3
                continue
            assert rc == 0 and done == site.nsteps, (trial, m)
            assert np.array_equal(out[:, :, m].T, o_out, equal_nan=True), (trial, m, f)


def test_full_size_ensemble_is_invariant_under_member_permutation_and_segmentation():
    """Size-independent properties at BASELINE's per-GPU size (131 072 members; two years keep it quick): a member's
    results do not depend on which block / lane / work item integrates it, nor on how the run is cut into segments.
    (a) the ensemble in reversed member order gives the reversed state, bit for bit; (b) duplicated members give
    duplicated results; (c) three run segments end in the state of one run."""
    M = 131072
    site = synth.synth_site(6, 2, "half-daily", with_events=True)
    P = synth.synth_params(M // 2, stream=77)
    P = np.concatenate([P, P], axis=1)                                   # every member twice
    kw = dict(outputs=A.OUT_MOMENTS | A.OUT_LOGLIK, summary_cols=[A.O["nee"]], math=A.MATH_FAST, nee_sigma=0.5)
    rng = np.random.default_rng(3)
    site.nee_obs = np.where(rng.uniform(size=site.nsteps) < 0.2, np.nan, rng.normal(0, 1.5, site.nsteps))
    with api.Ensemble([site], P, None, synth.SYNTH_FLAGS, **kw) as ens:
        ens.run()
        st, ll, mean = ens.state(), ens.loglik(), ens.mean()
    assert np.array_equal(st[:, : M // 2], st[:, M // 2:], equal_nan=True) and np.array_equal(ll[: M // 2], ll[M // 2:])
    with api.Ensemble([site], np.ascontiguousarray(P[:, ::-1]), None, synth.SYNTH_FLAGS, **kw) as ens:
        for t0, t1 in ((0, 500), (500, 501), (501, site.nsteps)):
            ens.run(t0, t1)
        st2, ll2 = ens.state(), ens.loglik()
    assert np.array_equal(st2[:, ::-1], st, equal_nan=True) and np.array_equal(ll2[::-1], ll)
    assert np.isfinite(mean).all()


@pytest.mark.parametrize("math", [A.MATH_VALIDATION, A.MATH_FAST])
def test_mixed_site_blocks_are_bit_identical(oracle, math):
    """128-member blocks share a block between the tail of one site and the head of the next (BlockDesc): sites with
    different forcing, event schedules, record counts (ragged: the second site of a block may end first or last),
    member counts that leave tails of every size, a site small enough to be swallowed whole, likelihoods against
    per-site observations, event records, segments.  Every member against the oracle, and the whole run against
    32-member blocks (one site per block)."""
    lens = [2, 1, 2, 1, 2]
    counts = [100, 37, 5, 150, 28]
    sites = [synth.synth_site(30 + i, lens[i], ["half-daily", "unequal"][i % 2], with_events=True) for i in range(5)]
    rng = np.random.default_rng(11)
    for s_ in sites:
        s_.nee_obs = np.where(rng.uniform(size=s_.nsteps) < 0.3, np.nan, rng.normal(0, 1.5, s_.nsteps))
    ms = np.repeat(np.arange(5, dtype=np.int32), counts)
    P = synth.synth_params(int(ms.size), stream=31)
    kw = dict(outputs=A.OUT_FULL | A.OUT_LOGLIK | A.OUT_EVENTS, math=math, nee_sigma=0.5, max_event_records=256)

    def run(block, segments):
        with api.Ensemble(sites, P, ms, synth.SYNTH_FLAGS, block_threads=block, out_steps_capacity=0, **kw) as ens:
            outs = []
            T = ens.max_steps
            cuts = [0, T] if not segments else [0, 300, 301, 1000, T]
            for a_, b_ in zip(cuts[:-1], cuts[1:]):
                ens.run(a_, b_)
                outs.append(ens.output())
            return dict(out=np.concatenate(outs, axis=1), ll=ens.loglik(), state=ens.state(), status=ens.status(),
                        counts=ens.event_counts(), recs=ens.event_records())

    wide, narrow, seg = run(128, False), run(32, False), run(128, True)
    for other in (narrow, seg):
        assert np.array_equal(wide["out"], other["out"], equal_nan=True)
        assert np.array_equal(wide["ll"], other["ll"]) and np.array_equal(wide["state"], other["state"], equal_nan=True)
        assert np.array_equal(wide["counts"], other["counts"]) and np.array_equal(wide["status"], other["status"])
    for m in list(range(0, ms.size, 7)) + [99, 100, 136, 137, 141, 142, 291, 292, ms.size - 1]:
        site = sites[ms[m]]
        rc, done, o_out, _, o_recs = oracle.run(synth.SYNTH_FLAGS, P[:, m], site, want_debug=False, max_event_records=256)
        assert rc == 0
        assert np.array_equal(wide["out"][:, :site.nsteps, m].T, o_out, equal_nan=True), m
        assert np.isnan(wide["out"][:, site.nsteps:, m]).all()
        assert [(r.step, r.type, r.variant, r.nval, tuple(r.val)) for r in wide["recs"][m]] == \
               [(r.step, r.type, r.variant, r.nval, tuple(r.val)) for r in o_recs], m
