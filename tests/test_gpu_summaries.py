"""GPU: on-device ensemble summaries.  The reference has no ensemble code (parity
unpinned for these outputs); they are defined in sip_reduce.cu and checked here against a
trivially-written numpy computation over the oracle-validated per-step output."""
import numpy as np
import pytest

from sipnet_b200 import _abi as A, api, synth

pytestmark = pytest.mark.gpu

COLS = [A.O["nee"], A.O["gpp"], A.O["soilWater"]]
QS = [0.05, 0.25, 0.5, 0.95, 1.0]


def test_loglik_matches_numpy(oracle):
    site = synth.synth_site(0, 3, "half-daily", with_events=True)
    P = synth.synth_params(70)
    rc, done, o_out, _, _ = oracle.run(synth.SYNTH_FLAGS, P[:, 0], site, want_debug=False)
    site.nee_obs = synth.synth_obs(o_out[:, A.O["nee"]].copy())
    sigma = 0.5
    for math in (A.MATH_FAST, A.MATH_VALIDATION):
        ens = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL | A.OUT_LOGLIK, nee_sigma=sigma, math=math)
        ens.run()
        nee = ens.output()[A.O["nee"]]                       # [T][M]
        ll, lln = ens.loglik(), ens.loglik_n()
        ens.close()
        mask = ~np.isnan(site.nee_obs)
        assert np.all(lln == mask.sum())
        d = (nee[mask, :] - site.nee_obs[mask, None]) / sigma
        want = (-0.5 * d * d - np.log(sigma) - 0.5 * np.log(2 * np.pi)).sum(axis=0)
        np.testing.assert_allclose(ll, want, rtol=1e-12)
    # segmented runs accumulate the same likelihood (same order of additions => same bits)
    ens = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_LOGLIK, nee_sigma=sigma)
    for t0 in range(0, site.nsteps, 700):
        ens.run(t0, min(site.nsteps, t0 + 700))
    assert np.array_equal(ens.loglik(), ll)
    ens.close()


def test_moments_and_quantiles_match_numpy():
    sites, P, ms, flags = synth.config_c3(nsites=3, members_per_site=37, nyears=2)
    P[A.P["leafAllocation"], 5] = 0.9                         # one failed member (bad allocation): excluded
    ens = api.Ensemble(sites, P, ms, flags, outputs=A.OUT_FULL | A.OUT_MOMENTS | A.OUT_QUANTILES, summary_cols=COLS,
                       quantiles=QS)
    ens.run()
    out = ens.output()
    mean, var, q = ens.mean(), ens.variance(), ens.quantiles()
    ens.close()
    assert np.isnan(out[:, :, 5]).all()
    for s in range(3):
        sel = ms == s
        for i, c in enumerate(COLS):
            x = out[c][:, sel]                                # [T][members of site]
            with np.errstate(invalid="ignore"):
                np.testing.assert_allclose(mean[s, i], np.nanmean(x, axis=1), rtol=1e-12, atol=1e-300)
                np.testing.assert_allclose(var[s, i], np.nanvar(x, axis=1), rtol=1e-9, atol=1e-300)
                np.testing.assert_allclose(q[s, i], np.nanquantile(x, QS, axis=1), rtol=4e-16, atol=1e-300)  # order statistics + lerp


def test_summary_only_mode_matches_full_mode():
    """Without OUT_FULL only the summary columns are kept on the device; same numbers."""
    site = synth.synth_site(2, 2, "unequal", with_events=True)
    P = synth.synth_params(300, stream=2)
    kw = dict(summary_cols=COLS, quantiles=QS)
    a = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL | A.OUT_MOMENTS | A.OUT_QUANTILES, **kw)
    b = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_MOMENTS | A.OUT_QUANTILES, **kw)
    a.run(); b.run()
    assert np.array_equal(a.mean(), b.mean(), equal_nan=True)
    assert np.array_equal(a.variance(), b.variance(), equal_nan=True)
    assert np.array_equal(a.quantiles(), b.quantiles(), equal_nan=True)
    a.close(); b.close()


def test_rows_summary_and_zero_copy_view():
    import torch
    from sipnet_b200 import distributed as D
    rng = np.random.default_rng(5)
    x = rng.normal(size=(300, 4097))
    x[7, 11] = np.nan
    x[:, 100:110] = 1.25                                       # ties
    t = torch.from_numpy(x).cuda()
    mean, var, q = D.rows_summary(t, QS)
    np.testing.assert_allclose(mean.cpu().numpy(), np.nanmean(x, axis=1), rtol=1e-12)
    np.testing.assert_allclose(var.cpu().numpy(), np.nanvar(x, axis=1), rtol=1e-10)
    np.testing.assert_allclose(q.cpu().numpy(), np.nanquantile(x, QS, axis=1), rtol=4e-16)
    assert np.array_equal(q.cpu().numpy()[-1], np.nanmax(x, axis=1)) and np.array_equal(q.cpu().numpy()[2], np.nanmedian(x, axis=1))
    # zero-copy torch view of the library's device output
    site = synth.synth_site(0, 1, "half-daily")
    P = synth.synth_params(48)
    ens = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL)
    ens.run()
    ens.sync()
    view = D.DeviceArray(ens.device_ptr(A.GATHER_FULL), (A.NOUT, site.nsteps, 48)).tensor()
    assert np.array_equal(view.cpu().numpy(), ens.output(), equal_nan=True)
    ens.close()


@pytest.mark.parametrize("kind", ["normal", "wide", "duplicates", "constant", "nan-heavy", "two-values"])
def test_rows_summary_long_rows(kind):
    """Rows longer than the in-shared-memory candidate buffer take the histogram levels of the row kernel:
    well-spread data (two levels), values spanning many binades and both signs, heavy duplication (levels
    until the key is complete), a constant row, mostly-NaN rows; more quantiles than one launch handles."""
    import torch
    from sipnet_b200 import distributed as D
    rng = np.random.default_rng(11)
    n, m = 6, 70001
    if kind == "normal":
        x = rng.normal(3.0, 2.0, size=(n, m))
    elif kind == "wide":
        x = rng.normal(size=(n, m)) * 10.0 ** rng.integers(-30, 30, size=(n, m))
        x[:, ::7] = 0.0
        x[:, 1::97] = -0.0
    elif kind == "duplicates":
        x = rng.integers(0, 4, size=(n, m)).astype(np.float64) * 0.1
    elif kind == "constant":
        x = np.full((n, m), -7.25)
    elif kind == "nan-heavy":
        x = rng.normal(size=(n, m))
        x[rng.uniform(size=(n, m)) < 0.9] = np.nan
        x[0, :] = np.nan                                      # a row with no finite value
        x[1, 5:] = np.nan                                     # five finite values
    else:
        x = np.where(rng.uniform(size=(n, m)) < 0.5, 1.0, np.nextafter(1.0, 2.0))
    qs = [0.0, 0.001, 0.05, 0.25, 0.5, 0.75, 0.999, 1.0]
    mean, var, q = D.rows_summary(torch.from_numpy(x).cuda(), qs)
    with np.errstate(invalid="ignore"), __import__("warnings").catch_warnings():
        __import__("warnings").simplefilter("ignore")
        want_mean, want_var = np.nanmean(x, axis=1), np.nanvar(x, axis=1)
        want_q = np.nanquantile(x, qs, axis=1)
    np.testing.assert_allclose(mean.cpu().numpy(), want_mean, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(var.cpu().numpy(), want_var, rtol=1e-9, atol=1e-300)
    got = q.cpu().numpy()
    np.testing.assert_allclose(got, want_q, rtol=4e-16, atol=0)
    assert np.array_equal(np.isnan(got), np.isnan(want_q))
    # the order statistics themselves are exact
    fin = np.where(np.isnan(x), np.inf, x)
    has = np.isfinite(fin).any(axis=1)
    assert np.array_equal(got[0][has], fin.min(axis=1)[has])


def test_rows_summary_is_deterministic():
    import torch
    from sipnet_b200 import distributed as D
    x = torch.from_numpy(np.random.default_rng(3).normal(size=(4, 100003))).cuda()
    a = [t.cpu().numpy() for t in D.rows_summary(x, QS)]
    b = [t.cpu().numpy() for t in D.rows_summary(x, QS)]
    assert all(np.array_equal(u, v) for u, v in zip(a, b))


def test_rows_summary_is_stream_ordered_on_a_side_stream():
    """On a non-default stream the call only enqueues work (no allocation, no synchronisation): the result is
    ordered on that stream and equals the synchronous call's."""
    import torch
    from sipnet_b200 import distributed as D
    x = torch.from_numpy(np.random.default_rng(8).normal(size=(64, 50001))).cuda()
    want = [t.cpu().numpy() for t in D.rows_summary(x, QS)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        y = x * 1.0                                          # work queued ahead of the summary on the same stream
        got = D.rows_summary(y, QS)
    side.synchronize()
    assert all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(want, got))
