"""The mass-balance tracker (reference src/sipnet/balance.c:13-148, called from
updatePoolsAndBalance(), sipnet.c:1769-1806).  It is diagnostic -- nothing feeds back into state -- and is evaluated
by the validation dump only.

CPU: the oracle's restatement gives the reference's deltaC / deltaN bit for bit (goldens, wide-prior ensembles with
the full event schedule, where clamping and mortality make some steps NOT balance), and passes the reference's own
regression bar (tests/sipnet/test_modeling/testBalance.c: |deltaC|, |deltaN| < 1e-8 on every one of the 40 steps of
balance.param / balance.clim, in its three flag / event variants -- our goldens balance_*).
GPU: the kernel's two balance rows equal the oracle's bit for bit, and SIPNET_GPU_ST_BALANCE marks exactly the
members for which the reference would print its warning."""
import numpy as np
import pytest

from conftest import Golden, golden_names, have_gpu
from sipnet_b200 import _abi as A, synth


def _first_member(g):
    return g.params if g.params.ndim == 1 else g.params[:, 0]


@pytest.mark.parametrize("name", golden_names())
def test_oracle_balance_equals_reference(oracle, refshim, name):
    g = Golden(name)
    rc1, d1, b1 = oracle.run_balance(g.flags, _first_member(g), g.site)
    rc2, d2, b2 = refshim.run_balance(g.flags, _first_member(g), g.site)
    assert (rc1, d1) == (rc2, d2)
    assert np.array_equal(b1[:d1], b2[:d2], equal_nan=True)


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("balance_")])
def test_reference_balance_regression_bar(oracle, name):
    """testBalance.c: the balance closes to 1e-8 on every step of the reference's fixture."""
    g = Golden(name)
    rc, done, bal = oracle.run_balance(g.flags, _first_member(g), g.site)
    assert rc == 0 and done == g.nsteps == 40
    assert np.all(np.abs(bal) < 1e-8)


def test_oracle_balance_equals_reference_wide_prior(oracle, refshim):
    site = synth.synth_site(4, 3, "unequal", with_events=True)
    P = synth.synth_params(12, stream=5)
    nonzero = 0
    for m in range(P.shape[1]):
        rc1, d1, b1 = oracle.run_balance(synth.SYNTH_FLAGS, P[:, m], site)
        rc2, d2, b2 = refshim.run_balance(synth.SYNTH_FLAGS, P[:, m], site)
        assert (rc1, d1) == (rc2, d2) and np.array_equal(b1[:d1], b2[:d2], equal_nan=True), m
        nonzero += int((b2[:d2] != 0).sum())
    assert nonzero > 0, "the ensemble should contain steps the reference itself flags as unbalanced"


@pytest.mark.gpu
@pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")
@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith(("balance_", "synth_")) or n == "smoke_russell_2"])
def test_gpu_balance_rows_equal_oracle(oracle, name):
    from sipnet_b200 import api
    g = Golden(name)
    P = np.ascontiguousarray(_first_member(g).reshape(A.NPARAMS, 1))
    rc, done, bal = oracle.run_balance(g.flags, P[:, 0], g.site)
    assert rc == 0
    with api.Ensemble([g.site], P, None, g.flags, outputs=A.OUT_FULL | A.OUT_DEBUG, math=A.MATH_VALIDATION) as ens:
        ens.run()
        got = ens.balance()                                   # [2][T][1]
        status = ens.status()
    assert np.array_equal(got[:, :done, 0].T, bal[:done])
    assert bool(status[0] & A.ST_BALANCE) == bool((bal[:done] != 0).any())


@pytest.mark.gpu
@pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")
def test_gpu_balance_ensemble_and_segments(oracle):
    from sipnet_b200 import api
    site = synth.synth_site(4, 3, "unequal", with_events=True)
    P = synth.synth_params(40, stream=5)
    with api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_DEBUG, math=A.MATH_VALIDATION,
                      out_steps_capacity=500) as ens:
        parts = []
        for t0 in range(0, site.nsteps, 500):
            ens.run(t0, min(site.nsteps, t0 + 500))
            parts.append(ens.balance())
        got = np.concatenate(parts, axis=1)
        status = ens.status()
    for m in range(0, 40, 3):
        rc, done, bal = oracle.run_balance(synth.SYNTH_FLAGS, P[:, m], site)
        assert np.array_equal(got[:, :, m].T, bal), m
        assert bool(status[m] & A.ST_BALANCE) == bool((bal != 0).any()), m


@pytest.mark.gpu
@pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")
def test_gpu_informational_bits_equal_oracle(oracle):
    """SIPNET_GPU_ST_LEAFON_LIMITED / _N_LIMITED / _MINN_LIMITED mark the members for which the reference would print
    its limitation messages or cap the mineral-N losses (limitations.c:46-62, 85-114, 119-130), in every numerics."""
    from sipnet_b200 import api
    site = synth.synth_site(4, 3, "unequal", with_events=True)
    P = synth.synth_params(64, stream=5)
    P[A.P["nVolatilizationFrac"], 4:20] = 5.0                 # losses above the mineral N pool, N limitation after it
    P[A.P["leafGrowth"], 20:24] = 0.0                         # no leaf-on flush: nothing to limit
    bits = A.ST_LEAFON_LIMITED | A.ST_N_LIMITED | A.ST_MINN_LIMITED
    want = []
    for m in range(P.shape[1]):
        oracle.run_balance(synth.SYNTH_FLAGS, P[:, m], site)
        want.append(oracle.last_info & bits)
    want = np.array(want, np.uint32)
    for b in (A.ST_LEAFON_LIMITED, A.ST_N_LIMITED, A.ST_MINN_LIMITED):
        assert (want & b).any() and not (want & b).all(), hex(b)
    for math in (A.MATH_VALIDATION, A.MATH_FAST):
        with api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL, math=math) as ens:
            ens.run()
            got = ens.status() & bits
        assert np.array_equal(got, want), math


@pytest.mark.gpu
@pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")
def test_gpu_message_counters_equal_oracle(oracle):
    """SIPNET_GPU_GATHER_COUNTERS (validation dump): how OFTEN the reference would have printed each informational
    message -- leaf-on limitation (limitations.c:48-61), N limitation (:98-102), the mineral-N cap (:119-130), the
    non-negative-stock warning (sipnet.c:1346-1356, once per clamped stock) -- per member, continued across segments
    and cleared by a reset."""
    from sipnet_b200 import api
    site = synth.synth_site(4, 3, "unequal", with_events=True)
    P = synth.synth_params(48, stream=5)
    P[A.P["nVolatilizationFrac"], 4:20] = 5.0
    P[A.P["leafGrowth"], 20:24] = 0.0
    P[A.P["baseSoilResp"], 24:30] *= 40.0                      # soil carbon runs out: clamp warnings
    want = []
    for m in range(P.shape[1]):
        oracle.run_balance(synth.SYNTH_FLAGS, P[:, m], site)
        want.append(oracle.last_counts.copy())
    want = np.array(want, np.uint32).T                          # [NCOUNTERS][members]
    for k in range(A.NCOUNTERS):
        assert want[k].max() > 1, f"counter {k} is never above one on this ensemble: pick harder parameters"
    with api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_DEBUG, out_steps_capacity=500) as ens:
        for t0 in range(0, site.nsteps, 500):
            ens.run(t0, min(site.nsteps, t0 + 500))
        got = ens.counters()
        status = ens.status()
        assert np.array_equal(got, want)
        assert np.array_equal(got[A.CNT_CLAMPED] > 0, (status & A.ST_CLAMPED) != 0)
        assert np.array_equal(got[A.CNT_N_LIMITED] > 0, (status & A.ST_N_LIMITED) != 0)
        ens.reset()
        ens.run(0, 500)
        assert (ens.counters() <= want).all() and not np.array_equal(ens.counters(), want)
