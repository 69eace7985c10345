"""CPU: restart checkpoints on the host side (sipnet_b200/host/sip_restart.c) against the reference's
behaviour (reference src/sipnet/restart.c).

Golden data came from the UNMODIFIED reference binary (tests/golden/make_golden.py: restart_golden):
a checkpoint it wrote after the first year of the russell_2 smoke case, and its exit code for every tampered
variant in tests/restart_cases.py.  When oracle/_ref is present the verdicts are re-checked live."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT
from host_util import ContextC, RestartC, SiteDataC, flags_c, host_lib, restart_protos, split_case, unpack_smoke
from restart_cases import CASES
from sipnet_b200 import _abi as A

CKPT = os.path.join(GOLDEN_DIR, "restart_russell_2.ckpt")
GOLD = json.load(open(os.path.join(GOLDEN_DIR, "restart_cases.json")))
RUSSELL_2 = dict(litterPool=1, nitrogenCycle=1, anaerobic=1)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sipnet_ref")


@pytest.fixture(scope="module")
def lib():
    return restart_protos(host_lib())


@pytest.fixture(scope="module")
def segments(tmp_path_factory):
    td = str(tmp_path_factory.mktemp("restart"))
    smoke = unpack_smoke(os.path.join(td, "smoke"))
    a, b = os.path.join(td, "a"), os.path.join(td, "b")
    split_case(os.path.join(smoke, "russell_2"), a, b, 2016)
    return a, b


def load_site(lib, clim):
    s = SiteDataC()
    assert lib.sip_read_clim(clim.encode(), 1, 1, C.byref(s)) == 0
    return s


def verdict(lib, path, site, flags=RUSSELL_2):
    """what `sipnet_gpu --restart-in path` does before touching the device: parse, then the load-time checks"""
    r = RestartC()
    rc = lib.sip_read_restart(path.encode(), C.byref(r))
    if rc:
        return rc
    ctx = ContextC()
    ctx.flags = flags_c(flags)
    ctx.quiet = 1
    return lib.sip_check_restart(path.encode(), C.byref(r), C.byref(ctx), C.byref(site))


def test_reference_checkpoint_round_trips_byte_for_byte(lib, tmp_path):
    r = RestartC()
    assert lib.sip_read_restart(CKPT.encode(), C.byref(r)) == 0, lib.sip_host_error()
    assert r.modelVersion == b"2.1.0" and r.processedSteps == 2928
    assert (r.boundaryYear, r.boundaryDay, r.boundaryTime, r.boundaryLength) == (2016, 366, 23.0, 0.125)
    assert r.meanLength == 250 and r.meanTotWeight == 5.0 and r.isAlive == 1
    assert [getattr(r.flags, n) for n in A.FLAG_NAMES] == [flags_c(RUSSELL_2).__getattribute__(n) for n in A.FLAG_NAMES]
    out = str(tmp_path / "copy.ckpt")
    assert lib.sip_write_restart(out.encode(), C.byref(r)) == 0
    assert open(out, "rb").read() == open(CKPT, "rb").read()      # %.17g / %d formats, key order, blank lines


@pytest.mark.parametrize("name", sorted(CASES))
def test_tampered_checkpoints_get_the_references_verdict(lib, segments, tmp_path, name):
    text = open(CKPT).read()
    path = str(tmp_path / "ck_in")
    with open(path, "w") as f:
        f.write(CASES[name](text))
    site = load_site(lib, os.path.join(segments[1], "sipnet.clim"))
    got = verdict(lib, path, site)
    assert got == GOLD["exit_codes"][name], (name, lib.sip_host_error())
    if os.path.exists(REF_BIN):                                   # live: the unmodified reference on the same file
        r = subprocess.run([REF_BIN, "-i", "sipnet.in", "--quiet", "--restart-in", path], cwd=segments[1],
                           stdout=subprocess.DEVNULL)
        assert r.returncode == got
    lib.sip_site_free(C.byref(site))


def test_missing_file_and_wrong_segment(lib, segments):
    site_b = load_site(lib, os.path.join(segments[1], "sipnet.clim"))
    assert verdict(lib, "/nonexistent/ck", site_b) == 6           # openFile(): EXIT_CODE_FILE_OPEN_OR_READ_ERROR
    site_a = load_site(lib, os.path.join(segments[0], "sipnet.clim"))
    assert verdict(lib, CKPT, site_a) == 9                        # resuming into the segment the checkpoint ended
    assert verdict(lib, CKPT, site_b, flags=dict(RUSSELL_2, gdd=0)) == 9
    assert verdict(lib, CKPT, site_b) == 0


def test_checkpoint_maps_onto_device_state_rows(lib):
    """sip_restart_to_state / sip_restart_from_device are inverse on everything the file holds."""
    lib.sip_restart_to_state.argtypes = [C.POINTER(RestartC), C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.c_int64]
    lib.sip_restart_from_device.argtypes = [C.POINTER(RestartC), C.POINTER(ContextC), C.POINTER(SiteDataC), C.c_longlong,
                                            C.c_longlong] + [C.POINTER(C.c_double), C.c_int64] * 2 + [
                                                C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64]
    r = RestartC()
    assert lib.sip_read_restart(CKPT.encode(), C.byref(r)) == 0
    state = np.full(A.NSTATE, np.nan)
    rv, rw = np.zeros(250), np.zeros(250)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lib.sip_restart_to_state(C.byref(r), dp(state), 1, dp(rv), dp(rw), 1)
    assert state[A.S["plantWoodC"]] == r.envi[0] and state[A.S["plantCAccountingDelta"]] == r.envi[12]
    assert state[A.S["totNee"]] == r.trackers[25] and state[A.S["trackersLastYear"]] == 2016
    assert state[A.S["meanStart"]] == r.meanStart and state[A.S["meanLast"]] == r.meanLast
    assert np.array_equal(rv, np.array(r.values[:])) and np.array_equal(rw, np.array(r.weights[:]))
    assert not np.isnan(state).any()
    # back: per-step trackers come from the last debug row
    dbg = np.zeros(A.NDEBUG)
    dbg[13 + 56:13 + 56 + 33] = np.array(r.trackers[:])
    dbg[A.NDEBUG - 1] = r.isAlive
    ctx = ContextC()
    ctx.flags = flags_c(RUSSELL_2)
    site = SiteDataC()
    year, day = (C.c_int32 * 1)(2016), (C.c_int32 * 1)(366)
    tm, ln = (C.c_double * 1)(23.0), (C.c_double * 1)(0.125)
    site.nsteps, site.year, site.day, site.time, site.length = 1, year, day, tm, ln
    back = RestartC()
    lib.sip_restart_from_device(C.byref(back), C.byref(ctx), C.byref(site), r.processedSteps, r.checkpointUtcEpoch,
                                dp(state), 1, dp(dbg), 1, dp(rv), dp(rw), 1)
    back.buildInfo = r.buildInfo
    assert bytes(back) == bytes(r)


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/sipnet_ref not built")
def test_random_checkpoint_mutants_get_the_references_verdict(lib, segments, tmp_path):
    """Seeded random edits of the reference's checkpoint (tests/test_reader_fuzz.py: mutate) -- the reference binary's
    exit code on `--restart-in mutant` must be what our reader + load-time checks return."""
    import random

    from test_reader_fuzz import mutate
    src = open(CKPT).readlines()
    rng = random.Random(505)
    site = load_site(lib, os.path.join(segments[1], "sipnet.clim"))
    path = str(tmp_path / "ck_mut")
    seen = set()
    for n in range(150):
        lines = src
        for _ in range(rng.choice([1, 1, 2])):
            lines = mutate(lines, rng, keep_first=0)
        open(path, "w").writelines(lines)
        r = subprocess.run([REF_BIN, "-i", "sipnet.in", "--quiet", "--restart-in", path], cwd=segments[1],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        got = verdict(lib, path, site)
        if r.returncode < 0:
            assert got != 0, n
            continue
        if r.returncode not in (0, 5, 6, 9):      # accepted by the loader, then the RUN failed on the edited state
            assert got == 0, (n, r.returncode)    # (e.g. a huge ring weight: exit 7 from the mean tracker)
            continue
        assert got == r.returncode, (n, [ln for ln in lines if ln not in src][:3], lib.sip_host_error())
        seen.add(got)
    assert seen == {0, 9}
    lib.sip_site_free(C.byref(site))
