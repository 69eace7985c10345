"""GPU: the drop-in driver end to end.  `sipnet_gpu -i sipnet.in` on the reference's own smoke
inputs must reproduce the reference's committed outputs BYTE FOR BYTE: sipnet.out (checked by md5 of
the reference's file) and events.out (checked against the reference's bytes)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import Golden
from host_util import DRIVER, unpack_smoke

pytestmark = pytest.mark.gpu
SMOKE = ["niwot", "russell_1", "russell_2", "russell_3"]


@pytest.fixture(scope="module")
def smoke_dir(tmp_path_factory):
    return unpack_smoke(str(tmp_path_factory.mktemp("smoke")))


@pytest.mark.parametrize("math", ["fast", "validation"])
@pytest.mark.parametrize("case", SMOKE)
def test_smoke_case_byte_identical(smoke_dir, case, math):
    g = Golden("smoke_" + case)
    d = os.path.join(smoke_dir, case)
    for f in ("sipnet.out", "events.out"):
        if os.path.exists(os.path.join(d, f)):
            os.remove(os.path.join(d, f))
    args = [DRIVER, "-i", "sipnet.in", "--quiet"] + (["--validation-math"] if math == "validation" else [])
    r = subprocess.run(args, cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(os.path.join(d, "sipnet.out"), "rb").read()
    assert hashlib.md5(out).hexdigest() == g.main_out_md5, f"{case}: sipnet.out differs from the reference's"
    assert open(os.path.join(d, "events.out"), "rb").read() == g.events_out


def test_ensemble_extension(smoke_dir, tmp_path):
    """--ensemble-params: member k's files equal a single-member run with that .param."""
    d = os.path.join(smoke_dir, "russell_2")
    base = open(os.path.join(d, "sipnet.param")).read()
    variants = [base, base.replace("aMax 53.2895432752984", "aMax 40.0"), base.replace("soilWHC 12", "soilWHC 9.5")]
    paths = []
    for k, txt in enumerate(variants):
        p = os.path.join(d, f"member{k}.param")
        open(p, "w").write(txt)
        paths.append(p)
    open(os.path.join(d, "ens.txt"), "w").write("\n".join(paths) + "\n")
    r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--ensemble-params", "ens.txt"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for k, txt in enumerate(variants):
        open(os.path.join(d, "sipnet.param"), "w").write(txt)
        r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0
        assert open(os.path.join(d, "sipnet.out"), "rb").read() == open(os.path.join(d, f"sipnet.out.{k}"), "rb").read()
        assert open(os.path.join(d, "events.out"), "rb").read() == open(os.path.join(d, f"events.out.{k}"), "rb").read()
    open(os.path.join(d, "sipnet.param"), "w").write(base)


def test_driver_exit_codes(smoke_dir):
    d = os.path.join(smoke_dir, "russell_1")
    base = open(os.path.join(d, "sipnet.param")).read()
    try:
        open(os.path.join(d, "sipnet.param"), "w").write(base.replace("leafAllocation 0.2", "leafAllocation 0.7"))
        r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 3            # ensureAllocation(), sipnet.c:1117-1122
    finally:
        open(os.path.join(d, "sipnet.param"), "w").write(base)
    ev = open(os.path.join(d, "events.in")).read()
    try:
        open(os.path.join(d, "events.in"), "w").write("2015 300 irrig 1.0 1\n" + ev)
        r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 5            # first event before the climate record (frontend.c:217-222)
    finally:
        open(os.path.join(d, "events.in"), "w").write(ev)


@pytest.mark.parametrize("case", SMOKE)
def test_debug_log_byte_identical(smoke_dir, case):
    """--debug-log: the three per-step logs (%.15g of every Envi/Fluxes/Trackers field) equal the reference's."""
    import json
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "debug_log_md5.json")))[case]
    d = os.path.join(smoke_dir, case)
    r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--debug-log", "dbg"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for k in ("envi", "fluxes", "trackers"):
        got = hashlib.md5(open(os.path.join(d, f"dbg_{k}.log"), "rb").read()).hexdigest()
        assert got == want[k], f"{case}: dbg_{k}.log differs from the reference's"


# ---- restart checkpoints: hand a segmented run back and forth with the reference -----------
import json  # noqa: E402

from conftest import GOLDEN_DIR, ROOT  # noqa: E402
from host_util import split_case, strip_volatile  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sipnet_ref")
RESTART_GOLD = json.load(open(os.path.join(GOLDEN_DIR, "restart_cases.json")))
RESTART_CKPT = os.path.join(GOLDEN_DIR, "restart_russell_2.ckpt")


def _run(binary, cwd, *extra):
    r = subprocess.run([binary, "-i", "sipnet.in", "--quiet", *extra], cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return r


def _md5(path, strip=False):
    data = open(path, "rb").read()
    return hashlib.md5(strip_volatile(data) if strip else data).hexdigest()


@pytest.mark.parametrize("math", ["fast", "validation"])
def test_restart_against_reference_goldens(smoke_dir, tmp_path, math):
    """russell_2 cut after 2016.  Our checkpoint equals the reference's byte for byte (bar the time stamp and
    build id lines); resuming from the REFERENCE's checkpoint reproduces the reference's second segment --
    sipnet.out, events.out and the next checkpoint -- byte for byte."""
    extra = ["--validation-math"] if math == "validation" else []
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    split_case(os.path.join(smoke_dir, "russell_2"), a, b, 2016)
    _run(DRIVER, a, "--restart-out", "ck", *extra)
    assert _md5(os.path.join(a, "sipnet.out")) == RESTART_GOLD["segment1_out_md5"]
    assert strip_volatile(open(os.path.join(a, "ck"), "rb").read()) == strip_volatile(open(RESTART_CKPT, "rb").read())
    for ck in (RESTART_CKPT, os.path.join(a, "ck")):               # the reference's checkpoint, then our own
        _run(DRIVER, b, "--restart-in", ck, "--restart-out", "ck2", *extra)
        assert _md5(os.path.join(b, "sipnet.out")) == RESTART_GOLD["segment2_out_md5"]
        assert _md5(os.path.join(b, "events.out")) == RESTART_GOLD["segment2_events_md5"]
        assert _md5(os.path.join(b, "ck2"), strip=True) == RESTART_GOLD["segment2_checkpoint_md5"]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/sipnet_ref not built")
@pytest.mark.parametrize("case,year", [("niwot", 2000), ("russell_3", 2016), ("russell_1", 2016)])
def test_restart_round_trip_with_live_reference(smoke_dir, tmp_path, case, year):
    """ours -> reference -> ours: every hand-over through a checkpoint file, each leg compared with the reference
    running the whole chain on its own."""
    legs = {}
    for who, binary in (("ref", REF_BIN), ("ours", DRIVER)):
        a, b = str(tmp_path / who / "a"), str(tmp_path / who / "b")
        split_case(os.path.join(smoke_dir, case), a, b, year)
        _run(binary, a, "--restart-out", "ck")
        legs[who] = (a, b)
    ra, rb = legs["ref"]
    oa, ob = legs["ours"]
    assert strip_volatile(open(os.path.join(oa, "ck"), "rb").read()) == strip_volatile(open(os.path.join(ra, "ck"), "rb").read())
    _run(REF_BIN, rb, "--restart-in", os.path.join(oa, "ck"), "--restart-out", "ck2")     # reference resumes from OUR file
    _run(DRIVER, ob, "--restart-in", os.path.join(ra, "ck"), "--restart-out", "ck2")      # we resume from the REFERENCE's
    for f in ("sipnet.out", "events.out"):
        assert open(os.path.join(ob, f), "rb").read() == open(os.path.join(rb, f), "rb").read(), f
    assert _md5(os.path.join(ob, "ck2"), strip=True) == _md5(os.path.join(rb, "ck2"), strip=True)
    # and the two segments together are the unsegmented run (testRestartMVP.c:253-297)
    whole = str(tmp_path / "whole")
    os.makedirs(whole)
    for fn in os.listdir(os.path.join(smoke_dir, case)):
        if fn.endswith((".in", ".param", ".clim")):
            import shutil
            shutil.copy(os.path.join(smoke_dir, case, fn), whole)
    _run(DRIVER, whole)
    header = 1 if b"year" in open(os.path.join(whole, "sipnet.out"), "rb").readline() else 0
    seg = open(os.path.join(oa, "sipnet.out"), "rb").read().splitlines()[header:] + \
        open(os.path.join(ob, "sipnet.out"), "rb").read().splitlines()[header:]
    assert open(os.path.join(whole, "sipnet.out"), "rb").read().splitlines()[header:] == seg


def test_restart_errors_exit_like_the_reference(smoke_dir, tmp_path):
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    split_case(os.path.join(smoke_dir, "russell_2"), a, b, 2016)
    run = lambda cwd, *extra: subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", *extra], cwd=cwd,  # noqa: E731
                                             capture_output=True, text=True).returncode
    assert run(b, "--restart-in", "does-not-exist") == 6
    assert run(a, "--restart-in", RESTART_CKPT) == 9                  # the checkpoint does not precede this segment
    assert run(b, "--restart-in", RESTART_CKPT, "--no-nitrogen-cycle") == 9
    assert run(b, "--restart-in", RESTART_CKPT) == 0


def test_site_list_runs_many_sites_in_one_launch(smoke_dir, tmp_path):
    """--site-list: several sites (own forcing, events, members; different record counts) in ONE launch.  Every
    output file equals the one a separate single-site run of the driver writes."""
    import shutil
    src = os.path.join(smoke_dir, "russell_2")
    work = str(tmp_path / "multi")
    os.makedirs(work)
    shutil.copy(os.path.join(src, "sipnet.in"), work)
    # site A: the whole smoke case; site B: its first year only; site C: whole case, two members with changed parameters
    a, b, c = (os.path.join(work, n) for n in ("siteA", "siteB", "siteC"))
    split_case(src, b, str(tmp_path / "unused"), 2016)
    for d in (a, c):
        os.makedirs(d)
        for fn in ("sipnet.param", "sipnet.clim", "events.in"):
            shutil.copy(os.path.join(src, fn), d)
    base = open(os.path.join(src, "sipnet.param")).read()
    members = []
    for k, txt in enumerate((base.replace("aMax 53.2895432752984", "aMax 44.0"), base.replace("soilWHC 12", "soilWHC 10.5"))):
        p = os.path.join(c, f"m{k}.param")
        open(p, "w").write(txt)
        members.append(p)
    open(os.path.join(c, "members.txt"), "w").write("\n".join(members) + "\n")
    with open(os.path.join(work, "sites.txt"), "w") as f:
        f.write("# file-prefix [events-prefix [member-list]]\n")
        f.write(f"{a}/sipnet\n")
        f.write(f"{b}/sipnet {b}/events   # explicit events prefix\n")
        f.write(f"{c}/sipnet {c}/events {c}/members.txt\n")
    r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--site-list", "sites.txt"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = {n: open(n, "rb").read() for n in (f"{a}/sipnet.out", f"{a}/events.out", f"{b}/sipnet.out", f"{b}/events.out",
                                             f"{c}/sipnet.out.0", f"{c}/events.out.0", f"{c}/sipnet.out.1", f"{c}/events.out.1")}
    # the same sites one at a time
    def single(d, param=None):
        if param:
            shutil.copy(param, os.path.join(d, "sipnet.param"))
        shutil.copy(os.path.join(work, "sipnet.in"), d)
        r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        return open(os.path.join(d, "sipnet.out"), "rb").read(), open(os.path.join(d, "events.out"), "rb").read()
    assert single(a) == (got[f"{a}/sipnet.out"], got[f"{a}/events.out"])
    assert hashlib.md5(got[f"{a}/sipnet.out"]).hexdigest() == Golden("smoke_russell_2").main_out_md5     # = the reference's file
    assert single(b) == (got[f"{b}/sipnet.out"], got[f"{b}/events.out"])
    assert hashlib.md5(got[f"{b}/sipnet.out"]).hexdigest() == RESTART_GOLD["segment1_out_md5"]
    for k in range(2):
        assert single(c, members[k]) == (got[f"{c}/sipnet.out.{k}"], got[f"{c}/events.out.{k}"])
    # bad lists
    open(os.path.join(work, "empty.txt"), "w").write("# nothing\n")
    assert subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--site-list", "empty.txt"], cwd=work).returncode == 5
    assert subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--site-list", "missing.txt"], cwd=work).returncode == 6
    assert subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--site-list", "sites.txt", "--restart-out", "x"], cwd=work).returncode == 8


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_ngpus() < 2", reason="needs two GPUs on the box")
def test_many_member_launches_on_all_gpus_equal_one_gpu(smoke_dir, tmp_path):
    """The driver fans a --site-list / --ensemble-params launch out over every visible GPU (sipnet_gpu_multi_*; the
    seam is the reference's runModelOutput(), sipnet.c:1954-1990).  Every output file must equal, byte for byte, the
    one written with --devices 1: whole sites per GPU for the site list (5 sites), one site's members split over the
    GPUs for the ensemble (7 members)."""
    import shutil
    src = os.path.join(smoke_dir, "russell_2")
    base = open(os.path.join(src, "sipnet.param")).read()

    def make(work):
        os.makedirs(work)
        shutil.copy(os.path.join(src, "sipnet.in"), work)
        sites = []
        for i in range(5):
            d = os.path.join(work, f"site{i}")
            os.makedirs(d)
            for fn in ("sipnet.clim", "events.in"):
                shutil.copy(os.path.join(src, fn), d)
            open(os.path.join(d, "sipnet.param"), "w").write(base.replace("aMax 53.2895432752984", f"aMax {40 + 3 * i}.5"))
            sites.append(d)
        open(os.path.join(work, "sites.txt"), "w").write("".join(f"{d}/sipnet\n" for d in sites))
        ens = os.path.join(work, "ens")
        os.makedirs(ens)
        for fn in ("sipnet.clim", "events.in", "sipnet.param"):
            shutil.copy(os.path.join(src, fn), ens)
        members = []
        for k in range(7):
            p = os.path.join(ens, f"m{k}.param")
            open(p, "w").write(base.replace("soilWHC 12", f"soilWHC {9 + 0.5 * k}"))
            members.append(p)
        open(os.path.join(ens, "members.txt"), "w").write("\n".join(members) + "\n")
        return sites, ens

    outs = {}
    for tag, extra in (("one", ["--devices", "1"]), ("all", [])):
        work = str(tmp_path / tag)
        sites, ens = make(work)
        r = subprocess.run([DRIVER, "-i", "sipnet.in", "--site-list", "sites.txt", *extra], cwd=work, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        if tag == "all":
            assert f"on {min(_ngpus(), 5)} GPU(s)" in r.stdout, r.stdout
        r = subprocess.run([DRIVER, "-i", "../sipnet.in", "-f", "sipnet", "-e", "events", "--ensemble-params", "members.txt", *extra],
                           cwd=ens, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        files = [f"{d}/sipnet.out" for d in sites] + [f"{d}/events.out" for d in sites]
        files += [f"{ens}/sipnet.out.{k}" for k in range(7)] + [f"{ens}/events.out.{k}" for k in range(7)]
        outs[tag] = [open(f, "rb").read() for f in files]
        assert all(len(b) > 1000 for b in outs[tag][:5])
    assert outs["one"] == outs["all"]


@pytest.mark.parametrize("case", SMOKE)
def test_single_variable_outputs_byte_identical(smoke_dir, case):
    """--do-single-outputs: <prefix>.NEE / .NEE_cum / .GPP / .GPP_cum equal the reference's files (md5 from
    tests/golden/make_golden.py); GPP_cum is accumulated on the host in the reference's own addition order."""
    want = json.load(open(os.path.join(GOLDEN_DIR, "single_outputs_md5.json")))[case]
    d = os.path.join(smoke_dir, case)
    r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--do-single-outputs", "--no-do-main-output"], cwd=d,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for k, md5 in want.items():
        assert hashlib.md5(open(os.path.join(d, "sipnet." + k), "rb").read()).hexdigest() == md5, k
