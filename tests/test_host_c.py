"""CPU: the host C library (readers, writers, configuration) against the reference's semantics.

Expected values are the committed golden vectors, which were produced by the reference's OWN
readers and writers (tests/golden/make_golden.py); when oracle/_ref is present the same checks are
repeated live against the unmodified reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import Golden, golden_names
from host_util import CFile, ContextC, DRIVER, flags_c, host_lib, read_site_c, unpack_smoke
from sipnet_b200 import _abi as A

SMOKE = ["niwot", "russell_1", "russell_2", "russell_3"]


@pytest.fixture(scope="module")
def lib():
    return host_lib()


@pytest.fixture(scope="module")
def smoke_dir(tmp_path_factory):
    return unpack_smoke(str(tmp_path_factory.mktemp("smoke")))


@pytest.mark.parametrize("case", SMOKE)
def test_readers_match_reference_parse(lib, smoke_dir, case):
    g = Golden("smoke_" + case)
    d = os.path.join(smoke_dir, case)
    fl = flags_c(g.flags)
    p = np.zeros(A.NPARAMS)
    assert lib.sip_read_params(os.path.join(d, "sipnet.param").encode(), C.byref(fl), 1,
                               p.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert np.array_equal(p, g.params)
    rc, site = read_site_c(lib, os.path.join(d, "sipnet.clim"), g.flags["gdd"], os.path.join(d, "events.in"), fl, p)
    assert rc == 0
    assert np.array_equal(site["year"], g.site.year) and np.array_equal(site["day"], g.site.day)
    for k in A.CLIM_COLS:
        assert np.array_equal(site[k], g.site.clim[k]), k          # bit-exact unit conversions / floors
    assert site["events"] == g.site.events


@pytest.mark.parametrize("name", golden_names())
def test_state_rows_are_byte_identical(lib, tmp_path, name):
    g = Golden(name)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    want = bytes(z["out_text"].tobytes())
    path = str(tmp_path / "rows.out")
    f = CFile(path)
    for i, r in enumerate(g.rows):
        row = np.ascontiguousarray(g.out32[i])
        lib.sip_write_state_row(f.fp, int(g.site.year[r]), int(g.site.day[r]), float(g.site.clim["time"][r]),
                                row.ctypes.data_as(C.POINTER(C.c_double)), 1)
    f.close()
    assert open(path, "rb").read() == want


@pytest.mark.parametrize("name", golden_names())
def test_events_out_is_byte_identical(lib, oracle, tmp_path, name):
    """oracle event records -> C writer == the reference's events.out bytes."""
    g = Golden(name)
    rc, done, _, _, recs = oracle.run(g.flags, g.params, g.site, want_debug=False, max_event_records=8192)
    path = str(tmp_path / "events.out")
    f = CFile(path)
    if g.print_header:
        lib.sip_write_events_header(f.fp)
    for r in recs:
        assert lib.sip_write_event_row(f.fp, int(g.site.year[r.step]), int(g.site.day[r.step]), C.byref(r)) == 0
    f.close()
    assert open(path, "rb").read() == g.events_out


def test_header_text(lib, tmp_path):
    path = str(tmp_path / "h.out")
    f = CFile(path)
    lib.sip_write_header(f.fp)
    f.close()
    txt = open(path).read()
    assert txt.startswith("year day  time plantWoodC plantLeafC woodCreation     soil coarseRootC fineRootC   litter")
    assert txt.endswith("nUptake      ch4  nppStorage\n") and txt.count("\n") == 1


def write(path, text):
    with open(path, "w") as f:
        f.write(text)
    return str(path)


def test_param_reader_errors(lib, tmp_path, smoke_dir):
    base = open(os.path.join(smoke_dir, "russell_2", "sipnet.param")).read()
    fl = flags_c(dict(litterPool=1, nitrogenCycle=1, anaerobic=1))
    p = np.zeros(A.NPARAMS)
    pp = p.ctypes.data_as(C.POINTER(C.c_double))
    rd = lambda txt: lib.sip_read_params(write(tmp_path / "x.param", txt).encode(), C.byref(fl), 1, pp)
    assert rd(base) == 0
    assert rd(base.replace("aMax 53.2895432752984", "aMax *")) == 3                 # '*' rejected
    assert rd(base.replace("aMax 53.2895432752984\n", "")) == 5                     # missing required
    assert rd(base + "aMax 1.0\n") == 5                                             # duplicate
    assert rd(base + "someUnknownParam 3\n! comment\n\n") == 0                       # unknown tolerated
    assert rd(base.replace("aMax ", "AMAX ")) == 0 and p[A.P["aMax"]] == 53.2895432752984   # case-insensitive
    assert rd(base.replace("soilWHC 12", "soilWHC 0")) == 0 and p[A.P["soilWHC"]] == 1e-6    # TINY floor
    assert rd(base.replace("kCN 80.0\n", "")) == 5                                  # required only because N cycle is on
    fl0 = flags_c({})
    assert lib.sip_read_params(write(tmp_path / "y.param", base.replace("kCN 80.0\n", "")).encode(), C.byref(fl0), 1, pp) == 0
    assert lib.sip_read_params(b"/nonexistent/x.param", C.byref(fl), 1, pp) == 6


def test_clim_reader_formats_and_errors(lib, tmp_path):
    row12 = "2016 1 0.0 0.5 10.0 9.0 7.5 1.0 700.0 600.0 500.0 1.5\n"
    row14 = "0 2016 1 0.0 0.5 10.0 9.0 7.5 1.0 700.0 600.0 500.0 1.5 0.0\n"
    rc, s = read_site_c(lib, write(tmp_path / "a.clim", row12 * 3), 1)
    assert rc == 0 and s["year"].size == 3 and s["par"][0] == 7.5 * (1.0 / 0.5) and s["precip"][0] == 1.0 * 0.1
    assert s["vpd"][0] == 700.0 * 0.001 and s["gdd"][0] == 5.0
    rc, s = read_site_c(lib, write(tmp_path / "b.clim", row14 * 2), 0)
    assert rc == 0 and s["tair"][1] == 10.0 and s["gdd"][0] == 0.0
    assert read_site_c(lib, write(tmp_path / "c.clim", "2016 1 0.0 0.5 10.0\n"), 1)[0] == 5        # wrong column count
    assert read_site_c(lib, write(tmp_path / "d.clim", ""), 1)[0] == 5                              # empty
    assert read_site_c(lib, write(tmp_path / "e.clim", row14 + row14.replace("0 2016", "1 2016", 1)), 1)[0] == 5  # 2 locations
    assert read_site_c(lib, write(tmp_path / "f.clim", row12 + "2016 1 0.5 0.5 abc\n"), 1)[0] == 5  # bad data
    rc, s = read_site_c(lib, write(tmp_path / "g.clim", row12.replace(" 0.5 10.0", " -43200 10.0")), 1)
    assert rc == 0 and s["length"][0] == 0.5                                                        # seconds -> days
    rc, s = read_site_c(lib, write(tmp_path / "h.clim", row12.replace("700.0", "0.0").replace(" 1.5\n", " 0.0\n")), 1)
    assert rc == 0 and s["vpd"][0] == 1e-6 and s["wspd"][0] == 1e-6                                 # TINY floors


def test_event_reader_errors(lib, tmp_path):
    clim = write(tmp_path / "a.clim", "2016 1 0.0 0.5 10.0 9.0 7.5 1.0 700.0 600.0 500.0 1.5\n")
    p = np.zeros(A.NPARAMS)
    fl = flags_c(dict(gdd=0))
    rd = lambda txt, f=fl: read_site_c(lib, clim, 0, write(tmp_path / "ev.in", txt), f, p)
    rc, s = rd("2016 5 irrig 2.8 1\n2016 5 fert 15 5 10\n2016 9 plant 1 2 3 4\n2017 1 harv 0.4 0.1 0.2 0.3 # note\n2017 2 till 0.2\n")
    assert rc == 0 and [e[2] for e in s["events"]] == [A.EV_IRRIGATION, A.EV_FERTILIZATION, A.EV_PLANTING, A.EV_HARVEST,
                                                       A.EV_TILLAGE]
    assert s["events"][3][4:] == (0.4, 0.1, 0.2, 0.3) and s["events"][0][3] == 1
    assert rd("2016 5 bogus 1\n")[0] == 4                         # unknown type
    assert rd("2016 5 irrig 2.8 1\n2016 4 irrig 2.8 1\n")[0] == 5   # out of order
    assert rd("2016 365 irrig 1 1\n2017 1 irrig 1 1\n")[0] == 0     # year boundary is in order
    assert rd("2016 5 harv 0.8 0.1 0.3 0.1\n")[0] == 3             # removed + transferred > 1
    assert rd("2016 5 harv 0.8 0.1\n")[0] == 5                     # all four harvest fractions are required
    assert rd("2016 5 irrig 2.8\n")[0] == 5
    assert rd("2016 irrig\n")[0] == 5
    assert rd("2016 5 plantdeath\n")[0] == 5
    assert rd("2016 5 leafon\n2016 200 leafoff\n")[0] == 0
    assert rd("2016 5 leafon 3\n")[0] == 5                         # leaf events take no numbers
    assert rd("2016 5 leafon\n", flags_c(dict(gdd=1)))[0] == 3      # conflicts with computed phenology
    pl = p.copy(); pl[A.P["leafOnDay"]] = 144
    assert read_site_c(lib, clim, 0, write(tmp_path / "ev2.in", "2016 5 leafoff\n"), fl, pl)[0] == 3
    assert read_site_c(lib, clim, 0, str(tmp_path / "missing.in"), fl, p)[0] == 0   # no file = no events


def test_config_precedence_validation_and_dump(tmp_path, smoke_dir):
    """The drop-in driver itself: sipnet.in + CLI precedence, validation exit codes and the
    <prefix>.config dump (runs up to the point where it needs a GPU)."""
    if not os.path.exists(DRIVER):
        pytest.skip("sipnet_gpu driver not built")
    d = os.path.join(smoke_dir, "russell_2")
    r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--no-snow"], cwd=d, capture_output=True, text=True)
    cfg = open(os.path.join(d, "sipnet.config")).read().splitlines()
    assert cfg[1].split() == ["Name", "Source", "Value"]
    rows = {ln.split()[0]: ln.split()[1:] for ln in cfg[2:]}
    assert rows["ANAEROBIC"] == ["INPUT_FILE", "1"] and rows["SNOW"] == ["COMMAND_LINE", "0"]
    assert rows["QUIET"] == ["COMMAND_LINE", "1"]             # CLI beats the file's QUIET = 0
    assert rows["FILE_PREFIX"] == ["INPUT_FILE", "sipnet"] and rows["PARAM_FILE"] == ["CALCULATED", "sipnet.param"]
    assert rows["GDD"] == ["DEFAULT", "1"] and rows["RESTART_IN"] == ["DEFAULT"]
    assert [ln.split()[0] for ln in cfg[2:]] == sorted(rows, key=lambda k: k.replace("_", "").lower().replace(
        "dosingleoutput", "dosingleoutputs"))
    # validation: nitrogen cycle needs litter pool + anaerobic -> exit 3 (context.c:195-223)
    assert subprocess.run([DRIVER, "-i", "sipnet.in", "--no-anaerobic"], cwd=d, capture_output=True).returncode == 3
    assert subprocess.run([DRIVER, "-i", "sipnet.in", "--soil-phenol"], cwd=d, capture_output=True).returncode == 3
    assert subprocess.run([DRIVER, "--bogus-flag"], cwd=d, capture_output=True).returncode == 8
    assert subprocess.run([DRIVER, "-i", "nope.in"], cwd=d, capture_output=True).returncode == 6
    bad = write(tmp_path / "bad.in", "RUNTYPE = mcmc\n")
    assert subprocess.run([DRIVER, "-i", bad], cwd=d, capture_output=True).returncode == 3
    bad2 = write(tmp_path / "bad2.in", "EVENTS = yes\n")
    assert subprocess.run([DRIVER, "-i", bad2], cwd=d, capture_output=True).returncode == 3


def test_readers_agree_with_live_reference(lib, refshim, smoke_dir):
    for case in SMOKE:
        g = Golden("smoke_" + case)
        d = os.path.join(smoke_dir, case)
        ref = refshim.read_clim(os.path.join(d, "sipnet.clim"), g.flags["gdd"])
        rc, mine = read_site_c(lib, os.path.join(d, "sipnet.clim"), g.flags["gdd"])
        assert rc == 0
        for k in A.CLIM_COLS:
            assert np.array_equal(mine[k], ref.clim[k])
        assert np.array_equal(refshim.read_params(os.path.join(d, "sipnet.param"), g.flags), g.params)


def test_driver_loads_many_sites_identically_on_any_thread_count(smoke_dir, tmp_path):
    """The drop-in driver reads the sites of a --site-list launch on several host threads (the readers are
    re-entrant).  SIPNET_GPU_TRACE_INPUTS prints a checksum of everything loaded: it must not depend on the thread
    count, and the first failing site in list order decides the exit code.  Without a GPU the driver stops at
    sipnet_gpu_init (exit 100), after the inputs: that is all this test needs."""
    import shutil
    import subprocess
    if not os.path.exists(DRIVER):
        pytest.skip("sipnet_gpu driver not built")
    src = os.path.join(smoke_dir, "russell_2")
    work = str(tmp_path / "multi")
    os.makedirs(work)
    shutil.copy(os.path.join(src, "sipnet.in"), work)
    clim = open(os.path.join(src, "sipnet.clim")).read().splitlines()
    base = open(os.path.join(src, "sipnet.param")).read()
    sites = []
    for n in range(10):
        d = os.path.join(work, f"site{n}")
        os.makedirs(d)
        open(os.path.join(d, "sipnet.clim"), "w").write("\n".join(clim[: 400 + 150 * n]) + "\n")   # ragged lengths
        open(os.path.join(d, "sipnet.param"), "w").write(base.replace("soilWHC 12", f"soilWHC {9 + n}.5"))
        shutil.copy(os.path.join(src, "events.in"), d)
        sites.append(d)
    members = []
    for k in range(3):
        p = os.path.join(sites[4], f"m{k}.param")
        open(p, "w").write(base.replace("aMax 53.2895432752984", f"aMax {40 + k}.0"))
        members.append(p)
    open(os.path.join(sites[4], "members.txt"), "w").write("\n".join(members) + "\n")
    with open(os.path.join(work, "sites.txt"), "w") as f:
        for n, d in enumerate(sites):
            f.write(f"{d}/sipnet {d}/events {d}/members.txt\n" if n == 4 else f"{d}/sipnet\n")

    def run(threads):
        env = dict(os.environ, SIPNET_GPU_TRACE_INPUTS="1", SIPNET_GPU_READER_THREADS=str(threads),
                   CUDA_VISIBLE_DEVICES="")                     # stop at init on a GPU box too
        r = subprocess.run([DRIVER, "-i", "sipnet.in", "--quiet", "--no-do-main-output", "--site-list", "sites.txt"], cwd=work,
                           capture_output=True, text=True, env=env)
        trace = [l for l in r.stderr.splitlines() if l.startswith("[TRACE  ] inputs:")]
        return r.returncode, trace, r.stdout

    rc1, t1, _ = run(1)
    assert rc1 == 100 and len(t1) == 1 and "10 site(s), 12 member(s)" in t1[0], (rc1, t1)
    for threads in (2, 4, 7):
        for _ in range(3):
            assert run(threads) == (rc1, t1, run(1)[2])
    # two broken sites: the first one in list order decides, whatever the thread count
    open(os.path.join(sites[6], "sipnet.clim"), "w").write("\n".join(clim[:50] + ["2016 x y z"] + clim[50:90]) + "\n")
    os.remove(os.path.join(sites[2], "sipnet.param"))
    for threads in (1, 4, 7):
        rc, trace, out = run(threads)
        assert rc == 6 and not trace and "site2/sipnet.param" in out, (threads, rc, out)
