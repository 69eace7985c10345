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
