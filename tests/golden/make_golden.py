#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle ref`):

    python tests/golden/make_golden.py

For every case it stores, in tests/golden/<case>.npz:
  inputs   flags[12], params[80] (as left by the reference's readParamData),
           year/day/clim[11][T] (as left by the reference's readClimData),
           events[n][8] (year, day, type, method, p0..p3 from readEventData)
  outputs  produced by the reference's own setupModel()/updateState() loop via
           oracle/ref_shim.c, at full double precision:
           rows      step indices kept (first 48, every 16th, last 8)
           out32     [len(rows)][32]  outputState() columns
           dbg       [len(rows)][106] every Envi/Fluxes/Trackers field
           colsum    [32]  sum over ALL steps of each output column (float64, step order)
           colabs    [32]  max |value| over all steps
           nsteps, rc
  text     the reference's events.out (exact bytes) and an md5 of its sipnet.out

Cases: the reference's four active smoke cases (tests/smoke/{niwot,russell_1..3})
and synthetic cases that reach the branches the smoke goldens miss
(planting/harvest/tillage/canopy irrigation/mortality/N limitation).
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyoracle import RefShim  # noqa: E402
from sipnet_b200 import _abi as A, synth  # noqa: E402
from sipnet_b200.api import flags_array  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

SMOKE = {
    "niwot": {},
    "russell_1": {},
    "russell_2": dict(litterPool=1, nitrogenCycle=1, anaerobic=1),
    "russell_3": dict(growthResp=1, leafWater=1, litterPool=1, waterHResp=0),
}


def keep_rows(T: int) -> np.ndarray:
    rows = set(range(min(48, T))) | set(range(0, T, 16)) | set(range(max(T - 8, 0), T))
    return np.array(sorted(rows), dtype=np.int64)


def events_matrix(events) -> np.ndarray:
    return np.array([list(map(float, e)) for e in events], dtype=np.float64).reshape(-1, 8)


def run_case(shim: RefShim, name: str, flags: dict, params: np.ndarray, site, print_header=1):
    with tempfile.TemporaryDirectory() as td:
        ev_out = os.path.join(td, "events.out")
        main_out = os.path.join(td, "sipnet.out")
        rc, done, out, dbg = shim.run(flags, params, site, events_out=ev_out, main_out=main_out,
                                      print_header=print_header)
        ev_text = open(ev_out, "rb").read() if os.path.exists(ev_out) else b""
        main_bytes = open(main_out, "rb").read()
        main_md5 = hashlib.md5(main_bytes).hexdigest()
        main_lines = main_bytes.split(b"\n")
    rows = keep_rows(done)
    hdr = 1 if print_header else 0
    out_text = b"\n".join(main_lines[hdr + int(r)] for r in rows) + b"\n"   # the reference's own text rows
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        flags=flags_array(flags), params=params, year=site.year, day=site.day,
        clim=np.stack([site.clim[k] for k in A.CLIM_COLS]), events=events_matrix(site.events),
        rows=rows, out32=out[rows], dbg=dbg[rows], colsum=out[:done].sum(axis=0),
        colabs=np.abs(out[:done]).max(axis=0), nsteps=np.int64(done), rc=np.int64(rc),
        print_header=np.int64(print_header),
        events_out=np.frombuffer(ev_text, dtype=np.uint8), main_out_md5=np.frombuffer(main_md5.encode(), dtype=np.uint8),
        out_text=np.frombuffer(out_text, dtype=np.uint8))
    print(f"{name}: rc={rc} steps={done} rows={rows.size} events.out={len(ev_text)}B")


def pack_smoke_inputs():
    """Raw input files of the reference's smoke cases (data, not code) so the drop-in CLI can be
    run end to end where /root/reference does not exist (the GPU box)."""
    import io
    import tarfile
    path = os.path.join(OUT, "smoke_inputs.tar.gz")
    with tarfile.open(path, "w:gz", compresslevel=9) as tar:
        for name in SMOKE:
            d = os.path.join(REF, "tests", "smoke", name)
            for fn in ("sipnet.in", "sipnet.param", "sipnet.clim", "events.in"):
                if fn == "sipnet.clim" and name in ("russell_2", "russell_3"):
                    continue  # byte-identical to russell_1/sipnet.clim; tests copy that one
                data = open(os.path.join(d, fn), "rb").read()
                ti = tarfile.TarInfo(f"{name}/{fn}")
                ti.size = len(data)
                ti.mtime = 0
                tar.addfile(ti, io.BytesIO(data))
    print("wrote", path, os.path.getsize(path), "bytes")


def debug_log_md5():
    """md5 of the reference binary's --debug-log files for the smoke cases -> tests/golden/debug_log_md5.json"""
    import json
    import shutil
    import subprocess
    res = {}
    for name in SMOKE:
        with tempfile.TemporaryDirectory() as td:
            for fn in ("sipnet.in", "sipnet.param", "sipnet.clim", "events.in"):
                shutil.copy(os.path.join(REF, "tests", "smoke", name, fn), td)
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "sipnet_ref"), "-i", "sipnet.in", "--quiet",
                                   "--debug-log", "dbg"], cwd=td, stdout=subprocess.DEVNULL)
            res[name] = {k: hashlib.md5(open(os.path.join(td, f"dbg_{k}.log"), "rb").read()).hexdigest()
                         for k in ("envi", "fluxes", "trackers")}
    json.dump(res, open(os.path.join(OUT, "debug_log_md5.json"), "w"), indent=1)
    print("debug log md5:", res)


def single_outputs_md5():
    """md5 of the reference binary's --do-single-outputs files (<prefix>.NEE, .NEE_cum, .GPP, .GPP_cum) for the smoke
    cases -> tests/golden/single_outputs_md5.json"""
    import json
    import shutil
    import subprocess
    res = {}
    for name in SMOKE:
        with tempfile.TemporaryDirectory() as td:
            for fn in ("sipnet.in", "sipnet.param", "sipnet.clim", "events.in"):
                shutil.copy(os.path.join(REF, "tests", "smoke", name, fn), td)
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "sipnet_ref"), "-i", "sipnet.in", "--quiet",
                                   "--do-single-outputs"], cwd=td, stdout=subprocess.DEVNULL)
            res[name] = {k: hashlib.md5(open(os.path.join(td, "sipnet." + k), "rb").read()).hexdigest()
                         for k in ("NEE", "NEE_cum", "GPP", "GPP_cum")}
    json.dump(res, open(os.path.join(OUT, "single_outputs_md5.json"), "w"), indent=1)
    print("single outputs md5:", res)


def restart_golden():
    """A checkpoint written by the reference binary (russell_2 cut after 2016), the reference's verdict (exit
    code) on every tampered variant in tests/restart_cases.py, and md5s of what the reference produces when it
    resumes from the checkpoint -> tests/golden/restart_russell_2.ckpt, restart_cases.json"""
    import json
    import shutil
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from host_util import split_case, strip_volatile, unpack_smoke
    from restart_cases import CASES
    ref = os.path.join(ROOT, "oracle", "_ref", "sipnet_ref")
    res = {"exit_codes": {}}
    with tempfile.TemporaryDirectory() as td:
        smoke = unpack_smoke(os.path.join(td, "smoke"))
        a, b = os.path.join(td, "a"), os.path.join(td, "b")
        split_case(os.path.join(smoke, "russell_2"), a, b, 2016)
        subprocess.check_call([ref, "-i", "sipnet.in", "--quiet", "--restart-out", "ck"], cwd=a, stdout=subprocess.DEVNULL)
        text = open(os.path.join(a, "ck")).read()
        shutil.copy(os.path.join(a, "ck"), os.path.join(OUT, "restart_russell_2.ckpt"))
        for name, mutate in CASES.items():
            with open(os.path.join(b, "ck_in"), "w") as f:
                f.write(mutate(text))
            r = subprocess.run([ref, "-i", "sipnet.in", "--quiet", "--restart-in", "ck_in"], cwd=b, stdout=subprocess.DEVNULL)
            res["exit_codes"][name] = r.returncode
        subprocess.check_call([ref, "-i", "sipnet.in", "--quiet", "--restart-in", "../a/ck", "--restart-out", "ck2"], cwd=b,
                              stdout=subprocess.DEVNULL)
        res["segment1_out_md5"] = hashlib.md5(open(os.path.join(a, "sipnet.out"), "rb").read()).hexdigest()
        res["segment2_out_md5"] = hashlib.md5(open(os.path.join(b, "sipnet.out"), "rb").read()).hexdigest()
        res["segment2_events_md5"] = hashlib.md5(open(os.path.join(b, "events.out"), "rb").read()).hexdigest()
        res["segment2_checkpoint_md5"] = hashlib.md5(strip_volatile(open(os.path.join(b, "ck2"), "rb").read())).hexdigest()
    json.dump(res, open(os.path.join(OUT, "restart_cases.json"), "w"), indent=1)
    print("restart golden:", res)


def balance_fixtures(shim):
    """The fixture of the reference's mass-balance regression test (tests/sipnet/test_modeling/testBalance.c:
    balance.param + balance.clim, 40 steps of 0.125 d; litter + nitrogen + anaerobic, GDD phenology off): as is
    (model-computed leaf-on/off on days 47 and 49), without the litter pool, and with the test's leaf-on/leaf-off
    EVENT file -- for which leafOnDay/leafOffDay are set to 0 in a copy of the parameter file, because the reference's
    reader rejects leaf events next to day-based phenology (events.c:251-261)."""
    import re
    d = os.path.join(REF, "tests", "sipnet", "test_modeling")
    on = dict(litterPool=1, nitrogenCycle=1, gdd=0, waterHResp=1, anaerobic=1)
    with tempfile.TemporaryDirectory() as td:
        text = open(os.path.join(d, "balance.param")).read()
        text0 = re.sub(r"^(leafOnDay|leafOffDay)\s+\S+", r"\1 0", text, flags=re.M)
        pfile0 = os.path.join(td, "balance_events.param")
        open(pfile0, "w").write(text0)
        for name, flags, pfile, evfile in (("balance_plain", on, os.path.join(d, "balance.param"), None),
                                           ("balance_no_litter_pool", dict(gdd=0, waterHResp=1), os.path.join(d, "balance.param"), None),
                                           ("balance_leaf_events", on, pfile0, "events_leaf.in")):
            full = dict(A.DEFAULT_FLAGS)
            full.update(flags)
            site = shim.read_clim(os.path.join(d, "balance.clim"), full["gdd"])
            params = shim.read_params(pfile, full)
            site.events = shim.read_events(os.path.join(d, evfile)) if evfile else []
            run_case(shim, name, full, params, site)


def main():
    pack_smoke_inputs()
    debug_log_md5()
    single_outputs_md5()
    restart_golden()
    shim = RefShim()
    for name, flags in SMOKE.items():
        d = os.path.join(REF, "tests", "smoke", name)
        full = dict(A.DEFAULT_FLAGS)
        full.update(flags)
        site = shim.read_clim(os.path.join(d, "sipnet.clim"), full["gdd"])
        site.events = shim.read_events(os.path.join(d, "events.in"))
        params = shim.read_params(os.path.join(d, "sipnet.param"), full)
        run_case(shim, "smoke_" + name, full, params, site, print_header=0 if name == "niwot" else 1)
    balance_fixtures(shim)
    # synthetic: C3-style event schedule, two members of the wide prior + the anchor, 3 years
    for variant in ("half-daily", "unequal"):
        site = synth.synth_site(3, 3, variant, with_events=True)
        P = synth.synth_params(8, stream=3)
        for m in (0, 1, 5):
            run_case(shim, f"synth_{variant.replace('-', '')}_m{m}", synth.SYNTH_FLAGS, P[:, m], site)


if __name__ == "__main__":
    main()
