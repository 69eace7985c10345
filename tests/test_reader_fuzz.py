"""CPU: differential fuzzing of the host readers against the LIVE reference readers (oracle/_ref).

Seeded random mutations of the reference's smoke inputs (.param, .clim, events.in): dropped / extra / garbled
tokens, swapped and duplicated lines, blank lines, comments, sign flips, out-of-range values.  For every
mutant both readers must agree on the exit code and -- when the file is accepted -- on every parsed value, bit
for bit (readParamData sipnet.c:290-427 + modelParams.c, readClimData sipnet.c:128-277, readEventData
events.c:263-367).  Some malformed files make the reference itself die of SIGSEGV (e.g. a .param line with a single
token); each mutant is therefore first tried on the reference BINARY in a scratch directory, and where that
crashes the only requirement on our reader is a clean rejection.  Skipped where the reference build is absent."""
import ctypes as C
import os
import random
import shutil
import subprocess

import numpy as np
import pytest

from conftest import Golden, ROOT
from host_util import flags_c, host_lib, read_site_c, unpack_smoke
from sipnet_b200 import _abi as A

N_MUTANTS = 150
GARBAGE = ["nan", "inf", "-", "1e999", "abc", "1.2.3", "0x10", "*", "!", "#", "", "1e-400", "-0", "+5", "1,5", "2016.0", "99999999999"]


def mutate(lines, rng, keep_first=0):
    """one random edit of a list of text lines"""
    lines = list(lines)
    if not lines:
        return lines
    op = rng.choice(["drop_tok", "add_tok", "garble", "swap", "dup", "blank", "comment", "negate", "drop_line", "scale", "trail"])
    i = rng.randrange(keep_first, len(lines)) if len(lines) > keep_first else 0
    tok = lines[i].split()
    if op == "drop_tok" and tok:
        tok.pop(rng.randrange(len(tok)))
        lines[i] = " ".join(tok) + "\n"
    elif op == "add_tok":
        tok.insert(rng.randrange(len(tok) + 1), rng.choice(["7", "x", "0.5", "-1"]))
        lines[i] = " ".join(tok) + "\n"
    elif op == "garble" and tok:
        tok[rng.randrange(len(tok))] = rng.choice(GARBAGE)
        lines[i] = " ".join(tok) + "\n"
    elif op == "swap" and len(lines) > 1:
        j = rng.randrange(len(lines))
        lines[i], lines[j] = lines[j], lines[i]
    elif op == "dup":
        lines.insert(i, lines[i])
    elif op == "blank":
        lines.insert(i, rng.choice(["\n", "   \n", "\t\n"]))
    elif op == "comment":
        lines[i] = rng.choice(["!", "#", "//", "; "]) + lines[i]
    elif op == "negate" and tok:
        k = rng.randrange(len(tok))
        tok[k] = tok[k][1:] if tok[k].startswith("-") else "-" + tok[k]
        lines[i] = " ".join(tok) + "\n"
    elif op == "drop_line":
        lines.pop(i)
    elif op == "scale" and tok:
        k = rng.randrange(len(tok))
        try:
            tok[k] = repr(float(tok[k]) * rng.choice([0.0, 1e-9, 1e9, -1.0, 366.0]))
        except ValueError:
            pass
        lines[i] = " ".join(tok) + "\n"
    elif op == "trail":
        lines[i] = lines[i].rstrip("\n") + rng.choice(["   ", " # note", " ! note", "\r"]) + "\n"
    return lines


@pytest.fixture(scope="module")
def lib():
    return host_lib()


@pytest.fixture(scope="module")
def smoke_dir(tmp_path_factory):
    return unpack_smoke(str(tmp_path_factory.mktemp("smoke")))


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sipnet_ref")


class Scratch:
    """A directory the reference binary can run in: the case's sipnet.in, a few climate records, its events."""

    def __init__(self, smoke_dir, case, root):
        self.dir = os.path.join(root, "scratch_" + case)
        os.makedirs(self.dir, exist_ok=True)
        src = os.path.join(smoke_dir, case)
        shutil.copy(os.path.join(src, "sipnet.in"), self.dir)
        shutil.copy(os.path.join(src, "sipnet.param"), self.dir)
        shutil.copy(os.path.join(src, "events.in"), self.dir)
        open(os.path.join(self.dir, "sipnet.clim"), "w").writelines(open(os.path.join(src, "sipnet.clim")).readlines()[:16])

    def crashes(self, **files):
        """does the reference binary die of a signal with these files swapped in?"""
        for name, path in files.items():
            shutil.copy(path, os.path.join(self.dir, name))
        r = subprocess.run([REF_BIN, "-i", "sipnet.in", "--quiet", "--no-dump-config"], cwd=self.dir, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL)
        return r.returncode < 0


def ref_rc(fn, *args):
    try:
        return 0, fn(*args)
    except RuntimeError as e:                       # "... exit code N"
        return int(str(e).rsplit(" ", 1)[1]), None


@pytest.mark.parametrize("case", ["niwot", "russell_2"])
def test_param_reader_agrees_with_reference_on_mutants(lib, refshim, smoke_dir, tmp_path, case):
    g = Golden("smoke_" + case)
    src = open(os.path.join(smoke_dir, case, "sipnet.param")).readlines()
    rng = random.Random(101)
    fl = flags_c(g.flags)
    path = str(tmp_path / "m.param")
    scratch = Scratch(smoke_dir, case, str(tmp_path))
    accepted = crashed = 0
    for n in range(N_MUTANTS):
        lines = src
        for _ in range(rng.choice([1, 1, 2, 3])):
            lines = mutate(lines, rng)
        open(path, "w").writelines(lines)
        got = np.zeros(A.NPARAMS)
        rc = lib.sip_read_params(path.encode(), C.byref(fl), 1, got.ctypes.data_as(C.POINTER(C.c_double)))
        if scratch.crashes(**{"sipnet.param": path}):
            crashed += 1
            assert rc != 0, (n, "the reference crashes on this file; it must at least be rejected")
            continue
        want_rc, want = ref_rc(refshim.read_params, path, g.flags)
        assert rc == want_rc, (n, [ln for ln in lines if ln not in src], lib.sip_host_error())
        if rc == 0:
            accepted += 1
            assert np.array_equal(got, want, equal_nan=True), n
    assert 10 < accepted < N_MUTANTS and crashed >= 0   # the mutants exercise both outcomes


@pytest.mark.parametrize("case,gdd", [("niwot", 1), ("russell_1", 1), ("russell_1", 0)])
def test_clim_reader_agrees_with_reference_on_mutants(lib, refshim, smoke_dir, tmp_path, case, gdd):
    src = open(os.path.join(smoke_dir, case, "sipnet.clim")).readlines()[:60]
    rng = random.Random(202)
    path = str(tmp_path / "m.clim")
    scratch = Scratch(smoke_dir, case, str(tmp_path))
    accepted = 0
    for n in range(N_MUTANTS):
        lines = src
        for _ in range(rng.choice([1, 1, 2])):
            lines = mutate(lines, rng)
        open(path, "w").writelines(lines)
        rc, got = read_site_c(lib, path, gdd)
        if scratch.crashes(**{"sipnet.clim": path}):
            assert rc != 0, n
            continue
        want_rc, want = ref_rc(refshim.read_clim, path, gdd)
        assert rc == want_rc, (n, [ln for ln in lines if ln not in src], lib.sip_host_error())
        if rc == 0:
            accepted += 1
            assert np.array_equal(got["year"], want.year) and np.array_equal(got["day"], want.day), n
            for k in A.CLIM_COLS:
                assert np.array_equal(got[k], want.clim[k], equal_nan=True), (n, k)
    assert 10 < accepted


@pytest.mark.parametrize("case", ["russell_1", "russell_2"])
def test_event_reader_agrees_with_reference_on_mutants(lib, refshim, smoke_dir, tmp_path, case):
    g = Golden("smoke_" + case)
    d = os.path.join(smoke_dir, case)
    src = open(os.path.join(d, "events.in")).readlines()
    extra = ["2016 200 plant 10 3 2 5\n", "2016 210 harv 0.8 0.0 0.2 1.0\n", "2016 100 till 0.2\n", "2017 5 irrig 3.0 0\n",
             "2016 120 leafon\n", "2016 300 leafoff\n"]
    rng = random.Random(303)
    fl = flags_c(g.flags)
    params = refshim.read_params(os.path.join(d, "sipnet.param"), g.flags)   # also sets the reference's flags / params
    clim = os.path.join(d, "sipnet.clim")
    path = str(tmp_path / "events.in")
    scratch = Scratch(smoke_dir, case, str(tmp_path))
    accepted = 0
    for n in range(N_MUTANTS):
        lines = list(src)
        if rng.random() < 0.5:
            lines.insert(rng.randrange(len(lines) + 1), rng.choice(extra))
        for _ in range(rng.choice([0, 1, 1, 2])):
            lines = mutate(lines, rng)
        open(path, "w").writelines(lines)
        rc, got = read_site_c(lib, clim, g.flags["gdd"], path, fl, params)
        if scratch.crashes(**{"events.in": path}):
            assert rc != 0, n
            continue
        want_rc, want = ref_rc(refshim.read_events, path)
        if want_rc == 0 and rc == 5:
            continue        # first event before the first climate record: the reference checks that later (frontend.c:217-222)
        assert rc == want_rc, (n, lines, lib.sip_host_error())
        if rc == 0:
            accepted += 1
            assert got["events"] == want, n
    assert 10 < accepted


def test_config_parser_agrees_with_reference_on_mutants(smoke_dir, tmp_path):
    """sipnet.in (frontend.c:35-128, context.c) through the two BINARIES: same exit code and, when the file is
    accepted, the same <prefix>.config dump (bar the time-stamp line).  Without a GPU our driver stops at device
    initialisation (exit 100) -- after the configuration has been parsed, validated and dumped."""
    from host_util import DRIVER
    if not (os.path.exists(DRIVER) and os.path.exists(REF_BIN)):
        pytest.skip("needs both binaries")
    src = [ln for ln in open(os.path.join(smoke_dir, "russell_2", "sipnet.in")) if ln.strip() and not ln.lstrip().startswith("!")]
    src += ["GDD = 0\n", "SOIL_PHENOL: 0\n", "growth resp 1\n", "Leaf-Water = 1\n", "FILE_PREFIX = sipnet\n", "RUNTYPE = standard\n",
            "UNKNOWN_KEY = 3\n", "EVENTS_PREFIX = events\n", "WATER_HRESP = 1 ! trailing comment\n"]
    rng = random.Random(404)
    dirs = {}
    for who in ("ref", "ours"):
        dirs[who] = str(tmp_path / who)
        os.makedirs(dirs[who])
        for fn in ("sipnet.param", "events.in"):
            shutil.copy(os.path.join(smoke_dir, "russell_2", fn), dirs[who])
        open(os.path.join(dirs[who], "sipnet.clim"), "w").writelines(
            open(os.path.join(smoke_dir, "russell_2", "sipnet.clim")).readlines()[:16])
    accepted = rejected = 0
    for n in range(120):
        lines = list(src)
        rng.shuffle(lines)
        for _ in range(rng.choice([0, 1, 1, 2, 3])):
            lines = mutate(lines, rng)
        extra = rng.choice([[], [], ["--no-snow"], ["--flooding"], ["--carbon-saturation"], ["--no-litter-pool"], ["--gdd", "--soil-phenol"],
                            ["-f", "sipnet"], ["--no-print-header"], ["--file-name", "sipnet"]])
        codes, dumps = {}, {}
        for who, binary in (("ref", REF_BIN), ("ours", DRIVER)):
            open(os.path.join(dirs[who], "sipnet.in"), "w").writelines(lines)
            cfg = os.path.join(dirs[who], "sipnet.config")
            if os.path.exists(cfg):
                os.remove(cfg)
            r = subprocess.run([binary, "-i", "sipnet.in", "--dump-config", *extra], cwd=dirs[who], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)
            codes[who] = r.returncode
            dumps[who] = open(cfg).read().splitlines()[1:] if os.path.exists(cfg) else None
        if codes["ref"] < 0:
            assert codes["ours"] not in (0, 100), (n, lines)
            continue
        if codes["ref"] == 0:
            accepted += 1
            assert codes["ours"] in (0, 100), (n, codes, lines, extra)
            assert dumps["ours"] == dumps["ref"], (n, lines, extra)
        else:
            rejected += 1
            assert codes["ours"] == codes["ref"], (n, codes, lines, extra)
    assert accepted > 10 and rejected > 3


def _read_clim_arrays(lib, path, gdd):
    from host_util import SiteDataC
    s = SiteDataC()
    rc = lib.sip_read_clim(path.encode(), gdd, 1, C.byref(s))
    out = None
    if rc == 0:
        n = s.nsteps
        out = [np.ctypeslib.as_array(s.year, (n,)).copy(), np.ctypeslib.as_array(s.day, (n,)).copy()]
        out += [np.ctypeslib.as_array(getattr(s, k), (n,)).copy() for k in A.CLIM_COLS]
        lib.sip_site_free(C.byref(s))
    return rc, out, lib.sip_host_error()


NUMBER_SPELLINGS = ["0", "-0", "+0", "0.0", "-0.000", "5", "+5", "-5", "5.", ".5", "-.5", "0.5", "12.345", "1e3", "1E3", "1e+3",
                    "1.5e-3", "-2.5E-07", "123456789012345", "1234567890123456", "0.000000000000001", "9007199254740993",
                    "1e22", "1e23", "1e-22", "1e-23", "123456789012345e7", "123456789012345e8", "4.9e-324", "1e-400",
                    "1e400", "1.7976931348623157e308", "0.1", "0.2", "0.3", "2.675", "1.005", "8.41e21", "00012.50",
                    "3.14159265358979323846264338327950288", "1e0005", "1e00005", "0.30000000000000004", "100000000000000000000000"]


def test_clim_fast_path_equals_the_scanf_loop(tmp_path, monkeypatch):
    """The one-record-per-line fast path of sip_read_clim against its own fscanf loop (SIPNET_HOST_SCANF_CLIM): same
    exit code, message and bits on files that spell numbers every way the fast path accepts or refuses, with
    records split over lines, glued lines, blank lines, CR LF, tabs, trailing blanks, a missing final newline, an
    over-long line, and seeded mutants."""
    lib = host_lib()
    rng = random.Random(20260117)

    def record(year, day):
        return [str(year), str(day)] + [rng.choice(NUMBER_SPELLINGS) if rng.random() < 0.5
                                        else repr(round(rng.uniform(-50, 3000), rng.randrange(0, 12))) for _ in range(10)]

    def both(text, gdd):
        path = str(tmp_path / "f.clim")
        with open(path, "w", newline="") as f:
            f.write(text)
        monkeypatch.delenv("SIPNET_HOST_SCANF_CLIM", raising=False)
        fast = _read_clim_arrays(lib, path, gdd)
        monkeypatch.setenv("SIPNET_HOST_SCANF_CLIM", "1")
        slow = _read_clim_arrays(lib, path, gdd)
        monkeypatch.delenv("SIPNET_HOST_SCANF_CLIM", raising=False)
        assert fast[0] == slow[0] and fast[2] == slow[2], (fast[0], slow[0], fast[2], slow[2])
        if fast[0] == 0:
            for a, b in zip(fast[1], slow[1]):
                assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))
        return fast[0]

    accepted = 0
    for trial in range(120):
        recs = [record(2016, 1 + i // 2) for i in range(rng.randrange(1, 60))]
        lines = [" ".join(r) + "\n" for r in recs]
        style = trial % 12
        if style == 1:
            lines = ["\t".join(r) + "  \t\n" for r in recs]                      # tabs, trailing blanks
        elif style == 2:
            lines = [" ".join(r) + "\r\n" for r in recs]                         # CR LF
        elif style == 3 and len(recs) > 2:
            k = rng.randrange(1, len(recs))
            lines[k] = " ".join(recs[k][:5]) + "\n" + " ".join(recs[k][5:]) + "\n"  # one record over two lines
        elif style == 4 and len(recs) > 2:
            k = rng.randrange(1, len(recs) - 1)
            lines[k] = lines[k].rstrip("\n") + " " + lines.pop(k + 1)           # two records on one line
        elif style == 5:
            lines.insert(rng.randrange(1, len(lines) + 1), rng.choice(["\n", "  \n", "\t\r\n"]))
        elif style == 6:
            lines[-1] = lines[-1].rstrip("\n")                                   # no final newline
        elif style == 7 and len(recs) > 1:
            k = rng.randrange(1, len(recs))
            lines[k] = " " * 1100 + lines[k]                                     # longer than the line buffer
        elif style == 8 and len(recs) > 1:
            k = rng.randrange(1, len(recs))
            tok = recs[k][:]
            tok[rng.randrange(12)] = rng.choice(GARBAGE + ["1e", "1e+", "0x1p3", "1d5", "5.5.5", "--5", "1_000", "2016.5"])
            lines[k] = " ".join(tok) + "\n"
        elif style >= 9:
            for _ in range(rng.randrange(1, 4)):
                lines = mutate(lines, rng, keep_first=0)
        accepted += both("".join(lines), trial % 2) == 0
    assert 40 < accepted < 120          # both outcomes are exercised
