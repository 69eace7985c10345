"""ctypes access to the CHECKERS -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this module.  The product (sipnet_b200/) never does.

Two checkers:
  * RefShim  -- oracle/_ref/libsipnet_refshim.so: the UNMODIFIED reference
                sources (compiled from /root/reference by oracle/Makefile) behind
                an in-memory harness (oracle/ref_shim.c).
  * Oracle   -- oracle/libsipnet_oracle.so: our plain-C restatement
                (oracle/sipnet_oracle.c), which lives in git and always exists.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sipnet_b200 import _abi as A
from sipnet_b200.api import SiteData, flags_array

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "sipnet_ref")
REF_SHIM = os.path.join(REF_DIR, "libsipnet_refshim.so")
ORACLE_LIB = os.path.join(HERE, "libsipnet_oracle.so")

_dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
_ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))


def build_oracle(force: bool = False) -> str:
    """Compile the restatement (gcc, seconds)."""
    src = os.path.join(HERE, "sipnet_oracle.c")
    if force or not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "all"])
    return ORACLE_LIB


def build_ref(reference: str = "/root/reference") -> bool:
    """Compile the reference checkers if the reference tree is present."""
    if not os.path.isdir(os.path.join(reference, "src")):
        return have_ref()
    subprocess.check_call(["make", "-C", HERE, "-s", "ref", f"REF={reference}"])
    return True


def have_ref() -> bool:
    return os.path.exists(REF_SHIM) and os.path.exists(REF_BIN)


class _Runner:
    """Shared argument marshalling for both checkers (same call shape)."""

    def _site_args(self, site: SiteData):
        arr, n = site.event_array()
        return ([C.c_int64(site.nsteps), _ip(site.year), _ip(site.day)]
                + [_dp(site.clim[k]) for k in A.CLIM_COLS]
                + [C.c_int64(n), C.cast(arr, C.POINTER(A.Event))], arr)


class RefShim(_Runner):
    def __init__(self, path: str = REF_SHIM):
        self.lib = C.CDLL(path)
        self.lib.sipref_run.restype = C.c_int
        self.lib.sipref_read_clim.restype = C.c_int64
        self.lib.sipref_read_events.restype = C.c_int64
        self.lib.sipref_read_params.restype = C.c_int
        self.lib.sipref_param_offset.restype = C.c_int
        self.lib.sipref_param_offset.argtypes = [C.c_char_p]

    def run(self, flags: dict, params80: np.ndarray, site: SiteData, want_debug=True,
            events_out: str | None = None, main_out: str | None = None, print_header=1):
        """-> (rc, steps_done, out32[T][32], dbg[T][106] or None)"""
        T = site.nsteps
        fl = flags_array(flags)
        p = np.ascontiguousarray(params80, dtype=np.float64)
        out = np.full((T, A.NOUT), np.nan)
        dbg = np.full((T, A.NDEBUG), np.nan) if want_debug else None
        done = C.c_int64(0)
        sargs, keep = self._site_args(site)
        rc = self.lib.sipref_run(
            _ip(fl), _dp(p), *sargs,
            events_out.encode() if events_out else None,
            main_out.encode() if main_out else None, C.c_int(print_header),
            _dp(out), _dp(dbg) if dbg is not None else None, C.byref(done))
        return rc, int(done.value), out, dbg

    def run_balance(self, flags: dict, params80: np.ndarray, site: SiteData):
        """-> (rc, steps_done, balance[T][2]): balanceTracker.deltaC / deltaN after every updateState()
        (checkBalance(), balance.c:129-148) of the unmodified reference."""
        bal = np.full((site.nsteps, 2), np.nan)
        self.lib.sipref_set_balance_out(_dp(bal), C.c_int64(site.nsteps))
        try:
            rc, done, _, _ = self.run(flags, params80, site, want_debug=False)
        finally:
            self.lib.sipref_set_balance_out(None, C.c_int64(0))
        return rc, done, bal

    def read_clim(self, path: str, gdd: int = 1, cap: int = 1 << 20) -> SiteData:
        n = sum(1 for _ in open(path))
        cap = max(n + 8, 16)
        year = np.zeros(cap, np.int32)
        day = np.zeros(cap, np.int32)
        cols = np.zeros((11, cap))
        got = self.lib.sipref_read_clim(path.encode(), C.c_int(gdd), C.c_int64(cap), _ip(year), _ip(day), _dp(cols))
        if got < 0:
            raise RuntimeError(f"reference readClimData exit code {-got}")
        return SiteData(year[:got].copy(), day[:got].copy(),
                        {k: cols[i, :got].copy() for i, k in enumerate(A.CLIM_COLS)})

    def read_params(self, path: str, flags: dict) -> np.ndarray:
        out = np.zeros(A.NPARAMS)
        rc = self.lib.sipref_read_params(path.encode(), _ip(flags_array(flags)), _dp(out))
        if rc:
            raise RuntimeError(f"reference readParamData exit code {rc}")
        return out

    def read_events(self, path: str) -> list:
        cap = sum(1 for _ in open(path)) + 4 if os.path.exists(path) else 4
        buf = (A.Event * cap)()
        n = self.lib.sipref_read_events(path.encode(), C.c_int64(cap), buf)
        if n < 0:
            raise RuntimeError(f"reference readEventData exit code {-n}")
        return [(e.year, e.day, e.type, e.method, e.p[0], e.p[1], e.p[2], e.p[3]) for e in buf[:n]]

    def param_offset(self, name: str) -> int:
        return self.lib.sipref_param_offset(name.encode())


class Oracle(_Runner):
    def __init__(self, path: str | None = None):
        self.lib = C.CDLL(path or build_oracle())
        self.lib.sipnet_oracle_run.restype = C.c_int

    def run(self, flags: dict, params80: np.ndarray, site: SiteData, want_debug=True,
            max_event_records: int = 0):
        """-> (rc, steps_done, out32[T][32], dbg[T][106] or None, records)"""
        T = site.nsteps
        fl = flags_array(flags)
        p = np.ascontiguousarray(params80, dtype=np.float64)
        out = np.full((T, A.NOUT), np.nan)
        dbg = np.full((T, A.NDEBUG), np.nan) if want_debug else None
        done = C.c_int64(0)
        nrec = C.c_int32(0)
        recs = (A.EventRecord * max(max_event_records, 1))()
        sargs, keep = self._site_args(site)
        rc = self.lib.sipnet_oracle_run(
            _ip(fl), _dp(p), *sargs, _dp(out), _dp(dbg) if dbg is not None else None,
            C.byref(done), recs, C.c_int32(max_event_records), C.byref(nrec))
        return rc, int(done.value), out, dbg, list(recs[:min(nrec.value, max_event_records)])

    def run_balance(self, flags: dict, params80: np.ndarray, site: SiteData):
        """-> (rc, steps_done, balance[T][2]): the mass-balance tracker's deltaC / deltaN per step."""
        T = site.nsteps
        fl = flags_array(flags)
        p = np.ascontiguousarray(params80, dtype=np.float64)
        bal = np.full((T, 2), np.nan)
        done = C.c_int64(0)
        clim = (C.POINTER(C.c_double) * 11)(*[_dp(site.clim[k]) for k in A.CLIM_COLS])
        arr, n = site.event_array()
        self.lib.sipnet_oracle_run_diag.restype = C.c_int
        info = C.c_uint32(0)
        rc = self.lib.sipnet_oracle_run_diag(_ip(fl), _dp(p), C.c_int64(T), _ip(site.year), _ip(site.day), clim,
                                             C.c_int64(n), C.cast(arr, C.POINTER(A.Event)), _dp(bal), C.byref(done),
                                             C.byref(info))
        self.last_info = int(info.value)         # SIPNET_GPU_ST_*_LIMITED bits of the run
        cnt = (C.c_uint32 * A.NCOUNTERS)()
        self.lib.sipnet_oracle_last_counts(cnt)
        self.last_counts = np.array(list(cnt), np.uint32)   # how often each message occurred (CNT_* order)
        return rc, int(done.value), bal
