"""ctypes access to the host C library (sipnet_b200/libsipnet_host.so) for tests."""
import ctypes as C
import os
import shutil
import subprocess
import tarfile

import numpy as np

from conftest import GOLDEN_DIR, ROOT
from sipnet_b200 import _abi as A

HOST_LIB = os.environ.get("SIPNET_HOST_LIB") or os.path.join(ROOT, "sipnet_b200", "libsipnet_host.so")   # override: sanitizer builds
DRIVER = os.path.join(ROOT, "sipnet_b200", "sipnet_gpu")
NAME_MAX = 256


class SiteDataC(C.Structure):
    _fields_ = ([("nsteps", C.c_int64), ("year", C.POINTER(C.c_int32)), ("day", C.POINTER(C.c_int32))]
                + [(n, C.POINTER(C.c_double)) for n in A.CLIM_COLS]
                + [("nevents", C.c_int64), ("events", C.POINTER(A.Event))])


class ContextC(C.Structure):
    _fields_ = ([("flags", A.Flags)]
                + [(n, C.c_int32) for n in ("doMainOutput", "doSingleOutputs", "dumpConfig", "printHeader", "quiet")]
                + [(n, C.c_char * NAME_MAX) for n in ("paramFile", "climFile", "outFile", "outConfigFile", "eventsPrefix",
                                                      "eventsInFile", "eventsOutFile", "inputFile", "restartIn",
                                                      "restartOut", "debugLogPrefix", "filePrefix")]
                + [("source", C.c_int32 * 32), ("ensembleParamList", C.c_char * NAME_MAX),
                   ("validationMath", C.c_int32), ("helpOrVersion", C.c_int32), ("siteList", C.c_char * NAME_MAX)])


def host_lib():
    if not os.path.exists(HOST_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "sipnet_b200", "host"), "-s",
                               os.path.join(ROOT, "sipnet_b200", "libsipnet_host.so")])
    lib = C.CDLL(HOST_LIB)
    lib.sip_host_error.restype = C.c_char_p
    lib.sip_read_params.argtypes = [C.c_char_p, C.POINTER(A.Flags), C.c_int, C.POINTER(C.c_double)]
    lib.sip_read_clim.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(SiteDataC)]
    lib.sip_read_events.argtypes = [C.c_char_p, C.POINTER(A.Flags), C.POINTER(C.c_double), C.c_int, C.POINTER(SiteDataC)]
    lib.sip_site_free.argtypes = [C.POINTER(SiteDataC)]
    lib.sip_write_state_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int64]
    lib.sip_write_event_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(A.EventRecord)]
    lib.sip_write_header.argtypes = [C.c_void_p]
    lib.sip_write_events_header.argtypes = [C.c_void_p]
    lib.sip_print_config.argtypes = [C.POINTER(ContextC), C.c_void_p, C.c_char_p]
    return lib


_libc = C.CDLL(None)
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


class CFile:
    def __init__(self, path, mode="w"):
        self.fp = _libc.fopen(path.encode(), mode.encode())
        assert self.fp

    def close(self):
        _libc.fclose(self.fp)


def unpack_smoke(dst: str):
    """Extract the committed raw smoke inputs into dst/<case>/ (russell_2/3 share russell_1's .clim)."""
    with tarfile.open(os.path.join(GOLDEN_DIR, "smoke_inputs.tar.gz")) as tar:
        tar.extractall(dst)
    for c in ("russell_2", "russell_3"):
        shutil.copy(os.path.join(dst, "russell_1", "sipnet.clim"), os.path.join(dst, c, "sipnet.clim"))
    return dst


def flags_c(flags: dict) -> A.Flags:
    f = A.Flags()
    merged = dict(A.DEFAULT_FLAGS)
    merged.update(flags)
    for n in A.FLAG_NAMES:
        setattr(f, n, int(merged[n]))
    return f


def read_site_c(lib, clim_path, gdd, events_path=None, flags=None, params=None):
    s = SiteDataC()
    rc = lib.sip_read_clim(clim_path.encode(), gdd, 1, C.byref(s))
    if rc:
        return rc, None
    if events_path is not None:
        p = np.ascontiguousarray(params, dtype=np.float64)
        rc = lib.sip_read_events(events_path.encode(), C.byref(flags), p.ctypes.data_as(C.POINTER(C.c_double)), 1, C.byref(s))
        if rc:
            lib.sip_site_free(C.byref(s))
            return rc, None
    n = s.nsteps
    out = dict(year=np.ctypeslib.as_array(s.year, (n,)).copy(), day=np.ctypeslib.as_array(s.day, (n,)).copy())
    for k in A.CLIM_COLS:
        out[k] = np.ctypeslib.as_array(getattr(s, k), (n,)).copy()
    out["events"] = [(e.year, e.day, e.type, e.method, e.p[0], e.p[1], e.p[2], e.p[3]) for e in s.events[:s.nevents]]
    lib.sip_site_free(C.byref(s))
    return 0, out


# ---- restart checkpoints ---------------------------------------------------------------------------
RESTART_RING = 250


class RestartC(C.Structure):
    """sip_restart (sipnet_b200/host/sip_host.h)."""
    _fields_ = [("modelVersion", C.c_char * 32), ("buildInfo", C.c_char * 96),
                ("checkpointUtcEpoch", C.c_longlong), ("processedSteps", C.c_longlong), ("flags", A.Flags),
                ("boundaryYear", C.c_int), ("boundaryDay", C.c_int), ("boundaryTime", C.c_double),
                ("boundaryLength", C.c_double), ("meanLength", C.c_int), ("meanStart", C.c_int), ("meanLast", C.c_int),
                ("meanTotWeight", C.c_double), ("meanSum", C.c_double), ("envi", C.c_double * 13),
                ("trackers", C.c_double * 33), ("didLeafGrowth", C.c_int), ("didLeafFall", C.c_int),
                ("phenLastYear", C.c_int), ("isAlive", C.c_int), ("dTillMod", C.c_double),
                ("harvestFracRemoved", C.c_double), ("harvestFracTransferred", C.c_double),
                ("values", C.c_double * RESTART_RING), ("weights", C.c_double * RESTART_RING)]


def restart_protos(lib):
    lib.sip_read_restart.argtypes = [C.c_char_p, C.POINTER(RestartC)]
    lib.sip_write_restart.argtypes = [C.c_char_p, C.POINTER(RestartC)]
    lib.sip_check_restart.argtypes = [C.c_char_p, C.POINTER(RestartC), C.POINTER(ContextC), C.POINTER(SiteDataC)]
    return lib


def split_case(src: str, dst1: str, dst2: str, last_year_of_first: int):
    """Cut a smoke case into two consecutive segments at a year boundary (climate AND events, as the
    reference's restart contract asks: docs/developer-guide/restart-checkpoint.md)."""
    for d in (dst1, dst2):
        os.makedirs(d, exist_ok=True)
        for fn in ("sipnet.in", "sipnet.param"):
            shutil.copy(os.path.join(src, fn), d)
    clim = open(os.path.join(src, "sipnet.clim")).read().splitlines(keepends=True)
    ycol = 1 if len(clim[0].split()) == 14 else 0          # legacy files lead with a location column
    with open(os.path.join(dst1, "sipnet.clim"), "w") as a, open(os.path.join(dst2, "sipnet.clim"), "w") as b:
        for ln in clim:
            (a if int(ln.split()[ycol]) <= last_year_of_first else b).write(ln)
    ev = os.path.join(src, "events.in")
    if os.path.exists(ev):
        with open(os.path.join(dst1, "events.in"), "w") as a, open(os.path.join(dst2, "events.in"), "w") as b:
            for ln in open(ev):
                tok = ln.split()
                if not tok or tok[0].startswith("#"):
                    a.write(ln); b.write(ln)
                    continue
                (a if int(tok[0]) <= last_year_of_first else b).write(ln)


def strip_volatile(restart_text: bytes) -> bytes:
    """Drop the two lines that legitimately differ between writers (wall-clock stamp, build id)."""
    return b"\n".join(ln for ln in restart_text.split(b"\n")
                      if not ln.startswith((b"meta_info.checkpoint_utc_epoch", b"meta_info.build_info")))
