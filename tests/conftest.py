"""pytest configuration: the `gpu` marker and shared fixtures.

CPU suite (`-m "not gpu"`): oracle vs golden vectors / vs the reference build,
host logic, ABI/export checks.  GPU suite (`-m gpu`): parity of the CUDA path
(through the C ABI) against the oracle, the goldens and -- when oracle/_ref
travelled with the snapshot -- the unmodified reference itself.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sipnet_b200 import _abi as A  # noqa: E402
from sipnet_b200.api import SiteData  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.flags = {n: int(v) for n, v in zip(A.FLAG_NAMES, z["flags"])}
        self.params = z["params"].astype(np.float64)
        clim = z["clim"]
        self.site = SiteData(z["year"], z["day"], {k: clim[i] for i, k in enumerate(A.CLIM_COLS)})
        ev = z["events"]
        self.site.events = [(int(r[0]), int(r[1]), int(r[2]), int(r[3]), r[4], r[5], r[6], r[7]) for r in ev]
        self.rows = z["rows"]
        self.out32 = z["out32"]
        self.dbg = z["dbg"]
        self.colsum = z["colsum"]
        self.colabs = z["colabs"]
        self.nsteps = int(z["nsteps"])
        self.rc = int(z["rc"])
        self.print_header = int(z["print_header"])
        self.events_out = bytes(z["events_out"].tobytes())
        self.main_out_md5 = bytes(z["main_out_md5"].tobytes()).decode()


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def refshim():
    from oracle import pyoracle
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return pyoracle.RefShim()


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_lib():
    from sipnet_b200.api import load_library
    return load_library()


# --- comparison helper: |a-b| <= rtol * max(|a|, |b|, scale) -------------------
# SURVEY 7 hard part 2: the parity metric is relative with a per-variable scale
# floor; `scale` is the magnitude of the column over the run (plantCAccountingDelta
# is a cancellation accumulator and is scaled by plantWoodC).
def rel_err(a, b, scale):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
    den = np.where(den == 0, 1.0, den)
    err = np.abs(a - b) / den
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0.0, err)
