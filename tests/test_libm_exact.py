"""The restated exp/pow (sipnet_b200/csrc/sip_libm.cuh) against the live glibc libm.

CPU: the product header is compiled in host mode by tests/native/libm_check.cpp
and compared bit for bit on ~10^7 inputs (model-shaped ranges, raw random bit
patterns, every special branch: zeros, infinities, NaN, subnormal bases,
negative bases, overflow / underflow / subnormal results).
GPU: the device build is compared bit for bit with numpy/glibc on the host."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT


def test_host_restatement_is_bit_exact():
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "libm_check")
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "sipnet_b200", "csrc"),
                               os.path.join(ROOT, "tests", "native", "libm_check.cpp"), "-lm", "-o", exe])
        out = subprocess.run([exe, "700000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "0 mismatches; pow" in out.stdout and out.stdout.strip().endswith("0 mismatches")


def test_tables_match_system_libm():
    """The committed table header equals what tools/gen_libm_tables.py extracts from this box's libm."""
    hdr = os.path.join(ROOT, "sipnet_b200", "csrc", "sip_libm_tables.h")
    before = open(hdr).read()
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "gen_libm_tables.py")], stdout=subprocess.DEVNULL)
    after = open(hdr).read()
    if before != after:
        open(hdr, "w").write(before)
    assert before == after, "system libm tables differ from the committed ones (different glibc?)"


@pytest.mark.gpu
def test_device_libm_is_bit_exact():
    from sipnet_b200 import api
    rng = np.random.default_rng(12345)
    n = 400000
    xs = np.concatenate([rng.uniform(-60, 20, n), rng.uniform(-750, 710, n), rng.uniform(-1100, 1100, n // 4),
                         np.array([0.0, -0.0, 1e-300, -1e-17, 709.78, -745.13, 512.0, -512.0, np.inf, -np.inf, np.nan])])
    got = api.device_libm("exp", xs)
    # numpy's exp may use its own SIMD kernels; the expected values come from the C library itself
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.exp.restype = ctypes.c_double
    libm.exp.argtypes = [ctypes.c_double]
    libm.pow.restype = ctypes.c_double
    libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
    idx = np.concatenate([rng.choice(xs.size - 11, 60000, replace=False), np.arange(xs.size - 11, xs.size)])
    want = np.array([libm.exp(float(xs[i])) for i in idx])
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(got[idx]), np.isnan(want))
    assert np.array_equal(got[idx][ok].view(np.uint64), want[ok].view(np.uint64))
    bx = np.concatenate([np.full(n, 2.0), rng.uniform(1, 6, n), rng.uniform(1e-6, 6, n), rng.uniform(0, 1, n),
                         rng.uniform(0, 2, n // 4), -rng.uniform(0, 10, n // 4)])
    by = np.concatenate([rng.uniform(-300, 10, n), rng.uniform(-6, 6, n), rng.uniform(0.5, 4, n), rng.uniform(0, 4, n),
                         rng.uniform(-1100, 1100, n // 4), np.round(rng.uniform(-40, 40, n // 4))])
    gotp = api.device_libm("pow", bx, by)
    idx = rng.choice(bx.size, 60000, replace=False)
    wantp = np.array([libm.pow(float(bx[i]), float(by[i])) for i in idx])
    ok = ~np.isnan(wantp)
    assert np.array_equal(np.isnan(gotp[idx]), np.isnan(wantp))
    assert np.array_equal(gotp[idx][ok].view(np.uint64), wantp[ok].view(np.uint64))
