"""The printf-free sipnet.out row formatter and the many-member block writer (host/sip_output.c) against the reference's
own printf statements (sipnet.c:455-472), byte for byte.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from host_util import CFile, host_lib
from sipnet_b200 import _abi as A

ROW_MAX = 12288
# (column, precision) of outputState()'s fields in sipnet_gpu.h column order
PREC = [2, 2, 2, 2, 2, 2, 2, 3, 3, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 8, 4, 4, 4, 4, 4, 6, 4, 4, 4, 4, 4]


@pytest.fixture(scope="module")
def lib():
    lib = host_lib()
    for f in (lib.sip_format_state_row, lib.sip_format_state_row_printf):
        f.restype = C.c_size_t
        f.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int64]
    lib.sip_write_state_block.restype = C.c_int
    lib.sip_write_state_block.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int64),
                                          C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32)),
                                          C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_double), C.c_int64, C.c_int64]
    return lib


def both(lib, year, day, time, row):
    row = np.ascontiguousarray(row, np.float64)
    a, b = C.create_string_buffer(ROW_MAX), C.create_string_buffer(ROW_MAX)
    p = row.ctypes.data_as(C.POINTER(C.c_double))
    na = lib.sip_format_state_row(a, year, day, time, p, 1)
    nb = lib.sip_format_state_row_printf(b, year, day, time, p, 1)
    return a.raw[:na], b.raw[:nb]


def adversarial_values(rng, n):
    """Values on and next to the decimal rounding boundaries of every precision in use, exact ties, signed zeros,
    values that round to zero, the magnitudes of the model's columns, huge and non-finite ones."""
    out = []
    for prec in (2, 3, 4, 6, 8):
        k = rng.integers(0, 10 ** 7, n).astype(np.float64)
        base = (k + 0.5) / 10.0 ** prec                      # decimal ties (inexact in binary: either side)
        out += [base, np.nextafter(base, np.inf), np.nextafter(base, -np.inf), -base]
        j = rng.integers(0, 4096, n).astype(np.float64)
        out.append((2 * j + 1) / 2.0 ** rng.integers(1, 12, n))  # dyadic values: exact ties for small precisions
    out.append(rng.uniform(-1, 1, n) * 10.0 ** rng.integers(-12, 16, n))
    out.append(rng.uniform(-1e-9, 1e-9, n))
    out.append(np.array([0.0, -0.0, 0.005, -0.005, 0.125, -0.125, 0.375, 2.5e-3, 1e15, -1e15, 9.999999999e14, 4.5e15,
                         1e16, 1e22, 1e300, -1e308, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 0.995, 9.995, 99.995,
                         0.9999999, 999999.995, 1e9 - 0.005, 123456789.125, 0.5, 1.5, 2.5, -0.5]))
    return np.concatenate(out)


def test_fast_rows_equal_printf_rows(lib):
    rng = np.random.default_rng(20260117)
    vals = adversarial_values(rng, 6000)
    rng.shuffle(vals)
    vals = vals[: (vals.size // A.NOUT) * A.NOUT].reshape(-1, A.NOUT)
    for i, row in enumerate(vals):
        a, b = both(lib, 2011 + i % 10, 1 + i % 366, [0.0, 12.0, 7.99, 23.999][i % 4], row)
        assert a == b, (row.tolist(), a, b)
    # odd "year day time" values take the same route
    for y, d, t in [(0, 0, 0.0), (-5, 7, -0.001), (12345, 1000, 123.456), (2020, 366, 12.005), (1999, 1, 0.125)]:
        a, b = both(lib, y, d, t, vals[0])
        assert a == b


def test_model_like_rows_equal_printf_rows(lib):
    """Rows shaped like the model's (pools 1e2..1e4, fluxes 1e-6..10, many exact zeros)."""
    rng = np.random.default_rng(7)
    scale = 10.0 ** np.array([4, 2, 0, 4, 3, 2, 3, 1, 0, 1] + [0] * 10 + [-1, -1, 0, 2, 1, 0, -4, -3, -3, -2, -3, 2], float)
    for i in range(20000):
        row = rng.uniform(-1, 1, A.NOUT) * scale
        row[rng.uniform(size=A.NOUT) < 0.2] = 0.0
        a, b = both(lib, 2015, 1 + i % 365, 12.0 * (i % 2), row)
        assert a == b


def test_block_writer_equals_row_writer(lib, tmp_path):
    """8- and 5-member blocks in the gathered [col][step][member] layout, sites of different length."""
    rng = np.random.default_rng(3)
    M, T = 13, 150
    buf = rng.normal(0, 50, (A.NOUT, T, M))
    buf[:, :, 3] = 0.0
    nsteps = np.array([T, T, 100, T, 64, 65, 1, T, T, 0, 99, T, 128], np.int64)
    year = np.repeat(np.arange(2011, 2011 + 3), 50).astype(np.int32)
    day = (np.arange(T) % 365 + 1).astype(np.int32)
    time = np.where(np.arange(T) % 2 == 0, 0.0, 12.0)
    for m0, count in ((0, 8), (8, 5)):
        files = [CFile(str(tmp_path / f"blk{m0 + k}.out")) for k in range(count)]
        fp = (C.c_void_p * count)(*[f.fp for f in files])
        ns = np.ascontiguousarray(nsteps[m0:m0 + count])
        yp = (C.POINTER(C.c_int32) * count)(*[year.ctypes.data_as(C.POINTER(C.c_int32))] * count)
        dp = (C.POINTER(C.c_int32) * count)(*[day.ctypes.data_as(C.POINTER(C.c_int32))] * count)
        tp = (C.POINTER(C.c_double) * count)(*[time.ctypes.data_as(C.POINTER(C.c_double))] * count)
        rc = lib.sip_write_state_block(fp, count, ns.ctypes.data_as(C.POINTER(C.c_int64)), yp, dp, tp,
                                       C.cast(buf.ctypes.data + 8 * m0, C.POINTER(C.c_double)), T * M, M)
        assert rc == 0
        for f in files:
            f.close()
    for m in range(M):
        want = b""
        for t in range(int(nsteps[m])):
            _, b = both(lib, int(year[t]), int(day[t]), float(time[t]), buf[:, t, m])
            want += b
        assert open(tmp_path / f"blk{m}.out", "rb").read() == want, m


def test_fast_formatter_is_fast(lib):
    """Not a benchmark gate: records the speed-up over printf on this host (and fails only if it is slower)."""
    import time as _t
    rng = np.random.default_rng(1)
    rows = rng.normal(0, 100, (20000, A.NOUT))
    a = C.create_string_buffer(ROW_MAX)
    ptrs = [r.ctypes.data_as(C.POINTER(C.c_double)) for r in rows]
    res = {}
    for name, fn in (("fast", lib.sip_format_state_row), ("printf", lib.sip_format_state_row_printf)):
        t0 = _t.perf_counter()
        for p in ptrs:
            fn(a, 2015, 100, 12.0, p, 1)
        res[name] = _t.perf_counter() - t0
    print(f"rows/s: fast {len(rows) / res['fast']:.3g}, printf {len(rows) / res['printf']:.3g} (ctypes call overhead included)")
    assert res["fast"] < res["printf"]


PATH_MAX = 256 + 32


@pytest.mark.parametrize("nthreads", [1, 3, 0])
def test_state_files_of_a_launch(lib, tmp_path, nthreads):
    """All members' files from the gathered [col][T][M] array: 21 members (two full blocks + one of five) of two
    sites with different lengths, on 1 / 3 / all threads; then a path that cannot be opened."""
    lib.sip_write_state_files.restype = C.c_int
    lib.sip_write_state_files.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_int32)),
                                          C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_double)), C.c_int64,
                                          C.POINTER(C.c_double), C.c_int, C.c_int]
    rng = np.random.default_rng(11)
    M, T = 21, 200
    buf = rng.normal(0, 20, (A.NOUT, T, M))
    siteT = [200, 131]
    years = [np.repeat(np.arange(2011, 2015), 50).astype(np.int32), np.full(131, 1999, np.int32)]
    days = [(np.arange(200) // 2 % 365 + 1).astype(np.int32), (np.arange(131) + 100).astype(np.int32)]
    times = [np.where(np.arange(200) % 2 == 0, 0.0, 12.0), np.linspace(0, 23.99, 131)]
    site_of = [0] * 12 + [1] * 9
    paths = C.create_string_buffer(M * PATH_MAX)
    for m in range(M):
        name = str(tmp_path / f"s{site_of[m]}.out.{m}").encode()
        paths[m * PATH_MAX: m * PATH_MAX + len(name)] = name
    ns = np.array([siteT[s] for s in site_of], np.int64)
    yp = (C.POINTER(C.c_int32) * M)(*[years[s].ctypes.data_as(C.POINTER(C.c_int32)) for s in site_of])
    dp = (C.POINTER(C.c_int32) * M)(*[days[s].ctypes.data_as(C.POINTER(C.c_int32)) for s in site_of])
    tp = (C.POINTER(C.c_double) * M)(*[times[s].ctypes.data_as(C.POINTER(C.c_double)) for s in site_of])
    call = lambda: lib.sip_write_state_files(paths, M, ns.ctypes.data_as(C.POINTER(C.c_int64)), yp, dp, tp, T,
                                             buf.ctypes.data_as(C.POINTER(C.c_double)), 1, nthreads)
    assert call() == 0
    hdr = str(tmp_path / "hdr")
    f = CFile(hdr)
    lib.sip_write_header(f.fp)
    f.close()
    header = open(hdr, "rb").read()
    for m in range(M):
        s = site_of[m]
        want = header
        for t in range(siteT[s]):
            want += both(lib, int(years[s][t]), int(days[s][t]), float(times[s][t]), buf[:, t, m])[1]
        assert open(tmp_path / f"s{s}.out.{m}", "rb").read() == want, m
    bad = str(tmp_path / "no_such_dir" / "x.out").encode()
    paths[17 * PATH_MAX: 17 * PATH_MAX + len(bad) + 1] = bad + b"\0"
    assert call() == 6                                                   # EXIT_CODE_FILE_OPEN_OR_READ_ERROR
    assert b"no_such_dir" in lib.sip_host_error()
