"""CPU: the C-ABI library loads, exports every symbol include/sipnet_gpu.h
declares, the ctypes mirror matches the C struct layouts, and -- without a GPU --
the product fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT, have_gpu
from sipnet_b200 import _abi as A, api, synth

HEADER = os.path.join(ROOT, "include", "sipnet_gpu.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sipnet_gpu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    syms = declared_symbols()
    assert {"sipnet_gpu_init", "sipnet_gpu_run", "sipnet_gpu_gather", "sipnet_gpu_destroy"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sipnet_gpu.h but not exported"
    assert lib.sipnet_gpu_abi_version() == A.ABI_VERSION


def test_ctypes_mirror_matches_c_layout():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "sipnet_gpu.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(sipnet_gpu_flags), sizeof(sipnet_gpu_event), sizeof(sipnet_gpu_site),
         sizeof(sipnet_gpu_event_record), sizeof(sipnet_gpu_config), offsetof(sipnet_gpu_config, stream));
  printf("%d %d %d %d\n", SIPNET_GPU_NPARAMS, SIPNET_GPU_NOUT, SIPNET_GPU_NDEBUG, SIPNET_GPU_NSTATE);
  printf("%d %d %d\n", SIPNET_P_soilCSaturation, SIPNET_O_nppStorage, SIPNET_S_meanLast);
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        lines = subprocess.check_output([exe]).decode().split("\n")
    sizes = list(map(int, lines[0].split()))
    assert sizes == [C.sizeof(A.Flags), C.sizeof(A.Event), C.sizeof(A.Site), C.sizeof(A.EventRecord),
                     C.sizeof(A.Config), A.Config.stream.offset]
    assert list(map(int, lines[1].split())) == [A.NPARAMS, A.NOUT, A.NDEBUG, A.NSTATE]
    assert list(map(int, lines[2].split())) == [A.P["soilCSaturation"], A.O["nppStorage"],
                                                A.STATE_NAMES.index("meanLast")]


def test_header_is_plain_c():
    """No torch / CUDA types in the boundary: the header compiles as C11 on its own."""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write('#include "sipnet_gpu.h"\nint x;\n')
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-c", "-I", os.path.join(ROOT, "include"),
                               src, "-o", os.path.join(td, "t.o")])


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    sites, params, ms, flags = synth.config_c2(nmembers=4, nyears=1)
    with pytest.raises(api.SipnetGpuError) as ei:
        api.Ensemble(sites, params, ms, flags)
    assert ei.value.code == A.ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """The shipped package and library never reference oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "sipnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "sipnet_oracle" not in txt and "oracle/" not in txt, (dirpath, f)
    out = subprocess.run(["nm", "-D", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sipnet_oracle" not in out and "sipref_" not in out


def test_tile_skips_only_unused_rows():
    """The parameter rows left out of the shared-memory tile (sip_types.cuh kTileSkip) are never read by
    the step, and the list is ascending (tile_slot() relies on it)."""
    types = open(os.path.join(ROOT, "sipnet_b200", "csrc", "sip_types.cuh")).read()
    step = open(os.path.join(ROOT, "sipnet_b200", "csrc", "sip_step.cuh")).read()
    body = re.search(r"#define SIP_TILE_SKIP_LIST(.*?)\nconstexpr", types, re.S).group(1)
    names = re.findall(r"SIPNET_P_(\w+)", body)
    assert len(names) >= 10
    idx = [A.P[n] for n in names]
    assert idx == sorted(idx)
    for n in names:
        assert f"SIP_P({n})" not in step and f"SIPNET_P_{n}" not in step, n
