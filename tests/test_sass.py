"""CPU: properties of the compiled device code, read from the in-tree library with cuobjdump (no GPU needed):
the build targets sm_100a only, the step kernel stages forcing with the TMA bulk copy + mbarrier, works in FP64
without contraction surprises, and the headline variant (C2: crop-N flags, 32-member blocks, all 32 columns, static
schedule) holds its state in registers with no local-memory spills."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "sipnet_b200", "libsipnet_gpu.so")
pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and shutil.which("cuobjdump")), reason="library or cuobjdump missing")


@pytest.fixture(scope="module")
def res_usage():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", out):
        usage[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)))
    return out, usage


def kernels(usage, *needles):
    return {k: v for k, v in usage.items() if all(n in k for n in needles)}


def test_only_sm_100a_code_is_shipped(res_usage):
    out, _ = res_usage
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs


def test_step_kernel_variants_exist_and_fit_two_blocks_per_sm(res_usage):
    _, usage = res_usage
    runs = kernels(usage, "run_kernel")
    # 3 flag policies x {32, 128, 128 with two sites per block} x {all columns, summary columns} x {static, dynamic},
    # for the optimistic and the throughput numerics, + 3 x 2 general
    assert len(kernels(runs, "FastNum")) == 36 and len(kernels(runs, "ThroughNum")) == 36 and len(kernels(runs, "ExactNum")) == 6
    for name, u in runs.items():
        assert u["reg"] <= 255
        if "Li128E" in name:
            assert u["reg"] * 128 * 2 <= 65536, name           # two 128-member blocks per SM (register file)
    setup = kernels(usage, "derive_params_kernel") | kernels(usage, "init_state_kernel") | kernels(usage, "row_summary_kernel")
    assert len(setup) == 3 and all(u["stack"] <= 64 for u in setup.values())
    assert next(iter(kernels(usage, "row_summary_kernel").values()))["reg"] <= 64   # 512 threads x 2 blocks per SM


def headline(usage):
    k = kernels(usage, "run_kernel", "StaticFlagsILj947EE", "FastNum", "Li32ELb0ELb1ELb0")   # BLOCK 32, !REPLAY, FULL, !DYN
    assert len(k) == 1, list(k)
    return next(iter(k.items()))


def test_headline_kernel_has_no_spills_and_uses_tma(res_usage):
    _, usage = res_usage
    name, u = headline(usage)
    assert u["stack"] == 0, "the C2 kernel must keep the member's state in registers"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, LIB], capture_output=True, text=True).stdout
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", sass, flags=re.M)
    assert len(ops) > 4000
    has = lambda prefix: any(o.startswith(prefix) for o in ops)  # noqa: E731
    assert has("UBLKCP"), "forcing chunks are staged by the TMA bulk copy (cp.async.bulk)"
    assert has("SYNCS"), "... completing on an mbarrier"
    assert has("DFMA") and has("DMUL") and has("DADD")
    assert not has("LDL") and not has("STL")
    assert not has("MUFU.EX2") and not has("MUFU.LG2"), "exp/pow are the glibc-exact restatement, not the fast intrinsics"
    assert sum(o.startswith("STG") for o in ops) >= 32          # one streaming store per output column per step
