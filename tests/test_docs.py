"""CPU: the documents do not point at files or tests that no longer exist."""
import os
import re

import pytest

from conftest import ROOT

DOCS = ["DESIGN.md", "README.md", "INTEGRATION.md", "profiles/README.md", "oracle/README.md", "tools/README.md"]
TOP = ("sipnet_b200/", "tests/", "tools/", "oracle/", "profiles/", "include/", "bench.py", "__graft_entry__.py")
SHORT_DIRS = ["sipnet_b200/csrc", "sipnet_b200/host", "sipnet_b200", "tests", "tools", "oracle", "profiles", "include", ""]


def candidates(text):
    for tok in re.findall(r"`([^`\n]+)`", text):
        tok = tok.strip()
        m = re.match(r"^([\w./\-]+\.(?:jsonl|json|cuh|cu|cpp|csv|ckpt|npz|inc|txt|py|md|sh|c|h))(?![\w/])(?:::(\w+))?", tok)
        if m and ("*" not in tok) and ("<" not in tok) and ("{" not in tok):
            yield m.group(1), m.group(2)


def resolve(path, doc_dir):
    if path.startswith("/") or path.startswith("src/") or "reference" in path:
        return "skip"
    for base in [doc_dir] + SHORT_DIRS:
        p = os.path.join(ROOT, base, path)
        if os.path.exists(p):
            return p
    return None


@pytest.mark.parametrize("doc", DOCS)
def test_paths_and_test_names_in_documents_exist(doc):
    text = open(os.path.join(ROOT, doc)).read()
    missing = []
    for path, test in candidates(text):
        if "/" not in path and not path.startswith(("sip", "test_", "bench", "kat_", "r01_", "make_golden", "libm_check")):
            continue                                    # a bare word with an extension (e.g. a reference file name)
        if os.path.basename(path) in ("sipnet.c", "events.c", "restart.c", "runmean.c", "context.c", "cli.c", "frontend.c",
                                      "nitrogen.c", "balance.c", "depeffects.c", "limitations.c", "outputItems.c",
                                      "debug_log.c", "modelParams.c", "state.h", "events.h", "context.h", "util.h",
                                      "exitCodes.h", "runmean.h", "version.h", "sipnet.h", "testBalance.c", "helpers.c"):
            continue                                    # reference sources, cited by name
        p = resolve(path, os.path.dirname(doc))
        if p == "skip":
            continue
        if p is None:
            if path.startswith(("gpurun_out", "baseline/", "sipnet.", "events.", "balance.", "libm.so", "sites.", "list.", "members.")):
                continue                                # scratch outputs / example file names
            missing.append(path)
        elif test and test.startswith("test_") and p.endswith(".py"):
            if not re.search(r"def %s\w*\(" % re.escape(test.rstrip("_*")), open(p).read()):
                missing.append(f"{path}::{test}")
    assert not missing, f"{doc} mentions {sorted(set(missing))}"
