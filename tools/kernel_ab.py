#!/usr/bin/env python
"""A/B timing of the step kernel's variants on one GPU (one child process per setting, because the library
reads its measurement switches once).

    python tools/kernel_ab.py [--members 131072] [--years 10] [--modes c4,c5] [--occ 2,3,4] [--libs a.so,b.so]

Per setting: kernel ms (CUDA events around the launch, mean of 3 after a warm-up pass), member-steps/s, and a
checksum of the results (log-likelihood bits for c5, NEE mean bits for c4) -- every variant must print the same
checksum: they are the same arithmetic.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(args):
    sys.path.insert(0, ROOT)
    import numpy as np
    from sipnet_b200 import _abi as A, api, synth
    M = args.members
    site = synth.synth_site(0, args.years, "half-daily")
    params = synth.synth_params(M, stream=100)
    flags = dict(synth.SYNTH_FLAGS)
    if args.mode == "c5":
        rng = np.random.default_rng(7)
        site.nee_obs = np.where(rng.uniform(size=site.nsteps) < 0.2, np.nan, rng.normal(0, 1.5, site.nsteps))
        kw = dict(outputs=A.OUT_LOGLIK, nee_sigma=0.5)
    elif args.mode == "c4":
        kw = dict(outputs=A.OUT_MOMENTS, summary_cols=[A.O["nee"], A.O["gpp"]])
    else:  # full output, segmented
        kw = dict(outputs=A.OUT_FULL, out_steps_capacity=64)
    math = {"fast": A.MATH_FAST, "throughput": A.MATH_THROUGHPUT, "validation": A.MATH_VALIDATION}[args.math]
    ens = api.Ensemble([site], params, None, flags, math=math, device=0, block_threads=args.block, **kw)
    T = ens.max_steps

    def one():
        ens.reset()
        if args.mode == "full":
            ms = 0.0
            for t0 in range(0, T, 64):
                ens.run(t0, min(T, t0 + 64))
                ms += ens.last_run_ms()
            return ms
        ens.run(0, T)
        return ens.last_run_ms()

    one()
    ms = [one() for _ in range(3)]
    if args.mode == "c5":
        digest = hashlib.sha1(ens.loglik().tobytes()).hexdigest()[:12]
    elif args.mode == "c4":
        digest = hashlib.sha1(ens.mean().tobytes()).hexdigest()[:12]
    else:
        digest = hashlib.sha1(ens.state().tobytes()).hexdigest()[:12]
    st = ens.status()
    ens.close()
    k = float(np.mean(ms))
    print(json.dumps({"mode": args.mode, "math": args.math, "block": args.block, "members": M, "occ": os.environ.get("SIPNET_GPU_OCC", "auto"),
                      "lib": os.path.basename(os.environ.get("SIPNET_GPU_LIB", "default")), "kernel_ms": round(k, 3),
                      "member_steps_per_s": M * T / (k * 1e-3), "checksum": digest,
                      "replayed": int((st & A.ST_REPLAY).astype(bool).sum())}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=131072)
    ap.add_argument("--years", type=int, default=10)
    ap.add_argument("--modes", default="c4,c5")
    ap.add_argument("--occ", default="2,3,4")
    ap.add_argument("--libs", default="")
    ap.add_argument("--mode", default="")
    ap.add_argument("--block", type=int, default=0, help="block_threads (0 = library default)")
    ap.add_argument("--math", default="fast", help="fast | throughput | validation (comma list in the parent)")
    args = ap.parse_args()
    if args.mode:
        return child(args)
    libs = [x for x in args.libs.split(",") if x] or [""]
    for lib in libs:
        for mode in args.modes.split(","):
            for occ in args.occ.split(","):
                for math in args.math.split(","):
                    env = dict(os.environ)
                    if occ != "auto":
                        env["SIPNET_GPU_OCC"] = occ
                    if lib:
                        env["SIPNET_GPU_LIB"] = os.path.abspath(lib)
                    subprocess.call([sys.executable, os.path.abspath(__file__), "--mode", mode, "--members", str(args.members),
                                     "--years", str(args.years), "--math", math, "--block", str(args.block)], env=env)


if __name__ == "__main__":
    main()
