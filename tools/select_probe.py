#!/usr/bin/env python
"""Micro-benchmark of the row summary (exact quantile select): python tools/select_probe.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sipnet_b200 import distributed as D

qs = [0.05, 0.5, 0.95]
for nrows, ncols in ((7306, 131072), (3653, 262144), (913, 1048576), (128, 1048576)):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((nrows, ncols), dtype=torch.float64, device="cuda", generator=g) * 3.0 + 1.0
    for moments in (False, True):
        D.rows_summary(x, qs, moments=moments)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            D.rows_summary(x, qs, moments=moments)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"rows {nrows} x {ncols} moments={moments}: {dt * 1e3:.2f} ms  ({x.numel() * 8 / dt / 1e9:.0f} GB/s of row data)", flush=True)
    del x
