#!/usr/bin/env python
"""bench.py -- ensemble member-timesteps/sec of the SIPNET integration loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2"): 1 synthetic site x 4096-member
parameter ensemble, 10 years of half-daily forcing (T = 7306), flags
litter+anaerobic+nitrogen, FULL per-step output (32 doubles per member-step).
One "step" = one pass of the hot path over that batch (state reset + all 7306
model steps of all members).  Per-GPU work is fixed as N grows (weak scaling:
every rank integrates its own 4096-member ensemble, no data-path collective).

value  : member-timesteps/s with inputs resident in HBM, timed with CUDA events
         on the launching stream, max over ranks.
e2e    : same metric through the C ABI with HOST buffers: parameter upload
         (pinned host -> device) + run + gather of the full output into pinned
         host memory, all inside the timed region.
roofline: dominant kernel = the fused step kernel; FP64-issue bound
         (~3.2 kFLOP of FP64 issue per member-step) against the FP64
         FMA peak measured live by sipnet_gpu_measure_fp64_peak, and the HBM
         figure (256 B/member-step) against MEASURED_PEAKS.json beside it.
cpu_baseline: the unmodified reference (oracle/_ref) or the oracle port timed on
         the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ensemble member-timesteps/sec"
UNIT = "member-timesteps/s"
FLOP_PER_MEMBER_STEP = 3200.0   # SURVEY 8d / BASELINE.md 4: ~630 FP64 ops + ~19 libm calls
BYTES_PER_MEMBER_STEP = 256.0   # 32 output doubles per member-step (full output)
WORKLOAD = "C2: 1 site x 4096 members x 10 yr half-daily (T=7306), litter+anaerobic+nitrogen, full per-step output"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members", type=int, default=4096)
    ap.add_argument("--years", type=int, default=10)
    ap.add_argument("--math", default="fast", choices=["fast", "validation"])
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 = the bench workload (default); c3/c4/c5 = the other BASELINE.json configs (extra lines)")
    ap.add_argument("--sites", type=int, default=0, help="c3: number of sites on this GPU (default 10000 / n_gpus)")
    ap.add_argument("--cap", type=int, default=0, help="c4: steps per run segment kept on the device (0 = the whole run)")
    ap.add_argument("--verify", action="store_true", help="c4 pipelined: compare its quantiles with the sequential pass")
    ap.add_argument("--pipeline", action="store_true",
                    help="c4 on several GPUs: segment-pipelined pass (exchange + select of segment i under the kernel of "
                         "segment i+1); measured slower than the default at 8 GPUs, faster at 2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-filled", action="store_true", help="skip the filled-GPU (C4 per-GPU share) roofline measurement")
    return ap.parse_args()


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------- CPU arms
def _ref_worker(args):
    """One process = a slice of members through the UNMODIFIED reference (oracle/_ref shim)."""
    flags, params, site_npz, idx = args
    from oracle.pyoracle import RefShim
    from sipnet_b200 import _abi as A
    from sipnet_b200.api import SiteData
    shim = RefShim()
    site = SiteData(site_npz["year"], site_npz["day"], {k: site_npz[k] for k in A.CLIM_COLS})
    t0 = time.perf_counter()
    steps = 0
    for m in idx:
        rc, done, _, _ = shim.run(flags, params[:, m], site, want_debug=False)
        steps += done
    return steps, time.perf_counter() - t0


class CpuReference:
    """The reference CPU implementation on all host cores (one process per core,
    each running members through oracle/_ref -- the unmodified reference -- or,
    if that build is absent, the oracle port on threads)."""

    def __init__(self, sites, params, flags, cores: int):
        from oracle import pyoracle
        from sipnet_b200 import _abi as A
        self.site = sites[0]
        self.params = params
        self.flags = flags
        self.cores = cores
        self.kind = "reference" if pyoracle.have_ref() else "port"
        self.pool = None
        if self.kind == "reference":
            self.site_npz = dict(year=self.site.year, day=self.site.day,
                                 **{k: self.site.clim[k] for k in A.CLIM_COLS})
            self.pool = mp.get_context("spawn").Pool(cores)
            # warm: import, load the .so, page in
            self.pool.map(_ref_worker, [(flags, params, self.site_npz, [i % params.shape[1]]) for i in range(cores)])
        else:
            self.oracle = pyoracle.Oracle()

    def sample(self, members_per_core: int):
        """-> (member-steps/s, description)"""
        from sipnet_b200 import _abi as A
        site, cores = self.site, self.cores
        nm = min(self.params.shape[1], members_per_core * cores)
        if self.kind == "reference":
            chunks = [list(range(i, nm, cores)) for i in range(cores)]
            t0 = time.perf_counter()
            res = self.pool.map(_ref_worker, [(self.flags, self.params, self.site_npz, c) for c in chunks])
            wall = time.perf_counter() - t0
            return sum(r[0] for r in res) / wall, (f"{nm} members x {site.nsteps} steps, one process per core, "
                                                    "unmodified reference (oracle/_ref), no text output")
        from sipnet_b200.api import flags_array
        out = np.zeros((nm, A.NOUT))
        fl = flags_array(self.flags)
        clim = (C.POINTER(C.c_double) * 11)(*[site.clim[k].ctypes.data_as(C.POINTER(C.c_double)) for k in A.CLIM_COLS])
        arr, nev = site.event_array()
        P = np.ascontiguousarray(self.params[:, :nm])
        lib = self.oracle.lib
        lib.sipnet_oracle_run_ensemble.restype = C.c_int
        t0 = time.perf_counter()
        lib.sipnet_oracle_run_ensemble(
            fl.ctypes.data_as(C.POINTER(C.c_int32)), P.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(nm),
            C.c_int64(nm), C.c_int64(site.nsteps), site.year.ctypes.data_as(C.POINTER(C.c_int32)),
            site.day.ctypes.data_as(C.POINTER(C.c_int32)), clim, C.c_int64(nev), arr, C.c_int(cores),
            out.ctypes.data_as(C.POINTER(C.c_double)))
        wall = time.perf_counter() - t0
        return nm * site.nsteps / wall, f"{nm} members x {site.nsteps} steps, {cores} threads, oracle port"

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sipnet_b200 import synth
    sites, params, ms, flags = synth.config_c2(nmembers=args.members, nyears=args.years)
    cores = os.cpu_count() or 1
    # bounded sample per step: ~1-2 s of work per core (0.35 M member-steps/s/core in the survey)
    per_core = max(1, min(48, args.members // cores))
    cpu = CpuReference(sites, params, flags, cores)
    kind = cpu.kind
    vals = []
    for i in range(args.warmup + args.steps):
        v, sample = cpu.sample(per_core)
        if i >= args.warmup:
            vals.append(v)
    cpu.close()
    value = float(np.mean(vals))
    T = sites[0].nsteps
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (min(params.shape[1], per_core * cores) * T) / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "members_per_gpu": args.members, "model_steps": T,
                   "note": "each step is a bounded sample of the workload on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_filled(local: int, years: int, fp64_peak_tflops: float):
    """Roofline of the same kernel when the GPU is FILLED: BASELINE.json's target configuration (C4: 1M members on
    8 GPUs) per-GPU share, 131072 members x 10 yr, with the on-device reductions the config asks for (mean, variance
    and exact 5/50/95 % quantiles of NEE and GPP per step).  C2's 4096 members are 128 warps on 592 warp schedulers:
    its roofline fraction measures the model's sequential time loop, not the kernel."""
    from sipnet_b200 import _abi as A, api, synth
    M = 131072
    site = synth.synth_site(0, years, "half-daily")
    params = synth.synth_params(M, stream=100)
    ens = api.Ensemble([site], params, None, dict(synth.SYNTH_FLAGS), math=A.MATH_FAST, device=local,
                       outputs=A.OUT_MOMENTS | A.OUT_QUANTILES, summary_cols=[A.O["nee"], A.O["gpp"]],
                       quantiles=[0.05, 0.5, 0.95])
    T = ens.max_steps

    def one_pass():
        ens.reset()
        ens.run(0, T)
        ens.device_ptr(A.GATHER_QUANTILES)      # launches the row-summary kernels (results stay on the device)
        ens.sync()

    one_pass()
    kern, total = [], []
    for _ in range(3):
        ens.sync()
        ens.timer_start()
        one_pass()
        total.append(ens.timer_stop_ms())
        kern.append(ens.last_run_ms())
    q = ens.quantiles()
    assert np.isfinite(q).all()
    ens.close()
    kern_ms, total_ms = float(np.mean(kern)), float(np.mean(total))
    ach = M * T / (kern_ms * 1e-3) * FLOP_PER_MEMBER_STEP / 1e12
    whole = M * T / (total_ms * 1e-3)
    return {"workload": f"C4 per-GPU share: {M} members x {T} steps, on-device mean/variance + exact 5/50/95% quantiles "
                        "of NEE and GPP",
            "bound": "fp64", "kernel": "sip::run_kernel", "kernel_ms": kern_ms, "achieved": ach, "peak": fp64_peak_tflops,
            "unit": "TFLOP/s", "frac": ach / fp64_peak_tflops if fp64_peak_tflops > 0 else None,
            "whole_job_ms": total_ms, "whole_job_value": whole,
            "whole_job_frac": whole * FLOP_PER_MEMBER_STEP / 1e12 / fp64_peak_tflops if fp64_peak_tflops > 0 else None}


# ---------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from sipnet_b200 import _abi as A, api, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # every rank owns an independent ensemble (weak scaling); different parameter stream per rank
    site = synth.synth_site(0, args.years, "half-daily")
    params = synth.synth_params(args.members, stream=rank)
    flags = dict(synth.SYNTH_FLAGS)
    T = site.nsteps
    M = args.members
    math = A.MATH_FAST if args.math == "fast" else A.MATH_VALIDATION
    lib = api.load_library()

    ens = api.Ensemble([site], params, None, flags, outputs=A.OUT_FULL, math=math, device=local,
                       block_threads=args.block)
    peak = C.c_double(0.0)
    lib.sipnet_gpu_measure_fp64_peak(local, C.byref(peak))
    fp64_peak_tflops = float(peak.value)

    def one_step():
        ens.reset()
        ens.run(0, T)

    for _ in range(max(args.warmup, 3)):
        one_step()
    ens.sync()

    sampler = ClockSampler(local)
    launches0 = ens.launch_count()
    barrier()
    sampler.start()
    ens.timer_start()
    kernel_ms = []
    for _ in range(args.steps):
        one_step()
    total_ms = ens.timer_stop_ms()
    barrier()
    clocks = sampler.stop()
    launches = ens.launch_count() - launches0
    # per-launch duration of the dominant kernel (CUDA events around the launch, on its stream)
    for _ in range(3):
        one_step()
        kernel_ms.append(ens.last_run_ms())
    kern_ms = float(np.mean(kernel_ms))
    ms_per_step = max_over_ranks(total_ms / args.steps)
    value = world * M * T / (ms_per_step * 1e-3)

    # ---- end to end: pinned host params -> device, run, full output -> pinned host
    e2e = None
    if not args.no_e2e:
        out_bytes = A.NOUT * T * M * 8
        par_bytes = A.NPARAMS * M * 8
        hout = lib.sipnet_gpu_host_alloc(out_bytes)
        hpar = lib.sipnet_gpu_host_alloc(par_bytes)
        if not hout or not hpar:
            raise SystemExit("bench.py: pinned host allocation failed")
        C.memmove(hpar, params.ctypes.data, par_bytes)

        def e2e_step():
            ens.set_params(hpar, M)                               # H2D of the ensemble + setupModel()
            ens.run_to_host(hout, 0, T, nbytes=out_bytes)         # run, D2H pipelined behind the next segment

        e2e_step()
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        ens.timer_start()
        for _ in range(n_e2e):
            e2e_step()
        e2e_ms = max_over_ranks(ens.timer_stop_ms() / n_e2e)
        barrier()
        # the result really is in host memory: compare one element with a device-side gather
        probe = np.ctypeslib.as_array((C.c_double * 8).from_address(hout + 8 * (A.O["nee"] * T * M + (T - 1) * M)))
        assert np.isfinite(probe).all()
        e2e = {"value": world * M * T / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": par_bytes,
               "d2h_bytes_per_step": out_bytes, "ms_per_step": e2e_ms}
        lib.sipnet_gpu_host_free(hout)
        lib.sipnet_gpu_host_free(hpar)
    ens.close()
    filled = None
    if world == 1 and not args.no_filled and args.members == 4096:
        filled = measure_filled(local, args.years, fp64_peak_tflops)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if M == 4096 and T == 7306:
                traffic = tj["traffic_bytes_per_launch"]      # from one ncu --set full capture of this workload
        except Exception:
            pass
        steps_per_s_kernel = M * T / (kern_ms * 1e-3)
        fp64_ach = steps_per_s_kernel * FLOP_PER_MEMBER_STEP / 1e12
        hbm_ach = steps_per_s_kernel * BYTES_PER_MEMBER_STEP / 1e9
        roof_fp64 = {"bound": "fp64", "achieved": fp64_ach, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                     "frac": fp64_ach / fp64_peak_tflops if fp64_peak_tflops > 0 else None, "traffic": traffic,
                     "peak_source": "measured live (sipnet_gpu_measure_fp64_peak, DFMA chains)",
                     "kernel": "sip::run_kernel", "kernel_ms": kern_ms,
                     "algorithmic": f"{FLOP_PER_MEMBER_STEP:.0f} FP64 flop/member-step x {M * T} member-steps"}
        roof_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_ach / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                    "algorithmic": f"{BYTES_PER_MEMBER_STEP:.0f} B/member-step x {M * T} member-steps"}
        primary, alt = (roof_fp64, roof_hbm) if (roof_fp64["frac"] or 0) >= roof_hbm["frac"] else (roof_hbm, roof_fp64)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "members_per_gpu": M, "model_steps": T, "math": args.math,
                       "block_threads": args.block or "auto",
                       "l2": "each step writes 7.66 GB of output (>> 126 MB L2), inputs are re-read from HBM"},
            "roofline": primary, "roofline_alt": alt, "clocks": clocks, "gpu_launches": int(launches),
        }
        if filled is not None:
            line["roofline_filled_gpu"] = filled
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            cpu = CpuReference([site], params, flags, cores)
            cpu.sample(2)
            v, sample = cpu.sample(max(1, min(96, M // cores)))
            cpu.close()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": cpu.kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_other_config(args):
    """BASELINE.json configs[2..4] (C3, C4, C5) on N GPUs: members sharded, no data-path collective,
    NCCL only for the final gather.  One JSON line; not the headline bench workload."""
    import torch
    import torch.distributed as dist

    from sipnet_b200 import _abi as A, api, distributed as D, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_build = time.perf_counter()
    if args.workload == "c3":      # 10k sites x 100 members, events; per-site ensemble mean/variance of NEE
        nsites_total = 10000
        nsites = args.sites or nsites_total // world
        sites, params, ms, flags = synth.config_c3(nsites=nsites, members_per_site=100, nyears=args.years,
                                                   site0=rank * nsites)
        kw = dict(outputs=A.OUT_MOMENTS, summary_cols=[A.O["nee"]], out_steps_capacity=256)
        name = f"C3: {nsites * world} sites x 100 members, {args.years} yr half-daily, events.in schedule, per-site NEE mean/variance"
    elif args.workload == "c4":    # 1M members, 1 site, ensemble mean + quantiles of NEE and GPP
        total = args.members if args.members != 4096 else 1 << 20
        M = total // world
        sites = [synth.synth_site(0, args.years, "half-daily")]
        params = synth.synth_params(M, stream=100 + rank)
        ms, flags = np.zeros(M, np.int32), dict(synth.SYNTH_FLAGS)
        # the two summary columns of the whole run stay on the device (15 GB at 131072 members): one launch per kernel
        kw = dict(outputs=A.OUT_MOMENTS, summary_cols=[A.O["nee"], A.O["gpp"]], out_steps_capacity=args.cap)
        name = f"C4: {total} members x {args.years} yr on {world} GPU(s), on-device mean/variance + exact quantiles (NEE, GPP)"
    else:                          # c5: 256k draws scored by NEE log-likelihood
        total = args.members if args.members != 4096 else 1 << 18
        M = total // world
        sites = [synth.synth_site(0, args.years, "half-daily")]
        rngobs = np.random.default_rng(7)
        sites[0].nee_obs = np.where(rngobs.uniform(size=sites[0].nsteps) < 0.2, np.nan, rngobs.normal(0, 1.5, sites[0].nsteps))
        params = synth.synth_params(M, stream=200 + rank)
        ms, flags = np.zeros(M, np.int32), dict(synth.SYNTH_FLAGS)
        kw = dict(outputs=A.OUT_LOGLIK, nee_sigma=0.5)
        name = f"C5: {total} parameter draws x {args.years} yr, on-device NEE log-likelihood, NCCL all_gather"
    t_build = time.perf_counter() - t_build
    pipelined = args.workload == "c4" and world > 1 and args.pipeline
    s_run = s_sum = None
    s_side = torch.cuda.Stream(device=local) if world > 1 else None
    if pipelined:                                        # library work on a torch-visible stream, summaries on another
        s_run, s_sum = torch.cuda.Stream(device=local), torch.cuda.Stream(device=local, priority=-1)
        kw["stream"] = s_run.cuda_stream
        kw["out_steps_capacity"] = args.cap or 1024
    ens = api.Ensemble(sites, params, ms, flags, math=A.MATH_FAST, device=local, **kw)
    T = ens.max_steps
    M_local = params.shape[1]
    cap = kw.get("out_steps_capacity", T) or T
    qs = [0.05, 0.5, 0.95]

    checks = []
    collected = []                                       # quantile tensors of the last pass (--verify)

    def c4_pipelined_pass():
        """C4 on several GPUs, segment-pipelined: while the step kernel integrates segment i+1 on the library's
        stream, the summary stream packs, exchanges (all-to-all over NVLink) and selects the quantiles of
        segment i.  The only host synchronisation is at the end of the pass."""
        ens.reset()
        packed = None
        keep = []
        ld = (M_local + 15) // 16 * 16
        for t0 in range(0, T, cap):
            t1 = min(T, t0 + cap)
            n = t1 - t0
            if packed is not None:
                s_run.wait_event(packed)                 # the previous segment's columns have been copied out
            ens.run(t0, t1)                              # asynchronous, on s_run
            mean_ptr, var_ptr = ens.device_ptr(A.GATHER_MEAN), ens.device_ptr(A.GATHER_VARIANCE)   # moments kernel, s_run
            ready = torch.cuda.Event()
            ready.record(s_run)
            with torch.cuda.stream(s_sum):
                s_sum.wait_event(ready)
                colbuf = D.DeviceArray(ens.device_ptr(A.GATHER_FULL), (2, n, M_local), (n * ld, ld, 1)).tensor(local)
                mom = torch.stack([D.DeviceArray(mean_ptr, (2, n)).tensor(local), D.DeviceArray(var_ptr, (2, n)).tensor(local)])
                sends = [D.pack_time_slices(colbuf[i], world, copy=True) for i in range(2)]
                packed = torch.cuda.Event()
                packed.record(s_sum)
                gathered = [torch.empty_like(mom) for _ in range(world)]
                dist.all_gather(gathered, mom)           # per-rank (mean, variance); combined in rank order below
                for i in range(2):
                    rows, _, _ = D.exchange_time_slices(sends[i], n, [M_local] * world)
                    keep.append(D.rows_summary(rows, qs, moments=False)[2])
                keep.append(gathered)
        s_sum.synchronize()
        s_run.synchronize()
        collected[:] = [k for k in keep if torch.is_tensor(k)]
        g = keep[-1]                                     # ordered (Chan) combination of the last segment's moments
        cnt = [np.full(g[0][0][0].shape, float(M_local))] * world
        D.combine_moments(cnt, [x[0][0].cpu().numpy() for x in g], [x[1][0].cpu().numpy() for x in g])

    def one_pass(prof=None):
        """prof: dict of phase -> seconds, filled with a device synchronize after every phase (an extra,
        untimed pass; the timed passes run without those synchronizes)."""
        def lap(name, t_prev):
            if prof is None:
                return 0.0
            torch.cuda.synchronize()
            now = time.perf_counter()
            prof[name] = prof.get(name, 0.0) + now - t_prev
            return now
        tp = lap("_", time.perf_counter())
        collected.clear()
        ens.reset()
        if args.workload == "c5":
            ens.run(0, T)
            tp = lap("run_kernel", tp)
            ll = D.DeviceArray(ens.device_ptr(A.GATHER_LOGLIK), (M_local,)).tensor(local)
            if world > 1:
                D.all_gather_members(ll, [M_local] * world)
            lap("gather", tp)
            return
        for t0 in range(0, T, cap):
            t1 = min(T, t0 + cap)
            ens.run(t0, t1)
            tp = lap("run_kernel", tp)
            n = t1 - t0
            if args.workload == "c3":                           # per-site moments stay on the device (1.2 GB at full size)
                mom = D.DeviceArray(ens.device_ptr(A.GATHER_MEAN), (ens.nsites, n)).tensor(local)
                var = D.DeviceArray(ens.device_ptr(A.GATHER_VARIANCE), (ens.nsites, n)).tensor(local)
                checks.append(float(mom[0, -1].item()) + float(var[-1, 0].item()))   # a result really is read back
                tp = lap("moments", tp)
                continue
            mean, var = ens.mean(), ens.variance()              # local shard's per-step moments
            tp = lap("moments", tp)
            if args.workload == "c4":
                if world > 1:
                    cnt = np.full_like(mean[0], float(M_local))
                    D.all_gather_moments(torch.from_numpy(cnt).cuda(), torch.from_numpy(mean[0]).cuda(),
                                         torch.from_numpy(var[0]).cuda())
                    tp = lap("moments_gather", tp)
                ld = (M_local + 15) // 16 * 16
                colbuf = D.DeviceArray(ens.device_ptr(A.GATHER_FULL), (2, n, M_local), (n * ld, ld, 1)).tensor(local)
                if world > 1 and prof is None:
                    # the select of column 0 runs on a side stream while column 1 is being exchanged
                    rows0, _, _ = D.time_transpose(colbuf[0], [M_local] * world)
                    exchanged = torch.cuda.Event()
                    exchanged.record()
                    with torch.cuda.stream(s_side):
                        s_side.wait_event(exchanged)
                        q0 = D.rows_summary(rows0, qs, moments=False)[2]
                        rows0.record_stream(s_side)
                    rows1, _, _ = D.time_transpose(colbuf[1], [M_local] * world)
                    q1 = D.rows_summary(rows1, qs, moments=False)[2]
                    torch.cuda.current_stream().wait_stream(s_side)
                    collected.extend([q0, q1])
                    continue
                for i in range(2):
                    rows = colbuf[i]
                    if world > 1:
                        rows, _, _ = D.time_transpose(rows, [M_local] * world)
                        tp = lap("time_transpose", tp)
                    collected.append(D.rows_summary(rows, qs, moments=False)[2])
                    tp = lap("quantile_select", tp)

    timed_pass = c4_pipelined_pass if pipelined else one_pass
    timed_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        timed_pass()
    barrier()
    dt = (time.perf_counter() - t0) / args.steps
    tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    prof = {}
    if args.verify and args.workload == "c4" and world > 1:   # same quantiles, bit for bit, as the one-after-the-other pass
        got = [t.clone() for t in collected]
        one_pass({})
        torch.cuda.synchronize()
        same = len(got) == len(collected) and all(torch.equal(a, b) for a, b in zip(got, collected))
        if not same:
            raise SystemExit("bench.py: overlapped C4 pass differs from the sequential pass")
    verified = args.verify and args.workload == "c4" and world > 1
    one_pass(prof)
    if pipelined:
        prof["pipelined"] = "segments of %d steps: all-to-all + select of segment i overlap the kernel of segment i+1" % cap
    if verified:
        prof["verified_against_sequential_pass"] = True
    prof.pop("_", None)
    status = ens.status()
    ens.close()
    if rank == 0:
        print(json.dumps({"metric": METRIC, "workload": name, "phases_s": {k: (round(v, 5) if isinstance(v, float) else v) for k, v in prof.items()}, "value": world * M_local * T / float(tmax.item()),
                          "unit": UNIT, "n_gpus": world, "steps": args.steps, "s_per_pass": float(tmax.item()),
                          "members_per_gpu": M_local, "model_steps": T, "input_build_s": t_build,
                          "replayed_members": int((status & A.ST_REPLAY).astype(bool).sum()),
                          "timing": "wall clock around barrier+synchronize (includes summaries and the NCCL gather)",
                          "dtype": "f64", "data": "synthetic"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload != "c2":
        run_other_config(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
