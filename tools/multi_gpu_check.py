#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--members 8192] [--years 2]

C4-shaped: one site, members sharded over ranks.  The product path -- sipnet_gpu_comm_summaries (C ABI, NCCL
all-reduce of key histograms, no member value moves) -- is cross-checked against an INDEPENDENT route through
torch.distributed (ordered all_gather combine of the moments; ONE all-to-all time-transpose + the one-GPU row
select for the quantiles): quantiles must agree bit for bit.
C5-shaped: NEE log-likelihood per draw, sipnet_gpu_comm_gather_loglik vs torch all_gather.
Every rank also integrates the whole ensemble alone and compares with numpy (moments to 1e-12, quantiles to 4e-16
relative -- same order statistics --, likelihoods bit for bit).  Prints one JSON line with timings."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sipnet_b200 import _abi as A, api, distributed as D, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=8192)
    ap.add_argument("--years", type=int, default=2)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M = args.members
    site = synth.synth_site(0, args.years, "half-daily", with_events=True)
    P = synth.synth_params(M)
    # observations: NEE of member 0 + noise (C5)
    with api.Ensemble([site], P[:, :1].copy(), None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL, device=local) as e0:
        e0.run()
        site.nee_obs = synth.synth_obs(e0.output()[A.O["nee"], :, 0].copy())
    ms = np.zeros(M, np.int32)
    _, mine = D.partition_members(ms, 1, world, rank)
    counts = [D.partition_members(ms, 1, world, r)[1].size for r in range(world)]
    cols = [A.O["nee"], A.O["gpp"]]
    qs = [0.05, 0.5, 0.95]
    T = site.nsteps
    ens = api.Ensemble([site], np.ascontiguousarray(P[:, mine]), None, synth.SYNTH_FLAGS,
                       outputs=A.OUT_MOMENTS | A.OUT_QUANTILES | A.OUT_LOGLIK, summary_cols=cols, quantiles=qs, nee_sigma=0.5,
                       device=local)
    cid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        cid.copy_(torch.frombuffer(bytearray(api.unique_comm_id()), dtype=torch.uint8))
    dist.broadcast(cid, src=0)
    ens.join_team(world, rank, bytes(cid.cpu().numpy().tobytes()))
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ens.run()
    ens.sync()
    t_run = time.perf_counter() - t0
    # ---- C5: gather of likelihoods
    ll_local = D.DeviceArray(ens.device_ptr(A.GATHER_LOGLIK), (len(mine),)).tensor(local)
    t0 = time.perf_counter()
    ll = D.all_gather_members(ll_local, counts)
    torch.cuda.synchronize()
    t_ll = time.perf_counter() - t0
    # ---- C4: moments (ordered combine) and exact quantiles (time transpose)
    mean_l, var_l = ens.mean()[0], ens.variance()[0]                  # [ncols][T] of the local shard
    cnt = np.full_like(mean_l, float(len(mine)))
    t0 = time.perf_counter()
    N, mean, var = D.all_gather_moments(torch.from_numpy(cnt).cuda(), torch.from_numpy(mean_l).cuda(),
                                        torch.from_numpy(var_l).cuda())
    colbuf = D.DeviceArray(ens.device_ptr(A.GATHER_FULL), (len(cols), T, len(mine)), (T * ens_ld(len(mine)), ens_ld(len(mine)), 1)).tensor(local)
    quant = []
    for i in range(len(cols)):
        rows, ta, tb = D.time_transpose(colbuf[i], counts)
        _, _, q = D.rows_summary(rows, qs)
        quant.append((ta, tb, q))
    torch.cuda.synchronize()
    t_sum = time.perf_counter() - t0
    # ---- the product path: the C ABI's team (after the reference route, which reads the handle's LOCAL moments)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ens.team_summaries()
    ens.sync()
    t_team = time.perf_counter() - t0
    team_q, team_mean, team_var = ens.quantiles()[0], ens.mean()[0], ens.variance()[0]     # [ncols][nq][T], [ncols][T]
    team_ll = ens.team_loglik()
    ok = True
    ok &= bool(np.array_equal(team_ll, ll.cpu().numpy()))
    ok &= bool(np.allclose(team_mean, mean, rtol=1e-13, atol=1e-300) and np.allclose(team_var, var, rtol=1e-9, atol=1e-24))
    for i in range(len(cols)):
        ta, tb, q = quant[i]
        ok &= bool(np.array_equal(team_q[i][:, ta:tb], q.cpu().numpy(), equal_nan=True))            # bit for bit
    if not args.no_check:
        # every rank checks its share of the quantiles against rank-local numpy on a full single-GPU run
        full = api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL | A.OUT_LOGLIK, nee_sigma=0.5, device=local)
        full.run()
        out = full.output()
        ll_full = full.loglik()
        full.close()
        ok &= bool(np.array_equal(ll.cpu().numpy(), ll_full))
        for i, c in enumerate(cols):
            x = out[c]
            ok &= bool(np.allclose(mean[i], x.mean(axis=1), rtol=1e-12, atol=1e-300))
            ok &= bool(np.allclose(var[i], x.var(axis=1), rtol=1e-9, atol=1e-300))
            ta, tb, q = quant[i]
            ok &= bool(np.allclose(q.cpu().numpy(), np.quantile(x[ta:tb], qs, axis=1), rtol=4e-16, atol=1e-300))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "multi_gpu", "world": world, "members": M, "steps": T, "ok": bool(flag.item()),
                          "run_s": t_run, "loglik_gather_s": t_ll, "all_to_all_route_s": t_sum, "team_summaries_s": t_team,
                          "member_steps_per_s": M * T / t_run}), flush=True)
    ens.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


def ens_ld(m):
    return (m + 15) // 16 * 16


if __name__ == "__main__":
    main()
