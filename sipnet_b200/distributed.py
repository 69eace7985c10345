"""Multi-GPU layer: one process per GPU, members sharded with no data-path collective.

Members are independent (no cross-member term anywhere in the model), so `run` needs no
communication.  Collectives appear only in the final gather of summaries:

  * log-likelihoods (config C5): all_gather of one double per member;
  * ensemble moments (C4): all_gather of per-rank (count, mean, M2) per (column, step), combined in
    RANK ORDER on every rank (Chan's parallel formula), so the result is bit-reproducible;
  * exact ensemble quantiles of a single-site ensemble spread over ranks (C4): ONE exchange step --
    an all-to-all time-transpose, after which every rank holds the complete ensemble for its share
    of the steps and selects the quantiles locally (sipnet_gpu_rows_summary).

Backends: "nccl" over NVLink/NVSwitch on the GPU box, "gloo" in the CPU tests (same code path; the
local row summary is then numpy).  Python/torch is plumbing here; the arithmetic is in the library.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np


def partition_members(member_site: np.ndarray, nsites: int, world: int, rank: int) -> Tuple[np.ndarray, np.ndarray]:
    """-> (site indices, member indices) owned by `rank`.

    Whole sites per rank when there are at least `world` sites (forcing is not replicated);
    otherwise every rank gets every site and an even, contiguous share of each site's members."""
    member_site = np.asarray(member_site)
    if nsites >= world:
        bounds = [(nsites * r) // world for r in range(world + 1)]
        sites = np.arange(bounds[rank], bounds[rank + 1])
        members = np.flatnonzero((member_site >= bounds[rank]) & (member_site < bounds[rank + 1]))
        return sites, members
    sites = np.arange(nsites)
    chunks = []
    for s in range(nsites):
        idx = np.flatnonzero(member_site == s)
        lo, hi = (idx.size * rank) // world, (idx.size * (rank + 1)) // world
        chunks.append(idx[lo:hi])
    return sites, np.concatenate(chunks) if chunks else np.zeros(0, np.int64)


def combine_moments(counts: Sequence[np.ndarray], means: Sequence[np.ndarray], variances: Sequence[np.ndarray]):
    """Chan et al. pairwise combination, applied in rank order.  Inputs are per-rank arrays of the
    same shape (population variance); returns (count, mean, variance)."""
    n = np.array(counts[0], dtype=np.float64)
    # a rank with no finite member in a row reports (0, NaN, NaN): it must not poison the combination
    mu = np.where(n > 0, np.array(means[0], dtype=np.float64), 0.0)
    m2 = np.where(n > 0, np.array(variances[0], dtype=np.float64) * n, 0.0)
    for nb, mub, varb in zip(counts[1:], means[1:], variances[1:]):
        nb = np.asarray(nb, dtype=np.float64)
        has = nb > 0
        mub = np.where(has, np.asarray(mub, dtype=np.float64), mu)
        varb = np.where(has, np.asarray(varb, dtype=np.float64), 0.0)
        tot = n + nb
        delta = mub - mu
        safe = np.where(tot > 0, tot, 1.0)
        mu = mu + delta * nb / safe
        m2 = m2 + varb * nb + delta * delta * n * nb / safe
        n = tot
    with np.errstate(invalid="ignore", divide="ignore"):
        return n, np.where(n > 0, mu, np.nan), np.where(n > 0, m2 / np.where(n > 0, n, 1.0), np.nan)


def all_gather_members(local: "torch.Tensor", local_counts: List[int], group=None) -> "torch.Tensor":
    """Gather one value per member (e.g. log-likelihoods) from every rank, rank-major.
    `local_counts[r]` = number of members on rank r (may differ between ranks)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    width = max(local_counts)
    pad = torch.full((width,), float("nan"), dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, local_counts)])


def all_gather_moments(count: "torch.Tensor", mean: "torch.Tensor", var: "torch.Tensor", group=None):
    """all_gather per-rank (count, mean, var) and combine them in rank order (deterministic)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    packed = torch.stack([count.to(mean.dtype), mean, var]).contiguous()
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    host = [o.cpu().numpy() for o in out]
    return combine_moments([h[0] for h in host], [h[1] for h in host], [h[2] for h in host])


def pack_time_slices(cols: "torch.Tensor", world: int, copy: bool = False) -> List["torch.Tensor"]:
    """cols [nsteps][m_local] (a view of the library's column buffer) -> contiguous per-destination blocks of steps.
    copy=True always copies (a slice of an unpadded buffer is already contiguous and would otherwise alias it), so
    that the column buffer may be overwritten while the exchange is still in flight."""
    nsteps = cols.shape[0]
    tb = [(nsteps * r) // world for r in range(world + 1)]
    if copy:
        return [cols[tb[r]:tb[r + 1], :].clone(memory_format=__import__("torch").contiguous_format) for r in range(world)]
    return [cols[tb[r]:tb[r + 1], :].contiguous() for r in range(world)]


def exchange_time_slices(send: List["torch.Tensor"], nsteps: int, local_counts: List[int], group=None):
    """all-to-all of packed blocks -> ([t1 - t0][M_total] rows of this rank's share of the steps, t0, t1)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tb = [(nsteps * r) // world for r in range(world + 1)]
    recv = [torch.empty((tb[rank + 1] - tb[rank], local_counts[r]), dtype=send[0].dtype, device=send[0].device)
            for r in range(world)]
    dist.all_to_all(recv, send, group=group) if dist.get_backend(group) != "gloo" else _all_to_all_gloo(recv, send, group)
    return torch.cat(recv, dim=1), tb[rank], tb[rank + 1]


def time_transpose(cols: "torch.Tensor", local_counts: List[int], group=None) -> Tuple["torch.Tensor", int, int]:
    """The one exchange step for exact cross-rank quantiles.

    cols: [nsteps][m_local] (this rank's members of ONE output column).  Returns ([t1 - t0][M_total],
    t0, t1): all members for this rank's contiguous share of the steps, members in rank-major order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    return exchange_time_slices(pack_time_slices(cols, world), cols.shape[0], local_counts, group)


def _all_to_all_gloo(recv, send, group=None):
    """gloo has no all_to_all for uneven splits on every build: emulate it with broadcasts of lists
    (CPU tests only; sizes are tiny there)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    for src in range(world):
        for dst in range(world):
            if src == dst:
                if rank == src:
                    recv[src].copy_(send[dst])
                continue
            if rank == src:
                dist.send(send[dst], dst, group=group)
            elif rank == dst:
                dist.recv(recv[src], src, group=group)


def rows_summary(rows: "torch.Tensor", probs: Sequence[float], moments: bool = True):
    """mean, variance, quantiles per row of a [nrows][ncols] tensor.  CUDA tensors go through the
    library's reducers (sipnet_gpu_rows_summary); CPU tensors (gloo tests) through numpy.
    moments=False skips mean/variance (returned as None) when only the quantiles are wanted."""
    import torch
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    if rows.is_cuda:
        from . import api
        lib = api.load_library()
        nrows, ncols = rows.shape
        mean = torch.empty(nrows, dtype=torch.float64, device=rows.device) if moments else None
        var = torch.empty_like(mean) if moments else None
        quant = torch.empty((max(probs.size, 1), nrows), dtype=torch.float64, device=rows.device)
        rc = lib.sipnet_gpu_rows_summary(rows.device.index or 0, C.c_void_p(rows.data_ptr()), nrows, ncols, rows.stride(0),
                                         C.c_void_p(probs.ctypes.data), int(probs.size),
                                         C.c_void_p(mean.data_ptr() if moments else None),
                                         C.c_void_p(var.data_ptr() if moments else None), C.c_void_p(quant.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise api.SipnetGpuError(rc, (lib.sipnet_gpu_last_error() or b"").decode())
        return mean, var, quant[: probs.size]
    a = rows.numpy()
    if not moments:
        with np.errstate(invalid="ignore"):
            fin = np.where(np.isfinite(a), a, np.nan)
            return None, None, torch.from_numpy(np.ascontiguousarray(np.nanquantile(fin, probs, axis=1)))
    with np.errstate(invalid="ignore"):
        fin = np.where(np.isfinite(a), a, np.nan)
        mean = np.nanmean(fin, axis=1)
        var = np.nanvar(fin, axis=1)
        q = np.nanquantile(fin, probs, axis=1) if probs.size else np.zeros((0, a.shape[0]))
    return torch.from_numpy(mean), torch.from_numpy(var), torch.from_numpy(np.ascontiguousarray(q))


class DeviceArray:
    """Zero-copy torch view of a library-owned device buffer (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], strides_elems: Tuple[int, ...] = None, typestr: str = "<f8"):
        itemsize = 8 if typestr == "<f8" else 4
        self.__cuda_array_interface__ = {
            "shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
            "strides": None if strides_elems is None else tuple(int(x) * itemsize for x in strides_elems)}

    def tensor(self, device: int = 0):
        import torch
        return torch.as_tensor(self, device=f"cuda:{device}")
