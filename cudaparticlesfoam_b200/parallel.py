"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed).

The path shards trivially: tracers do not interact, so particles are partitioned by
contiguous index range, the mesh is replicated, and the only per-step communication is
  * the solver's cell velocity field: rank 0 -> everyone (NCCL broadcast into a device buffer that
    `cpf_update_velocity(on_device=1)` consumes on the same stream), and
  * a fixed-size statistics vector: everyone -> rank 0 (gather) / all (all-reduce).
The reference funnels everything to the MPI master and a single GPU
(/root/reference/src/advect.H:59-89, Pstream::gatherList).  Works with the gloo backend on CPU
tensors too, which is how the CPU test-suite exercises it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

STAT_KEYS = ("n_particles", "n_active", "n_negative_tet", "n_escaped", "n_reflections", "n_exact", "n_hops", "n_substeps",
             "kinetic_energy")


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition(n_total: int, n_ranks: int, rank: int) -> tuple[int, int]:
    """Contiguous index range [start, start+count) of `rank`; the first n_total % n_ranks ranks get one extra."""
    if not (0 <= rank < n_ranks):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_total), int(n_ranks))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def broadcast_field(buf: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Per-step velocity field fan-out; `buf` is [nCells,3] float64 on every rank (device for NCCL)."""
    if buf.dtype != torch.float64 or buf.dim() != 2 or buf.shape[1] != 3:
        raise ValueError("cell field must be float64 [nCells,3]")
    if world()[1] > 1:
        dist.broadcast(buf, src=src)
    return buf


def stats_vector(stats: dict, device=None) -> torch.Tensor:
    return torch.tensor([float(stats.get(k, 0)) for k in STAT_KEYS], dtype=torch.float64, device=device)


def reduce_stats(stats: dict, device=None) -> dict:
    """Sum of every rank's counters, available on all ranks."""
    v = stats_vector(stats, device)
    if world()[1] > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    out = {k: (float(x) if k == "kinetic_energy" else int(round(float(x)))) for k, x in zip(STAT_KEYS, v.tolist())}
    return out


def gather_stats(stats: dict, device=None, dst: int = 0):
    """Per-rank counters gathered on `dst` (list of dicts there, None elsewhere)."""
    rank, n = world()
    v = stats_vector(stats, device)
    if n == 1:
        return [dict(zip(STAT_KEYS, v.tolist()))]
    bucket = [torch.empty_like(v) for _ in range(n)] if rank == dst else None
    dist.gather(v, bucket, dst=dst)
    if rank != dst:
        return None
    return [dict(zip(STAT_KEYS, b.tolist())) for b in bucket]


def max_over_ranks(x: float, device=None) -> float:
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
