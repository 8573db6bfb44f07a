"""Host-side helpers for one-rank-per-GPU runs.

The multi-GPU DATA plane lives behind the C ABI (csrc/cpf_comm.cu: ncclBroadcast of the cell field, ncclAllReduce of the
statistics slot, both inside libcpf).  What is left for the host is what the reference's host does with Pstream
(/root/reference/src/advect.H:59-89, src/initCuda.H:207-270): decide which index range of the particle cloud a rank
tracks, and carry the 128-byte NCCL id from rank 0 to the others.  torch.distributed (gloo works, no GPU needed) is
used for that control plane only, which is also how the CPU test-suite exercises it with world_size 2.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


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


def cell_slices(cells_per_rank) -> list[tuple[int, int]]:
    """(offset, count) of every rank's cells in the merged global numbering of a decomposed solver run: ranks in order,
    each rank's cells in its local order (the numbering src/initCuda.H builds when it merges the processor meshes)."""
    out, off = [], 0
    for c in cells_per_rank:
        out.append((off, int(c)))
        off += int(c)
    return out


def share_unique_id(make_id=None, src: int = 0) -> bytes:
    """rank `src` calls make_id() (cpf_comm_unique_id) and every rank returns the same 128 bytes."""
    rank, n = world()
    if n == 1:
        return make_id()
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == src:
        raw = make_id()
        if len(raw) != 128:
            raise ValueError("the NCCL unique id is 128 bytes")
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    dist.broadcast(buf, src=src)
    return bytes(buf.tolist())


def max_over_ranks(x: float) -> float:
    """Timings are reported as the max over ranks (a CPU tensor: works on gloo; also serves as a barrier)."""
    t = torch.tensor([float(x)], dtype=torch.float64)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
