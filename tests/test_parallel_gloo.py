"""world_size-2 gloo test of the host-side control plane of a one-rank-per-GPU run (index partition, NCCL id
hand-over, cell slices, max-over-ranks timing): the same helpers bench.py uses; the data plane (NCCL inside libcpf)
is covered on two GPUs by tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cudaparticlesfoam_b200 import parallel, synth

    start, count = parallel.partition(n_total, world, rank)
    # every rank seeds ITS slice of the one global cloud (bench.py build_inputs does exactly this)
    cloud = synth.seed_box(n_total, (0, 0, 0), (1, 1, 1))[start:start + count]
    # rank 0 creates the communicator id, everyone gets the same 128 bytes (here a stand-in for ncclGetUniqueId: no GPU)
    uid = parallel.share_unique_id((lambda: bytes(range(128))) if rank == 0 else None)
    # cells of a decomposed solver run: rank r owns 400 + 100 r cells
    counts = [400 + 100 * r for r in range(world)]
    sl = parallel.cell_slices(counts)
    tmax = parallel.max_over_ranks(1.0 + rank)
    out[rank] = dict(start=start, count=count, cloud_sum=float(cloud[:, :3].sum()), uid=uid, slice=sl[rank], tmax=tmax)
    dist.destroy_process_group()


def test_partition_properties():
    from cudaparticlesfoam_b200 import parallel

    for n, w in ((10, 3), (100_000_000, 8), (7, 8), (0, 4)):
        parts = [parallel.partition(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (s0, c0), (s1, _) in zip(parts, parts[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        parallel.partition(10, 2, 2)


def test_two_rank_gloo_plumbing():
    from cudaparticlesfoam_b200 import synth

    world, n_total = 2, 10_001
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert (r0["start"], r0["count"], r1["start"], r1["count"]) == (0, 5001, 5001, 5000)
    full = synth.seed_box(n_total, (0, 0, 0), (1, 1, 1))[:, :3].sum()
    assert abs(r0["cloud_sum"] + r1["cloud_sum"] - full) < 1e-9 * abs(full)
    assert r0["uid"] == r1["uid"] == bytes(range(128))               # the id reached rank 1 unchanged
    assert r0["slice"] == (0, 400) and r1["slice"] == (400, 500)      # merged cell numbering of a decomposed run
    assert r0["tmax"] == r1["tmax"] == 2.0
