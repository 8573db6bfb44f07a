"""world_size-2 gloo test of the multi-GPU host plumbing (partition / field broadcast / statistics),
run on CPU tensors: the same code path bench.py drives over NCCL."""
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
    # every rank seeds ITS slice of the one global cloud
    cloud = synth.seed_box(n_total, (0, 0, 0), (1, 1, 1))[start:start + count]
    # rank 0 owns the solver field; everyone receives it
    ncell = 1000
    U = torch.zeros((ncell, 3), dtype=torch.float64)
    if rank == 0:
        U[:] = torch.from_numpy(synth.field_uniform_vortex(np.random.default_rng(1).random((ncell, 3))))
    parallel.broadcast_field(U)
    stats = {"n_particles": count, "n_active": count - rank, "n_reflections": 10 * (rank + 1), "kinetic_energy": 0.5 * (rank + 1)}
    red = parallel.reduce_stats(stats)
    gat = parallel.gather_stats(stats)
    tmax = parallel.max_over_ranks(1.0 + rank)
    out[rank] = dict(start=start, count=count, cloud_sum=float(cloud[:, :3].sum()), usum=float(U.sum()), red=red,
                     gat=gat, tmax=tmax)
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
    assert r0["usum"] == r1["usum"] and r0["usum"] != 0.0          # broadcast delivered rank 0's field
    assert r0["red"] == r1["red"]
    assert r0["red"]["n_particles"] == n_total and r0["red"]["n_active"] == n_total - 1
    assert r0["red"]["n_reflections"] == 30 and abs(r0["red"]["kinetic_energy"] - 1.5) < 1e-12
    assert r1["gat"] is None and len(r0["gat"]) == 2 and r0["gat"][1]["n_reflections"] == 20
    assert r0["tmax"] == r1["tmax"] == 2.0
