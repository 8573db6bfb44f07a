"""One rank per GPU behind the C ABI (cpf_comm.cu), asynchronous statistics, global particle ids.

The reference runs its whole CUDA path on the MPI master with one GPU (src/initCuda.H:207-270, src/advect.H:59-89);
here every rank tracks an index range of the cloud on its own GPU.  The property that makes that safe to use: N ranks
reproduce the single-GPU run bit for bit -- positions, tet ids, velocities and the summed statistics."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, make_case, n_gpus

pytestmark = pytest.mark.gpu


def _run(api, pm, U_list, p, nsub, dt, *, base=0, **cfg):
    tr = api.ParticleTracker(**cfg)
    tr.upload_poly(pm)
    tr.update_velocity(U_list[0])
    tr.set_particles(p)
    tr.set_particle_id_base(base)
    if tr.cfg.rng == api.RNG_XORWOW:
        tr.init_rng()
    tr.locate_initial()
    for k, U in enumerate(U_list):
        tr.update_velocity(U)
        tr.substeps(nsub, dt)
    out = tr.download()
    st = tr.stats()
    tr.close()
    return out, st


@pytest.mark.parametrize("rng", ["philox", "xorwow"])
def test_index_range_shards_reproduce_the_single_gpu_run(synth, orc, rng):
    """the random-walk streams are keyed by GLOBAL particle id, so two contexts that track the two halves of a
    cloud (cpf_set_particle_id_base) give exactly the particles the single context gives."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(9, 8, 7), jitter=0.2, n=30000)
    fields = [U, synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=7.0)]
    cfg = dict(rng=api.RNG_PHILOX if rng == "philox" else api.RNG_XORWOW, diffusion_coeff=2e-3, fuse_substeps=5, sort_interval=7)
    (pa, va, ta), sa = _run(api, pm, fields, p, 12, 0.01, **cfg)
    h = 13001
    (p0, v0, t0), s0 = _run(api, pm, fields, p[:h], 12, 0.01, base=0, **cfg)
    (p1, v1, t1), s1 = _run(api, pm, fields, p[h:], 12, 0.01, base=h, **cfg)
    assert np.array_equal(np.concatenate([t0, t1]), ta)
    assert np.array_equal(np.concatenate([p0, p1]).view(np.uint64), pa.view(np.uint64))
    assert np.array_equal(np.concatenate([v0, v1])[:, :3].view(np.uint64), va[:, :3].view(np.uint64))
    for k in ("n_substeps", "n_reflections", "n_hops", "n_active", "n_escaped"):
        assert s0[k] + s1[k] == sa[k], k
    # and the halves differ from each other's streams: without the base the second half would repeat the first half's noise
    (p1b, _, _), _ = _run(api, pm, fields, p[h:], 12, 0.01, base=0, **cfg)
    assert not np.array_equal(p1b.view(np.uint64), p1.view(np.uint64))


def test_async_statistics_equal_the_blocking_scan(synth, orc):
    """cpf_stats_request / cpf_stats_collect: a request is a snapshot at its place in the stream; light requests derive
    n_active from the escape and freeze counters."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(10, 6, 6), jitter=0.15, n=40000, field=(1.5, 0.1, 0.0))
    kind = np.zeros(len(pm.patch_starts) - 1, dtype=np.int32)
    kind[1] = api.PATCH_ESCAPE  # x-max: through-flow leaves
    tr = api.ParticleTracker(rng=api.RNG_PHILOX, diffusion_coeff=1e-3, fuse_substeps=6, sort_interval=12)
    tr.upload_poly(pm, patch_kind=kind)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    tr.substeps(6, 0.02)
    tr.stats_request(full=True)
    tr.substeps(6, 0.02)          # enqueued behind the request: must not leak into it
    tr.stats_request(full=False)
    tr.substeps(6, 0.02)
    tr.stats_request(full=False)
    a = tr.stats_collect()
    b = tr.stats_collect()
    c = tr.stats_collect()
    ref = tr.stats()              # blocking full scan of the final state
    assert a["full"] and not b["full"] and not c["full"]
    assert a["n_substeps"] < b["n_substeps"] < c["n_substeps"] == ref["n_substeps"]
    assert a["n_escaped"] <= b["n_escaped"] <= c["n_escaped"] == ref["n_escaped"] and ref["n_escaped"] > 100
    assert c["n_active"] == ref["n_active"], "light n_active = last scan - escapes - freezes since"
    assert a["n_active"] > b["n_active"] > c["n_active"]
    for k in ("n_reflections", "n_hops", "n_exact"):
        assert c[k] == ref[k], k
    with pytest.raises(api.CpfError):
        tr.stats_collect()        # nothing outstanding
    for _ in range(4):
        tr.stats_request()
    with pytest.raises(api.CpfError):
        tr.stats_request()        # ring of four
    tr.close()


def test_communicator_of_one_rank_needs_no_nccl(synth, orc):
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(6, 5, 4), jitter=0.1, n=5000)
    outs = []
    for mode in ("plain", "bcast", "slices"):
        tr = api.ParticleTracker(rng=api.RNG_NONE)
        tr.upload_poly(pm)
        if mode != "plain":
            tr.comm_init(None, 0, 1)
            assert tr.comm_info()[:2] == (0, 1)
        if mode == "plain":
            tr.update_velocity(U)
        elif mode == "bcast":
            tr.update_velocity_bcast(U, root=0)
        else:
            tr.update_velocity_slices(0, U)
        tr.set_particles(p)
        tr.locate_initial()
        tr.substeps(10, 0.02)
        outs.append(tr.download())
        tr.close()
    for o in outs[1:]:
        assert np.array_equal(o[2], outs[0][2]) and np.array_equal(o[0].view(np.uint64), outs[0][0].view(np.uint64))


@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("rng", ["philox", "xorwow"])
def test_two_ranks_over_nccl_reproduce_one_gpu(tmp_path, synth, orc, rng):
    """Two processes, one GPU each, NCCL inside libcpf: broadcast field, sliced field, summed statistics."""
    from cudaparticlesfoam_b200 import api

    n, nsub, dt = 24000, 8, 0.01
    worker = os.path.join(ROOT, "tests", "multi_rank_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(tmp_path), rng, str(n), str(nsub), str(dt)], cwd=ROOT,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [pr.communicate(timeout=600) for pr in procs]
    for pr, (so, se) in zip(procs, outs):
        assert pr.returncode == 0, se[-3000:]
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    meta = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    # the same job on one GPU, through the single-rank calls
    pm = synth.box_mesh(9, 8, 7, jitter=0.2)
    lo, hi = pm.lo + 0.02, pm.hi - 0.02
    fields = [synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=5.0 + k) for k in range(4)]
    tr = api.ParticleTracker(rng=api.RNG_PHILOX if rng == "philox" else api.RNG_XORWOW, diffusion_coeff=2e-3, fuse_substeps=4, sort_interval=6)
    tr.upload_poly(pm)
    tr.update_velocity(fields[0])
    tr.seed_box(n, lo, hi)
    if rng == "xorwow":
        tr.init_rng()
    tr.locate_initial()
    for U in fields:
        tr.update_velocity(U)
        tr.substeps(nsub, dt)
    p1, v1, t1 = tr.download()
    s1 = tr.stats()
    tr.close()
    assert np.array_equal(np.concatenate([res[0]["t"], res[1]["t"]]), t1)
    assert np.array_equal(np.concatenate([res[0]["p"], res[1]["p"]]).view(np.uint64), p1.view(np.uint64))
    assert np.array_equal(np.concatenate([res[0]["v"], res[1]["v"]])[:, :3].view(np.uint64), v1[:, :3].view(np.uint64))
    for m in meta:  # the statistics every rank holds are the sums over both ranks
        for k in ("n_particles", "n_active", "n_substeps", "n_reflections", "n_hops", "n_escaped"):
            assert m["stats"][k] == s1[k], (k, m["stats"][k], s1[k])
        assert abs(m["stats"]["kinetic_energy"] - s1["kinetic_energy"]) <= 1e-9 * abs(s1["kinetic_energy"])
        assert m["nranks"] == 2 and m["nccl"] > 20000
