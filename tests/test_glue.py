"""The drop-in glue (src/initCuda.H, src/advect.H) compiled over the OpenFOAM shim exactly the way
the two reference solvers include it, then run end to end on the GPU and checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT, make_case, n_gpus


def _build_driver(tmp_path, convex=True):
    exe = str(tmp_path / ("glue_driver_convex" if convex else "glue_driver_bary"))
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror=return-type", f"-I{ROOT}/include", f"-I{ROOT}/src", f"-I{ROOT}/tests/glue"]
    if convex:
        cmd.append("-DConvexPoly")  # applications/*/Make/options:1-5 (RTX=false)
    cmd += [f"{ROOT}/tests/glue/glue_driver.C", "-o", exe, f"-L{ROOT}/cudaparticlesfoam_b200", "-lcpf",
            f"-Wl,-rpath,{ROOT}/cudaparticlesfoam_b200"]
    subprocess.check_call(cmd)
    return exe


def _write_case(path, pm, fields, n, save_interval, rng, deltaT, dt, D, lo, hi):
    with open(path, "wb") as f:
        f.write(struct.pack("8i", pm.n_points, pm.n_faces, pm.n_internal, pm.n_cells, len(pm.patch_starts) - 1, n, save_interval, rng))
        f.write(struct.pack("9d", deltaT, dt, D, *lo, *hi))
        for a in (pm.points, pm.face_offsets, pm.face_verts, pm.owner, pm.neighbour, pm.cell_centres, pm.patch_starts):
            f.write(np.ascontiguousarray(a).tobytes())
        for U in fields:
            f.write(np.ascontiguousarray(U, dtype=np.float64).tobytes())


def test_glue_compiles_and_fails_loudly_without_a_device(tmp_path, synth, orc):
    exe = _build_driver(tmp_path)
    exe2 = _build_driver(tmp_path, convex=False)
    assert os.path.exists(exe) and os.path.exists(exe2)
    if HAVE_GPU:
        return
    pm, mesh, U, p = make_case(synth, orc, dims=(3, 3, 3), jitter=0.0, n=10)
    _write_case(tmp_path / "case.bin", pm, [U], 10, 10, 0, 0.01, 0.005, 0.0, (0.1, 0.1, 0.1), (0.9, 0.9, 0.9))
    r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / "out.bin"), "1"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("convex", [True, False], ids=["ConvexPoly", "RTX"])
def test_glue_end_to_end_matches_oracle(tmp_path, synth, orc, convex):
    pm, mesh, U, _ = make_case(synth, orc, dims=(9, 8, 7), jitter=0.15, n=1)
    n, dt, deltaT, nsteps, save = 6000, 0.004, 0.03, 3, 5
    lo, hi = (0.05, 0.05, 0.05), (0.95, 0.95, 0.95)
    fields = [synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=2 * np.pi * (1 + 0.2 * k)) for k in range(nsteps)]
    _write_case(tmp_path / "case.bin", pm, fields, n, save, 0, deltaT, dt, 0.0, lo, hi)
    exe = _build_driver(tmp_path, convex)
    r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / "out.bin"), str(nsteps)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "ADVECT_MODE: " + ("ConvexPoly" if convex else "RTX") in r.stdout and "nCycles: 8" in r.stdout
    raw = open(tmp_path / "out.bin", "rb").read()
    (m,) = struct.unpack_from("q", raw, 0)
    assert m == n
    pp = np.frombuffer(raw, dtype=np.float64, count=4 * n, offset=8).reshape(n, 4)
    vv = np.frombuffer(raw, dtype=np.float64, count=4 * n, offset=8 + 32 * n).reshape(n, 4)
    tt = np.frombuffer(raw, dtype=np.int32, count=n, offset=8 + 64 * n)
    (step,) = struct.unpack_from("i", raw, 8 + 68 * n)
    ncyc = int(max(np.ceil(deltaT / dt), 1))
    assert step == nsteps * ncyc
    # oracle: same seeding stream, same sub-cycling rule (src/advect.H:36-37)
    p = synth.seed_box(n, lo, hi)
    cl = orc.Cloud.make(p, orc.locate_brute(mesh, p))
    for k in range(nsteps):
        orc.substeps(mesh, cl, orc.expand_velocity(mesh, fields[k]), ncyc, deltaT / ncyc, convex=convex)
    assert np.array_equal(tt, cl.tet)
    assert np.array_equal(pp.view(np.uint64), cl.p.view(np.uint64))
    assert np.array_equal(vv[:, :3].view(np.uint64), cl.vel[:, :3].view(np.uint64))
    # VTU cadence of the original: file 0 from init, then step+1 whenever step % saveInterval == 0
    got = sorted(f for f in os.listdir(tmp_path) if f.startswith("particle_"))
    want = ["particle_0000.vtu"] + [f"particle_{s + 1:04d}.vtu" for s in range(step) if s % save == 0]
    assert got == sorted(want)
    head = open(tmp_path / got[1]).read()
    for name in ("Position", "ParticleType", "ParticleID", "ParticleTetID"):
        assert f"Name='{name}'" in head


def _read_out(path, n):
    raw = open(path, "rb").read()
    (m,) = struct.unpack_from("q", raw, 0)
    assert m == n, (m, n)
    pp = np.frombuffer(raw, dtype=np.float64, count=4 * n, offset=8).reshape(n, 4)
    vv = np.frombuffer(raw, dtype=np.float64, count=4 * n, offset=8 + 32 * n).reshape(n, 4)
    tt = np.frombuffer(raw, dtype=np.int32, count=n, offset=8 + 64 * n)
    (step,) = struct.unpack_from("i", raw, 8 + 68 * n)
    cc = np.frombuffer(raw, dtype=np.int32, count=n, offset=8 + 68 * n + 4)
    return pp, vv, tt, step, cc


@pytest.mark.gpu
def test_glue_with_the_stock_dictionary_runs_the_default_configuration(tmp_path, synth, orc):
    """A cudaParticlesDict without any of the optional keys: XORWOW random walk (the reference's stream), library-default
    fusing, sort every 50 sub-steps -- the configuration a user who changes nothing gets.  It must equal the Python
    mirror of the two snippets with a default-constructed config, which tests/test_gpu_reference_pin.py pins to the
    reference's own kernels."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, _ = make_case(synth, orc, dims=(9, 8, 7), jitter=0.15, n=1)
    n, dt, deltaT, nsteps, save, D = 8000, 0.004, 0.05, 3, 10, 2e-3
    lo, hi = (0.05, 0.05, 0.05), (0.95, 0.95, 0.95)
    fields = [synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=2 * np.pi * (1 + 0.2 * k)) for k in range(nsteps)]
    _write_case(tmp_path / "case.bin", pm, fields, n, save, -1, deltaT, dt, D, lo, hi)   # -1: no randomWalk key at all
    exe = _build_driver(tmp_path, True)
    r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / "out.bin"), str(nsteps)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "#adv: Random Seed=1591593751" in r.stdout
    pp, vv, tt, step, cc = _read_out(tmp_path / "out.bin", n)
    tr = api.ParticleTracker(dt=dt, diffusion_coeff=D, save_interval=save)   # everything else: cpf_default_config
    assert (tr.cfg.rng, tr.cfg.fuse_substeps, tr.cfg.sort_interval) == (api.RNG_XORWOW, 0, 50)
    tr.upload_poly(pm)
    tr.update_velocity(fields[0])
    tr.seed_box(n, lo, hi)
    tr.init_rng()
    tr.locate_initial()
    for k in range(nsteps):
        tr.advect(fields[k], deltaT)
    p2, v2, t2 = tr.download()
    st = tr.stats()
    tr.close()
    assert step == tr.step and np.array_equal(tt, t2)
    assert np.array_equal(pp.view(np.uint64), p2.view(np.uint64)) and np.array_equal(vv[:, :3].view(np.uint64), v2[:, :3].view(np.uint64))
    assert st["n_exact"] < 0.2 * st["n_substeps"], "the default configuration must run on the filtered pipeline"


@pytest.mark.gpu
@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("rw", ["none", "xorwow"])
def test_glue_decomposed_run_one_rank_per_gpu(tmp_path, synth, orc, rw):
    """The nProcs() > 1 branch of the two snippets, executed: two processes over the shim's file-based Pstream, each
    with the processor mesh decomposePar would give it, each on its own GPU, NCCL inside libcpf for the field slices
    and the statistics.  Positions, velocities and containing cells must equal the serial run of the same case."""
    pm, mesh, U, _ = make_case(synth, orc, dims=(8, 7, 6), jitter=0.15, n=1)
    n, dt, deltaT, nsteps, save, D = 9001, 0.004, 0.03, 3, 5, (2e-3 if rw == "xorwow" else 0.0)
    lo, hi = (0.05, 0.05, 0.05), (0.95, 0.95, 0.95)
    fields = [synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=2 * np.pi * (1 + 0.2 * k)) for k in range(nsteps)]
    code = 1 if rw == "xorwow" else 0
    exe = _build_driver(tmp_path, True)
    ser = tmp_path / "serial"
    ser.mkdir()
    _write_case(ser / "case.bin", pm, fields, n, save, code, deltaT, dt, D, lo, hi)
    r = subprocess.run([exe, str(ser / "case.bin"), str(ser / "out.bin"), str(nsteps)], cwd=ser, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p1, v1, t1, step1, c1 = _read_out(ser / "out.bin", n)
    par = tmp_path / "parallel"
    par.mkdir()
    subs = synth.processor_meshes(pm, 2)
    procs = []
    for rank, sub in enumerate(subs):
        (par / f"processor{rank}").mkdir()
        c0 = rank * pm.n_cells // 2
        _write_case(par / f"case{rank}.bin", sub, [f[c0:c0 + sub.n_cells] for f in fields], n, save, code, deltaT, dt, D, lo, hi)
        env = dict(os.environ, CPF_SHIM_NPROCS="2", CPF_SHIM_RANK=str(rank), CPF_SHIM_DIR=str(par))
        procs.append(subprocess.Popen([exe, str(par / f"case{rank}.bin"), str(par / f"out{rank}.bin"), str(nsteps)], cwd=par, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [pr.communicate(timeout=600) for pr in procs]
    for pr, (so, se) in zip(procs, outs):
        assert pr.returncode == 0, se[-3000:]
    cnt = [n // 2 + (1 if r < n % 2 else 0) for r in range(2)]
    res = [_read_out(par / f"out{r}.bin", cnt[r]) for r in range(2)]
    pp = np.concatenate([x[0] for x in res]); vv = np.concatenate([x[1] for x in res]); cc = np.concatenate([x[4] for x in res])
    assert res[0][3] == res[1][3] == step1
    assert np.array_equal(pp.view(np.uint64), p1.view(np.uint64)), "positions of the decomposed run differ from the serial run"
    assert np.array_equal(vv[:, :3].view(np.uint64), v1[:, :3].view(np.uint64))
    assert np.array_equal(cc, c1), "containing cells (global ids) differ"
    # every rank wrote its own VTU series into its processor directory, with GLOBAL particle ids
    for rank in range(2):
        files = sorted(f for f in os.listdir(par / f"processor{rank}") if f.startswith("particle_"))
        assert files and files[0] == "particle_0000.vtu"


@pytest.mark.gpu
def test_glue_optional_keys_reach_the_library(tmp_path, synth, orc):
    """relocateLost / integrator / fuseSubSteps from the dictionary: RK2 through the glue equals the oracle extension, and
    relocateLost = true (cpf_relocate_lost after every chunk) changes nothing when no particle loses its tet."""
    pm, mesh, U, _ = make_case(synth, orc, dims=(8, 7, 6), jitter=0.15, n=1)
    n, dt, deltaT, nsteps, save = 5000, 0.004, 0.02, 2, 5
    lo, hi = (0.05, 0.05, 0.05), (0.95, 0.95, 0.95)
    fields = [synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=2 * np.pi * (1 + 0.2 * k)) for k in range(nsteps)]
    _write_case(tmp_path / "case.bin", pm, fields, n, save, 0, deltaT, dt, 0.0, lo, hi)
    exe = _build_driver(tmp_path, True)
    outs = {}
    for tag, extra in (("rk2", "integrator=rk2;fuseSubSteps=3"), ("rk2_relocate", "integrator=rk2;fuseSubSteps=3;relocateLost=1")):
        env = dict(os.environ, CPF_SHIM_DICT=extra)
        r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / f"{tag}.bin"), str(nsteps)], cwd=tmp_path, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs[tag] = _read_out(tmp_path / f"{tag}.bin", n)
    pp, vv, tt, step, cc = outs["rk2"]
    ncyc = int(max(np.ceil(deltaT / dt), 1))
    p = synth.seed_box(n, lo, hi)
    cl = orc.Cloud.make(p, orc.locate_brute(mesh, p))
    for k in range(nsteps):
        orc.ext_substeps(mesh, cl, orc.expand_velocity(mesh, fields[k]), ncyc, deltaT / ncyc, integrator=1)
    assert np.array_equal(tt, cl.tet) and np.array_equal(pp.view(np.uint64), cl.p.view(np.uint64))
    p2, v2, t2, _, _ = outs["rk2_relocate"]
    assert np.array_equal(t2, tt) and np.array_equal(p2.view(np.uint64), pp.view(np.uint64))
