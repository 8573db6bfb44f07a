"""Parity at the shapes BASELINE.json names (configs[0..4]).  Tracers do not interact, so at full
size the oracle follows a SAMPLE of the very same cloud (bit-exact for those particles) and the rest
is covered by size-independent properties: filtered == exact-only policy (checksum over all
particles), partition invariance (two half-clouds == one cloud), containment, conservation of
particle count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def _structured_tets(pm, p):
    nx, ny, nz = pm.dims
    h = (pm.hi - pm.lo) / np.array([nx, ny, nz])
    ijk = np.clip(((p[:, :3] - pm.lo) / h).astype(np.int64), 0, np.array([nx, ny, nz]) - 1)
    return ijk[:, 0] + nx * (ijk[:, 1] + ny * ijk[:, 2])


def test_config0_pitzdaily_standin_1e5_euler(synth, orc):
    """configs[0]: 12 225 hex cells one cell thick, 1e5 tracers, Euler, 100 sub-steps -- full size vs oracle."""
    from cudaparticlesfoam_b200 import api

    pm = synth.backward_step_mesh()
    assert pm.n_cells == 12225
    mesh = orc.tet_mesh_from_poly(pm)
    U = np.zeros((pm.n_cells, 3))
    y = (pm.cell_centres[:, 1] - pm.lo[1]) / (pm.hi[1] - pm.lo[1])
    U[:, 0] = 10.0 * 4 * y * (1 - y)
    U[:, 1] = 0.5 * np.sin(40 * pm.cell_centres[:, 0])
    span = pm.hi - pm.lo
    p = synth.seed_box(100_000, pm.lo + 0.01 * span, pm.lo + np.array([0.3, 0.99, 0.99]) * span)
    tr = api.ParticleTracker(rng=api.RNG_NONE, dt=1e-4, sort_interval=20, fuse_substeps=10)
    tr.init_cuda(pm, U, particles=p)
    _, _, tet0 = tr.download(pos=False, vel=False)
    cl = orc.Cloud.make(p, tet0)
    assert (tet0 >= 0).all()
    # the initial tets are the containing ones (lowest id on ties): check against the oracle's own walk
    chk = orc.Cloud.make(p, (12 * _structured_tets(pm, p)).astype(np.int32))
    orc.bary_query(mesh, chk)
    assert (chk.tet != tet0).mean() < 1e-4
    orc.substeps(mesh, cl, orc.expand_velocity(mesh, U), 100, 1e-4)
    n = tr.advect(None, 0.01)  # src/advect.H: deltaT 0.01 / dt 1e-4 -> 100 cycles
    assert n == 100
    pp, vv, tt = tr.download()
    assert np.array_equal(tt, cl.tet) and _same(pp, cl.p) and _same(vv[:, :3], cl.vel[:, :3])
    assert tr.stats()["n_reflections"] > 0
    tr.close()


def test_config1_box100_1e6_rk2(synth, orc):
    """configs[1]: 100^3 hex (12e6 tets, past the reference's 2^20-vertex limit), 1e6 tracers, RK2, frozen
    uniform+vortex field.  Oracle on a 50k-particle sample of the same cloud; all particles stay contained."""
    from cudaparticlesfoam_b200 import api

    pm = synth.box_mesh(100, 100, 100, jitter=0.1)
    U = synth.field_uniform_vortex(pm.cell_centres, U0=(0.3, 0.0, 0.0), R=0.2)
    p = synth.seed_box(1_000_000, pm.lo + 0.02, pm.hi - 0.02)
    tr = api.ParticleTracker(rng=api.RNG_NONE, integrator=api.RK2, sort_interval=10, fuse_substeps=5)
    tr.init_cuda(pm, U, particles=p)
    info = tr.mesh_info()
    assert info["n_tets"] == 12_000_000 and info["n_verts"] == 2_030_301
    _, _, tet0 = tr.download(pos=False, vel=False)
    assert (tet0 >= 0).all()
    tr.substeps(10, 0.004)
    pp, vv, tt = tr.download()
    # sample: every 20th particle through the oracle extension
    mesh = orc.tet_mesh_from_poly(pm)
    idx = np.arange(0, p.shape[0], 20)
    cl = orc.Cloud.make(p[idx], tet0[idx])
    orc.ext_substeps(mesh, cl, orc.expand_velocity(mesh, U), 10, 0.004, integrator=1)
    assert np.array_equal(tt[idx], cl.tet) and _same(pp[idx], cl.p) and _same(vv[idx, :3], cl.vel[:, :3])
    # containment of ALL particles in the tet they report
    w = orc.bary_of(mesh, pp, tt)
    assert w.min() > -1e-9 and (pp[:, 3] == 1).all()
    tr.close()


def test_config2_channel_1e7_filtered_equals_exact_and_oracle_sample(synth, orc):
    """configs[2] (the bench workload): 1e6-cell channel, 1e7 tracers, field refreshed every step, Philox random
    walk.  The filtered pipeline and the exact-only policy must agree on every particle; a sample is replayed by the
    oracle with the deviates read back from the library."""
    from cudaparticlesfoam_b200 import api

    import bench

    w = bench.WORKLOADS["channel1M_1e7"]
    pm, p, fields, _ = bench.build_inputs(w, 0, 1)
    res = []
    xis = None
    for path in (api.PATH_FILTERED, api.PATH_EXACT):
        tr = api.ParticleTracker(rng=api.RNG_PHILOX, diffusion_coeff=w["D"], dt=w["dt"], sort_interval=20, fuse_substeps=10, path=path)
        tr.upload_poly(pm)
        tr.update_velocity(fields[0])
        tr.set_particles(p)
        tr.locate_initial()
        if path == api.PATH_FILTERED:
            _, _, tet0 = tr.download(pos=False, vel=False)
        for k in range(2):
            tr.advect(fields[k], w["dt"] * w["ncycles"])
        res.append(tr.download())
        st = tr.stats()
        assert st["n_active"] == p.shape[0] and st["n_substeps"] == 2 * w["ncycles"] * p.shape[0]
        tr.close()
    (pa, va, ta), (pb, vb, tb) = res
    assert np.array_equal(ta, tb) and _same(pa, pb) and _same(va[:, :3], vb[:, :3])
    # oracle on a sample, no random walk (deviates are an oracle input; Philox is covered at small size)
    tr = api.ParticleTracker(rng=api.RNG_NONE, dt=w["dt"], sort_interval=20, fuse_substeps=10)
    tr.upload_poly(pm)
    tr.update_velocity(fields[0])
    tr.set_particles(p)
    tr.set_tets(tet0)
    for k in range(2):
        tr.advect(fields[k], w["dt"] * w["ncycles"])
    pp, vv, tt = tr.download()
    tr.close()
    mesh = orc.tet_mesh_from_poly(pm)
    idx = np.arange(0, p.shape[0], 100)
    cl = orc.Cloud.make(p[idx], tet0[idx])
    for k in range(2):
        orc.substeps(mesh, cl, orc.expand_velocity(mesh, fields[k]), w["ncycles"], w["dt"])
    assert np.array_equal(tt[idx], cl.tet) and _same(pp[idx], cl.p) and _same(vv[idx, :3], cl.vel[:, :3])


def test_config3_polyhedral_rk4_partitioned(synth, orc):
    """configs[3] code path at reduced size: polyhedral cells (hexagonal faces, 20 tets per cell, non-trivial
    tetBasePtIs), RK4, particles partitioned by index over two contexts with the mesh replicated."""
    from cudaparticlesfoam_b200 import api, parallel

    pm = synth.polyhex_mesh(14, 12, 10, jitter=0.1)
    mesh = orc.tet_mesh_from_poly(pm)
    assert mesh.n_tets == 20 * pm.n_cells
    U = synth.field_uniform_vortex(pm.cell_centres, U0=(0.2, 0.0, 0.05), R=0.25)
    p = synth.seed_box(60_000, pm.lo + 0.02, pm.hi - 0.02)
    parts = []
    for r in range(2):
        s, c = parallel.partition(p.shape[0], 2, r)
        tr = api.ParticleTracker(rng=api.RNG_NONE, integrator=api.RK4, sort_interval=5, fuse_substeps=5)
        tr.init_cuda(pm, U, particles=p[s:s + c])
        tv, tc = tr.download_tets()
        assert np.array_equal(tv, mesh.idx) and np.array_equal(tc, mesh.tet_cell)
        tr.substeps(20, 0.01)
        parts.append(tr.download())
        tr.close()
    pp = np.concatenate([a[0] for a in parts]); vv = np.concatenate([a[1] for a in parts]); tt = np.concatenate([a[2] for a in parts])
    cl = orc.Cloud.make(p, orc.locate_brute(mesh, p))
    orc.ext_substeps(mesh, cl, orc.expand_velocity(mesh, U), 20, 0.01, integrator=4)
    assert np.array_equal(tt, cl.tet) and _same(pp, cl.p) and _same(vv[:, :3], cl.vel[:, :3])


def test_config4_wall_bounded_dispersion_with_escape(synth, orc):
    """configs[4] physics at reduced size: channel, random walk, reflecting walls, outlet escape, compaction by
    the sort; particle count is conserved (active + escaped) and the oracle agrees bit for bit."""
    from cudaparticlesfoam_b200 import api

    pm = synth.box_mesh(40, 10, 10, lo=(0, 0, 0), hi=(4.0, 1.0, 1.0), jitter=0.1)
    mesh = orc.tet_mesh_from_poly(pm)
    U = synth.field_channel(pm.cell_centres, lo=pm.lo, hi=pm.hi)
    p = synth.seed_box(200_000, (2.5, 0.02, 0.02), (3.98, 0.98, 0.98))
    kinds = [0, 1, 0, 0, 0, 0]
    fk = orc.face_kinds(pm, mesh, kinds)
    tr = api.ParticleTracker(rng=api.RNG_PHILOX, diffusion_coeff=2e-3, sort_interval=4, fuse_substeps=4)
    tr.upload_poly(pm, patch_kind=kinds)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    _, _, tet0 = tr.download(pos=False, vel=False)
    cl = orc.Cloud.make(p, tet0)
    Utet = orc.expand_velocity(mesh, U)
    esc = 0
    for s in range(24):
        xi = tr.next_normals()
        esc += orc.ext_substeps(mesh, cl, Utet, 1, 0.05, face_kind=fk, xi=xi[None], D=2e-3)
        tr.substeps(1, 0.05)
    pp, vv, tt = tr.download()
    st = tr.stats()
    assert esc > 10_000 and st["n_escaped"] == esc and st["n_active"] + esc == p.shape[0]
    assert np.array_equal(tt, cl.tet) and _same(pp, cl.p)
    tr.close()
