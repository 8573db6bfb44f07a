"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical
seeded inputs.  Bit-exact for tet ids and (fp64) positions/velocities -- the stated tolerance of the
north star (1e-6 relative) is met with margin 0."""
import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu


def _tracker(**kw):
    from cudaparticlesfoam_b200 import api

    kw.setdefault("rng", api.RNG_NONE)
    return api.ParticleTracker(**kw)


def _assert_same_state(p, v, t, cl, what=""):
    bad = np.flatnonzero(t != cl.tet)
    assert bad.size == 0, f"{what}: {bad.size} tet ids differ, first {bad[:5]}: {t[bad[:5]]} vs {cl.tet[bad[:5]]}"
    assert np.array_equal(p.view(np.uint64), cl.p.view(np.uint64)), f"{what}: positions not bit-identical"
    if v is not None:
        live = cl.p[:, 3] != 0
        assert np.array_equal(v[live, :3].view(np.uint64), cl.vel[live, :3].view(np.uint64)), f"{what}: velocities differ"


def test_mesh_builder_matches_oracle(synth, orc):
    """A1+A2: tets and neighbour information equal the oracle's decomposition and face tables."""
    pm, mesh, U, p = make_case(synth, orc, dims=(9, 7, 5), jitter=0.25)
    tr = _tracker()
    tr.upload_poly(pm)
    tv, tc = tr.download_tets()
    assert np.array_equal(tv, mesh.idx) and np.array_equal(tc, mesh.tet_cell)
    nb = tr.download_neighbours()
    # oracle neighbour across reference face slot k
    f = mesh.tetfacets
    front, back = mesh.finfo[f, 0], mesh.finfo[f, 1]
    me = np.arange(mesh.n_tets)[:, None]
    other = np.where(back == me, front, back)
    assert np.array_equal(nb >= 0, other >= 0)
    assert np.array_equal(nb[nb >= 0], other[other >= 0])
    info = tr.mesh_info()
    assert info["n_boundary_faces"] == mesh.n_boundary and info["n_tets"] == mesh.n_tets
    # boundary links carry the patch of the polyMesh face
    npatch = len(pm.patch_starts) - 1
    assert set(np.unique(-nb[nb < 0] - 1)) == set(range(npatch))
    tr.close()


def test_invalid_meshes_are_rejected(synth, orc):
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(3, 3, 3), jitter=0.0)
    tr = _tracker()
    bad = mesh.idx.copy()
    bad[5, [1, 2]] = bad[5, [2, 1]]  # inverted tet
    with pytest.raises(api.CpfError) as e:
        tr.upload_tets(mesh.pos, bad, mesh.tet_cell, pm.n_cells)
    assert e.value.code == 3 and "volume" in str(e.value)
    bad = mesh.idx.copy()
    bad[7, 2] = mesh.pos.shape[0] + 5  # vertex id out of range: must be refused, not dereferenced
    with pytest.raises(api.CpfError) as e:
        tr.upload_tets(mesh.pos, bad, mesh.tet_cell, pm.n_cells)
    assert e.value.code == 3 and "out of range" in str(e.value)
    bad_cell = mesh.tet_cell.copy()
    bad_cell[3] = pm.n_cells
    with pytest.raises(api.CpfError) as e:
        tr.upload_tets(mesh.pos, mesh.idx, bad_cell, pm.n_cells)
    assert e.value.code == 1
    with pytest.raises(api.CpfError):
        tr.mesh_info()  # no half-built mesh is left behind by a failed upload
    tr.upload_tets(mesh.pos, mesh.idx, mesh.tet_cell, pm.n_cells)  # the handle is still usable
    tr.close()


def test_initial_location_matches_brute_force(synth, orc):
    """A4: BVH location == lowest containing tet id (brute force), outside -> -1."""
    pm, mesh, U, p = make_case(synth, orc, dims=(10, 9, 8), jitter=0.2, n=20000, margin=-0.05)
    # put some particles exactly on mesh vertices / faces / cell centres (documented tie class:
    # both sides return the lowest id, so even these must agree)
    p[:200, :3] = mesh.pos[:200]
    p[200:400, :3] = 0.5 * (mesh.pos[mesh.idx[:200, 1]] + mesh.pos[mesh.idx[:200, 2]])
    tr = _tracker()
    tr.upload_poly(pm)
    tr.set_particles(p)
    tr.locate_initial()
    _, _, t = tr.download(pos=False, vel=False)
    ref = orc.locate_brute(mesh, p)
    assert (ref < 0).sum() > 100
    assert np.array_equal(t, ref)
    tr.close()


@pytest.mark.parametrize("path", [1, 0], ids=["exact", "filtered"])
@pytest.mark.parametrize("jitter", [0.0, 0.2])
def test_convex_substeps_bit_exact(synth, orc, path, jitter):
    """A5+A7+A8+A9 fused (default ConvexPoly build, no random walk): 60 sub-steps with reflections."""
    pm, mesh, U, p = make_case(synth, orc, dims=(12, 10, 8), jitter=jitter, n=30000)
    Utet = orc.expand_velocity(mesh, U)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    tr = _tracker(path=path)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    dt = 0.02
    for chunk in (1, 7, 52):
        orc.substeps(mesh, cl, Utet, chunk, dt)
        tr.substeps(chunk, dt)
        pp, vv, tt = tr.download()
        _assert_same_state(pp, vv, tt, cl, f"after chunk {chunk}")
    st = tr.stats()
    assert st["n_reflections"] > 0, "the case must exercise wall reflection"
    assert st["n_active"] == int((cl.p[:, 3] != 0).sum())
    tr.close()


def test_filtered_path_rarely_needs_exact_arithmetic(synth, orc):
    """Closed swirling flow (few wall contacts): the guard-band filter must let almost every
    particle-sub-step through the cheap path, and still match the oracle bit for bit."""
    pm, mesh, U, p = make_case(synth, orc, dims=(12, 12, 6), jitter=0.2, n=40000, field="swirl", margin=0.1)
    Utet = orc.expand_velocity(mesh, U)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    tr = _tracker(path=0)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    orc.substeps(mesh, cl, Utet, 30, 0.02)
    tr.substeps(30, 0.02)
    pp, vv, tt = tr.download()
    _assert_same_state(pp, vv, tt, cl, "swirl")
    st = tr.stats()
    assert st["n_hops"] > 1.5 * st["n_substeps"], "the case must cross tets"
    assert st["n_exact"] < 0.03 * st["n_substeps"], (st["n_exact"], st["n_substeps"])
    tr.close()


@pytest.mark.parametrize("fuse", [1, 6], ids=["unfused", "fused6"])
def test_wall_contacts_after_several_hops_stay_in_the_fast_kernels(synth, orc, fuse):
    """Uniform flow into a corner with steps of ~2.5 cells: particles cross many tets before they reach a wall,
    bounce in nearly every sub-step and meet two or three walls at once near edges and the corner.  The wall-capable
    queue pass replays the crossed faces in the reference's arithmetic (cpf_geom.cuh wall_reflect_on_path); the
    second wall of a sub-step, and a reflection on the sub-step whose velocity is reported, go through the exact
    kernel.  Everything must equal the oracle bit for bit -- positions, tets and the reflected velocities."""
    pm, mesh, U, p = make_case(synth, orc, dims=(10, 9, 8), jitter=0.15, n=30000, field=(1.0, 0.8, 0.6))
    Utet = orc.expand_velocity(mesh, U)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    tr = _tracker(path=0, fuse_substeps=fuse, sort_interval=7)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    dt = 0.25
    for chunk in (1, 6, 13):
        orc.substeps(mesh, cl, Utet, chunk, dt)
        tr.substeps(chunk, dt)
        pp, vv, tt = tr.download()
        _assert_same_state(pp, vv, tt, cl, f"corner flow after chunk {chunk}")
    st = tr.stats()
    assert st["n_reflections"] > 2 * p.shape[0], "the case must be dominated by wall contacts"
    assert st["n_hops"] > 3 * st["n_substeps"], "and by long walks"
    tr.close()


def test_filtered_path_handles_degenerate_starts(synth, orc):
    """Particles sitting exactly on vertices, edges, faces and cell centres must take the exact
    path and still reproduce the oracle bit for bit."""
    pm, mesh, U, p = make_case(synth, orc, dims=(6, 6, 6), jitter=0.0, n=6000, field=(0.3, 0.2, 0.1))
    k = 1500
    p[:k, :3] = mesh.pos[np.arange(k) % mesh.pos.shape[0]]                      # vertices (incl. centres)
    a = mesh.pos[mesh.idx[:k, 1]]; b = mesh.pos[mesh.idx[:k, 2]]; c = mesh.pos[mesh.idx[:k, 3]]
    p[k:2 * k, :3] = 0.5 * (a + b)                                              # edge midpoints
    p[2 * k:3 * k, :3] = (a + b + c) / 3.0                                      # face centroids
    tet0 = orc.locate_brute(mesh, p)
    Utet = orc.expand_velocity(mesh, U)
    cl = orc.Cloud.make(p, tet0)
    tr = _tracker(path=0)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.set_tets(tet0)
    orc.substeps(mesh, cl, Utet, 25, 0.05)
    tr.substeps(25, 0.05)
    pp, vv, tt = tr.download()
    _assert_same_state(pp, vv, tt, cl, "degenerate starts")
    tr.close()


def test_fused_and_sorted_runs_equal_plain_run(synth, orc):
    """Fusing sub-steps into one launch and re-sorting particles by cell must not change a bit."""
    pm, mesh, U, p = make_case(synth, orc, dims=(10, 10, 10), jitter=0.15, n=20000, field="channel")
    Utet = orc.expand_velocity(mesh, U)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    orc.substeps(mesh, cl, Utet, 40, 0.03)
    for kw in (dict(fuse_substeps=8), dict(sort_interval=5), dict(fuse_substeps=4, sort_interval=6)):
        tr = _tracker(**kw)
        tr.upload_poly(pm)
        tr.update_velocity(U)
        tr.set_particles(p)
        tr.locate_initial()
        tr.substeps(40, 0.03)
        pp, vv, tt = tr.download()
        _assert_same_state(pp, vv, tt, cl, str(kw))
        tr.close()


@pytest.mark.parametrize("path,rng", [(0, 0), (1, 0), (0, 2), (0, 1)], ids=["filtered", "exact", "filtered-philox", "filtered-xorwow"])
def test_bary_mode_bit_exact(synth, orc, path, rng):
    """A7'+A8' (RTX=true build): barycentric walk + RTreflection.  filtered: the fp32 walk towards the end point
    (visit_bary32) with the reference arithmetic for what it refuses -- walls, unclear minima, points placed exactly on
    vertices, edges and faces; exact: k_exact<BARY> only.  Both against the oracle, fused chunks, with and without random walk."""
    pm, mesh, U, p = make_case(synth, orc, dims=(10, 8, 6), jitter=0.2, n=30000)
    Utet = orc.expand_velocity(mesh, U)
    # a share of the particles sits exactly on mesh features: the walk must refuse or decide like the reference
    rs = np.random.default_rng(3)
    t = rs.integers(0, mesh.idx.shape[0], size=3000)
    a, b, c = mesh.pos[mesh.idx[t, 0]], mesh.pos[mesh.idx[t, 1]], mesh.pos[mesh.idx[t, 2]]
    p[:1000, :3] = a[:1000]
    p[1000:2000, :3] = 0.5 * (a[1000:2000] + b[1000:2000])
    p[2000:3000, :3] = (a[2000:] + b[2000:] + c[2000:]) / 3.0
    # ... features of the INTERIOR: on the domain boundary the start tet is undefined and the reference's RTX kernels read
    # out of bounds for a negative id
    onb = ((np.abs(p[:, :3] - pm.lo) < 1e-12) | (np.abs(p[:, :3] - pm.hi) < 1e-12)).any(axis=1)
    p[onb, :3] = 0.5 * (pm.lo + pm.hi) + 0.3 * (rs.random((int(onb.sum()), 3)) - 0.5)
    tet0 = orc.locate_brute(mesh, p)
    lostp = tet0 < 0
    p[lostp, :3] = 0.5 * (pm.lo + pm.hi) + 0.3 * (rs.random((int(lostp.sum()), 3)) - 0.5)
    tet0 = orc.locate_brute(mesh, p)
    assert (tet0 >= 0).all()
    cl = orc.Cloud.make(p, tet0)
    from cudaparticlesfoam_b200 import api

    D = 2e-3 if rng else 0.0
    tr = _tracker(locator=api.LOCATOR_BARY, path=path, rng=rng, diffusion_coeff=D, fuse_substeps=7, sort_interval=9)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    if rng == api.RNG_XORWOW:
        tr.init_rng()
    tr.set_tets(tet0)
    for chunk in (1, 20, 29):
        xi = tr.normals(chunk) if rng else None
        orc.substeps(mesh, cl, Utet, chunk, 0.02, convex=False, xi=xi, D=D)
        tr.substeps(chunk, 0.02)
        pp, vv, tt = tr.download()
        _assert_same_state(pp, vv, tt, cl, f"bary mode, chunk {chunk}")
    st = tr.stats()
    assert st["n_reflections"] > 0
    if path == 0:
        # a deferred particle finishes its chunk in the exact kernel, and this small box with a strong vortex sends many
        # particles into walls (always deferred in this build); on the bench workload 0.06 % of the sub-steps are exact
        assert 0 < st["n_exact"] < 0.7 * st["n_substeps"], "the filtered barycentric walk must carry its share of the sub-steps"
    tr.close()


@pytest.mark.parametrize("rng", [1, 2], ids=["xorwow", "philox"])
def test_random_walk_matches_oracle_given_the_deviates(synth, orc, rng):
    """A6: disp += sqrt(2 D dt) * xi.  The deviates are read back from the library and fed to the
    oracle, so the comparison stays bit-exact for either generator."""
    pm, mesh, U, p = make_case(synth, orc, dims=(8, 8, 8), jitter=0.2, n=10000, field="channel")
    Utet = orc.expand_velocity(mesh, U)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    tr = _tracker(rng=rng, diffusion_coeff=2e-3, sort_interval=4)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    dt = 0.02
    xs = []
    for s in range(12):
        xi = tr.next_normals()
        xs.append(xi)
        orc.substeps(mesh, cl, Utet, 1, dt, xi=xi[None], D=2e-3)
        tr.substeps(1, dt)
    pp, vv, tt = tr.download()
    _assert_same_state(pp, vv, tt, cl, "random walk")
    x = np.concatenate(xs).ravel()
    assert abs(x.mean()) < 0.02 and abs(x.std() - 1.0) < 0.02, "deviates must be standard normal"
    tr.close()


def test_velocity_refresh_and_advect_cycle_count(synth, orc):
    """A10 + src/advect.H:36-37: nCycles = max(ceil(deltaT/dt),1), cycleDt = deltaT/nCycles, new
    cell field picked up by the next call."""
    pm, mesh, U, p = make_case(synth, orc, dims=(8, 8, 8), jitter=0.1, n=8000)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    tr = _tracker(dt=3e-3)
    tr.init_cuda(pm, U, particles=p)
    t = 0.0
    for step in range(4):
        Ut = synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=2 * np.pi * (1 + 0.3 * step))
        n = tr.advect(Ut, 0.01)
        assert n == 4
        orc.substeps(mesh, cl, orc.expand_velocity(mesh, Ut), 4, 0.01 / 4)
    pp, vv, tt = tr.download()
    _assert_same_state(pp, vv, tt, cl, "coupled refresh")
    tr.close()


def test_field_upload_one_step_ahead_of_running_substeps(synth, orc):
    """cpf_update_velocity with a HOST field uploads on the copy stream into the idle half of the double buffer:
    issuing the refresh for step k+1 while the sub-steps of step k are still running (no synchronisation in
    between, page-locked source) must neither disturb step k nor leak the old field into step k+1."""
    import torch

    pm, mesh, U, p = make_case(synth, orc, dims=(10, 10, 10), jitter=0.1, n=200000)
    fields = [synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=2 * np.pi * (1 + 0.5 * k)) for k in range(5)]
    pinned = [torch.from_numpy(np.ascontiguousarray(f)).pin_memory() for f in fields]
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    tr = _tracker(sort_interval=0)
    tr.upload_poly(pm)
    tr.set_particles(p)
    tr.locate_initial()
    tr.update_velocity_ptr(pinned[0].data_ptr(), False)
    for k in range(4):
        tr.substeps(12, 2e-3)                                     # asynchronous: still running ...
        tr.update_velocity_ptr(pinned[k + 1].data_ptr(), False)   # ... while the next field is uploaded
        orc.substeps(mesh, cl, orc.expand_velocity(mesh, fields[k]), 12, 2e-3)
    pp, vv, tt = tr.download()
    _assert_same_state(pp, vv, tt, cl, "overlapped refresh")
    tr.close()


@pytest.mark.parametrize("mode", ["filtered", "exact", "bary"])
def test_reflect_wall_off_freezes_particles_like_the_reference(synth, orc, mode):
    """reflectWall = false (src/initCuda.H:67): a particle whose walk meets a wall keeps the id -(tet+1), is still moved
    by S5 and is frozen by S1 of the next sub-step (particles.cu:334-338).  No reflection code runs in any kernel."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(8, 8, 6), jitter=0.15, n=20000, field=(0.9, 0.5, -0.4))
    Utet = orc.expand_velocity(mesh, U)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    convex = mode != "bary"
    tr = _tracker(reflect_wall=0, path=api.PATH_EXACT if mode == "exact" else api.PATH_FILTERED,
                  locator=api.LOCATOR_CONVEX if convex else api.LOCATOR_BARY, fuse_substeps=5, sort_interval=6)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.set_tets(tet0)
    for chunk in (1, 5, 14):
        orc.substeps(mesh, cl, Utet, chunk, 0.05, convex=convex, reflect=False)
        tr.substeps(chunk, 0.05)
        pp, vv, tt = tr.download()
        _assert_same_state(pp, vv, tt, cl, f"reflectWall off, {mode}, chunk {chunk}")
    st = tr.stats()
    assert st["n_reflections"] == 0
    assert 0 < st["n_active"] < p.shape[0], "some, not all, particles must have left through a wall"
    assert st["n_active"] == int((cl.p[:, 3] != 0).sum())
    tr.close()


def test_empty_and_inactive_inputs(synth, orc):
    pm, mesh, U, p = make_case(synth, orc, dims=(4, 4, 4), jitter=0.0, n=64)
    tr = _tracker()
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(np.zeros((0, 4)))
    tr.locate_initial()
    tr.substeps(3, 0.01)
    assert tr.stats()["n_particles"] == 0
    # outside / inactive particles are frozen exactly like the reference does (w := 0, never moved)
    p[:10, 0] += 5.0
    p[10:20, 3] = 0.0
    tr.set_particles(p)
    tr.locate_initial()
    tr.substeps(5, 0.01)
    pp, _, tt = tr.download()
    assert np.all(tt[:10] == -1) and np.all(pp[:10, 3] == 0) and np.array_equal(pp[:20, :3], p[:20, :3])
    assert np.all(pp[20:, 3] == 1)
    tr.close()
