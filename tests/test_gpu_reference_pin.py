"""Pins the oracle (and the product) to the UNMODIFIED reference kernels, compiled from
/root/reference into oracle/_ref/libref_rtxadvect.so and run here on the same B200.
Everything is compared bit for bit."""
import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(synth, orc):
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libref_rtxadvect.so not built (needs /root/reference at build time)")
    pm, mesh, U, p = make_case(synth, orc, dims=(10, 8, 6), jitter=0.2, n=20000)
    Utet = orc.expand_velocity(mesh, U)
    tet0 = orc.locate_brute(mesh, p)
    return pm, mesh, U, Utet, p, tet0


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def test_each_reference_kernel_convex(case, orc):
    pm, mesh, U, Utet, p, tet0 = case
    rr = orc.RefRun(mesh, Utet, p, tet0)
    cl = orc.Cloud.make(p, tet0)
    dt = 0.03
    R = rr.R
    import ctypes as C
    for step in range(30):
        R.ref_advect(rr.mh, rr.ph, C.c_double(dt), C.c_int(0)); orc.advect(mesh, cl, Utet, dt)
        g = rr.download()
        assert _same(g.disp[:, :3], cl.disp[:, :3]) and _same(g.vel[:, :3], cl.vel[:, :3]), f"advect step {step}"
        R.ref_locate_convex(rr.mh, rr.ph); orc.locate_convex(mesh, cl)
        g = rr.download()
        assert np.array_equal(g.tet, cl.tet), f"locator step {step}"
        nwall = int((cl.tet < 0).sum())
        R.ref_reflect_convex(rr.mh, rr.ph); orc.reflect_convex(mesh, cl)
        g = rr.download()
        assert np.array_equal(g.tet, cl.tet), f"reflector ids step {step}"
        assert _same(g.p, cl.p) and _same(g.disp[:, :3], cl.disp[:, :3]) and _same(g.vel[:, :3], cl.vel[:, :3]), f"reflector step {step} ({nwall} wall hits)"
        R.ref_move(rr.ph); orc.move(cl)
        g = rr.download()
        assert _same(g.p, cl.p), f"move step {step}"
    rr.close()


def test_reference_loop_convex_and_bary(case, orc):
    pm, mesh, U, Utet, p, tet0 = case
    for convex in (True, False):
        rr = orc.RefRun(mesh, Utet, p, tet0)
        cl = orc.Cloud.make(p, tet0)
        rr.substeps(80, 0.025, convex=convex)
        orc.substeps(mesh, cl, Utet, 80, 0.025, convex=convex)
        g = rr.download()
        assert np.array_equal(g.tet, cl.tet), f"convex={convex}"
        assert _same(g.p, cl.p) and _same(g.vel[:, :3], cl.vel[:, :3]), f"convex={convex}"
        rr.close()


def test_reference_vertex_velocity_kernel(case, orc, synth):
    """cuda/particles.cu:244-313 (VertexVelocity mode, unreachable from the glue but in the library)."""
    pm, mesh, U, Utet, p, tet0 = case
    Uv = np.zeros((mesh.n_tets, 3))
    Uv[: mesh.pos.shape[0]] = synth.field_uniform_vortex(mesh.pos, R=0.3)
    rr = orc.RefRun(mesh, Uv, p, tet0)
    cl = orc.Cloud.make(p, tet0)
    rr.substeps(40, 0.02, vertex_velocity=True)
    orc.substeps(mesh, cl, Uv, 40, 0.02, vertex_velocity=True)
    g = rr.download()
    assert np.array_equal(g.tet, cl.tet) and _same(g.p, cl.p) and _same(g.vel[:, :3], cl.vel[:, :3])
    rr.close()


def test_reference_brownian_stream_and_product_xorwow(case, orc):
    """The product's XORWOW mode must replay the reference's cuRAND stream bit for bit
    (seed 1591593751, subsequence = particle id), through the whole fused sub-step."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, Utet, p, tet0 = case
    D, dt, nsteps = 1e-3, 0.02, 16
    draws = orc.RefRun(mesh, Utet, p, tet0, init_rng=True)
    xi = np.stack([draws.draw_normals() for _ in range(nsteps)])
    draws.close()
    rr = orc.RefRun(mesh, Utet, p, tet0, init_rng=True)
    rr.substeps(nsteps, dt, brownian=True, D=D)
    g = rr.download()
    rr.close()
    cl = orc.Cloud.make(p, tet0)
    orc.substeps(mesh, cl, Utet, nsteps, dt, xi=xi, D=D)
    assert np.array_equal(g.tet, cl.tet) and _same(g.p, cl.p), "oracle vs reference with random walk"
    # dict(): the drop-in's default configuration (library default fuse, sort every 50); all on the filtered 3-pass pipeline
    for kw in (dict(), dict(fuse_substeps=4, sort_interval=5), dict(fuse_substeps=10, sort_interval=7), dict(fuse_substeps=14, sort_interval=0), dict(path=api.PATH_EXACT)):
        tr = api.ParticleTracker(rng=api.RNG_XORWOW, diffusion_coeff=D, **kw)
        tr.upload_poly(pm)
        tr.update_velocity(U)
        tr.set_particles(p)
        tr.init_rng()
        tr.locate_initial()
        tr.substeps(nsteps, dt)
        pp, vv, tt = tr.download()
        assert np.array_equal(tt, g.tet), kw
        assert _same(pp, g.p) and _same(vv[:, :3], g.vel[:, :3]), kw
        tr.close()


def test_product_equals_reference_on_the_default_path(case, orc):
    """End to end: the reference's 5-kernel loop vs one fused launch, ids and positions identical."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, Utet, p, tet0 = case
    rr = orc.RefRun(mesh, Utet, p, tet0)
    rr.bary_query()  # the reference's narrow phase on the seeded ids
    g0 = rr.download()
    assert np.array_equal(g0.tet, tet0)
    rr.substeps(100, 0.02)
    g = rr.download()
    rr.close()
    tr = api.ParticleTracker(rng=api.RNG_NONE, fuse_substeps=10, sort_interval=20)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    tr.substeps(100, 0.02)
    pp, vv, tt = tr.download()
    assert np.array_equal(tt, g.tet) and _same(pp, g.p) and _same(vv[:, :3], g.vel[:, :3])
    tr.close()
