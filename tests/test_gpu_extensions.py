"""Features beyond the reference (RK2/RK4, vertex/cellPoint-style interpolation, outlet escape):
the CUDA path against the oracle extension (oracle/cpf_oracle_ext.c) -- bit-exact.  Parity against
the reference itself is unpinned for these (it does not implement them)."""
import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.mark.parametrize("integrator", [0, 1, 4], ids=["euler", "rk2", "rk4"])
@pytest.mark.parametrize("vertex", [False, True], ids=["cell", "vertex"])
def test_integrators_and_interpolation(synth, orc, integrator, vertex):
    from cudaparticlesfoam_b200 import api

    if integrator == 0 and not vertex:
        pytest.skip("covered by test_gpu_parity")
    pm, mesh, U, p = make_case(synth, orc, dims=(9, 8, 7), jitter=0.2, n=12000, field="swirl", margin=0.02)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    Uo = orc.point_values(pm, U) if vertex else orc.expand_velocity(mesh, U)
    orc.ext_substeps(mesh, cl, Uo, 25, 0.02, vertex_velocity=vertex, integrator=integrator)
    tr = api.ParticleTracker(rng=api.RNG_NONE, integrator=integrator, interp=api.INTERP_VERTEX if vertex else api.INTERP_TET,
                             sort_interval=7, fuse_substeps=5)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    tr.substeps(25, 0.02)
    pp, vv, tt = tr.download()
    assert np.array_equal(tt, cl.tet)
    assert _same(pp, cl.p) and _same(vv[:, :3], cl.vel[:, :3])
    tr.close()


def test_point_interpolation_matches_oracle(synth, orc):
    """volPointInterpolation-style point values computed on the device == oracle; linear fields are
    reproduced exactly in the interior."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(6, 6, 6), jitter=0.0, n=2000, field=(0.0, 0.0, 0.0))
    U = np.ascontiguousarray(1.0 + 2.0 * pm.cell_centres[:, [1, 2, 0]])  # linear field
    Uv = orc.point_values(pm, U)
    interior = np.all((pm.points > 1e-9) & (pm.points < 1 - 1e-9), axis=1)
    assert np.abs(Uv[: pm.n_points][interior] - (1.0 + 2.0 * pm.points[interior][:, [1, 2, 0]])).max() < 1e-12
    # through the particle path: with a linear field, vertex interpolation + Euler gives v(P) exactly
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    orc.ext_substeps(mesh, cl, Uv, 1, 1e-3, vertex_velocity=True)
    tr = api.ParticleTracker(rng=api.RNG_NONE, interp=api.INTERP_VERTEX)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    tr.substeps(1, 1e-3)
    pp, vv, tt = tr.download()
    assert _same(pp, cl.p) and _same(vv[:, :3], cl.vel[:, :3])
    inner = np.all((p[:, :3] > 0.2) & (p[:, :3] < 0.8), axis=1)
    assert np.abs(vv[inner, :3] - (1.0 + 2.0 * p[inner][:, [1, 2, 0]])).max() < 1e-12
    tr.close()


@pytest.mark.parametrize("rng", [0, 2], ids=["none", "philox"])
def test_outlet_escape(synth, orc, rng):
    """x+ is an ESCAPE patch, everything else reflects: escaped particles are parked on the outlet
    plane, deactivated, counted; the rest matches the oracle bit for bit."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(10, 6, 6), jitter=0.15, n=15000, field=(1.0, 0.15, -0.1))
    kinds = [api.PATCH_REFLECT] * 6
    kinds[1] = api.PATCH_ESCAPE  # synth.PATCH_NAMES: x-, x+, y-, y+, z-, z+
    fk = orc.face_kinds(pm, mesh, kinds)
    tet0 = orc.locate_brute(mesh, p)
    cl = orc.Cloud.make(p, tet0)
    tr = api.ParticleTracker(rng=rng, diffusion_coeff=1e-3, sort_interval=6, fuse_substeps=4)
    tr.upload_poly(pm, patch_kind=kinds)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    Utet = orc.expand_velocity(mesh, U)
    n_esc = 0
    for s in range(30):
        xi = tr.next_normals() if rng else None
        n_esc += orc.ext_substeps(mesh, cl, Utet, 1, 0.02, face_kind=fk, xi=None if xi is None else xi[None], D=1e-3 if rng else 0.0)
        tr.substeps(1, 0.02)
    pp, vv, tt = tr.download()
    st = tr.stats()
    assert n_esc > 1000 and st["n_escaped"] == n_esc
    assert np.array_equal(tt, cl.tet) and _same(pp, cl.p)
    gone = pp[:, 3] == 0
    assert gone.sum() == n_esc and np.abs(pp[gone, 0] - 1.0).max() < 1e-12 and (tt[gone] < 0).all()
    assert st["n_active"] == (~gone).sum()
    tr.close()


def test_lost_particle_relocation(synth, orc):
    """Lost-particle fallback: active particles with a negative tet id are re-located by the BVH
    (lowest containing tet, as brute force), instead of being frozen on the next sub-step."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(8, 8, 8), jitter=0.2, n=5000)
    tet_true = orc.locate_brute(mesh, p)
    lost = np.arange(0, 5000, 7)
    tet = tet_true.copy()
    tet[lost] = -1                      # pretend these were lost
    p2 = p.copy()
    p2[:3, 0] += 4.0                    # really outside: must stay lost
    tet[:3] = -1
    tr = api.ParticleTracker(rng=api.RNG_NONE)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p2)
    tr.set_tets(tet)
    tr.relocate_lost()
    _, _, tt = tr.download(pos=False, vel=False)
    want = tet_true.copy()
    want[:3] = -1
    assert np.array_equal(tt, want)
    tr.close()


def test_lost_only_pass_equals_full_location_at_scale(synth, orc):
    """The lost-only pass (compacted, Morton-sorted work list; one resident wave of CTAs in a grid-stride loop) must give
    the answer of the all-particles pass for exactly the particles it is asked about, and touch nobody else: 3e5 random
    points -- more than one resident wave of threads, so the grid-stride loop runs -- a third of them marked lost,
    inactive particles and out-of-domain points mixed in; a sample is checked against brute force."""
    from cudaparticlesfoam_b200 import api

    n = 300_000
    pm, mesh, U, p = make_case(synth, orc, dims=(12, 10, 9), jitter=0.2, n=n)
    rng = np.random.default_rng(7)
    p[rng.choice(n, 1000, replace=False), 3] = 0.0   # inactive slots
    outside = rng.choice(n, 50, replace=False)
    p[outside, 0] += 50.0                             # outside the mesh (some of them inactive as well)
    tr = api.ParticleTracker(rng=api.RNG_NONE)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    _, _, t_full = tr.download(pos=False, vel=False)
    active = p[:, 3] != 0
    assert (t_full[~active] == -1).all() and (t_full[outside] == -1).all()
    sample = rng.choice(np.flatnonzero(active), 400, replace=False)
    assert np.array_equal(t_full[sample], orc.locate_brute(mesh, p[sample]))
    lost = rng.random(n) < 1.0 / 3.0
    marker_tet = np.where(lost, -1, t_full).astype(np.int32)
    marker_tet[~active] = -7                           # an inactive slot keeps whatever id it carries
    tr.set_tets(marker_tet)
    tr.relocate_lost()
    _, _, t_again = tr.download(pos=False, vel=False)
    want = t_full.copy()
    want[~active] = -7
    assert np.array_equal(t_again, want)
    tr.relocate_lost()                                 # nothing left to do except the out-of-domain points: a no-op
    _, _, t_third = tr.download(pos=False, vel=False)
    assert np.array_equal(t_third, want)
    tr.close()


@pytest.mark.parametrize("path", [0, 1], ids=["filtered", "exact"])
def test_per_patch_restitution_matches_the_oracle(synth, orc, path):
    """Rebound model: a restitution coefficient per boundary patch scales the mirrored part of the end point
    and of the velocity.  Product (in-place wall pass + exact finisher, or exact only) against the oracle extension, bit for
    bit, with different coefficients on the six walls of a box and a flow that drives particles into edges and corners;
    e = 1 on every patch must leave the reference's results untouched."""
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(8, 7, 6), jitter=0.15, n=30000, field=(0.9, 0.7, -0.5))
    tet0 = orc.locate_brute(mesh, p)
    Utet = orc.expand_velocity(mesh, U)
    e = np.array([1.0, 0.5, 0.8, 0.25, 1.0, 0.6])
    out = {}
    for name, coeff in (("specular", np.ones(6)), ("rebound", e)):
        tr = api.ParticleTracker(rng=api.RNG_NONE, fuse_substeps=5, sort_interval=10, path=path)
        tr.upload_poly(pm)
        tr.set_patch_restitution(coeff)
        tr.update_velocity(U)
        tr.set_particles(p)
        tr.set_tets(tet0)
        cl = orc.Cloud.make(p, tet0)
        gains = orc.face_gains(pm, mesh, coeff)
        for chunk in (7, 20, 13):
            orc.ext_substeps(mesh, cl, Utet, chunk, 0.02, face_gain=gains)
            tr.substeps(chunk, 0.02)
            pp, vv, tt = tr.download()
            assert np.array_equal(tt, cl.tet), (name, chunk, int((tt != cl.tet).sum()))
            assert _same(pp, cl.p) and _same(vv[:, :3], cl.vel[:, :3]), (name, chunk)
        st = tr.stats()
        assert st["n_reflections"] > 20000
        out[name] = (pp.copy(), cl)
        tr.close()
    # e = 1 is the reference: the plain oracle (no extension) gives the same bits
    ref = orc.Cloud.make(p, tet0)
    orc.substeps(mesh, ref, Utet, 40, 0.02)
    assert _same(out["specular"][0], ref.p)
    assert not _same(out["rebound"][0], ref.p)
    with pytest.raises(api.CpfError):
        tr2 = api.ParticleTracker(rng=api.RNG_NONE)
        tr2.upload_poly(pm)
        tr2.set_patch_restitution(np.array([1.0, 0.0, 1.0, 1.0, 1.0, 1.0]))   # e = 0 would park particles ON the wall plane
