"""BASELINE config 4 code path on a genuinely polyhedral mesh (hexagonal prisms, zigzag side walls): device mesh builder,
BVH location and the filtered advection against the oracle, through the C ABI.

The CPU side of this case (decomposition, topology, oracle tracking, filter model) is covered by
tests/test_oracle_cpu.py and tests/test_filter_soundness_cpu.py; this file is the GPU half."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _assert_same(p, v, t, cl, what):
    assert np.array_equal(t, cl.tet), f"{what}: {(t != cl.tet).sum()} tet ids differ"
    assert np.array_equal(p.view(np.uint64), cl.p.view(np.uint64)), f"{what}: positions not bit-identical"
    live = cl.p[:, 3] != 0
    assert np.array_equal(v[live, :3].view(np.uint64), cl.vel[live, :3].view(np.uint64)), f"{what}: velocities differ"


@pytest.mark.parametrize("integrator,gen", [(0, "honeycomb_mesh"), (1, "honeycomb_mesh"), (4, "honeycomb_mesh_fast")], ids=["euler", "rk2", "rk4-C4"])
def test_honeycomb_prisms_match_the_oracle(synth, orc, integrator, gen):
    """rk4-C4: the oracle sample of BASELINE config 4 (bench.py poly10M_1e8_rk4: the same generator, cell shape, integrator
    and field family at a size the oracle finishes in seconds)."""
    from cudaparticlesfoam_b200 import api

    pm = getattr(synth, gen)(9, 8, 6)
    mesh = orc.tet_mesh_from_poly(pm)
    rng = np.random.default_rng(5)
    n = 30000
    p = np.ones((n, 4))
    p[:, :3] = pm.lo + rng.random((n, 3)) * (pm.hi - pm.lo)
    tet0 = orc.locate_brute(mesh, p)
    c = 0.5 * (pm.lo + pm.hi)
    U = np.zeros((pm.n_cells, 3))
    U[:, 0] = -(pm.cell_centres[:, 1] - c[1]) * 3.0 + 0.4
    U[:, 1] = (pm.cell_centres[:, 0] - c[0]) * 3.0
    U[:, 2] = 0.3
    tr = api.ParticleTracker(rng=api.RNG_NONE, integrator=integrator, sort_interval=6, fuse_substeps=5)
    tr.upload_poly(pm)
    info = tr.mesh_info()
    assert info["n_tets"] == 20 * pm.n_cells and info["n_cells"] == pm.n_cells
    tr.update_velocity(U)
    tr.set_particles(p)
    tr.locate_initial()
    _, _, t_dev = tr.download()
    inside = tet0 >= 0
    assert np.array_equal(t_dev[~inside] < 0, np.ones((~inside).sum(), dtype=bool)), "points outside the zigzag boundary must not be located"
    w = orc.bary_of(mesh, p[inside], np.maximum(t_dev[inside], 0))
    assert (t_dev[inside] >= 0).all() and w.min() > -1e-9, "every inside point must be located in a tet that contains it"
    # same start tets on both sides (on-face ties of the two locators aside), then 40 sub-steps with wall contacts
    tr.set_particles(p)
    tr.set_tets(tet0)
    cl = orc.Cloud.make(p, tet0)
    Utet = orc.expand_velocity(mesh, U)
    for chunk in (1, 9, 30):
        if integrator == 0:
            orc.substeps(mesh, cl, Utet, chunk, 0.01)
        else:
            orc.ext_substeps(mesh, cl, Utet, chunk, 0.01, integrator=integrator)
        tr.substeps(chunk, 0.01)
        pp, vv, tt = tr.download()
        _assert_same(pp, vv, tt, cl, f"honeycomb after chunk {chunk}")
    assert tr.stats()["n_reflections"] > 0
    tr.close()
