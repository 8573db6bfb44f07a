"""Soundness of the fp32 guarded walk, tested on the CPU.

oracle/cpf_filter_model.c restates the product's filter (cpf_geom.cuh visit_fast32 + the 64-byte records of
cpf_mesh.cu k_build_fast) over the oracle's mesh tables.  The property: whenever the filter certifies a walk, it ends
in the tet the reference's fp64 segment walk ends in -- over random segments and over segments built to graze
vertices, edges and faces, on regular, skewed, anisotropic and far-from-the-origin meshes.  Refusals cost time, never
correctness; their rate is bounded where the geometry is benign."""
import numpy as np
import pytest

from conftest import make_case  # noqa: F401  (fixtures synth / orc come from conftest)


def _cases(synth):
    yield "regular", synth.box_mesh(9, 8, 7, jitter=0.0), 0.93
    yield "jittered", synth.box_mesh(9, 8, 7, jitter=0.25), 0.93
    yield "anisotropic 1:50", synth.box_mesh(8, 8, 8, lo=(0, 0, 0), hi=(1.0, 1.0, 0.02), jitter=0.2), 0.90
    yield "far from the origin", synth.box_mesh(8, 8, 8, lo=(4096.0, -2048.0, 1024.0), hi=(4097.0, -2047.0, 1025.0), jitter=0.2), 0.90
    yield "tiny cells", synth.box_mesh(8, 8, 8, lo=(0, 0, 0), hi=(1e-5, 1e-5, 1e-5), jitter=0.2), 0.90
    yield "honeycomb prisms", synth.honeycomb_mesh(9, 8, 8), 0.85


def _segments(rng, pm, mesh, orc, n):
    """start points (random + exactly on vertices / edge midpoints / face centroids) and displacements (random lengths
    from 1e-6 to ~2 cells, plus segments that end on, or pass exactly through, a vertex / edge midpoint / face centroid)"""
    span = pm.hi - pm.lo
    h = span / np.array([8.0, 8.0, 8.0])
    p = np.ones((n, 4))
    p[:, :3] = pm.lo + (0.02 + 0.96 * rng.random((n, 3))) * span
    k = n // 8
    nt = mesh.idx.shape[0]
    t = rng.integers(0, nt, size=3 * k)
    a, b, c = mesh.pos[mesh.idx[t, 0]], mesh.pos[mesh.idx[t, 1]], mesh.pos[mesh.idx[t, 2]]
    p[:k, :3] = a[:k]
    p[k:2 * k, :3] = 0.5 * (a[k:2 * k] + b[k:2 * k])
    p[2 * k:3 * k, :3] = (a[2 * k:] + b[2 * k:] + c[2 * k:]) / 3.0
    inside = np.all((p[:, :3] > pm.lo) & (p[:, :3] < pm.hi), axis=1)
    p[~inside, :3] = pm.lo + 0.5 * span
    length = 10.0 ** rng.uniform(-6.0, 0.3, size=n)
    d = rng.normal(size=(n, 3))
    d *= (length / np.linalg.norm(d, axis=1))[:, None]
    d *= h
    # aimed segments: end exactly on a feature of a nearby tet, or pass through it (twice the distance)
    tet0 = orc.locate_brute(mesh, p)
    aim = np.arange(3 * k, 6 * k)
    tt = np.clip(tet0[aim], 0, nt - 1)
    va, vb, vc = mesh.pos[mesh.idx[tt, 1]], mesh.pos[mesh.idx[tt, 2]], mesh.pos[mesh.idx[tt, 3]]
    target = np.concatenate([va[:k], 0.5 * (va[k:2 * k] + vb[k:2 * k]), (va[2 * k:] + vb[2 * k:] + vc[2 * k:]) / 3.0])
    through = rng.random(3 * k) < 0.5
    d[aim] = (target - p[aim, :3]) * np.where(through, 2.0, 1.0)[:, None]
    disp = np.zeros((n, 4))
    disp[:, :3] = d
    return p, disp, tet0


def test_fp32_filter_certifies_only_what_the_reference_decides(synth, orc):
    rng = np.random.default_rng(1591593751)
    total = certified = 0
    for name, pm, min_rate in _cases(synth):
        mesh = orc.tet_mesh_from_poly(pm)
        fm = orc.FilterModel(mesh)
        for rep in range(3):
            p, disp, tet0 = _segments(rng, pm, mesh, orc, 60000)
            ok = tet0 >= 0
            p, disp, tet0 = p[ok], disp[ok], tet0[ok]
            out, vis = fm.walk(p, disp, tet0)
            cl = orc.Cloud.make(p, tet0)
            cl.disp[:] = disp
            orc.locate_convex(mesh, cl)
            cert = out >= 0
            wrong = cert & (out != cl.tet)
            assert not wrong.any(), (name, int(wrong.sum()), p[wrong][:3], disp[wrong][:3], out[wrong][:3], cl.tet[wrong][:3])
            # benign part of the sample (random starts, random directions): the filter must let almost everything through
            # that does not end at a wall
            n = p.shape[0]
            benign = np.zeros(n, dtype=bool)
            benign[6 * (60000 // 8):] = True
            benign = benign[:n] & (cl.tet >= 0)
            assert cert[benign].mean() > min_rate, (name, cert[benign].mean())
            assert vis[cert].max() <= 48
            total += n
            certified += int(cert.sum())
    assert total > 800000 and certified > 0.25 * total  # three quarters of the sample are built to be refused or to hit walls


def test_filter_model_walks_like_the_exact_walk_visits(synth, orc):
    """The model is a faithful stand-in: on a benign case the number of tets it visits equals the exact walk's."""
    pm = synth.box_mesh(8, 8, 8, jitter=0.2)
    mesh = orc.tet_mesh_from_poly(pm)
    fm = orc.FilterModel(mesh)
    rng = np.random.default_rng(7)
    p = np.ones((20000, 4))
    p[:, :3] = 0.1 + 0.8 * rng.random((20000, 3))
    disp = np.zeros((20000, 4))
    disp[:, :3] = 0.08 * rng.normal(size=(20000, 3))
    tet0 = orc.locate_brute(mesh, p)
    out, vis = fm.walk(p, disp, tet0)
    assert (out >= 0).mean() > 0.95
    assert 1.5 < vis[out >= 0].mean() < 6.0
    assert 1e-8 < fm.guard < 1e-3 and fm.hmin > 0


def test_the_guard_band_is_what_makes_the_filter_sound(synth, orc):
    """The same adversarial sample with the band switched off, or with only its geometric part G*V6 (no rounding-error
    term), produces wrong certifications -- so the property test above can fail, and both parts of g are needed."""
    rng = np.random.default_rng(1591593751)
    pm = synth.box_mesh(9, 8, 7, jitter=0.0)
    mesh = orc.tet_mesh_from_poly(pm)
    fm = orc.FilterModel(mesh)
    p, disp, tet0 = _segments(rng, pm, mesh, orc, 60000)
    ok = tet0 >= 0
    p, disp, tet0 = p[ok], disp[ok], tet0[ok]
    cl = orc.Cloud.make(p, tet0)
    cl.disp[:] = disp
    orc.locate_convex(mesh, cl)

    def wrong(**kw):
        out, _ = fm.walk(p, disp, tet0, **kw)
        return int(((out >= 0) & (out != cl.tet)).sum())

    assert wrong() == 0
    assert wrong(guard=0.0, err_scale=0.0) > 100   # no band at all
    assert wrong(err_scale=0.0) > 10               # G*V6 only: fp32 rounding decides some exits
    assert wrong(err_scale=0.25) == 0              # the derived bound (0.66 of the term) has margin in practice


def _wall_segments(rng, pm, mesh, orc, n):
    """particles in the cells next to the boundary, moving mostly outwards: plain wall hits after 0..3 hops, hits aimed at
    the edges and corners of the box (two or three walls at once), grazing hits, and hits on the vertices / edge
    midpoints / centroids of boundary faces"""
    span = pm.hi - pm.lo
    h = span / 8.0
    p = np.ones((n, 4))
    u = rng.random((n, 3))
    side = rng.integers(0, 6, size=n)
    ax, hi_side = side % 3, side >= 3
    depth = 10.0 ** rng.uniform(-4.0, 0.2, size=n) * h[ax]          # distance from the wall, 1e-4 .. 1.6 cells
    p[:, :3] = pm.lo + (0.03 + 0.94 * u) * span
    rows = np.arange(n)
    p[rows, ax] = np.where(hi_side, pm.hi[ax] - depth, pm.lo[ax] + depth)
    d = rng.normal(size=(n, 3)) * h * 0.6
    d[rows, ax] = np.where(hi_side, 1.0, -1.0) * np.abs(d[rows, ax]) * 2.0 + np.where(hi_side, depth, -depth)
    k = n // 6
    # into box corners and edges from at most ~1.5 cells away: two or three walls within one sub-step
    sgn = np.where(rng.random((2 * k, 3)) < 0.5, 0.0, 1.0)
    corner = pm.lo + sgn * span
    off = 10.0 ** rng.uniform(-3.0, 0.2, size=(2 * k, 3)) * h * np.where(sgn > 0, -1.0, 1.0)
    p[:2 * k, :3] = corner + off
    p[k:2 * k, 0] = pm.lo[0] + (0.1 + 0.8 * rng.random(k)) * span[0]      # second group: an edge, not a corner
    aim = corner + rng.normal(size=(2 * k, 3)) * h * 10.0 ** rng.uniform(-3.0, -0.3, size=(2 * k, 1))  # near, not at, the corner
    d[:2 * k] = (aim - p[:2 * k, :3]) * rng.uniform(1.1, 2.0, size=(2 * k, 1))
    d[k:2 * k, 0] = rng.normal(size=k) * h[0] * 0.3
    # exactly at features of boundary faces
    bfaces = np.flatnonzero((mesh.finfo[:, 0] < 0) | (mesh.finfo[:, 1] < 0))
    if bfaces.size:
        fsel = bfaces[rng.integers(0, bfaces.size, size=k)]
        tri = mesh.pos[mesh.facets[fsel, :3]]
        wgt = rng.dirichlet((0.3, 0.3, 0.3), size=k)                      # mass near vertices and edges
        target = (tri * wgt[:, :, None]).sum(axis=1)
        d[2 * k:3 * k] = (target - p[2 * k:3 * k, :3]) * rng.uniform(1.0, 2.0, size=(k, 1))
    disp = np.zeros((n, 4))
    disp[:, :3] = d
    return p, disp


def test_in_place_wall_reflection_is_the_references_reflection(synth, orc):
    """The wall-capable fast pass (cpf_advect.cu k_fast<.., WALL = 1>, cpf_geom.cuh wall_reflect_on_path) replays only the
    certified crossings in fp64 and mirrors the end point about the certified wall face.  Whenever the model of that
    pass certifies a sub-step -- with or without a wall contact -- position, velocity and tet after S5 must equal the
    reference's locate + reflect + move bit for bit; edges, corners and grazing hits must be refused or right."""
    rng = np.random.default_rng(20261017)
    n_wall = 0
    for name, pm, _ in _cases(synth):
        mesh = orc.tet_mesh_from_poly(pm)
        fm = orc.FilterModel(mesh)
        for rep in range(2):
            p, disp = _wall_segments(rng, pm, mesh, orc, 60000)
            tet0 = orc.locate_brute(mesh, p)
            ok = tet0 >= 0
            p, disp, tet0 = p[ok], disp[ok], tet0[ok]
            v = np.zeros_like(p)
            v[:, :3] = disp[:, :3] / 0.01
            ref = orc.Cloud.make(p, tet0)
            ref.disp[:] = disp
            ref.vel[:] = v
            orc.locate_convex(mesh, ref)
            hit = ref.tet < 0
            orc.reflect_convex(mesh, ref)
            orc.move(ref)
            mod = orc.Cloud.make(p, tet0)
            mod.vel[:] = v
            st = fm.substep(mod, disp)
            c = st > 0
            assert np.array_equal(mod.tet[c], ref.tet[c]), (name, int((mod.tet[c] != ref.tet[c]).sum()))
            assert np.array_equal(mod.p[c, :3].view(np.uint64), ref.p[c, :3].view(np.uint64)), name
            assert np.array_equal(mod.vel[c, :3].view(np.uint64), ref.vel[c, :3].view(np.uint64)), name
            assert not (st[~hit] == 2).any() and not (st[hit] == 1).any(), "wall contacts must be recognised as such"
            assert hit.mean() > 0.5, "the sample must be dominated by wall contacts"
            assert (st[hit] == 2).mean() > 0.35, (name, (st[hit] == 2).mean())
            n_wall += int((st == 2).sum())
            if rep == 0 and name == "jittered":
                # a handler without the exact replay of the crossed faces gets the hit point's last bits wrong
                bad = orc.Cloud.make(p, tet0)
                bad.vel[:] = v
                sb = fm.substep(bad, disp, skip_replay=True)
                cb = sb == 2
                assert (bad.p[cb, :3] != ref.p[cb, :3]).any(axis=1).sum() > 5  # rare (last bits), but bit-exactness is the bar
    assert n_wall > 100000


def test_start_points_certified_by_the_previous_substep_need_no_start_check(synth, orc):
    """inside a launch the product checks a particle's start point ONCE (C1) and afterwards relies on
    "the end point C2 certified is the next start point".  Chains of sub-steps through the model with the start check
    dropped wherever the previous sub-step was certified -- half of the displacements aimed to END within 1e-9 .. 1e-2 tet
    sizes of a face, an edge or a vertex of the tet they end in, so that the next sub-step starts as close to a feature as
    C2 lets it -- must still only certify what the reference's walk decides."""
    rng = np.random.default_rng(20261018)
    n, K = 40000, 6
    checked = skipped = 0
    for name, pm, _ in _cases(synth):
        mesh = orc.tet_mesh_from_poly(pm)
        fm = orc.FilterModel(mesh)
        span = pm.hi - pm.lo
        h = span / 8.0
        p = np.ones((n, 4))
        p[:, :3] = pm.lo + (0.05 + 0.9 * rng.random((n, 3))) * span
        tet = orc.locate_brute(mesh, p)
        ok = tet >= 0
        p, tet = p[ok], tet[ok]
        certified = np.zeros(p.shape[0], dtype=np.uint8)
        for k in range(K):
            m = p.shape[0]
            d = rng.normal(size=(m, 3)) * h * 10.0 ** rng.uniform(-1.5, -0.3, size=(m, 1))
            # aimed half: end next to a feature of the CURRENT tet (vertex / edge midpoint / face centroid), pulled towards the
            # tet's centroid by a tiny fraction: the end point sits just inside, the next start is as borderline as C2 allows
            aim = np.flatnonzero(rng.random(m) < 0.5)
            V = mesh.pos[mesh.idx[tet[aim]]]                       # [a, 4, 3]
            kind = rng.integers(0, 3, size=aim.size)
            wgt = np.zeros((aim.size, 4))
            for q in range(aim.size):
                sel = rng.permutation(4)[: kind[q] + 1]
                wgt[q, sel] = 1.0 / (kind[q] + 1)
            feat = (V * wgt[:, :, None]).sum(axis=1)
            cen = V.mean(axis=1)
            eps = 10.0 ** rng.uniform(-9.0, -2.0, size=(aim.size, 1))
            d[aim] = feat + eps * (cen - feat) - p[aim, :3]
            disp = np.zeros((m, 4))
            disp[:, :3] = d
            out, vis = fm.walk(p, disp, tet, skip_c1_first=certified)
            cl = orc.Cloud.make(p, tet)
            cl.disp[:] = disp
            orc.locate_convex(mesh, cl)
            cert = out >= 0
            wrong = cert & (out != cl.tet)
            assert not wrong.any(), (name, k, int(wrong.sum()), int(certified[wrong].sum()))
            checked += int(cert.sum())
            skipped += int((cert & (certified > 0)).sum())
            # S5 for everything that stayed inside; what hit a wall starts again from a fresh interior point
            inside = cl.tet >= 0
            pn = p.copy()
            pn[:, :3] = p[:, :3] + disp[:, :3]
            fresh = ~inside
            pn[fresh, :3] = pm.lo + (0.3 + 0.4 * rng.random((int(fresh.sum()), 3))) * span
            tn = cl.tet.copy()
            if fresh.any():
                tn[fresh] = orc.locate_brute(mesh, pn[fresh])
            keep = tn >= 0
            p, tet = pn[keep], tn[keep]
            certified = (cert & inside)[keep].astype(np.uint8)
    assert checked > 400000 and skipped > 150000, (checked, skipped)


def test_fp32_barycentric_walk_certifies_only_what_baryTetSearch_decides(synth, orc):
    """RTX=true build (CPF_LOCATOR_BARY): the product's fp32 walk towards the end point (cpf_geom.cuh visit_bary32, modelled
    by orc_filter_bary_walk) against the reference's baryTetSearch (query/RTQuery.cu:35-90, oracle s3_locate_bary), over the
    same random and feature-grazing segments as the convex walk: a certified result must be the reference's tet; end points
    exactly on vertices, edges and faces, equal minima and wall contacts must be refused."""
    rng = np.random.default_rng(77)
    total = certified = 0
    for name, pm, min_rate in _cases(synth):
        mesh = orc.tet_mesh_from_poly(pm)
        fm = orc.FilterModel(mesh)
        for rep in range(2):
            p, disp, tet0 = _segments(rng, pm, mesh, orc, 60000)
            ok = tet0 >= 0
            p, disp, tet0 = p[ok], disp[ok], tet0[ok]
            out, vis = fm.bary_walk(p, disp, tet0)
            cl = orc.Cloud.make(p, tet0)
            cl.disp[:] = disp
            orc.locate_bary(mesh, cl)
            cert = out >= 0
            wrong = cert & (out != cl.tet)
            assert not wrong.any(), (name, int(wrong.sum()), p[wrong][:3], disp[wrong][:3], out[wrong][:3], cl.tet[wrong][:3])
            n = p.shape[0]
            benign = np.zeros(n, dtype=bool)
            benign[6 * (60000 // 8):] = True
            benign = benign[:n] & (cl.tet >= 0)
            assert cert[benign].mean() > min_rate, (name, cert[benign].mean())
            total += n
            certified += int(cert.sum())
        # the band is what makes it sound here too
        if name == "regular":
            out0, _ = fm.bary_walk(p, disp, tet0, guard=0.0, err_scale=0.0)
            assert int(((out0 >= 0) & (out0 != cl.tet)).sum()) > 20
    assert total > 500000 and certified > 0.25 * total
