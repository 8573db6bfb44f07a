"""The random walk the bench actually runs: the stateless Philox stream, FUSED.

(i)  Fused chunks against the ORACLE: the deviates of the next k sub-steps are read back (cpf_debug_normals) and fed to
     oracle/cpf_oracle.c, so a 10-sub-step launch sequence -- all-particles pass, wall pass, exact finisher -- is compared
     bit for bit with the reference algorithm, not with the product's own exact-only mode.
(ii) The stream itself: the reference draws curand_normal_double from XORWOW (cuda/particles.cu:565-567); Philox4x32-10 +
     fp32 Box-Muller on the MUFU approximations is a different generator with narrower arithmetic, so it has to pass the
     tests a stand-in for N(0,1) deviates must pass: Kolmogorov-Smirnov per component, moments, covariance across
     components / sub-steps / neighbouring particles, tail counts, and the diffusion law <|dx|^2> = 6 D t in free space.
"""
import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.mark.parametrize("rng,fuse", [(2, 10), (2, 16), (1, 10)], ids=["philox-fuse10", "philox-fuse16", "xorwow-fuse10"])
def test_fused_random_walk_chunks_equal_the_oracle(synth, orc, rng, fuse):
    from cudaparticlesfoam_b200 import api

    pm, mesh, U, p = make_case(synth, orc, dims=(9, 8, 8), jitter=0.2, n=40000, field="channel")
    Utet = orc.expand_velocity(mesh, U)
    D, dt = 2e-3, 0.02
    tr = api.ParticleTracker(rng=rng, diffusion_coeff=D, fuse_substeps=fuse, sort_interval=13)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    if rng == api.RNG_XORWOW:
        tr.init_rng()
    tr.locate_initial()
    _, _, tet0 = tr.download(pos=False, vel=False)
    cl = orc.Cloud.make(p, tet0)
    for chunk in (fuse, 3, 2 * fuse + 1):
        xi = tr.normals(chunk)                     # [chunk, n, 3], original order, stream not advanced
        orc.substeps(mesh, cl, Utet, chunk, dt, xi=xi, D=D)
        tr.substeps(chunk, dt)                     # fused: ceil(chunk / fuse) launch sequences
        pp, vv, tt = tr.download()
        assert np.array_equal(tt, cl.tet), (chunk, int((tt != cl.tet).sum()))
        assert _same(pp, cl.p), chunk
        live = cl.p[:, 3] != 0
        assert _same(vv[live, :3], cl.vel[live, :3]), chunk
    st = tr.stats()
    assert st["n_reflections"] > 1000 and 0 < st["n_exact"] < 0.2 * st["n_substeps"]
    tr.close()


def test_philox_deviates_are_standard_normal(synth, orc):
    from scipy import stats

    from cudaparticlesfoam_b200 import api

    pm, mesh, U, _ = make_case(synth, orc, dims=(4, 4, 4), jitter=0.0, n=1)
    n, k = 400_000, 8
    p = np.ones((n, 4))
    p[:, :3] = 0.4
    tr = api.ParticleTracker(rng=api.RNG_PHILOX, diffusion_coeff=1e-3)
    tr.upload_poly(pm)
    tr.update_velocity(U)
    tr.set_particles(p)
    x = tr.normals(k)                              # [k, n, 3]
    N = x.size
    # moments
    assert abs(x.mean()) < 5.0 / np.sqrt(N)
    assert abs(x.var() - 1.0) < 5.0 * np.sqrt(2.0 / N)
    assert abs(stats.skew(x.ravel())) < 5.0 * np.sqrt(6.0 / N)
    assert abs(stats.kurtosis(x.ravel())) < 5.0 * np.sqrt(24.0 / N) + 2e-3   # 24-bit uniforms cut the tail at 5.9 sigma
    # Kolmogorov-Smirnov against N(0,1), per component and per sub-step
    for c in range(3):
        d, pv = stats.kstest(x[:, :, c].ravel()[:1_000_000], "norm")
        assert pv > 1e-3, (c, d, pv)
    for q in range(k):
        d, pv = stats.kstest(x[q].ravel()[:600_000], "norm")
        assert pv > 1e-3, (q, d, pv)
    # covariance: between the three components of a draw, between consecutive sub-steps of a particle, between
    # neighbouring particle ids (counter-based streams must not correlate along either counter axis)
    lim = 5.0 / np.sqrt(n * k)
    flat = x.reshape(-1, 3)
    cov = np.cov(flat.T)
    assert np.abs(cov - np.eye(3)).max() < 5.0 * np.sqrt(2.0 / flat.shape[0]) + lim
    for c in range(3):
        for c2 in range(3):
            assert abs(np.mean(x[:-1, :, c] * x[1:, :, c2])) < 5.0 / np.sqrt(n * (k - 1)), ("sub-step lag", c, c2)
            assert abs(np.mean(x[:, :-1, c] * x[:, 1:, c2])) < 5.0 / np.sqrt((n - 1) * k), ("particle lag", c, c2)
    # the two normals of one Box-Muller pair (components 0 and 1) share a radius: their squares must still be uncorrelated
    assert abs(np.mean((x[..., 0] ** 2 - 1) * (x[..., 1] ** 2 - 1))) < 5.0 * 2.0 / np.sqrt(n * k)
    # tails: P(|x| > 4) = 6.334e-5, P(|x| > 5) = 5.73e-7
    n4, n5 = int((np.abs(x) > 4).sum()), int((np.abs(x) > 5).sum())
    e4, e5 = 6.334e-5 * N, 5.73e-7 * N
    assert abs(n4 - e4) < 5.0 * np.sqrt(e4), (n4, e4)
    assert n5 <= e5 + 5.0 * np.sqrt(e5) + 3, (n5, e5)
    assert np.abs(x).max() < 5.9                    # sqrt(-2 ln 2^-25): the documented truncation of the fp32 stream
    # a different seed and a different sub-step index give different, equally distributed streams
    tr.set_config(seed=12345)
    y = tr.normals(2)
    assert not np.array_equal(y, x[:2]) and abs(np.mean(x[:2] * y)) < 5.0 / np.sqrt(y.size)
    tr.close()


def test_free_space_diffusion_law(synth, orc):
    """U = 0, no wall within reach: after T = k dt the mean squared displacement is 6 D T (2 D T per component), for the
    Philox stream as for the reference's XORWOW stream."""
    from cudaparticlesfoam_b200 import api

    pm = synth.box_mesh(12, 12, 12, jitter=0.1)
    U = np.zeros((pm.n_cells, 3))
    n, k, dt, D = 300_000, 40, 1e-3, 2.5e-3        # sigma per component after T: sqrt(2 D T) = 0.014 << 0.4
    p = np.ones((n, 4))
    p[:, :3] = 0.5 + 0.1 * (synth.seed_box(n, (0, 0, 0), (1, 1, 1))[:, :3] - 0.5)
    out = {}
    for name, rng in (("philox", api.RNG_PHILOX), ("xorwow", api.RNG_XORWOW)):
        tr = api.ParticleTracker(rng=rng, diffusion_coeff=D, fuse_substeps=10)
        tr.upload_poly(pm)
        tr.update_velocity(U)
        tr.set_particles(p)
        if rng == api.RNG_XORWOW:
            tr.init_rng()
        tr.locate_initial()
        tr.substeps(k, dt)
        pp, _, tt = tr.download()
        st = tr.stats()
        tr.close()
        assert (tt >= 0).all() and st["n_reflections"] == 0
        d = pp[:, :3] - p[:, :3]
        out[name] = d
        var = (d ** 2).mean(axis=0)
        assert np.allclose(var, 2.0 * D * k * dt, rtol=5.0 * np.sqrt(2.0 / n)), (name, var / (2.0 * D * k * dt))
        assert abs((d ** 2).sum(axis=1).mean() / (6.0 * D * k * dt) - 1.0) < 5.0 * np.sqrt(2.0 / (3 * n))
        assert np.abs(d.mean(axis=0)).max() < 5.0 * np.sqrt(2.0 * D * k * dt / n)
    assert not np.array_equal(out["philox"], out["xorwow"])
