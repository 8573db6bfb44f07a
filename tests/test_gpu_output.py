"""SURVEY 8f N2 / N3 on the GPU: asynchronous binary VTU output and checkpoint / restart, through the C ABI.

The binary file must hold exactly what cpf_download returns at the moment of the call (the reference's
writeParticles2VTU arrays, cuda/utils.cpp:144-283), whatever is enqueued behind it; a restarted run must
continue bit-identically."""
import os
import re

import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu

_DT = {"Float64": "<f8", "Float32": "<f4", "Int32": "<i4", "UInt8": "u1"}


def read_vtu_appended(path):
    """Minimal reader for VTK XML with one raw appended section and UInt64 block headers."""
    raw = open(path, "rb").read()
    mark = raw.index(b"<AppendedData encoding='raw'>")
    start = raw.index(b"_", mark) + 1
    head = raw[:mark].decode()
    out = {}
    for m in re.finditer(r"<DataArray ([^>]*)/>", head):
        attr = dict(re.findall(r"(\w+)='([^']*)'", m.group(1)))
        off = start + int(attr["offset"])
        nbytes = int(np.frombuffer(raw[off:off + 8], dtype="<u8")[0])
        a = np.frombuffer(raw[off + 8:off + 8 + nbytes], dtype=_DT[attr["type"]])
        nc = int(attr.get("NumberOfComponents", "1"))
        out[attr["Name"]] = a.reshape(-1, nc) if nc > 1 else a
    npts = int(re.search(r"NumberOfPoints='(\d+)'", head).group(1))
    assert raw.rstrip().endswith(b"</VTKFile>")
    return npts, out


def _tracker(**kw):
    from cudaparticlesfoam_b200 import api

    kw.setdefault("rng", api.RNG_NONE)
    return api.ParticleTracker(**kw)


def _check_file(path, p, v, t, stride):
    npts, a = read_vtu_appended(path)
    sel = np.arange(0, p.shape[0], stride)
    assert npts == sel.size
    assert np.array_equal(a["Position"].view(np.uint64), np.ascontiguousarray(p[sel, :3]).view(np.uint64)), "positions must be the exact doubles"
    assert np.array_equal(a["ParticleType"], p[sel, 3].astype(np.int32))
    assert np.array_equal(a["ParticleID"], sel.astype(np.int32))
    assert np.array_equal(a["ParticleTetID"], t[sel]) and np.array_equal(a["ConvexTetID"], t[sel])
    vv = np.where(np.isnan(v[sel, :1]), 0.0, v[sel, :3])
    assert np.array_equal(a["vels"], vv.astype(np.float32))
    ke = 0.5 * (vv ** 2).sum(axis=1)
    assert np.allclose(a["KEs"], ke.astype(np.float32), rtol=2e-7, atol=0.0)
    assert np.array_equal(a["connectivity"], np.arange(sel.size, dtype=np.int32))
    assert np.array_equal(a["offsets"], np.arange(1, sel.size + 1, dtype=np.int32))
    assert a["types"].size == sel.size and np.all(a["types"] == 1)


def test_async_binary_vtu_holds_the_state_at_the_call(synth, orc, tmp_path):
    pm, mesh, U, p = make_case(synth, orc, dims=(10, 10, 10), jitter=0.15, n=60000, field="channel")
    tr = _tracker(sort_interval=4, fuse_substeps=4)
    tr.init_cuda(pm, U, particles=p)
    tr.substeps(12, 0.02)  # particles are in sorted (permuted) storage by now
    d = str(tmp_path)
    snaps = []
    for step, stride in ((1, 1), (2, 7), (3, 1)):
        snaps.append((step, stride, tr.download()))
        tr.write_vtu_async(d, step, stride)
        tr.substeps(8, 0.02)  # enqueued right behind the output: must not leak into the file
    tr.output_wait()
    for step, stride, (pp, vv, tt) in snaps:
        _check_file(os.path.join(d, f"particle_{step:04d}.vtu"), pp, vv, tt, stride)
    # the ASCII writer of the reference format sees the same particles (names and counts)
    tr.write_vtu(d, 9)
    txt = open(os.path.join(d, "particle_0009.vtu")).read(4000)
    assert f"NumberOfPoints='{p.shape[0]}'" in txt and "Name='Position'" in txt
    tr.close()


def test_async_output_reports_io_errors(synth, orc, tmp_path):
    pm, mesh, U, p = make_case(synth, orc, dims=(4, 4, 4), jitter=0.0, n=500)
    tr = _tracker()
    tr.init_cuda(pm, U, particles=p)
    tr.write_vtu_async(str(tmp_path / "no_such_dir"), 0, 1)
    with pytest.raises(RuntimeError):
        tr.output_wait()
    tr.write_vtu_async(str(tmp_path), 1, 1)  # the writer keeps working after an error
    tr.output_wait()
    assert os.path.exists(tmp_path / "particle_0001.vtu")
    tr.close()


@pytest.mark.parametrize("rng", [0, 1, 2], ids=["none", "xorwow", "philox"])
def test_checkpoint_restart_is_bit_identical(synth, orc, tmp_path, rng):
    pm, mesh, U, p = make_case(synth, orc, dims=(8, 8, 8), jitter=0.2, n=20000, field="channel")
    kw = dict(rng=rng, diffusion_coeff=2e-3 if rng else 0.0, sort_interval=5, fuse_substeps=4)
    a = _tracker(**kw)
    a.init_cuda(pm, U, particles=p)
    a.substeps(30, 0.02)
    pa, va, ta = a.download()
    sa = a.stats()
    a.close()

    b = _tracker(**kw)
    b.init_cuda(pm, U, particles=p)
    b.substeps(13, 0.02)
    ck = str(tmp_path / "cloud.ckpt")
    b.checkpoint_save(ck)
    b.close()

    c = _tracker(**kw)
    c.upload_poly(pm)
    c.update_velocity(U)
    c.checkpoint_load(ck)
    c.substeps(17, 0.02)
    pc, vc, tc = c.download()
    sc = c.stats()
    c.close()
    assert np.array_equal(ta, tc)
    assert np.array_equal(pa.view(np.uint64), pc.view(np.uint64))
    assert np.array_equal(va.view(np.uint64), vc.view(np.uint64))
    for k in ("n_reflections", "n_substeps", "n_active", "n_escaped"):
        assert sa[k] == sc[k], k
    assert sa["n_reflections"] > 0


def test_checkpoint_rejects_foreign_files(synth, orc, tmp_path):
    pm, mesh, U, p = make_case(synth, orc, dims=(4, 4, 4), jitter=0.0, n=100)
    tr = _tracker()
    tr.init_cuda(pm, U, particles=p)
    bad = tmp_path / "bad.ckpt"
    bad.write_bytes(b"not a checkpoint at all" * 10)
    with pytest.raises(RuntimeError):
        tr.checkpoint_load(str(bad))
    ck = str(tmp_path / "ok.ckpt")
    tr.checkpoint_save(ck)
    pm2, _, U2, _ = make_case(synth, orc, dims=(5, 4, 4), jitter=0.0, n=10)
    other = _tracker()
    other.upload_poly(pm2)
    with pytest.raises(RuntimeError):
        other.checkpoint_load(ck)  # another mesh
    other.close()
    tr.close()
