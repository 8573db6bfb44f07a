import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    """Ask the product itself (cpf_create probes the CUDA runtime): no PyTorch needed to see a device."""
    import ctypes as C

    lib_path = os.path.join(ROOT, "cudaparticlesfoam_b200", "libcpf.so")
    if os.path.exists(lib_path):
        try:
            lib = C.CDLL(lib_path)
            h = C.c_void_p()
            lib.cpf_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
            rc = lib.cpf_create(None, C.byref(h))
            if rc == 0:
                lib.cpf_destroy.argtypes = [C.c_void_p]
                lib.cpf_destroy(h)
                return True
            return False
        except OSError:
            pass
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def n_gpus() -> int:
    import ctypes as C

    for name in ("libcudart.so", "libcudart.so.12", "libcudart.so.13"):
        try:
            rt = C.CDLL(name)
            n = C.c_int(0)
            if rt.cudaGetDeviceCount(C.byref(n)) == 0:
                return n.value
        except OSError:
            continue
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 1 if HAVE_GPU else 0


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    expr = (config.getoption("-m") or "").strip()
    if expr == "gpu":
        # the GPU suite was asked for explicitly: a green run of 0 tests would hide that nothing was checked
        raise pytest.UsageError("-m gpu requested but libcpf sees no CUDA device (cpf_create -> CPF_ERR_NO_DEVICE)")
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o

    o.build()
    return o


@pytest.fixture(scope="session")
def synth():
    from cudaparticlesfoam_b200 import synth as s

    return s


def make_case(synth, orc, dims=(8, 7, 6), jitter=0.2, n=4000, field="vortex", seed=1591593751, margin=0.01):
    """A small seeded case: polyMesh, oracle tet mesh, cell field, cloud with brute-force start tets."""
    pm = synth.box_mesh(*dims, jitter=jitter)
    mesh = orc.tet_mesh_from_poly(pm)
    if field == "vortex":
        U = synth.field_uniform_vortex(pm.cell_centres, R=0.3)
    elif field == "swirl":
        U = synth.field_uniform_vortex(pm.cell_centres, U0=(0.0, 0.0, 0.0), R=0.3)
    elif field == "channel":
        U = synth.field_channel(pm.cell_centres, lo=pm.lo, hi=pm.hi)
    else:
        U = np.tile(np.asarray(field, dtype=np.float64), (pm.n_cells, 1))
    lo = pm.lo + margin * (pm.hi - pm.lo)
    hi = pm.hi - margin * (pm.hi - pm.lo)
    p = synth.seed_box(n, lo, hi, seed=seed)
    return pm, mesh, np.ascontiguousarray(U), p
