"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/cpf.h declares, and the product fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from conftest import HAVE_GPU, ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cpf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from cudaparticlesfoam_b200 import _lib

    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from cudaparticlesfoam_b200 import _lib

    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.cpf_abi_version() == 2


def test_config_struct_layout():
    from cudaparticlesfoam_b200 import _lib

    lib = _lib.load()
    c = _lib.CpfConfig()
    lib.cpf_default_config(C.byref(c))
    # src/initCuda.H:50-72 defaults
    assert (c.dt, c.diffusion_coeff, c.save_interval, c.seed) == (1e-4, 5.7e-6, 10, 1591593751)
    assert (c.interp, c.locator, c.integrator, c.rng, c.reflect_wall) == (0, 0, 0, 1, 1)
    assert C.sizeof(_lib.CpfStats) == 8 * 8 + 8 * 4


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful without a device")
def test_no_device_fails_loudly():
    from cudaparticlesfoam_b200 import api

    with pytest.raises(api.CpfError) as e:
        api.ParticleTracker()
    assert e.value.code == 5 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cudaparticlesfoam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "cpf_oracle" not in txt, f
