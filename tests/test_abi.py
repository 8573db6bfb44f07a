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
    assert lib.cpf_abi_version() == 3


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


def test_bench_reference_arm_contract_on_cpu():
    """bench.py --impl reference must print ONE JSON line with the contract's keys; without a GPU (or with --ref-arm cpu)
    it times the oracle port on the host cores.  Runs the small workload so that the CPU suite stays short."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-arm", "cpu", "--workload", "channel_small",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
