"""Generates tests/golden/ref_*.npz from the UNMODIFIED reference kernels (oracle/_ref) on a GPU:
    gpurun -- 'python tests/golden/make_golden.py'   (writes into gpurun_out/golden/, copy back here)
Inputs are regenerated from seeds by cudaparticlesfoam_b200.synth, so only outputs (+ an input
checksum) are stored."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_case  # noqa: E402
from cudaparticlesfoam_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402

CASES = {
    "convex_vortex": dict(dims=(7, 6, 5), jitter=0.2, n=3000, field="vortex", steps=(1, 60), dt=0.03, convex=True),
    "bary_vortex": dict(dims=(7, 6, 5), jitter=0.2, n=3000, field="vortex", steps=(1, 60), dt=0.03, convex=False),
    "convex_channel_brownian": dict(dims=(8, 5, 5), jitter=0.1, n=3000, field="channel", steps=(1, 24), dt=0.02, convex=True, D=1e-3),
}


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, c in CASES.items():
        pm, mesh, U, p = make_case(synth, orc, dims=c["dims"], jitter=c["jitter"], n=c["n"], field=c["field"])
        Utet = orc.expand_velocity(mesh, U)
        tet0 = orc.locate_brute(mesh, p)
        D = c.get("D", 0.0)
        out = dict(input_sha=np.frombuffer(hashlib.sha256(mesh.pos.tobytes() + mesh.idx.tobytes() + p.tobytes() + U.tobytes()).digest(), dtype=np.uint8))
        if D:
            dr = orc.RefRun(mesh, Utet, p, tet0, init_rng=True)
            out["xi"] = np.stack([dr.draw_normals() for _ in range(sum(c["steps"]))]).astype(np.float64)
            dr.close()
        rr = orc.RefRun(mesh, Utet, p, tet0, init_rng=bool(D))
        for i, k in enumerate(c["steps"]):
            rr.substeps(k, c["dt"], convex=c["convex"], brownian=bool(D), D=D)
            g = rr.download()
            out[f"p{i}"] = g.p.copy(); out[f"tet{i}"] = g.tet.copy(); out[f"vel{i}"] = g.vel.copy()
        rr.close()
        np.savez_compressed(os.path.join(outdir, f"ref_{name}.npz"), **out)
        print(name, "ok", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
