#!/usr/bin/env python
"""SURVEY 8f N2 measurement: how long the host is blocked by particle output, per file.
  ascii  = cpf_write_vtu (the reference's writeParticles2VTU format, blocking D2H + fprintf)
  async  = cpf_write_vtu_async (device pack, copy stream, writer thread, raw appended binary)
Prints one JSON line.  Usage: python tools/output_probe.py [--n 1000000] [--files 4]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--files", type=int, default=4)
    ap.add_argument("--substeps", type=int, default=10)
    a = ap.parse_args()
    from cudaparticlesfoam_b200 import api, synth

    pm = synth.channel_mesh(100, 25, 25)
    U = synth.field_channel(pm.cell_centres, lo=pm.lo, hi=pm.hi)
    p = synth.seed_box(a.n, pm.lo + 0.02 * (pm.hi - pm.lo), pm.hi - 0.02 * (pm.hi - pm.lo))
    tr = api.ParticleTracker(rng=api.RNG_PHILOX, diffusion_coeff=1.5e-5, dt=2e-4, sort_interval=50, fuse_substeps=10)
    tr.init_cuda(pm, U, particles=p)
    tr.substeps(a.substeps, 2e-4)
    tr.sync()
    out = {"n_particles": a.n, "files": a.files, "substeps_between_files": a.substeps}
    with tempfile.TemporaryDirectory() as d:
        # compute only
        t0 = time.perf_counter()
        for k in range(a.files):
            tr.substeps(a.substeps, 2e-4)
        tr.sync()
        out["compute_only_s"] = time.perf_counter() - t0
        # blocking ASCII
        t0 = time.perf_counter()
        for k in range(a.files):
            tr.substeps(a.substeps, 2e-4)
            tr.write_vtu(d, k)
        tr.sync()
        out["ascii_total_s"] = time.perf_counter() - t0
        out["ascii_bytes_per_file"] = os.path.getsize(os.path.join(d, "particle_0000.vtu"))
        # asynchronous binary
        blocked = 0.0
        t0 = time.perf_counter()
        for k in range(a.files):
            tr.substeps(a.substeps, 2e-4)
            t1 = time.perf_counter()
            tr.write_vtu_async(d, 100 + k, 1)
            blocked += time.perf_counter() - t1
        t_issue = time.perf_counter() - t0
        tr.sync()  # drains the writer as well
        out["async_total_s"] = time.perf_counter() - t0
        out["async_issue_s"] = t_issue
        out["async_host_blocked_in_call_s"] = blocked
        out["async_bytes_per_file"] = os.path.getsize(os.path.join(d, "particle_0100.vtu"))
    out["ascii_stall_per_file_s"] = (out["ascii_total_s"] - out["compute_only_s"]) / a.files
    out["async_stall_per_file_s"] = (out["async_total_s"] - out["compute_only_s"]) / a.files
    tr.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
