#!/usr/bin/env python
"""bench.py -- particle-steps/s of the fused particle hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one body of the reference's src/advect.H: velocity-field refresh + nCycles fused
Lagrangian sub-steps over every particle of the rank.  Weak scaling: every rank tracks its own
particles on a replicated mesh; with N > 1 the per-step velocity field is broadcast from rank 0 over
NCCL and a statistics vector is reduced back.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (U already in HBM), `e2e` =
the same through the C ABI with the cell field coming from pinned host memory every step and a
statistics read-back, `roofline` = the fused kernel against the measured HBM copy bandwidth,
`cpu_baseline` = the oracle port on the host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = 72  # algorithmic bytes per particle-step (double4 position+flag and int32 tet id, read + written)

WORKLOADS = {
    # BASELINE.json configs[2] / north-star target: 1M-cell channel, 1e7 tracers, field refreshed every step
    "channel1M_1e7": dict(mesh="channel", dims=(400, 50, 50), jitter=0.1, n=10_000_000, dt=0.005, ncycles=10, D=1.5e-5,
                          field="channel", desc="channel 400x50x50 hex (1e6 cells, 12e6 tets), 1e7 tracers, Euler, cell-constant U "
                          "refreshed every step, convex walk + specular walls, random walk D=1.5e-5, 10 sub-steps/step"),
    # BASELINE.json configs[1]
    "box100_1e6": dict(mesh="box", dims=(100, 100, 100), jitter=0.1, n=1_000_000, dt=0.004, ncycles=10, D=0.0, field="vortex",
                       desc="box 100^3 hex (1e6 cells), 1e6 tracers, frozen uniform+vortex field"),
    # small smoke-sized variant (CI / CPU-side sanity)
    "channel_small": dict(mesh="channel", dims=(80, 20, 20), jitter=0.1, n=200_000, dt=0.02, ncycles=10, D=1.5e-5, field="channel",
                          desc="channel 80x20x20 hex, 2e5 tracers"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if t0 <= t <= t1 + 0.15 and len(r) >= 7:
                try:
                    sm.append(float(r[0])); mx = max(mx, float(r[1]))
                except ValueError:
                    continue
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        if not sm:
            for t, r in self.rows[-3:]:
                try:
                    sm.append(float(r[0])); mx = max(mx, float(r[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


def build_inputs(w, rank, world=1):
    from cudaparticlesfoam_b200 import parallel, synth

    if w["mesh"] == "channel":
        nx, ny, nz = w["dims"]
        pm = synth.box_mesh(nx, ny, nz, lo=(0, 0, 0), hi=(4.0, 1.0, 1.0), jitter=w["jitter"])
    else:
        pm = synth.box_mesh(*w["dims"], jitter=w["jitter"])
    span = pm.hi - pm.lo
    # weak scaling: ONE global cloud of world*n particles, partitioned by contiguous index range; every rank
    # generates only its own slice (same stream as a single big seed_box call would give)
    start, count = parallel.partition(w["n"] * world, world, rank)
    lo, hi = pm.lo + 0.02 * span, pm.hi - 0.02 * span
    p = np.empty((count, 4))
    for ax in range(3):
        p[:, ax] = lo[ax] + synth.uniform01(1591593751, start + count, stream=ax)[start:] * (hi[ax] - lo[ax])
    p[:, 3] = 1.0
    if w["field"] == "channel":
        fields = [synth.field_channel(pm.cell_centres, t=0.05 * k, lo=pm.lo, hi=pm.hi) for k in range(4)]
    else:
        fields = [synth.field_uniform_vortex(pm.cell_centres, omega=2 * np.pi * (1 + 0.02 * k)) for k in range(4)]
    return pm, p, fields


def structured_seed_tets(pm, p):
    """First tet of the hex cell that contains each particle on the un-jittered lattice (start guess
    for the reference's own narrow phase baryQuery)."""
    nx, ny, nz = pm.dims
    h = (pm.hi - pm.lo) / np.array([nx, ny, nz])
    ijk = np.clip(((p[:, :3] - pm.lo) / h).astype(np.int64), 0, np.array([nx, ny, nz]) - 1)
    return (12 * (ijk[:, 0] + nx * (ijk[:, 1] + ny * ijk[:, 2]))).astype(np.int32)


# --------------------------------------------------------------------------------------------------
def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from cudaparticlesfoam_b200 import api

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pm, p, fields = build_inputs(w, rank, world)
    integ = {"euler": api.EULER, "rk2": api.RK2, "rk4": api.RK4}[args.integrator]
    tr = api.ParticleTracker(device=local_rank, rng=api.RNG_PHILOX if w["D"] > 0 else api.RNG_NONE, diffusion_coeff=w["D"], dt=w["dt"],
                             sort_interval=args.sort_interval, fuse_substeps=args.fuse, path=api.PATH_EXACT if args.exact else api.PATH_FILTERED,
                             integrator=integ, interp=api.INTERP_VERTEX if args.interp == "vertex" else api.INTERP_TET)
    # one explicit non-default stream shared by torch (copies, NCCL, timing events) and the library
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    tr.set_stream(stream.cuda_stream)
    t0 = time.time()
    tr.upload_poly(pm)
    t_mesh = time.time() - t0
    tr.update_velocity(fields[0])
    tr.set_particles(p)
    t0 = time.time()
    tr.locate_initial()
    tr.sync()
    t_loc = time.time() - t0
    if not args.shuffled:
        tr.sort()  # seeded positions are uniformly random, i.e. shuffled with respect to the cells
    ncell = pm.n_cells
    u_host = [torch.from_numpy(f).pin_memory() for f in fields]
    u_dev = [torch.from_numpy(f).to(dev) for f in fields]
    u_stage = torch.empty((ncell, 3), dtype=torch.float64, device=dev)
    stat_dev = torch.zeros(4, dtype=torch.float64, device=dev)
    deltaT = w["dt"] * w["ncycles"]
    h2d = ncell * 24
    d2h = 0

    from cudaparticlesfoam_b200 import parallel

    def bcast(t):
        parallel.broadcast_field(t, src=0)  # NCCL over NVLink when world > 1, no-op otherwise

    def step_resident(k):
        src = u_dev[k % len(u_dev)]
        if world > 1:
            if rank == 0:
                u_stage.copy_(src)
            bcast(u_stage)
            src = u_stage
        tr.update_velocity_ptr(src.data_ptr(), True)
        tr.advect(None, deltaT)

    side = torch.cuda.Stream(device=dev)
    u_stage2 = [torch.empty((ncell, 3), dtype=torch.float64, device=dev) for _ in range(2)]
    ev_stage = torch.cuda.Event()
    ev_taken = [torch.cuda.Event(), torch.cuda.Event()]  # the library has copied u_stage2[b] into its own buffer

    def stage_next(k):
        b = k % 2
        side.wait_event(ev_taken[b])
        with torch.cuda.stream(side):
            if rank == 0:
                u_stage2[b].copy_(u_host[k % len(u_host)], non_blocking=True)  # H2D from pinned memory
            bcast(u_stage2[b])
            ev_stage.record(side)

    def step_e2e(k):
        nonlocal d2h
        if world == 1:
            # through the C ABI with a HOST buffer: cpf_update_velocity uploads on the library's copy stream, so the
            # field of step k+1 (pinned memory, 24 MB) crosses PCIe while the sub-steps of step k run; every step
            # still carries exactly one H2D of U and one D2H of the statistics inside the timed region
            tr.advect(None, deltaT)
            tr.update_velocity_ptr(u_host[(k + 1) % len(u_host)].data_ptr(), False)
        else:
            # same one-step-ahead pipeline across ranks: while the sub-steps of step k run on the compute stream, rank 0
            # uploads U(k+1) and NCCL broadcasts it on a side stream; the library then takes it from the device buffer
            tr.advect(None, deltaT)
            stage_next(k + 1)
            stream.wait_event(ev_stage)
            tr.update_velocity_ptr(u_stage2[(k + 1) % 2].data_ptr(), True)
            ev_taken[(k + 1) % 2].record(stream)
        st = tr.stats()  # D2H of the counters (synchronises the stream)
        d2h = 8 * 11
        if world > 1:
            st = parallel.reduce_stats(st, device=dev)  # particle statistics come back over NCCL
        return st

    def timed(fn, K):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_a = time.time()
        e0.record(stream)
        for k in range(K):
            fn(k)
        e1.record(stream)
        torch.cuda.synchronize()
        t_b = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t_a, t_b

    clocks = ClockSampler(local_rank) if rank == 0 else None
    t_load0 = time.time()
    for k in range(args.warmup):
        step_resident(k)
    tr.sync()
    st0 = tr.stats()
    launches0 = tr.launch_count()
    tr.profile(True)
    ms, ta, tb = timed(step_resident, args.steps)
    nl, prof_ms, prof_max = tr.profile_read()
    tr.profile(False)
    launches = tr.launch_count() - launches0
    st1 = tr.stats()
    psteps = st1["n_substeps"] - st0["n_substeps"]  # active particle-sub-steps actually executed on this rank
    if world == 1:
        tr.update_velocity_ptr(u_host[0].data_ptr(), False)  # primes the one-step-ahead upload of step_e2e
    else:
        stage_next(0)
        stream.wait_event(ev_stage)
        tr.update_velocity_ptr(u_stage2[0].data_ptr(), True)
        ev_taken[0].record(stream)
    for k in range(min(args.warmup, 2)):
        step_e2e(k)
    st2 = tr.stats()
    ms_e2e, _, tb2 = timed(step_e2e, args.steps)
    st3 = tr.stats()
    ck = None
    if clocks:
        # nvidia-smi samples every 100 ms; the timed regions are tens of ms, so keep the GPU under the same load until
        # at least ~1 s of samples exist and report the window that covers warm-up, both timed regions and this tail
        t_hold = time.time()
        k = 0
        while time.time() - t_load0 < 1.2 or time.time() - t_hold < 0.4:
            tr.advect(None, deltaT); k += 1  # rank-local load only: NO collective here (only rank 0 samples clocks)
            if k % 8 == 0:
                tr.sync()
        tr.sync()
        ck = clocks.window(t_load0, time.time())
        clocks.stop()
    tot = torch.tensor([float(psteps), float(st3["n_substeps"] - st2["n_substeps"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    value = float(tot[0].item()) / (ms * 1e-3)
    e2e_value = float(tot[1].item()) / (ms_e2e * 1e-3)
    peak, peak_src = _peaks()
    avg_launch_ms = prof_ms / max(nl, 1)
    sub_per_launch = (w["ncycles"] * args.steps) / max(nl, 1)
    # achieved GB/s = B_alg x particle-steps/s of the fused kernel(s); with k sub-steps fused per launch the
    # particle state actually crosses HBM once per launch, so measured DRAM traffic (`traffic`) is BELOW the algorithmic bytes
    achieved = B_ALG * psteps / (prof_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    out = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": w["desc"], "particles_per_gpu": w["n"], "cells": pm.n_cells,
                   "substeps_per_step": w["ncycles"], "substeps_per_launch": sub_per_launch, "integrator": args.integrator, "interpolation": args.interp, "sort_interval": args.sort_interval, "initial_order": "shuffled" if args.shuffled else "sorted by cell",
                   "path": "exact" if args.exact else "filtered",
                   "e2e_path": ("host U -> cpf_update_velocity (copy stream, one step ahead of the sub-steps) -> cpf_advect -> cpf_stats_get" if world == 1
                                else "rank 0 host U -> H2D -> ncclBroadcast on a side stream, one step ahead of the sub-steps -> cpf_update_velocity(device) -> cpf_advect -> cpf_stats_get + NCCL reduce"),
                   "l2_hygiene": "working set (particle state + mesh) > 126 MB L2, no flush",
                   "parallelism": f"particles partitioned over {world} GPU(s), mesh replicated"},
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": ck,
        "roofline": {"bound": "hbm", "kernel": "cpf::k_fast (+ k_exact rounds)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_particle_step": B_ALG,
                     "algorithmic_bytes_per_launch": B_ALG * psteps / max(nl, 1), "avg_launch_ms": avg_launch_ms, "launches_timed": nl,
                     "substeps_per_launch": sub_per_launch, "kernel_share_of_step": prof_ms / ms},
        "stats": {"exact_fraction": (st1["n_exact"] - st0["n_exact"]) / max(psteps, 1), "hops_per_substep": (st1["n_hops"] - st0["n_hops"]) / max(psteps, 1),
                  "reflections": st1["n_reflections"] - st0["n_reflections"], "active": st1["n_active"], "mesh_build_s": t_mesh, "initial_locate_s": t_loc},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(w, pm, p, fields, tr)
    tr.close()
    if world > 1:
        dist.destroy_process_group()
    return out


def cpu_baseline(w, pm, p, fields, tr=None, n_sample=1_000_000, n_sub=20):
    """The oracle port (reference algorithm restated in C, OpenMP over all host cores) on a bounded
    sample of the same workload: the first n_sample particles, n_sub sub-steps, random walk excluded
    (the deviates are an input of the oracle)."""
    from oracle import orc

    t0 = time.time()
    mesh = orc.tet_mesh_from_poly(pm)
    t_build = time.time() - t0
    n = min(n_sample, p.shape[0])
    ps = np.ascontiguousarray(p[:n])
    cl = orc.Cloud.make(ps, structured_seed_tets(pm, ps))
    orc.bary_query(mesh, cl)
    Utet = orc.expand_velocity(mesh, fields[0])
    orc.substeps(mesh, cl, Utet, 1, w["dt"])  # warm-up
    t0 = time.time()
    orc.substeps(mesh, cl, Utet, n_sub, w["dt"])
    dt = time.time() - t0
    out = {"value": n * n_sub / dt, "unit": "particle-steps/s", "cores": os.cpu_count(), "kind": "port",
           "sample": f"first {n} particles x {n_sub} sub-steps of the same mesh/field, no random walk; oracle/cpf_oracle.c with OpenMP; "
                     f"topology build {t_build:.1f}s not timed"}
    # second CPU column of SURVEY 8(d): OpenFOAM-style barycentric tracking, restated (OpenFOAM itself is not available)
    try:
        cf = orc.Cloud.make(ps.copy(), cl.tet.copy())
        cf.p[:, :3] = cl.p[:, :3]
        live = cf.tet >= 0
        cf = orc.Cloud.make(np.ascontiguousarray(cf.p[live]), np.ascontiguousarray(cf.tet[live]))
        orc.foamtrack_substeps(mesh, cf, Utet, 1, w["dt"])
        t0 = time.time()
        orc.foamtrack_substeps(mesh, cf, Utet, n_sub, w["dt"])
        t_all = time.time() - t0
        n1 = min(cf.n, 100_000)
        c1 = orc.Cloud.make(np.ascontiguousarray(cf.p[:n1]), np.ascontiguousarray(cf.tet[:n1]))
        t0 = time.time()
        orc.foamtrack_substeps(mesh, c1, Utet, n_sub, w["dt"], threads=1)
        t_one = time.time() - t0
        out["foamtrack"] = {"label": "OpenFOAM-style CPU tracking (restated; OpenFOAM not available)", "unit": "particle-steps/s",
                            "all_cores": cf.n * n_sub / t_all, "cores": os.cpu_count(), "one_core": n1 * n_sub / t_one,
                            "sample": f"{cf.n} particles x {n_sub} sub-steps on all cores, {n1} x {n_sub} on one core; oracle/cpf_foamtrack.c"}
    except Exception as e:  # the baseline column must never take the bench line down
        out["foamtrack"] = {"unavailable": repr(e)}
    return out


# --------------------------------------------------------------------------------------------------
def run_reference(args, w, rank, world, local_rank):
    """The reference arm.  The reference's only implementation of this path is CUDA: its three .cu
    files are compiled UNMODIFIED into oracle/_ref/libref_rtxadvect.so and driven in the exact call
    order of src/advect.H (5 kernels + 5 device-wide syncs per sub-step, XORWOW random walk, per-tet
    velocity refreshed from a 12x expanded host vector).  If that library or a GPU is missing, the
    CPU oracle port is timed instead."""
    if rank != 0:
        return None
    from oracle import orc

    pm, p, fields = build_inputs(w, 0)
    deltaT = w["dt"] * w["ncycles"]
    have_gpu = False
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    base = {"metric": "particle-steps/s", "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": args.workload, "description": w["desc"], "cells": pm.n_cells, "substeps_per_step": w["ncycles"]}}
    if args.ref_arm == "cuda" and have_gpu and orc.ref_available():
        mesh = orc.tet_mesh_from_poly(pm)
        Utets = [orc.expand_velocity(mesh, f) for f in fields[:2]]
        rr = orc.RefRun(mesh, Utets[0], p, structured_seed_tets(pm, p), init_rng=w["D"] > 0)
        rr.bary_query()
        import torch

        def step(k):
            Ut = orc.expand_velocity(mesh, fields[k % len(fields)])  # the glue's 12x host expansion (src/advect.H:44-54)
            rr.update_velocity(Ut)                                   # cudaUpdateVelocity: vector by value + H2D + D2D
            rr.substeps(w["ncycles"], deltaT / w["ncycles"], convex=True, brownian=w["D"] > 0, D=w["D"], reflect=True)

        for k in range(args.warmup):
            step(k)
        torch.cuda.synchronize()
        t0 = time.time()
        for k in range(args.steps):
            step(k)
        torch.cuda.synchronize()
        sec = time.time() - t0
        g = rr.download()
        active = int((g.p[:, 3] != 0).sum())
        rr.close()
        val = active * w["ncycles"] * args.steps / sec
        base.update({"value": val, "ms_per_step": sec / args.steps * 1e3,
                     "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": 0, "kind": "reference",
                                      "sample": "full workload; the reference's path exists only as CUDA kernels, so its unmodified kernels "
                                                "(oracle/_ref, sm_100a) run on cuda:0 in the src/advect.H call order, host velocity expansion "
                                                "included; wall-clock around the blocking reference calls"},
                     "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        base["config"]["particles"] = int(w["n"])
        return base
    cb = cpu_baseline(w, pm, p, fields, n_sample=min(w["n"], 2_000_000), n_sub=w["ncycles"] * max(1, min(args.steps, 3)))
    base.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                 "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-arm", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--workload", default="channel1M_1e7", choices=sorted(WORKLOADS))
    ap.add_argument("--sort-interval", type=int, default=50)
    ap.add_argument("--integrator", choices=["euler", "rk2", "rk4"], default="euler", help="BASELINE configs[1] is RK2, configs[3] RK4 (extensions; reference is Euler)")
    ap.add_argument("--interp", choices=["cell", "vertex"], default="cell", help="vertex = cellPoint-style interpolation (extension; reference default is the cell value)")
    ap.add_argument("--shuffled", action="store_true", help="locality probe (SURVEY 8d): no initial sort by cell; combine with --sort-interval 0")
    ap.add_argument("--fuse", type=int, default=10)
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        out = run_reference(args, w, rank, world, local_rank)
    else:
        out = run_ours(args, w, rank, world, local_rank)
    if rank == 0 and out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
