#!/usr/bin/env python
"""bench.py -- particle-steps/s of the fused particle hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--rng philox|xorwow|none]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one body of the reference's src/advect.H: velocity-field refresh + nCycles fused Lagrangian sub-steps over
every particle of the rank.  Weak scaling: every rank tracks its own index range of one global cloud on a replicated
mesh.  With N > 1 the data plane is the library's own (cpf_comm.cu: NCCL broadcast of the field, NCCL sum of the
statistics slot, both behind the C ABI); torch.distributed (gloo) only carries the 128-byte NCCL id and the final
max-over-ranks of the timings.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (U already in HBM), `e2e` = the same through the
C ABI with the cell field coming from pinned host memory every step and a statistics read-back, `roofline` = the
advect launch sequence against the measured HBM copy bandwidth, `cpu_baseline` = the oracle port on the host cores over
a bounded sample, `extra` = the other workloads / configurations of BASELINE.json measured the same way (short runs).
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

_CH = dict(mesh="channel", dims=(400, 50, 50), jitter=0.1, n=10_000_000, dt=0.005, ncycles=10, D=1.5e-5, integrator="euler")
WORKLOADS = {
    # BASELINE.json configs[2] / north-star target: 1M-cell channel mesh, 1e7 tracers, field refreshed every step.
    # Closed recirculating flow: statistically steady under the reference's all-reflecting walls (both arms run it).
    "channel1M_1e7": dict(_CH, field="recirc",
                          desc="channel mesh 400x50x50 hex (1e6 cells, 12e6 tets), 1e7 tracers, Euler, cell-constant U refreshed every "
                               "step (closed recirculating flow: steady state), convex walk + specular walls, random walk D=1.5e-5, "
                               "10 sub-steps/step"),
    # round 1's workload: through-flow parks the cloud on the reflecting outlet (throughput drifts down over the run)
    "channel1M_1e7_throughflow": dict(_CH, field="channel",
                                      desc="as channel1M_1e7 with a Poiseuille through-flow and an all-reflecting outlet (not a steady state)"),
    # through-flow with an ESCAPE outlet and continuous re-injection behind the inlet (features the reference lacks)
    "channel1M_1e7_escape": dict(_CH, field="channel", escape=("x+",), reseed_every=5,
                                 desc="as channel1M_1e7 with a Poiseuille through-flow, outlet patch ESCAPE, escaped tracers "
                                      "re-seeded behind the inlet every 5 steps (steady state)"),
    # BASELINE.json configs[4]: 1e8 particles, 1M cells, random walk, reflecting walls, outlet escape + compaction
    "channel1M_1e8_escape": dict(_CH, field="channel", escape=("x+",), n=12_500_000, n_single=100_000_000,
                                 desc="C5: channel 1e6 cells, 1e8 tracers over 8 GPUs (1.25e7 per GPU), through-flow, outlet patch "
                                      "ESCAPE, escaped tracers compacted away by the counting sort"),
    # BASELINE.json configs[1]
    # (outlet ESCAPE + re-injection: under the reference's all-reflecting walls the uniform part of the field parks the
    # whole cloud on the x+ wall within one box transit -- round 2 measured 0.64 reflections per particle-step there)
    "box100_1e6": dict(mesh="box", dims=(100, 100, 100), jitter=0.1, n=1_000_000, dt=0.004, ncycles=10, D=0.0, field="vortex",
                       integrator="rk2", escape=("x+",), reseed_every=5,
                       desc="C2: box 100^3 hex (1e6 cells), 1e6 tracers, RK2, frozen uniform+vortex field, outlet patch ESCAPE, "
                            "escaped tracers re-seeded behind the inlet every 5 steps (steady state)"),
    "box100_1e6_allreflect": dict(mesh="box", dims=(100, 100, 100), jitter=0.1, n=1_000_000, dt=0.004, ncycles=10, D=0.0, field="vortex",
                                  integrator="rk2", desc="as box100_1e6 with the reference's all-reflecting walls (the cloud piles up on the x+ wall)"),
    # BASELINE.json configs[0]: pitzDaily stand-in, launch-latency regime (1e5 tracers; a step = one save interval of 10 sub-steps)
    "pitz_1e5": dict(mesh="pitz", n=100_000, dt=1e-5, ncycles=10, D=5.7e-6, field="pitz", integrator="euler",
                     desc="C1: pitzDaily-sized block (12 225 hex cells), 1e5 tracers, Euler, frozen flow, random walk, "
                          "10 sub-steps (one save interval) per step"),
    # BASELINE.json configs[3]: polyhedral mesh, RK4, 1e8 tracers over 8 GPUs
    "poly10M_1e8_rk4": dict(mesh="honeycomb", dims=(232, 232, 186), n=12_500_000, dt=0.002, ncycles=10, D=0.0, field="vortex",
                            integrator="rk4", desc="C4: honeycomb of hexagonal prisms, 1.0e7 polyhedral cells (2.0e8 tets), 1e8 tracers "
                                                   "over 8 GPUs (1.25e7 per GPU), RK4, frozen uniform+vortex field"),
    "poly1M_rk4": dict(mesh="honeycomb", dims=(108, 108, 86), n=2_000_000, dt=0.002, ncycles=10, D=0.0, field="vortex",
                       integrator="rk4", desc="honeycomb of hexagonal prisms, 1.0e6 polyhedral cells (2.0e7 tets), 2e6 tracers, RK4"),
    # small smoke-sized variant (CI / CPU-side sanity)
    "channel_small": dict(mesh="channel", dims=(80, 20, 20), jitter=0.1, n=200_000, dt=0.02, ncycles=10, D=1.5e-5, field="recirc",
                          integrator="euler", desc="channel 80x20x20 hex, 2e5 tracers"),
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


def partition(n_total, n_ranks, rank):
    from cudaparticlesfoam_b200 import parallel

    return parallel.partition(n_total, n_ranks, rank)


def build_mesh(w):
    from cudaparticlesfoam_b200 import synth

    if w["mesh"] == "channel":
        return synth.box_mesh(*w["dims"], lo=(0, 0, 0), hi=(4.0, 1.0, 1.0), jitter=w["jitter"])
    if w["mesh"] == "pitz":
        return synth.backward_step_mesh()
    if w["mesh"] == "honeycomb":
        return synth.honeycomb_mesh_fast(*w["dims"])
    return synth.box_mesh(*w["dims"], jitter=w["jitter"])


def build_fields(w, pm, nf=4):
    from cudaparticlesfoam_b200 import synth

    f = w["field"]
    if f == "recirc":
        return [synth.field_recirculation(pm.cell_centres, t=0.05 * k, lo=pm.lo, hi=pm.hi) for k in range(nf)]
    if f == "channel":
        return [synth.field_channel(pm.cell_centres, t=0.05 * k, lo=pm.lo, hi=pm.hi) for k in range(nf)]
    if f == "pitz":  # frozen flow of the tutorial: ~10 m/s bulk velocity through a 0.3 m domain, one snapshot
        return [synth.field_channel(pm.cell_centres, t=0.0, Umax=10.0, eps=0.02, lo=pm.lo, hi=pm.hi)]
    span = float((pm.hi - pm.lo)[:2].min())
    return [synth.field_uniform_vortex(pm.cell_centres, omega=2 * np.pi * (1 + 0.02 * k), R=0.2 * span) for k in range(nf)]


def build_inputs(w, rank, world=1, n_override=None):
    from cudaparticlesfoam_b200 import synth

    pm = build_mesh(w)
    span = pm.hi - pm.lo
    n = int(n_override or w["n"])
    # weak scaling: ONE global cloud of world*n particles, partitioned by contiguous index range; every rank
    # generates only its own slice (same stream as a single big seed_box call would give)
    start, count = partition(n * world, world, rank)
    lo, hi = pm.lo + 0.02 * span, pm.hi - 0.02 * span
    p = np.empty((count, 4))
    for ax in range(3):
        p[:, ax] = lo[ax] + synth.uniform01(1591593751, start + count, stream=ax)[start:] * (hi[ax] - lo[ax])
    p[:, 3] = 1.0
    return pm, p, build_fields(w, pm), start


def common_config(args, w, pm, n_per_gpu):
    """The keys BOTH arms print (the driver compares them)."""
    return {"workload": args.workload, "description": w["desc"], "particles_per_gpu": int(n_per_gpu), "cells": int(pm.n_cells),
            "substeps_per_step": int(w["ncycles"]), "integrator": args.integrator or w["integrator"], "interpolation": args.interp,
            "random_walk_D": w["D"]}


def structured_seed_tets(pm, p):
    """First tet of the hex cell that contains each particle on the un-jittered lattice (start guess
    for the reference's own narrow phase baryQuery)."""
    nx, ny, nz = pm.dims
    h = (pm.hi - pm.lo) / np.array([nx, ny, nz])
    ijk = np.clip(((p[:, :3] - pm.lo) / h).astype(np.int64), 0, np.array([nx, ny, nz]) - 1)
    return (12 * (ijk[:, 0] + nx * (ijk[:, 1] + ny * ijk[:, 2]))).astype(np.int32)


# --------------------------------------------------------------------------------------------------
class Rig:
    """One rank's tracker + the pinned / device copies of the field snapshots + the step functions."""

    def __init__(self, args, w, rank, world, local_rank, uid, rng_name, n_override=None):
        import torch

        from cudaparticlesfoam_b200 import api

        self.torch, self.api, self.w, self.rank, self.world = torch, api, w, rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.pm, p, fields, start = build_inputs(w, rank, world, n_override)
        self.p, self.fields = p, fields
        integ = {"euler": api.EULER, "rk2": api.RK2, "rk4": api.RK4}[args.integrator or w["integrator"]]
        rng = {"philox": api.RNG_PHILOX, "xorwow": api.RNG_XORWOW, "none": api.RNG_NONE}[rng_name] if w["D"] > 0 else api.RNG_NONE
        self.tr = tr = api.ParticleTracker(device=local_rank, rng=rng, diffusion_coeff=w["D"], dt=w["dt"], sort_interval=args.sort_interval,
                                           fuse_substeps=args.fuse, path=api.PATH_EXACT if args.exact else api.PATH_FILTERED, integrator=integ,
                                           interp=api.INTERP_VERTEX if args.interp == "vertex" else api.INTERP_TET,
                                           locator=api.LOCATOR_BARY if getattr(args, "locator", "convex") == "bary" else api.LOCATOR_CONVEX)
        # one explicit non-default stream shared by torch (pinned copies, timing events) and the library
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        tr.set_stream(self.stream.cuda_stream)
        kind = None
        if w.get("escape"):
            kind = np.zeros(len(self.pm.patch_starts) - 1, dtype=np.int32)
            for name in w["escape"]:
                kind[list(self.pm.patch_names).index(name)] = api.PATCH_ESCAPE
        t0 = time.time()
        tr.upload_poly(self.pm, patch_kind=kind)
        tr.sync()
        self.t_mesh = time.time() - t0
        if world > 1:
            tr.comm_init(uid, rank, world)
        tr.update_velocity(fields[0])
        tr.set_particles(p)
        tr.set_particle_id_base(start)  # random-walk streams keyed by GLOBAL particle id
        if rng == api.RNG_XORWOW:
            tr.init_rng()
        t0 = time.time()
        tr.locate_initial()
        tr.sync()
        self.t_loc = time.time() - t0
        if not args.shuffled:
            tr.sort()  # seeded positions are uniformly random, i.e. shuffled with respect to the cells
        self.u_host = [torch.from_numpy(f).pin_memory() for f in fields]
        self.u_dev = [torch.from_numpy(f).to(self.dev) for f in fields] if rank == 0 or world == 1 else []
        self.deltaT = w["dt"] * w["ncycles"]
        span = self.pm.hi - self.pm.lo
        self.slab = (self.pm.lo + np.array([0.02, 0.05, 0.05]) * span, self.pm.lo + np.array([0.06, 0.95, 0.95]) * span)
        self.reseed_every = int(w.get("reseed_every", 0))
        self.pending = 0
        self.h2d = self.pm.n_cells * 24
        self.d2h = 0

    # U already resident in HBM (on rank 0; the library broadcasts it over NCCL when world > 1, one step ahead: the
    # exchange of U(k+1) runs on the library's copy stream beside the sub-steps of step k)
    def step_resident(self, k):
        tr = self.tr
        if self.world == 1:
            tr.update_velocity_ptr(self.u_dev[k % len(self.u_dev)].data_ptr(), True)
            tr.advect(None, self.deltaT)
        else:
            tr.advect(None, self.deltaT)
            tr.update_velocity_bcast(self.u_dev[(k + 1) % len(self.fields)].data_ptr() if self.rank == 0 else None, root=0, on_device=2)
        if self.reseed_every and (k + 1) % self.reseed_every == 0:
            tr.reseed_inactive(*self.slab)

    # through the C ABI with HOST buffers: U(k+1) from pinned memory (uploaded / broadcast on the library's copy stream
    # while the sub-steps of step k run) and a statistics read-back per step, collected one step late (no host stall)
    def step_e2e(self, k):
        tr = self.tr
        tr.advect(None, self.deltaT)
        nxt = self.u_host[(k + 1) % len(self.u_host)].data_ptr()
        if self.world == 1:
            tr.update_velocity_ptr(nxt, False)
        else:
            tr.update_velocity_bcast(nxt if self.rank == 0 else None, root=0, on_device=False)
        if self.reseed_every and (k + 1) % self.reseed_every == 0:
            tr.reseed_inactive(*self.slab)
        tr.stats_request(full=False)  # summed over the ranks on the device (NCCL) when world > 1
        self.pending += 1
        self.d2h = 16 * 8
        st = None
        if self.pending > 1:
            st = tr.stats_collect()
            self.pending -= 1
        return st

    def drain(self):
        st = None
        while self.pending:
            st = self.tr.stats_collect()
            self.pending -= 1
        return st

    def timed(self, fn, K, max_over_ranks):
        torch = self.torch
        if self.world > 1:
            max_over_ranks(0.0)  # barrier
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for k in range(K):
            fn(k)
        e1.record(self.stream)
        torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1))


def measure(rig, steps, warmup, max_over_ranks, sum_over_ranks, e2e=True):
    """-> dict(value, ms, e2e_value, ms_e2e, prof...) for one rig: warm-up, K resident steps, K end-to-end steps."""
    tr = rig.tr
    for k in range(warmup):
        rig.step_resident(k)
    tr.sync()
    st0 = tr.stats()
    launches0 = tr.launch_count()
    tr.profile(True)
    ms = rig.timed(rig.step_resident, steps, max_over_ranks)
    nl, prof_ms, prof_max = tr.profile_read()
    tr.profile(False)
    launches = tr.launch_count() - launches0
    st1 = tr.stats()
    psteps = sum_over_ranks(st1["n_substeps"] - st0["n_substeps"]) if rig.world == 1 else st1["n_substeps"] - st0["n_substeps"]
    out = {"ms": ms, "psteps": psteps, "value": psteps / (ms * 1e-3), "launches": launches, "nl": nl, "prof_ms": prof_ms, "st0": st0, "st1": st1}
    if e2e:
        if rig.world == 1:
            tr.update_velocity_ptr(rig.u_host[0].data_ptr(), False)  # primes the one-step-ahead upload
        else:
            tr.update_velocity_bcast(rig.u_host[0].data_ptr() if rig.rank == 0 else None, root=0, on_device=False)
        for k in range(min(warmup, 2)):
            rig.step_e2e(k)
        rig.drain()
        st2 = tr.stats()
        ms_e2e = rig.timed(lambda k: rig.step_e2e(k), steps, max_over_ranks)
        rig.drain()
        st3 = tr.stats()
        ps2 = st3["n_substeps"] - st2["n_substeps"]
        out.update({"ms_e2e": ms_e2e, "e2e_value": ps2 / (ms_e2e * 1e-3)})
    return out


def run_ours(args, w, rank, world, local_rank):
    import torch

    from cudaparticlesfoam_b200 import api, parallel

    uid = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("gloo")  # control plane only: the NCCL id and the max-over-ranks of the timings
        uid = parallel.share_unique_id(api.ParticleTracker.comm_unique_id if rank == 0 else None)

    def max_over_ranks(x):
        return parallel.max_over_ranks(x)

    def sum_over_ranks(x):
        return x  # statistics are already summed over the ranks inside the library (cpf_stats_get is collective)

    n_over = args.particles or (w.get("n_single") if world == 1 and args.full_single else None)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    t_load0 = time.time()
    rig = Rig(args, w, rank, world, local_rank, uid, args.rng, n_over)
    tr, pm = rig.tr, rig.pm
    m = measure(rig, args.steps, args.warmup, max_over_ranks, sum_over_ranks)
    ck = None
    if clocks:
        # nvidia-smi samples every 100 ms; the timed regions are tens of ms, so keep the GPU under the same load until
        # at least ~1 s of samples exist and report the window that covers warm-up, both timed regions and this tail
        t_hold = time.time()
        k = 0
        while time.time() - t_load0 < 1.2 or time.time() - t_hold < 0.4:
            tr.advect(None, rig.deltaT); k += 1  # rank-local load only: NO collective here (only rank 0 samples clocks)
            if k % 8 == 0:
                tr.sync()
        tr.sync()
        ck = clocks.window(t_load0, time.time())
        clocks.stop()
    st0, st1 = m["st0"], m["st1"]
    psteps = m["psteps"]
    peak, peak_src = _peaks()
    nl, prof_ms = m["nl"], m["prof_ms"]
    avg_launch_ms = prof_ms / max(nl, 1)
    sub_per_launch = (w["ncycles"] * args.steps) / max(nl, 1)
    psteps_rank = psteps / world  # the launch timings are this rank's
    # achieved GB/s = B_alg x particle-steps/s of the advect launch sequence; with k sub-steps fused per launch
    # the particle state actually crosses HBM once per launch, so the measured DRAM traffic (`traffic`) is BELOW that
    achieved = B_ALG * psteps_rank / (prof_ms * 1e-3) / 1e9
    traffic = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            traffic = None
    hops = (st1["n_hops"] - st0["n_hops"]) / max(psteps, 1)
    cfg = common_config(args, w, pm, p_count := rig.p.shape[0])
    out = {
        "metric": "particle-steps/s", "value": m["value"], "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "details": {"rng": args.rng if w["D"] > 0 else "none", "substeps_per_launch": sub_per_launch, "sort_interval": args.sort_interval,
                    "initial_order": "shuffled" if args.shuffled else "sorted by cell", "path": "exact" if (args.exact or args.locator == "bary") else "filtered", "locator": args.locator,
                    "e2e_path": ("pinned host U -> cpf_update_velocity (H2D + repack on the copy stream, one step ahead of the sub-steps) -> "
                                 "cpf_advect -> cpf_stats_request / cpf_stats_collect (one step late)" if world == 1 else
                                 "rank 0 pinned host U -> cpf_update_velocity_bcast (H2D + ncclBroadcast + repack on the copy stream, one step ahead) "
                                 "-> cpf_advect -> cpf_stats_request (ncclAllReduce of the slot on the device) / cpf_stats_collect"),
                    "l2_hygiene": "working set (particle state + mesh) > 126 MB L2, no flush",
                    "parallelism": f"particles partitioned over {world} GPU(s) by index range, mesh replicated; NCCL inside libcpf (cpf_comm.cu)"},
        "e2e": {"value": m["e2e_value"], "unit": "particle-steps/s", "h2d_bytes_per_step": rig.h2d, "d2h_bytes_per_step": rig.d2h,
                "ms_per_step": m["ms_e2e"] / args.steps},
        "gpu_launches": int(m["launches"]),
        "clocks": ck,
        "roofline": {"bound": "hbm", "kernel": "advect launch sequence: cpf::k_lean (all particles) + cpf::k_fast<FIN> (finishing pass over its refusals)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_particle_step": B_ALG,
                     "algorithmic_bytes_per_launch": B_ALG * psteps_rank / max(nl, 1), "avg_launch_ms": avg_launch_ms, "launches_timed": nl,
                     "substeps_per_launch": sub_per_launch, "kernel_share_of_step": prof_ms / m["ms"],
                     # what the kernel is REALLY limited by (ncu, profiles/r2_summary.md): dependent-load latency, instruction issue and the L1 data pipe;
                     # HBM is nearly idle because the particle state stays in registers across the fused sub-steps
                     "measured_limiter": "latency of the dependent chain visit -> exit face -> next record at 28 warps/SM: issue slots 61 % busy, L1 data pipe 72 %, DRAM 10 % (ncu, profiles/r2_summary.md); not HBM",
                     "real_dram_frac": (traffic / (avg_launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                     # mesh-inclusive figure: 72 B of state + one 64-byte record per visited tet + the 32-byte
                     # origin position on ~half of the hops + the 32-byte cell velocity per sub-step
                     "bytes_per_particle_step_with_mesh": B_ALG + 64.0 * hops + 16.0 * max(hops - 1.0, 0.0) + 32.0},
        "stats": {"exact_fraction": (st1["n_exact"] - st0["n_exact"]) / max(psteps, 1), "hops_per_substep": hops,
                  "reflections_per_particle_step": (st1["n_reflections"] - st0["n_reflections"]) / max(psteps, 1),
                  "escaped": st1["n_escaped"] - st0["n_escaped"], "active": st1["n_active"], "mesh_build_s": rig.t_mesh,
                  "initial_locate_s": rig.t_loc},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(w, pm, rig.p, rig.fields, tr)
    tr.close()
    del rig
    torch.cuda.empty_cache()
    if world == 1 and not args.no_extra and args.workload == "channel1M_1e7" and not args.exact:
        out["extra"] = extras(args, local_rank, max_over_ranks, sum_over_ranks)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return out


def extras(args, local_rank, max_over_ranks, sum_over_ranks):
    """Short runs of the other configurations BASELINE.json names, measured exactly like the headline (N = 1 only)."""
    import copy

    import torch

    res = {}

    def one(name, workload, rng, steps, warmup, **over):
        a = copy.copy(args)
        a.workload, a.integrator, a.rng = workload, over.get("integrator"), rng
        try:
            rig = Rig(a, WORKLOADS[workload], 0, 1, local_rank, None, rng)
            m = measure(rig, steps, warmup, max_over_ranks, sum_over_ranks)
            res[name] = {"workload": workload, "rng": rng if WORKLOADS[workload]["D"] > 0 else "none", "integrator": a.integrator or WORKLOADS[workload]["integrator"],
                         "value": m["value"], "e2e": m["e2e_value"], "ms_per_step": m["ms"] / steps, "steps": steps, "warmup": warmup,
                         "roofline_frac": B_ALG * m["psteps"] / (m["prof_ms"] * 1e-3) / 1e9 / _peaks()[0],
                         "substeps_per_launch": WORKLOADS[workload]["ncycles"] * steps / max(m["nl"], 1)}
            rig.tr.close()
            del rig
            torch.cuda.empty_cache()
        except Exception as e:  # an extra must never take the headline down
            res[name] = {"unavailable": repr(e)[:300]}

    # the drop-in's DEFAULT configuration: the reference's XORWOW stream, library-default fusing
    a0 = args.fuse
    args.fuse = 0
    one("xorwow_default_config", "channel1M_1e7", "xorwow", 8, 3)
    args.fuse = a0
    one("throughflow_all_reflect", "channel1M_1e7_throughflow", "philox", 8, 3)
    one("throughflow_escape_reseed", "channel1M_1e7_escape", "philox", 10, 3)
    one("C2_box100_1e6_rk2", "box100_1e6", "none", 8, 3)
    one("C1_pitz_1e5", "pitz_1e5", "xorwow", 20, 5)
    return res


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

    pm, p, fields, _ = build_inputs(w, 0)
    deltaT = w["dt"] * w["ncycles"]
    have_gpu = False
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    base = {"metric": "particle-steps/s", "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": common_config(args, w, pm, p.shape[0])}
    if args.ref_arm == "cuda" and have_gpu and orc.ref_available() and w["integrator"] == "euler" and w["mesh"] != "honeycomb":
        mesh = orc.tet_mesh_from_poly(pm)
        rr = orc.RefRun(mesh, orc.expand_velocity(mesh, fields[0]), p, structured_seed_tets(pm, p), init_rng=w["D"] > 0)
        rr.bary_query()
        import torch

        t_refresh = t_kernel = 0.0

        def step(k, timed=False):
            nonlocal t_refresh, t_kernel
            t0 = time.perf_counter()
            Ut = orc.expand_velocity(mesh, fields[k % len(fields)])  # the glue's 12x host expansion (src/advect.H:44-54)
            rr.update_velocity(Ut)                                   # cudaUpdateVelocity: vector by value + H2D + D2D
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            # the five kernels per sub-step, CUDA-event time around the whole loop (ref_shim: ref_substeps_timed)
            ms = rr.substeps_timed(w["ncycles"], deltaT / w["ncycles"], convex=True, brownian=w["D"] > 0, D=w["D"], reflect=True)
            if timed:
                t_refresh += t1 - t0
                t_kernel += ms * 1e-3

        for k in range(args.warmup):
            step(k)
        torch.cuda.synchronize()
        t0 = time.time()
        for k in range(args.steps):
            step(k, True)
        torch.cuda.synchronize()
        sec = time.time() - t0
        g = rr.download()
        active = int((g.p[:, 3] != 0).sum())
        rr.close()
        val = active * w["ncycles"] * args.steps / sec
        base.update({"value": val, "ms_per_step": sec / args.steps * 1e3,
                     # where the reference's step goes: host-side 12x expansion + upload vs its five kernels per sub-step
                     "refresh_ms_per_step": t_refresh / args.steps * 1e3, "kernel_ms_per_step": t_kernel / args.steps * 1e3,
                     "kernel_only_value": active * w["ncycles"] * args.steps / max(t_kernel, 1e-9),
                     "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": 0, "kind": "reference",
                                      "sample": "full workload; the reference's path exists only as CUDA kernels, so its unmodified kernels "
                                                "(oracle/_ref, sm_100a) run on cuda:0 in the src/advect.H call order, host velocity expansion "
                                                "included; wall-clock around the blocking reference calls"},
                     "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
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
    ap.add_argument("--rng", default="philox", choices=["philox", "xorwow", "none"],
                    help="random walk stream: philox (stateless) or xorwow (the reference's cuRAND stream, the drop-in's default)")
    ap.add_argument("--sort-interval", type=int, default=50)
    ap.add_argument("--integrator", choices=["euler", "rk2", "rk4"], default=None, help="default: the workload's (C2 is RK2, C4 RK4; reference is Euler)")
    ap.add_argument("--interp", choices=["cell", "vertex"], default="cell", help="vertex = cellPoint-style interpolation (extension; reference default is the cell value)")
    ap.add_argument("--locator", choices=["convex", "bary"], default="convex", help="bary = the reference's RTX=true build (barycentric point walk + RTreflection): reference arithmetic only")
    ap.add_argument("--shuffled", action="store_true", help="locality probe (SURVEY 8d): no initial sort by cell; combine with --sort-interval 0")
    ap.add_argument("--fuse", type=int, default=10, help="sub-steps per launch sequence (0 = library default)")
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the workload's)")
    ap.add_argument("--full-single", action="store_true", help="N=1: run the workload's whole multi-GPU cloud on the one GPU")
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE configurations")
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
