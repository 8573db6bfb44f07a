"""TEST SCAFFOLDING for tests/test_gpu_multi.py: one rank of a 2-rank job, one GPU per rank, NCCL inside libcpf.
usage: multi_rank_worker.py rank nranks dir rng n nsub dt -- the unique id travels through a file (the host's transport)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, nranks, d, rng, n, nsub, dt = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6]), float(sys.argv[7])
    from cudaparticlesfoam_b200 import api, synth

    idf = os.path.join(d, "nccl_id.bin")
    if rank == 0:
        uid = api.ParticleTracker.comm_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            if time.time() - t0 > 120:
                raise RuntimeError("no unique id from rank 0")
            time.sleep(0.05)
        uid = open(idf, "rb").read()
    pm = synth.box_mesh(9, 8, 7, jitter=0.2)
    lo, hi = pm.lo + 0.02, pm.hi - 0.02
    fields = [synth.field_uniform_vortex(pm.cell_centres, R=0.3, omega=5.0 + k) for k in range(4)]
    tr = api.ParticleTracker(device=rank, rng=api.RNG_PHILOX if rng == "philox" else api.RNG_XORWOW, diffusion_coeff=2e-3, fuse_substeps=4,
                             sort_interval=6)
    tr.upload_poly(pm)
    tr.comm_init(uid, rank, nranks)
    first = rank * n // nranks
    count = (rank + 1) * n // nranks - first
    tr.seed_box_slice(first, count, lo, hi)
    # cells owned by this rank in a decomposed solver run: a contiguous range of the global cell ids
    c0 = rank * pm.n_cells // nranks
    c1 = (rank + 1) * pm.n_cells // nranks
    tr.update_velocity_bcast(fields[0] if rank == 0 else None, root=0)
    if rng == "xorwow":
        tr.init_rng()
    tr.locate_initial()
    for k, U in enumerate(fields):
        if k % 2 == 0:
            tr.update_velocity_bcast(U if rank == (k // 2) % nranks else None, root=(k // 2) % nranks)
        else:
            tr.update_velocity_slices(c0, U[c0:c1])
        tr.substeps(nsub, dt)
    p, v, t = tr.download()
    st = tr.stats()  # collective: summed over the ranks on the device
    r, nr, ver = tr.comm_info()
    np.savez(os.path.join(d, f"rank{rank}.npz"), p=p, v=v, t=t)
    json.dump({"stats": st, "rank": r, "nranks": nr, "nccl": ver}, open(os.path.join(d, f"rank{rank}.json"), "w"))
    tr.close()


if __name__ == "__main__":
    main()
