import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench
from cudaparticlesfoam_b200 import api
w=bench.WORKLOADS['channel1M_1e7']
pm,p,fields=bench.build_inputs(w,0,1)
dev=torch.device('cuda',0)
tr=api.ParticleTracker(rng=api.RNG_PHILOX, diffusion_coeff=w['D'], dt=w['dt'], sort_interval=20, fuse_substeps=10)
stream=torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); tr.set_stream(stream.cuda_stream)
tr.upload_poly(pm); tr.update_velocity(fields[0]); tr.set_particles(p); tr.locate_initial(); tr.sort()
uh=[torch.from_numpy(f).pin_memory() for f in fields]; us=torch.empty((pm.n_cells,3),dtype=torch.float64,device=dev)
def T(fn,n=6):
    torch.cuda.synchronize(); t=time.time()
    for k in range(n): fn(k)
    torch.cuda.synchronize(); return (time.time()-t)/n*1e3
for _ in range(3): tr.advect(None, 0.05)
print("advect only", T(lambda k: tr.advect(None,0.05)))
print("h2d only", T(lambda k: us.copy_(uh[k%4], non_blocking=True)))
print("update_velocity_ptr", T(lambda k: tr.update_velocity_ptr(us.data_ptr(), True)))
print("stats only", T(lambda k: tr.stats()))
print("sort only", T(lambda k: tr.sort()))
def e2e(k):
    us.copy_(uh[k%4], non_blocking=True); tr.update_velocity_ptr(us.data_ptr(), True); tr.advect(None,0.05); tr.stats()
print("e2e", T(e2e))
tr.set_config(sort_interval=0)
print("advect no sort", T(lambda k: tr.advect(None,0.05)))
