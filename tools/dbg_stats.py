import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench
from img_env_b200.spec import build_spec
from img_env_b200.lib import BatchedSim
for wn, S in (('c4',2),('c2',16),('c5',4),('c1',64)):
    w = bench.WORKLOADS[wn]; spec = build_spec(bench.make_cfg(w))
    sim = BatchedSim(spec, S, ped_yaw_mode=1); sim.reset(bench.make_resets(spec, w, S, 1))
    st = sim.debug_stats().reshape(-1,4)
    print(wn, 'tiles mean/max', st[:,0].mean(), st[:,0].max(), 'boundary mean/max', st[:,1].mean(), st[:,1].max(), 'heavy', st[:,2].mean(), st[:,2].max(), 'fallback', st[:,3].mean())
    sim.close()
