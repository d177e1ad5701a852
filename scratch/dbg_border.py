import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import *
from test_parity_gpu import _variant
from oracle.pyref import RefEnv, PyPost
from img_env_b200.lib import BatchedSim
cfg = _variant(R=4, P=2, scene="rvoscene", n_obj=1); spec = build_spec(cfg); R=4
rng = np.random.default_rng(28)
sim = BatchedSim(spec, 1, ped_yaw_mode=1); ref = RefEnv(spec); post = PyPost(spec)
rs = make_reset(spec, rng, lo=0.3, hi=2.0)
sim.reset([rs]); st = ref.reset(rs); post.on_reset(); post.get_states(st)
dones = np.zeros(R, int)
for t in range(3):
    acts = random_actions(R, rng); alive = (dones==0).astype(np.uint8)
    rb,pd = ref.get_internal(); rb=rb.copy(); rb[:,15] = post.tmp_distances if post.tmp_distances is not None else np.nan
    sim.set_internal(rb[None], pd[None], ref.rvo_get().astype(np.float64)[None])
    out = sim.step(torch.from_numpy(acts[None]).cuda(), torch.from_numpy(alive[None]).cuda()); torch.cuda.synchronize()
    vm = sim.debug_view_maps()[0]; vm2 = sim.debug_view_maps()[0]
    st = ref.step(acts*alive[:,None], alive); want = post.get_states(st)
    rb2,_ = ref.get_internal()
    for j in range(R):
        d = vm[j] != st['view_map'][j]
        print(t, j, 'alive', alive[j], 'coll/arr after', rb2[j,12], rb2[j,13], 'diff', int(d.sum()), 'dbg stable', int((vm[j]!=vm2[j]).sum()),
              'pose', np.round(rb2[j,:3],3), 'sm eq', bool((out['sensor_maps'][0,j].cpu().numpy().view(np.uint16) == want['sensor_maps'][j].view(np.uint16)).all()))
        if d.sum():
            ii,jj = np.nonzero(d); print('   bbox', ii.min(), ii.max(), jj.min(), jj.max(), 'pairs', {(int(a),int(b)) for a,b in zip(vm[j][d][:2000], st['view_map'][j][d][:2000])})
    dones = np.clip(np.clip(want['is_collisions'],-1,1)+want['is_arrives'],0,1)
print(sim.debug_stats())
