import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import *
from oracle.pyref import RefEnv, PyPost
from img_env_b200.lib import BatchedSim
def run(tag, **kw):
    lo=kw.pop('lo',3.0); hi=kw.pop('hi',8.0); seed=kw.pop('seed',7)
    cfg = base_cfg(**kw); spec = build_spec(cfg); R=spec['R']
    rng = np.random.default_rng(seed)
    sim = BatchedSim(spec, 1, ped_yaw_mode=1); ref = RefEnv(spec)
    rs = make_reset(spec, rng, lo=lo, hi=hi)
    sim.reset([rs]); ref.reset(rs)
    acts = random_actions(R, rng)
    rb,pd = ref.get_internal(); rb[:,15]=np.nan
    sv = ref.sfm_get()
    sim.set_internal(rb[None], pd[None], sv[None][:, :sim.solver_agents])
    sim.step(torch.from_numpy(acts[None]).cuda(), torch.ones(1,R,dtype=torch.uint8,device='cuda')); torch.cuda.synchronize()
    ref.step(acts, np.ones(R))
    a = sim.get_internal()[2][0]; b = ref.sfm_get()[:len(a)]
    d = np.abs(a[:,3:5]-b[:,3:5]).max(1)
    print(tag, 'NA', len(a), 'in_tree pre', sv[:,10].astype(int).tolist(), 'vel diff per agent', np.round(d,5).tolist(), 'dest', a[:,7].tolist(), b[:,7].tolist())
run('P6R2', R=2,P=6,scene='pedscene',n_obj=3)
run('P6R3', R=3,P=6,scene='pedscene',n_obj=3)
run('P9R1', R=1,P=9,scene='pedscene',n_obj=0, max_ped=10)
run('P10norel', R=1,P=10,scene='pedscene',n_obj=0, relation=0, max_ped=10)
run('P8R1', R=1,P=8,scene='pedscene',n_obj=0, max_ped=10)
