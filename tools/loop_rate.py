import sys, time, random, os, yaml, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import ROOT, base_cfg
from img_env_b200.envs import make_env
def cfg_(S):
    sampler = yaml.load(open(os.path.join(ROOT, "tests", "golden", "cfg", "test.yaml")), Loader=yaml.FullLoader)
    cfg = base_cfg(R=1, P=4, scene="rvoscene", n_obj=4); cfg.update(sampler)
    cfg.update(discrete_action=True, discrete_actions=[[0.0,-0.9],[0.2,0.0],[0.6,0.3],[0.4,0.9]], agent_num_per_env=1, image_batch=1, state_batch=3, laser_batch=0, time_max=50, continuous_actions=[[0,0.6],[-0.9,0.9]])
    cfg["wrapper"] = ["VelActionWrapper","TimeLimitWrapper","SensorsPaperRewardWrapper","InfoLogWrapper","MultiRobotCleanWrapper","StateBatchWrapper","ObsLaserStateTmp","NeverStopWrapper"]
    return cfg
for S in (64, 1024):
    random.seed(0)
    env = make_env(cfg_(S), num_scenes=S)
    t0=time.time(); env.reset(); torch.cuda.synchronize(); t_reset=time.time()-t0
    acts = torch.randint(0,4,(S,))
    n=100; resets=0
    torch.cuda.synchronize(); t0=time.time()
    for t in range(n):
        obs,r,d,info = env.step(acts); resets += int(info['all_down'].sum())
    torch.cuda.synchronize(); dt=time.time()-t0
    print('S',S,'full reset %.3fs'%t_reset,'loop %.1f robot-steps/s'%(S*n/dt),'ms/step %.2f'%(1e3*dt/n),'scene resets',resets)
    env.close()
import cProfile, pstats
S=1024; random.seed(0)
env = make_env(cfg_(S), num_scenes=S); env.reset(); acts = torch.randint(0,4,(S,))
pr = cProfile.Profile(); pr.enable()
for t in range(100): env.step(acts)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
