"""RL-loop rate through the reference's wrapper stack (envs/cfg/test.yaml's list) on batched scenes:
    python tools/loop_rate.py [scenes ...]
Actions are a device tensor (what a policy network would output); nothing but the wrappers' own bookkeeping touches the host."""
import os, random, sys, time
import torch, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from img_env_b200.envs import make_env, GraphedStep
from img_env_b200.scenarios import MAP_DIR


def cfg_():
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "cfg", "test_full.yaml")))
    cfg["time_max"] = 50
    cfg["wrapper"] = ["VelActionWrapper", "TimeLimitWrapper", "SensorsPaperRewardWrapper", "InfoLogWrapper", "MultiRobotCleanWrapper",
                      "TestEpisodeWrapper", "StateBatchWrapper", "ObsLaserStateTmp", "NeverStopWrapper"]
    return cfg


def run(S, n=200, profile=False, graphed=False):
    random.seed(0)
    env = make_env(cfg_(), num_scenes=S, map_dir=MAP_DIR)
    env.reset()
    acts = torch.randint(0, 28, (S,), device="cuda")
    for t in range(20):
        env.step(acts)
    step = GraphedStep(env, acts).step if graphed else env.step
    torch.cuda.synchronize()
    pr = None
    if profile:
        import cProfile
        pr = cProfile.Profile(); pr.enable()
    t0 = time.time()
    for t in range(n):
        obs, r, d, info = step(acts)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if pr:
        import pstats
        pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
    print("scenes %d %s: %.3f M robot-steps/s, %.3f ms/step, empty-queue resets %d" %
          (S, "CUDA-graph replay" if graphed else "eager", S * n / dt / 1e6, 1e3 * dt / n, env.sim.debug_counters()[2]), flush=True)
    env.close()


if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:] if x.isdigit()] or [1024, 8192]
    for S in sizes:
        run(S)
        run(S, graphed=True)
    if "--profile" in sys.argv:
        run(sizes[0], n=100, profile=True)
