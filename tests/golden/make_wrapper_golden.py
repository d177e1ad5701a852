"""Golden vectors for the wrapper stack: runs the REAL reference wrappers (/root/reference/envs/wrapper/base.py,
filter_states.py, imported by path with a stub `gym`) around tests/fake_env.FakeEnv in the order of
envs/cfg/test.yaml and stores what an RL loop would see. tests/test_wrappers_cpu.py replays the same seeds
through img_env_b200.envs.wrappers."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference"
ORDER = ["VelActionWrapper", "TimeLimitWrapper", "SensorsPaperRewardWrapper", "InfoLogWrapper", "MultiRobotCleanWrapper",
         "StateBatchWrapper", "ObsLaserStateTmp", "NeverStopWrapper"]


def wrapper_cfg(discrete):
    return dict(discrete_action=discrete, discrete_actions=[[0.0, -0.9], [0.2, 0.0], [0.6, 0.3], [0.4, 0.9, 1]],
                continuous_actions=[[0, 0.6], [-0.9, 0.9]], time_max=7, robot=dict(total=4), ped_sim=dict(total=3),
                env_type="robot_nav", agent_num_per_env=4, ped_safety_space=0.7, image_batch=2, state_batch=3, laser_batch=0)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_wrappers():
    gym = types.ModuleType("gym")

    class W:
        def __init__(self, env):
            self.env = env

        def __getattr__(self, n):
            return getattr(self.env, n)

        def reset(self, **kw):
            return self.env.reset(**kw)

        def step(self, a):
            return self.env.step(a)

    class OW(W):
        def reset(self, **kw):
            return self.observation(self.env.reset(**kw))

        def step(self, a):
            s, r, d, i = self.env.step(a)
            return self.observation(s), r, d, i
    gym.Wrapper, gym.ObservationWrapper = W, OW
    sys.modules["gym"] = gym
    envs = types.ModuleType("envs"); sys.modules["envs"] = envs
    state = _load("envs.state", os.path.join(REF, "envs/state/state.py"))
    action = _load("envs.action", os.path.join(REF, "envs/action/action.py"))
    utils = types.ModuleType("envs.utils"); utils.BagRecorder = object; sys.modules["envs.utils"] = utils
    base = _load("ref_wrapper_base", os.path.join(REF, "envs/wrapper/base.py"))
    filt = _load("ref_wrapper_filter", os.path.join(REF, "envs/wrapper/filter_states.py"))
    d = {n: getattr(base, n) for n in ORDER if hasattr(base, n)}
    d["ObsLaserStateTmp"] = filt.ObsLaserStateTmp
    return d, state.ImageState


def run_stack(wrappers, state_cls, cfg, seed, steps, list_actions):
    from fake_env import FakeEnv
    env = base = FakeEnv(state_cls, 4, seed, list_actions)
    for n in ORDER:
        env = wrappers[n](env, cfg)
    arng = np.random.default_rng(seed + 1)
    out = {}
    obs = env.reset()
    for k, o in enumerate(obs):
        out["r_obs%d" % k] = np.array(o)
    for t in range(steps):
        a = arng.integers(0, 4, 4) if cfg["discrete_action"] else np.stack([arng.uniform(-0.2, 0.9, 4), arng.uniform(-1.2, 1.2, 4)], 1)
        obs, rew, done, info = env.step(a)
        for k, o in enumerate(obs):
            out["s%d_obs%d" % (t, k)] = np.array(o)
        out["s%d_reward" % t] = np.array(rew, dtype=np.float64); out["s%d_done" % t] = np.array(done)
        for key in ("dones_info", "all_down", "is_clean", "speeds", "bool_get_close_to_human"):
            out["s%d_%s" % (t, key)] = np.array(info[key])
        out["s%d_actions" % t] = base.seen_actions[-1]
    return out


if __name__ == "__main__":
    wr, st = load_reference_wrappers()
    res = {}
    for disc in (True, False):
        o = run_stack(wr, st, wrapper_cfg(disc), 5 if disc else 6, 25, True)
        res.update({("d_" if disc else "c_") + k: v for k, v in o.items()})
    np.savez_compressed(os.path.join(HERE, "wrappers.npz"), **res)
    print("wrote wrappers.npz", len(res))
