"""The vectorised wrapper stack reproduces the reference's wrappers (golden vectors generated from the real
/root/reference/envs/wrapper/base.py by tests/golden/make_wrapper_golden.py), on numpy arrays and on torch
tensors, and handles several batched scenes with per-scene auto-reset."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_wrapper_golden import ORDER, wrapper_cfg  # noqa: E402
from fake_env import FakeEnv  # noqa: E402


def _stack(env, cfg):
    from img_env_b200.envs.wrappers import wrapper_dict
    for n in ORDER:
        env = wrapper_dict[n](env, cfg)
    return env


class TorchFake(FakeEnv):
    """FakeEnv handing out torch tensors like the CUDA env does (CPU tensors here)."""

    @staticmethod
    def _t(s):
        import torch
        for f in s.FIELDS:
            setattr(s, f, torch.from_numpy(np.ascontiguousarray(getattr(s, f))))
        return s

    def reset(self, **kw):
        return self._t(super().reset(**kw))

    def step(self, actions):
        import torch
        s, r, d, i = super().step(actions)
        return self._t(s), torch.from_numpy(r), torch.from_numpy(d), {"dones_info": torch.from_numpy(i["dones_info"])}


@pytest.mark.parametrize("discrete", [True, False])
@pytest.mark.parametrize("use_torch", [False, True])
def test_stack_matches_reference_golden(discrete, use_torch):
    from img_env_b200.envs import ImageState
    g = np.load(os.path.join(ROOT, "tests", "golden", "wrappers.npz"))
    pre = "d_" if discrete else "c_"
    cfg = wrapper_cfg(discrete)
    seed = 5 if discrete else 6
    base = (TorchFake if use_torch else FakeEnv)(ImageState, 4, seed, False)
    env = _stack(base, cfg)
    arng = np.random.default_rng(seed + 1)

    def np_(x):
        return x.numpy() if hasattr(x, "numpy") and not isinstance(x, np.ndarray) else np.asarray(x)
    obs = env.reset()
    for k, o in enumerate(obs):
        assert np.allclose(np_(o), g[pre + "r_obs%d" % k])
    for t in range(25):
        a = arng.integers(0, 4, 4) if discrete else np.stack([arng.uniform(-0.2, 0.9, 4), arng.uniform(-1.2, 1.2, 4)], 1)
        obs, rew, done, info = env.step(a)
        for k, o in enumerate(obs):
            assert np.allclose(np_(o), g[pre + "s%d_obs%d" % (t, k)]), (t, k)
        assert np.allclose(np_(rew), g[pre + "s%d_reward" % t]), t
        assert np.array_equal(np_(done), g[pre + "s%d_done" % t]), t
        for key in ("dones_info", "all_down", "is_clean", "bool_get_close_to_human"):
            assert np.array_equal(np_(info[key]).astype(np.int64), g[pre + "s%d_%s" % (t, key)].astype(np.int64)), (t, key)
        assert np.allclose(info["speeds"], g[pre + "s%d_speeds" % t]), t
        assert np.allclose(base.seen_actions[-1], g[pre + "s%d_actions" % t]), t


def test_per_scene_auto_reset_with_batched_scenes():
    """Two scenes in one env: a scene resets (frame stacks cleared, clean flags, timers) exactly when all of ITS robots are done."""
    from img_env_b200.envs import ImageState
    cfg = wrapper_cfg(True); cfg["robot"] = dict(total=2)

    class TwoScene(FakeEnv):
        def __init__(self):
            super().__init__(ImageState, 4, 3, False)
            self.resets = []

        def reset(self, **kw):
            self.resets.append(kw.get("scene_ids"))
            return super().reset(**kw)
    base = TwoScene()
    env = _stack(base, cfg)
    env.reset()
    seen_partial = False
    for t in range(60):
        obs, rew, done, info = env.step(np.zeros(4, np.int64))
        ad = np.asarray(info["all_down"]).reshape(2, 2)
        assert (ad[:, 0] == ad[:, 1]).all()
        if ad.any() and not ad.all():
            seen_partial = True
            s = int(np.argmax(ad[:, 0]))
            assert base.resets[-1] == [s]
            rows = slice(2 * s, 2 * s + 2)
            lasers, vec = obs[0], obs[1]
            assert np.all(vec[rows][:, :6] == 0) and not np.all(vec[rows][:, 6:] == 0)     # state_batch=3: two empty frames + the reset state
            other = slice(2 * (1 - s), 2 * (1 - s) + 2)
            assert not np.all(vec[other][:, :6] == 0)
    assert seen_partial
