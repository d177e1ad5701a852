"""Host logic of the Gym-style API (no GPU): the episode sampler reproduces the reference's EnvPos draw
for draw (golden vectors from the real reset_helper.py, tests/golden/make_reset_golden.py)."""
import os
import random

import numpy as np
import pytest
import yaml

from helpers import ROOT

CFG_DIR = os.path.join(ROOT, "tests", "golden", "cfg")


@pytest.mark.parametrize("name", ["test", "circle", "random", "10obs_5ped_baseline"])
def test_envpos_matches_reference_sampler(name):
    from img_env_b200.envs.reset_helper import EnvPos
    g = np.load(os.path.join(ROOT, "tests", "golden", "reset_helper.npz"))
    cfg = yaml.load(open(os.path.join(CFG_DIR, name + ".yaml")), Loader=yaml.FullLoader)
    for seed in (0, 1, 2):
        random.seed(seed)
        r = EnvPos(cfg).reset()
        key = "%s_%d" % (name, seed)
        obs, robots, peds = g[key + "_obs"], g[key + "_robots"], g[key + "_peds"]
        assert r["obs"].shape[0] == obs.shape[0]
        if obs.shape[0]:
            assert np.array_equal(r["obs"][:, 5:7], obs[:, 0:2]) and np.array_equal(r["obs"][:, 9:11], obs[:, 2:4])
            circ = r["obs"][:, 0] == 0
            assert np.array_equal(r["obs"][circ][:, 1:4], obs[circ][:, 4:7])
            assert np.array_equal(r["obs"][~circ][:, 1:5], obs[~circ][:, 4:8])
        assert np.array_equal(r["robots"][:, [0, 1, 4, 5, 6, 7]], robots)
        if peds.shape[0]:
            assert np.array_equal(r["peds"][:, [0, 1, 4, 5, 6, 7]], peds[:, :6])
            assert np.array_equal(r["traj_len"], peds[:, 6].astype(np.int32))
            assert np.array_equal(r["traj"][:, :, :2].reshape(-1, 4), peds[:, 7:11])


def test_state_and_actions_api():
    from img_env_b200.envs import ImageState, ContinuousAction, DiscreteActions
    s = ImageState(*[np.zeros((3, 2))] * 9)
    assert len(s) == 3 and "vector_states" in str(s)
    d = DiscreteActions([(0.2, 0.3), (0.4, -0.3, 1)])
    assert len(d) == 2 and d[1].reverse() == [0.4, -0.3, 1] and isinstance(d[0], ContinuousAction)
