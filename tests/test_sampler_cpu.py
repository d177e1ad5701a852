"""The native episode sampler (include/imgenv.h imgenv_sampler_*; reset_helper.py:115-345) on the host, no GPU:
its generator is CPython's `random` bit for bit, and its episodes are the ones the python EnvPos restatement
(itself pinned to the real reset_helper.py by tests/golden/reset_helper.npz) draws from the same seed."""
import copy
import os
import random

import numpy as np
import pytest
import yaml

from helpers import ROOT

CFG_DIR = os.path.join(ROOT, "tests", "golden", "cfg")


def _desc_min():
    from img_env_b200.envs.reset_helper import sampler_desc
    cfg = yaml.load(open(os.path.join(CFG_DIR, "test.yaml")), Loader=yaml.FullLoader)
    return sampler_desc(cfg)


@pytest.mark.parametrize("seed", [0, 1, 12345, 2 ** 32 - 1, 2 ** 32, 2 ** 61 + 12345])
def test_generator_is_cpython_random(seed):
    from img_env_b200.lib import NativeSampler
    s = NativeSampler(_desc_min(), num_scenes=2, seed=seed)
    s.seed(1, seed)                      # scene 0 is seeded with `seed`, scene 1 with seed+1 until re-seeded
    for scene in (0, 1):
        r = random.Random(seed)
        pick = random.Random(99)
        for _ in range(3000):
            k = pick.randrange(4)
            if k == 0:
                assert s.draw(scene, "random") == r.random()
            elif k == 1:
                a, b = pick.uniform(-5, 5), pick.uniform(-5, 5)
                assert s.draw(scene, "uniform", a, b) == r.uniform(a, b)
            elif k == 2:
                assert s.draw(scene, "gauss", 0.0, 0.5) == r.gauss(0.0, 0.5)
            else:
                hi = pick.randrange(0, 40)
                assert s.draw(scene, "randint", 0, hi) == r.randint(0, hi)
    s.close()


def _variants():
    base = yaml.load(open(os.path.join(CFG_DIR, "10obs_5ped_baseline.yaml")), Loader=yaml.FullLoader)
    out = {}
    for name in ["test", "circle", "random", "10obs_5ped_baseline"]:
        out[name] = yaml.load(open(os.path.join(CFG_DIR, name + ".yaml")), Loader=yaml.FullLoader)
    c = copy.deepcopy(base)             # go_back random + multi-range begin + view target
    c["ped_sim"]["go_back"] = "random"
    nr = c["robot"]["total"]
    c["robot"]["begin_poses_type"] = ["range_multi"] * nr
    c["robot"]["begin_poses"] = [[[1.0, 4.0, 1.0, 4.0], [5.0, 8.0, 5.0, 8.0], [1.0, 4.0, 5.0, 8.0]]] * nr
    c["robot"]["target_poses_type"] = ["range_view"] * nr
    c["robot"]["target_poses"] = [[0.5, 9.5, 0.5, 9.5]] * nr
    out["multi_view"] = c
    c = copy.deepcopy(base)             # crowded circle: exercises the 50-failure restart of all circle agents
    n_p = c["ped_sim"]["total"]
    c["circle_ranges"] = [1.2, 1.6]
    c["ped_sim"]["begin_poses_type"] = ["range_circle"] * n_p
    c["ped_sim"]["begin_poses"] = [[5.0, 5.0]] * n_p
    c["ped_sim"]["target_poses_type"] = ["range_circle"] * n_p
    c["ped_sim"]["target_poses"] = [[5.0, 5.0]] * n_p
    c["target_min_dist"] = 0.5
    out["crowded_circle"] = c
    c = copy.deepcopy(base)             # fixed / rand_angle / circle_fix mix with 6-value ranges
    nr = c["robot"]["total"]
    c["robot"]["begin_poses_type"] = ["range_circle_fix"] * nr
    c["robot"]["begin_poses"] = [[5.0, 5.0]] * nr
    c["robot"]["target_poses_type"] = ["circle_fix"] * nr
    c["robot"]["target_poses"] = [[5.0, 5.0]] * nr
    n_p = c["ped_sim"]["total"]
    c["ped_sim"]["begin_poses_type"] = ["range"] * n_p
    c["ped_sim"]["begin_poses"] = [[0.5, 9.5, 0.5, 9.5, -1.0, 1.0]] * n_p
    c["ped_sim"]["target_poses_type"] = (["fix", "rand_angle"] * n_p)[:n_p]
    c["ped_sim"]["target_poses"] = ([[2.0 + i, 9.0, 0.3] if i % 2 == 0 else [2.0 + i, 1.0, -1.0, 1.0] for i in range(n_p)])
    c["ped_sim"]["go_back"] = "no"
    out["fix_mix"] = c
    return out


VARIANTS = _variants()


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_native_sampler_matches_python_envpos(name):
    from img_env_b200.envs.reset_helper import EnvPos, sampler_desc
    from img_env_b200.lib import NativeSampler
    cfg = VARIANTS[name]
    n_obj = cfg.get("object", {"total": 0})["total"]
    for seed in (0, 7, 2 ** 40 + 3):
        s = NativeSampler(sampler_desc(cfg), num_scenes=1, seed=seed, max_obs=n_obj + 1, max_traj=3)
        random.seed(seed)
        env_pos = EnvPos(cfg)
        for _ in range(6):              # successive episodes continue the same stream
            want = env_pos.reset()
            got = s.sample([0])
            assert got["n_obs"][0] == want["obs"].shape[0]
            assert np.array_equal(got["obs"][0, :n_obj], want["obs"])
            assert np.array_equal(got["robots"][0], want["robots"])
            assert np.array_equal(got["peds"][0], want["peds"])
            assert np.array_equal(got["traj_len"][0], want["traj_len"])
            assert np.array_equal(got["traj"][0, :, :2], want["traj"])
        s.close()


def test_scene_streams_are_independent_and_reseedable():
    from img_env_b200.envs.reset_helper import sampler_desc
    from img_env_b200.lib import NativeSampler
    cfg = VARIANTS["10obs_5ped_baseline"]
    a = NativeSampler(sampler_desc(cfg), num_scenes=4, seed=100)
    first = a.sample([0, 1, 2, 3])
    assert not np.array_equal(first["robots"][0], first["robots"][1])
    b = NativeSampler(sampler_desc(cfg), num_scenes=1, seed=102)   # scene 2 of `a` == scene 0 of a sampler seeded 100+2
    assert np.array_equal(b.sample([0])["robots"][0], first["robots"][2])
    a.seed(3, 102)
    assert np.array_equal(a.sample([3])["robots"][0], first["robots"][2])


def test_rejects_configurations_the_reference_never_finishes():
    from img_env_b200.envs.reset_helper import sampler_desc
    from img_env_b200.lib import NativeSampler
    cfg = copy.deepcopy(VARIANTS["10obs_5ped_baseline"])
    nr = cfg["robot"]["total"]
    cfg["robot"]["begin_poses_type"] = ["fix"] * nr          # fixed start + sampled goal: reset_init is never cleared
    cfg["robot"]["begin_poses"] = [[1.0 + i, 1.0, 0.0] for i in range(nr)]
    with pytest.raises(ValueError, match="loops forever"):
        NativeSampler(sampler_desc(cfg), num_scenes=1)
    with pytest.raises(ValueError, match="descriptor size"):
        NativeSampler(sampler_desc(cfg)[:-1], num_scenes=1)


def test_batched_sampling_equals_scene_by_scene_sampling():
    """One generator per scene: a batched call in any scene order gives the episodes of scene-by-scene sampling."""
    from img_env_b200.envs.reset_helper import sampler_desc
    from img_env_b200.lib import NativeSampler
    cfg = VARIANTS["10obs_5ped_baseline"]
    a = NativeSampler(sampler_desc(cfg), num_scenes=300, seed=9)
    b = NativeSampler(sampler_desc(cfg), num_scenes=300, seed=9)
    ids = np.random.default_rng(0).permutation(300)[:257]
    all_at_once = a.sample(ids)
    for k, sc in enumerate(ids):
        one = b.sample([int(sc)])
        for f in ("n_obs", "obs", "robots", "peds", "traj_len", "traj"):
            assert np.array_equal(all_at_once[f][k], one[f][0]), (f, sc)
    with pytest.raises(ValueError, match="duplicate"):
        a.sample([3, 5, 3])
