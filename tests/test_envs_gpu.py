"""GPU: the Gym-style API (make_env / ImageEnv.reset / step) runs the reference's test.yaml scenario
through the CUDA library and returns the nine ImageState fields with the reference's shapes."""
import os
import random

import numpy as np
import pytest
import yaml

from helpers import ROOT, base_cfg

pytestmark = pytest.mark.gpu


def _cfg():
    sampler = yaml.load(open(os.path.join(ROOT, "tests", "golden", "cfg", "test.yaml")), Loader=yaml.FullLoader)
    cfg = base_cfg(R=1, P=4, scene="rvoscene", n_obj=4)
    cfg.update(sampler)
    cfg["discrete_action"] = True
    cfg["discrete_actions"] = [[0.0, -0.9], [0.2, 0.0], [0.6, 0.3], [0.4, 0.9]]
    cfg["wrapper"] = ["VelActionWrapper", "TimeLimitWrapper", "SensorsPaperRewardWrapper"]
    return cfg


def test_make_env_reset_step_shapes():
    import torch
    from img_env_b200.envs import make_env
    random.seed(3)
    env = make_env(_cfg(), num_scenes=3)
    s = env.reset()
    assert len(s) == 3
    assert tuple(s.sensor_maps.shape) == (3, 48, 48) and s.sensor_maps.dtype == torch.float16
    assert tuple(s.ped_maps.shape) == (3, 3, 48, 48) and tuple(s.lasers.shape) == (3, 1000)
    assert tuple(s.ped_vector_states.shape) == (3, 71) and float(s.ped_vector_states[0, 0]) == 4.0
    assert torch.all(s.step_ds == 0)
    d0 = torch.linalg.norm(s.vector_states[:, :2], dim=1)
    for t in range(5):
        s, r, done, info = env.step(torch.tensor([2, 1, 3]))
        assert r.shape == done.shape == (3,) and "dones_info" in info
    d1 = torch.linalg.norm(s.vector_states[:, :2], dim=1)
    assert torch.isfinite(s.lasers).all() and float(s.lasers.max()) <= 1.0 + 1e-6
    assert not torch.equal(d0, d1)
    env.close()


def test_numpy_state_has_reference_dtypes():
    from img_env_b200.envs import ImageEnv, ContinuousAction
    random.seed(4)
    env = ImageEnv(_cfg(), num_scenes=1, numpy_state=True)
    s = env.reset()
    assert s.vector_states.dtype == np.float64 and s.sensor_maps.dtype == np.float16 and s.is_collisions.dtype == np.int64
    assert s.is_arrives.dtype == bool and s.lasers.dtype == np.float64 and s.ped_maps.dtype == np.float32
    s, r, d, info = env.step([ContinuousAction(0.3, 0.1)])
    assert r.shape == (1,) and d.dtype == np.int64
    env.close()


def test_full_reference_wrapper_stack_with_auto_reset():
    """envs/cfg/test.yaml's wrapper list on 4 batched scenes: the RL loop never leaves the GPU except for actions."""
    import torch
    from img_env_b200.envs import make_env
    random.seed(5)
    cfg = _cfg()
    cfg.update(agent_num_per_env=1, image_batch=1, state_batch=3, laser_batch=0, time_max=6, continuous_actions=[[0, 0.6], [-0.9, 0.9]])
    cfg["wrapper"] = ["VelActionWrapper", "TimeLimitWrapper", "SensorsPaperRewardWrapper", "InfoLogWrapper", "MultiRobotCleanWrapper",
                      "TestEpisodeWrapper", "StateBatchWrapper", "ObsLaserStateTmp", "NeverStopWrapper"]
    env = make_env(cfg, num_scenes=4)
    obs = env.reset()
    lasers, vec, pm = obs
    assert tuple(lasers.shape) == (4, 1, 1000) and tuple(vec.shape) == (4, 9) and tuple(pm.shape) == (4, 3, 48, 48)
    resets = 0
    for t in range(20):
        obs, r, done, info = env.step(torch.tensor([2, 2, 1, 3]))
        assert r.shape == (4,) and r.dtype == torch.float64
        resets += int(info["all_down"].any())
        assert torch.isfinite(obs[0]).all()
    assert resets >= 2          # time_max=6 forces episodes to end and scenes to restart
    env.close()


def test_native_sampler_reset_equals_python_sampler_reset():
    """ImageEnv with the native EnvPos port (seeded s) == ImageEnv with the python EnvPos drawing from random.seed(s):
    same episodes, so the same first observation; partial re-resets keep the other scenes untouched."""
    import torch
    from img_env_b200.envs import ImageEnv
    cfg = _cfg()
    a = ImageEnv(dict(cfg, native_sampler=True, sampler_seed=11), num_scenes=1)
    b = ImageEnv(dict(cfg, native_sampler=False), num_scenes=1)
    assert a.sampler is not None and b.sampler is None
    random.seed(11)
    for _ in range(3):
        sa, sb = a.reset(), b.reset()
        for f in ("vector_states", "sensor_maps", "lasers", "ped_vector_states", "ped_maps", "is_collisions", "is_arrives"):
            assert torch.equal(getattr(sa, f), getattr(sb, f)), f
    a.close(); b.close()
    e = ImageEnv(dict(cfg, sampler_seed=5), num_scenes=4)
    s0 = e.reset()
    keep = s0.vector_states.clone()
    s1 = e.reset(scene_ids=[1, 3])
    assert torch.equal(s1.vector_states[[0, 2]], keep[[0, 2]]) and not torch.equal(s1.vector_states[[1, 3]], keep[[1, 3]])
    e.close()


def test_episode_record_matches_stepwise_state():
    """EpRes-style episode record (img_env.cpp:355-357, 397-408): the recorded poses / speeds of a scene equal the state read
    back after every step, restart at a reset of that scene only, and rows of robots that are not alive are flagged."""
    import torch
    from helpers import build_spec, make_reset, random_actions
    from img_env_b200.lib import BatchedSim
    spec = build_spec(base_cfg(R=2, P=3, scene="rvoscene", n_obj=2))
    rng = np.random.default_rng(3)
    sim = BatchedSim(spec, num_scenes=2, ped_yaw_mode=1)
    sim.record_enable(6)
    sim.reset([make_reset(spec, rng) for _ in range(2)])
    want_rb, want_pd, acts = [], [], []
    alive = np.ones((2, 2), np.uint8); alive[1, 1] = 0
    for t in range(8):                           # two more steps than the record holds
        a = np.stack([random_actions(2, rng) for _ in range(2)]).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda(), torch.from_numpy(alive).cuda())
        rb, pd, _ = sim.get_internal()
        want_rb.append(rb.copy()); want_pd.append(pd.copy()); acts.append(a)
    rec = sim.record_fetch(1)
    assert rec["robots"].shape == (6, 2, 6) and rec["peds"].shape == (6, 3, 5)
    for t in range(6):
        assert np.array_equal(rec["robots"][t, :, :3], want_rb[t][1][:, :3])
        assert np.array_equal(rec["robots"][t, :, 3:5], acts[t][1][:, :2].astype(np.float64))
        assert rec["robots"][t, :, 5].tolist() == [1.0, 0.0]
        assert np.array_equal(rec["peds"][t, :, :2], want_pd[t][1][:, :2]) and np.array_equal(rec["peds"][t, :, 3:5], want_pd[t][1][:, 6:8])
    sim.reset([make_reset(spec, rng)], scene_ids=[0])          # scene 0 restarts its record, scene 1 keeps it
    assert sim.record_fetch(0)["robots"].shape[0] == 0 and sim.record_fetch(1)["robots"].shape[0] == 6
    sim.step(torch.from_numpy(acts[0]).cuda())
    assert sim.record_fetch(0)["robots"].shape[0] == 1
    sim.record_enable(0)
    with pytest.raises(RuntimeError):
        sim.record_fetch(0)
    sim.close()


def test_state_is_not_aliased_across_steps():
    """a State kept by the caller must not change under it when the env steps again (the reference returns
    fresh numpy arrays); copy_state=False opts into zero-copy views of the library's output buffers."""
    import torch
    from img_env_b200.envs import ImageEnv
    random.seed(6)
    env = ImageEnv(_cfg(), num_scenes=2)
    s0 = env.reset()
    keep = {f: getattr(s0, f).clone() for f in s0.FIELDS}
    a = torch.tensor([[0.5, 0.3], [0.4, -0.5]], device="cuda")
    s1, *_ = env.step(a)
    s2, *_ = env.step(a)
    for f in s0.FIELDS:
        assert torch.equal(getattr(s0, f), keep[f]), "%s of the reset State changed after step()" % f
    assert not torch.equal(s1.ped_maps, s2.ped_maps) or not torch.equal(s1.lasers, s2.lasers)
    assert s1.ped_maps.data_ptr() != s2.ped_maps.data_ptr()
    env.close()
    env = ImageEnv(_cfg(), num_scenes=2, copy_state=False)
    s0 = env.reset(); s1, *_ = env.step(a)
    assert s0.ped_maps.data_ptr() == s1.ped_maps.data_ptr()       # documented: views of the bound output tensors
    env.close()


def test_ped_vector_wrapper_with_partial_auto_reset_normalises_once():
    """StatePedVectorWrapper under NeverStopWrapper's per-scene reset: rows of scenes that were not reset hold
    values normalised exactly once."""
    import torch
    from img_env_b200.envs import make_env
    from img_env_b200.envs.wrappers import StatePedVectorWrapper
    random.seed(7)
    cfg = _cfg()
    cfg["wrapper"] = ["StatePedVectorWrapper"]
    env = make_env(cfg, num_scenes=3)
    env.reset()
    a = torch.tensor([[0.3, 0.1]] * 3, device="cuda")
    s, *_ = env.step(a)
    raw = env.env.sim.out["ped_vector_states"].reshape(3, -1).clone()          # what the library wrote (un-normalised)
    s2 = env.reset(scene_ids=[1])                                               # partial reset re-observes scene 1 only
    raw_after = env.env.sim.out["ped_vector_states"].reshape(3, -1)
    assert torch.equal(raw_after[[0, 2]], raw[[0, 2]]), "library rows of untouched scenes must stay raw"
    k = (raw.shape[1] - 1) // 7
    avg = torch.tensor(StatePedVectorWrapper.avg, dtype=torch.float64, device="cuda"); std = torch.tensor(StatePedVectorWrapper.std, dtype=torch.float64, device="cuda")
    want = raw_after.clone()
    body = want[:, 1:].reshape(3, k, 7)
    idx = torch.arange(k, device="cuda")[None, :] < want[:, :1]
    want[:, 1:] = torch.where(idx[..., None], ((body.double() - avg) / std).float(), body).reshape(3, 7 * k)
    assert torch.allclose(s2.ped_vector_states, want)
    env.close()


def test_make_env_rejects_unknown_wrapper_and_reset_rejects_duplicate_scenes():
    from img_env_b200.envs import make_env, ImageEnv
    cfg = _cfg(); cfg["wrapper"] = ["VelActionWrapper", "NoSuchWrapper"]
    with pytest.raises(ValueError, match="NoSuchWrapper"):
        make_env(cfg, num_scenes=1)
    random.seed(8)
    cfg = _cfg(); cfg["device_autoreset"] = False          # (the device-mask path cannot list a scene twice)
    env = ImageEnv(cfg, num_scenes=3)
    env.reset()
    with pytest.raises(RuntimeError, match="duplicate scene id"):
        env.reset(scene_ids=[1, 1])
    env.close()


def test_dataset_scenes_through_the_gym_api():
    """ped_sim.type 'dataset': reset(cur_ped_pos_v_datas=...) as PedTrajectoryDatasetWrapper calls it (yaml_env.py:245-247)."""
    import torch
    from img_env_b200.envs import ImageEnv
    random.seed(9)
    cfg = _cfg()
    cfg["ped_sim"]["type"] = "dataset"; cfg["ped_sim"]["max_traj"] = 6
    env = ImageEnv(cfg, num_scenes=2, numpy_state=True)
    with pytest.raises(ValueError, match="cur_ped_pos_v_datas"):
        env.reset()
    P, T = cfg["ped_sim"]["total"], 6
    rng = np.random.default_rng(1)
    datas = np.zeros((P, T, 5))
    for p in range(P):
        pos = rng.uniform(3, 7, 2)
        for t in range(T):
            v = rng.uniform(-0.5, 0.5, 2)
            datas[p, t] = [pos[0], pos[1], np.arctan2(v[1], v[0]), v[0], v[1]]
            pos = pos + 0.4 * v
    env.reset(cur_ped_pos_v_datas=datas)
    for t in range(3):
        env.step(np.array([[0.2, 0.0], [0.3, 0.1]], np.float32))
        rb, pd, _ = env.sim.get_internal()
        assert np.allclose(pd[0][:, :2], datas[:, t, :2]) and np.allclose(pd[1][:, 6:8], datas[:, t, 3:5])
    env.close()
    del torch


def test_device_side_auto_reset_equals_host_driven_resets():
    """The episode-queue / device-mask reset path (imgenv_reset_masked, no host sync) gives every scene exactly the episodes and
    observations the host-driven per-scene reset path gives it: same wrapper stack, same seeds, two envs, bit-equal for 60 steps."""
    import torch
    from img_env_b200.envs import make_env
    outs = []
    for device_autoreset in (True, False):
        random.seed(5)
        cfg = _cfg()
        cfg.update(agent_num_per_env=1, image_batch=1, state_batch=3, laser_batch=0, time_max=5, continuous_actions=[[0, 0.6], [-0.9, 0.9]],
                   sampler_seed=1234, device_autoreset=device_autoreset)
        cfg["wrapper"] = ["VelActionWrapper", "TimeLimitWrapper", "SensorsPaperRewardWrapper", "InfoLogWrapper", "MultiRobotCleanWrapper",
                          "TestEpisodeWrapper", "StateBatchWrapper", "ObsLaserStateTmp", "NeverStopWrapper"]
        env = make_env(cfg, num_scenes=6)
        assert env.masked_reset == device_autoreset
        rec = [[x.clone() for x in env.reset()]]
        g = torch.Generator(device="cpu"); g.manual_seed(3)
        n_resets = 0
        for t in range(60):
            a = torch.randint(0, 4, (6,), generator=g).cuda()
            obs, r, done, info = env.step(a)
            rec.append([x.clone() for x in obs] + [r.clone(), done.clone()])
            n_resets += int(info["all_down"].sum())
        assert n_resets >= 20
        if device_autoreset:
            assert env.sim.debug_counters()[2] == 0, "an episode queue ran empty"
        outs.append(rec)
        env.close()
    for t, (a, b) in enumerate(zip(*outs)):
        for k, (x, y) in enumerate(zip(a, b)):
            assert torch.equal(x, y), "step %d, output %d differs between device-side and host-driven resets" % (t, k)


@pytest.mark.parametrize("scene", ["rvoscene", "pedscene"])
def test_graphed_step_equals_eager_step(scene):
    """GraphedStep (the whole wrapper stack + simulator + device-side auto-reset replayed as one CUDA graph) returns what the eager
    loop returns, step for step.  pedscene: the SFM quadtree update runs on its own stream and is normally joined at the NEXT
    call -- inside a capture it has to be joined before the call ends."""
    import torch
    from img_env_b200.envs import make_env, GraphedStep
    outs = []
    for graphed in (False, True):
        random.seed(6)
        cfg = _cfg()
        cfg.update(agent_num_per_env=1, image_batch=1, state_batch=3, laser_batch=0, time_max=5, continuous_actions=[[0, 0.6], [-0.9, 0.9]], sampler_seed=77)
        cfg["ped_sim"] = dict(cfg["ped_sim"], type=scene)
        cfg["wrapper"] = ["VelActionWrapper", "TimeLimitWrapper", "SensorsPaperRewardWrapper", "InfoLogWrapper", "MultiRobotCleanWrapper",
                          "TestEpisodeWrapper", "StateBatchWrapper", "ObsLaserStateTmp", "NeverStopWrapper"]
        env = make_env(cfg, num_scenes=5)
        env.reset()
        g = torch.Generator(device="cpu"); g.manual_seed(4)
        acts = [torch.randint(0, 4, (5,), generator=g).cuda() for _ in range(43)]
        rec = []
        step = env.step
        for t in range(3):                       # eager steps first: the wrappers create their per-row state lazily
            env.step(acts[t])
        if graphed:
            step = GraphedStep(env, acts[0], warmup=0).step
        for t in range(3, 43):
            obs, r, done, info = step(acts[t])
            rec.append([x.clone() for x in obs] + [r.clone(), done.clone(), info["all_down"].clone()])
        assert env.sim.debug_counters()[2] == 0
        outs.append(rec)
        env.close()
    for t, (a, b) in enumerate(zip(*outs)):
        for k, (x, y) in enumerate(zip(a, b)):
            assert torch.equal(x, y), "step %d, output %d differs between the graphed and the eager loop" % (t, k)
