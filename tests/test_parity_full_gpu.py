"""GPU parity at the sizes and on the inputs the bench measures, lock-step against the UNMODIFIED
reference node (oracle/_ref):

* BASELINE config C4 at FULL size (200 robots + 200 ervoscene pedestrians + 200 objects, 7333^2 grid), exactly
  bench.py's `make_cfg(WORKLOADS['c4'])` / `make_resets`: codes, arrivals, every float16 pixel, the 400x400 rasters,
  pedestrian state; on the reference's room_10.png and on the round-1 synthetic room with blocks (robots inside walls);
* the reference's real maps and yaml: envs/map/room_10.png + envs/cfg/test.yaml through ImageEnv (GridMap::read_image,
  grid_map.cpp:28-38), room_16_empty.png with SFM pedestrians;
* branches the round-1 suite never ran: has_jerk_limits (speed_limit.cpp:153-173), two robot types in one scene,
  relation_ped_robo = 0 with SFM and with ORCA.
"""
import os
import random

import numpy as np
import pytest
import yaml

from helpers import ROOT, base_cfg, build_spec, compare_state, random_actions
from img_env_b200.scenarios import MAP_DIR, WORKLOADS, make_cfg, make_resets
from test_parity_gpu import run_lockstep

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("synthetic", [False, True], ids=["room_10", "synthetic_blocks"])
def test_c4_full_size_lockstep(synthetic):
    w = WORKLOADS["c4"]
    cfg = make_cfg(w, synthetic)
    # 1 scene x 2 steps: the node needs ~7 s and ~12 GB per C4 scene-step (200 private 53.8 MB map clones)
    run_lockstep(cfg, seed=41, steps=2, resets=lambda spec, rng: make_resets(spec, w, 1, 1234))


def test_c5_shape_room_16_empty_sfm():
    # C5's map and solver at a size the node can run (it overflows its stack with >= 9 SFM robots)
    cfg = base_cfg(R=8, P=24, scene="pedscene", n_obj=0, max_ped=24)
    cfg["global_map"] = dict(resolution=0.1, map_file="room_16_empty.png")
    run_lockstep(cfg, seed=42, steps=4, lo=2.5, hi=13.5, map_dir=MAP_DIR)


def test_room_10_png_lockstep():
    cfg = base_cfg(R=3, P=6, scene="rvoscene", n_obj=4)
    cfg["global_map"] = dict(resolution=0.1, map_file="room_10.png")
    run_lockstep(cfg, seed=43, steps=5, lo=1.2, hi=9.8, map_dir=MAP_DIR)       # poses reach the walls at 1 m / 10 m


def test_test_yaml_on_room_10_through_image_env():
    """envs/cfg/test.yaml (all keys) + envs/map/room_10.png through the Gym API against the node, free running."""
    import torch
    from img_env_b200.envs import ImageEnv, ContinuousAction
    from oracle.pyref import RefEnv, PyPost, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "cfg", "test_full.yaml")))
    cfg["native_sampler"] = False          # the reference's own python EnvPos draws the episode
    random.seed(17)
    env = ImageEnv(cfg, num_scenes=1, numpy_state=True, map_dir=MAP_DIR)
    assert env.spec["grid"].shape == (733, 733)
    log = []
    orig = env.env_pose[0].reset
    env.env_pose[0].reset = lambda: (log.append(orig()), log[-1])[1]
    spec = env.spec
    R = spec["R"]
    ref = RefEnv(spec); post = PyPost(spec)
    rng = np.random.default_rng(3)
    errs = []

    def as_dict(s):
        return dict(vector_states=s.vector_states, sensor_maps=s.sensor_maps, is_collisions=s.is_collisions, is_arrives=s.is_arrives,
                    lasers=s.lasers, ped_vector_states=s.ped_vector_states, ped_maps=s.ped_maps, step_ds=s.step_ds, ped_min_dists=s.ped_min_dists)
    for ep in range(2):
        s = env.reset()
        post.on_reset()
        want = post.get_states(ref.reset(log[-1]))
        errs += compare_state(as_dict(s), want, spec, where="episode %d reset: " % ep)
        dones = np.zeros(R, np.int64)
        for t in range(4):
            acts = random_actions(R, rng)
            alive = (dones == 0).astype(np.uint8)
            s, rew, d, info = env.step([ContinuousAction(float(a[0]), float(a[1])) for a in acts])
            want = post.get_states(ref.step(acts * alive[:, None], alive))
            errs += compare_state(as_dict(s), want, spec, where="episode %d step %d: " % (ep, t))
            dones = np.clip(np.clip(want["is_collisions"], -1, 1) + want["is_arrives"], 0, 1)
            assert np.array_equal(d, dones)
    assert not errs, "\n".join(errs[:20])
    env.close()
    del torch


def _jerk_cfg():
    cfg = base_cfg(R=3, P=0, n_obj=2, control_hz=0.25)
    cfg["speed_limiter_v"] = dict(has_velocity_limits=True, has_acceleration_limits=True, has_jerk_limits=True, min_velocity=0,
                                  max_velocity=0.55, min_acceleration=-1.6, max_acceleration=1.2, min_jerk=0.9, max_jerk=2.0)
    # (no jerk limit on w: with the node's unassigned min_jerk ~ 1e-310 a negative-going w request is clamped to a denormal,
    #  and cmd's arc branch then divides by it -> the node itself returns NaN poses)
    cfg["speed_limiter_w"] = dict(has_velocity_limits=True, has_acceleration_limits=True, has_jerk_limits=False, min_velocity=-0.8,
                                  max_velocity=0.8, min_acceleration=-0.6, max_acceleration=2, min_jerk=1.5, max_jerk=3.0)
    return cfg


def test_jerk_limiter():
    """limit_jerk (speed_limit.cpp:153-173) incl. the constructor quirk: max_jerk := msg.min_jerk, min_jerk is never assigned
    (speed_limit.cpp:56-65).  The product is given the value the node's unassigned member actually holds."""
    import torch
    from img_env_b200.lib import BatchedSim
    from img_env_b200.scenarios import make_reset
    from oracle.pyref import RefEnv, PyPost, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    spec = build_spec(_jerk_cfg())
    R = spec["R"]
    rng = np.random.default_rng(51)
    sim = BatchedSim(spec, 1, ped_yaw_mode=1); ref = RefEnv(spec); post = PyPost(spec)
    jl = ref.jerk_limits()
    assert np.allclose(jl[:, 1], 0.9), "max_jerk must hold msg.min_jerk (ctor quirk)"
    sim.debug_set_min_jerk(jl[:, [0, 2]])
    rs = make_reset(spec, rng, lo=3.5, hi=7.5)
    out = sim.reset([rs]); post.on_reset(); post.get_states(ref.reset(rs))
    errs = []
    for t in range(10):      # free running: the limiter state (last two commands) must stay equal over several steps
        acts = random_actions(R, rng)
        rbr = ref.get_internal()[0]
        rbr[:, 12] = 0; rbr[:, 13] = 0                      # keep every robot alive so the limiter runs every step
        ref.set_internal(rbr, None)
        rb = rbr.copy(); rb[:, 15] = post.tmp_distances if post.tmp_distances is not None else np.nan
        sim.set_internal(rb[None], None, None)
        out = sim.step(torch.from_numpy(acts[None]).cuda(), torch.ones(1, R, dtype=torch.uint8, device="cuda"))
        torch.cuda.synchronize()
        want = post.get_states(ref.step(acts, np.ones(R)))
        errs += compare_state({k: v[0].cpu().numpy() for k, v in out.items()}, want, spec, where="step %d: " % t)
        got_rb = sim.get_internal()[0][0]
        if not np.allclose(got_rb[:, :12], ref.get_internal()[0][:, :12], rtol=1e-9, atol=1e-12):
            errs.append("step %d: limiter / pose state differs\n%s\n%s" % (t, got_rb[:, 6:10], ref.get_internal()[0][:, 6:10]))
    assert not errs, "\n".join(errs[:10])
    sim.close()


def test_two_robot_types_in_one_scene():
    cfg = base_cfg(R=4, P=6, scene="rvoscene", n_obj=3)
    cfg["robot"]["shape"] = ["circle", "rectangle", "circle", "rectangle"]
    cfg["robot"]["size"] = [[0, 0, 0.17], [-0.2, 0.2, -0.15, 0.15], [0, 0, 0.3], [-0.1, 0.25, -0.12, 0.12]]
    cfg["robot"]["sensor_cfgs"] = [[0.0, 0.0], [0.14, 0.0], [0.0, 0.0], [0.1, 0.05]]
    run_lockstep(cfg, seed=44, steps=6, lo=3.5, hi=7.0)


def test_sfm_relation_zero():
    # relation_ped_robo = 0: robots are not solver agents (img_env.cpp:411-417), pedestrians ignore them
    run_lockstep(base_cfg(R=3, P=8, scene="pedscene", n_obj=2, relation=0), seed=45, steps=6, lo=3.0, hi=8.0)


def test_orca_relation_zero():
    run_lockstep(base_cfg(R=3, P=10, scene="rvoscene", n_obj=3, relation=0), seed=46, steps=6, lo=3.5, hi=7.0)


def test_orca_dense_obstacle_field_no_truncation():
    """Many reset objects around slow pedestrians: every agent sees dozens of obstacle edges within (5*maxSpeed + 0.5) m.
    The reference keeps ALL of them (Agent.cpp:820-838); the solver must not truncate."""
    cfg = base_cfg(R=2, P=16, scene="rvoscene", n_obj=40, max_ped=16)
    run_lockstep(cfg, seed=47, steps=5, lo=3.5, hi=7.5)


@pytest.mark.parametrize("n_obj,seed", [(4, 61), (40, 62), (200, 63)])
def test_device_built_rvo_obstacle_tree(n_obj, seed):
    """The RVO vertex ring + BSP built by the reset kernel (rvotree.cuh) == the host restatement of RVOSimulator::addObstacle /
    KdTree::buildObstacleTreeRecursive bit for bit (split vertices, pre-order node ids, links), and its vertex ring == the
    reference node's own obstacles_ after processObstacles."""
    from img_env_b200.lib import BatchedSim
    from img_env_b200.scenarios import make_reset
    from oracle.pyref import RefEnv, have_ref
    cfg = base_cfg(R=1, P=2, scene="rvoscene", n_obj=n_obj, map_px=110 if n_obj < 100 else 400)
    spec = build_spec(cfg)
    rng = np.random.default_rng(seed)
    sim = BatchedSim(spec, 2, ped_yaw_mode=1)
    hi = 8.5 if n_obj < 100 else 30.0
    resets = [make_reset(spec, rng, lo=2.5, hi=hi), make_reset(spec, rng, lo=2.5, hi=hi)]
    sim.reset(resets)
    for sc in range(2):
        dev = sim.debug_rvo_tree(sc)
        host = sim.host_rvo_tree(dev["corners"], n_obj)
        assert dev["n"] == host["n"] and dev["root"] == host["root"] and dev["n"] >= 4 * n_obj
        assert np.array_equal(dev["verts"].view(np.uint32), host["verts"].view(np.uint32)), "vertex ring differs from the host restatement"
        assert np.array_equal(dev["nodes"], host["nodes"]), "BSP differs from the host restatement"
        if have_ref() and n_obj <= 40:
            ref = RefEnv(spec); ref.reset(resets[sc])
            rv = ref.rvo_obstacles()
            assert rv.shape[0] == dev["n"] and np.array_equal(rv.view(np.uint32), dev["verts"].view(np.uint32)), "vertex ring differs from the node's obstacles_"
    assert sim.debug_counters()[1] == 0
    sim.close()


def test_ped_order_ties():
    """Nearest-first order of ped_vector_states / ped_maps (yaml_env.py:446-466: a stable python sort on float64 keys) when keys
    collide: k_ped_obs sorts float32(key) << 32 | index and re-sorts runs of equal float32 keys by (float64 key, index).
    Robot at (4, 4) with yaw 0, so pedestrian offsets are exact in the robot frame:
    peds 0/1: equal float32 keys, different float64 keys, the farther one has the lower index (the run must be re-sorted);
    peds 2/3: exactly equal float64 keys (mirror images): index order; peds 4/5: the same nearer to the robot."""
    import torch
    from img_env_b200.lib import BatchedSim
    from img_env_b200.spec import rpy_to_q
    from helpers import make_reset
    from oracle.pyref import RefEnv, PyPost, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    cfg = base_cfg(R=1, P=6, scene="rvoscene", n_obj=0, max_ped=6)
    spec = build_spec(cfg)
    rng = np.random.default_rng(5)
    e = 2.0 ** -6
    offs = [(e * (1 + 2.0 ** -23), 1.0), (e, 1.0), (1.5, 0.75), (1.5, -0.75), (0.5, -0.5), (0.5, 0.5)]
    rs = make_reset(spec, rng, n_obj=0, robots_xy=[(4.0, 4.0)], peds_xy=[(4.0 + a, 4.0 + b) for a, b in offs])
    rs["robots"][0, 2:6] = rpy_to_q(0.0)
    sim = BatchedSim(spec, num_scenes=1, ped_yaw_mode=1)
    out = sim.reset([rs])
    torch.cuda.synchronize()
    ref = RefEnv(spec); post = PyPost(spec); post.on_reset()
    want = post.get_states(ref.reset(rs))
    got = {k: v[0].cpu().numpy() for k, v in out.items()}
    pv = want["ped_vector_states"].reshape(-1)[1:].reshape(6, 7)
    keys32 = (pv[:, 0].astype(np.float64) ** 2 + pv[:, 1].astype(np.float64) ** 2).astype(np.float32)
    assert (np.diff(keys32) == 0).sum() >= 3, "the scenario must produce runs of equal float32 keys"
    assert abs(pv[2, 0] - e) < 1e-9 and pv[3, 0] > pv[2, 0], "the node orders the float32-equal pair by its float64 keys"
    errs = compare_state(got, want, spec, where="ties: ")
    assert not errs, "\n".join(errs)
    assert np.array_equal(got["ped_vector_states"], want["ped_vector_states"].astype(np.float32).reshape(got["ped_vector_states"].shape))
    sim.close()


@pytest.mark.parametrize("shape,size", [("rectangle", [-0.45, 0.45, -0.3, 0.3]), ("circle", [0, 0, 0.5]), ("rectangle", [-0.2, 0.2, -0.15, 0.15])])
def test_big_robots_against_walls_and_greys(shape, size):
    """Collision codes of robots whose footprint box spans three 32-cell blocks, alone next to walls / grey cells: the
    observation kernel skips the collision lattice when nothing is under the robot's own box, so EVERY block under that box
    has to be looked at (a fuzz case found the 2 x 2 corner check of an earlier version wanting)."""
    from test_parity_gpu import _variant
    cfg = _variant(R=4, P=0, scene="rvoscene", n_obj=0, robot_shape=shape, grey_map=True, view=(0.02, 4.0) if size[-1] < 0.3 else (0.015, 6.0))
    cfg["robot"]["size"] = [list(size) for _ in range(4)]
    # poses spread over the whole room incl. the 0.5 m wall band (robots inside / touching walls collide with code 3)
    run_lockstep(cfg, seed=71, steps=6, S=3, lo=0.3, hi=10.7)
