"""Synthetic scenarios in the reference's yaml schema: cfg dicts, seeded ResetEnv requests and the bench workloads
(the shapes of BASELINE.json configs C1-C5).  Used by bench.py, the tests and smoke()."""
import math
import os

import numpy as np

from .spec import rpy_to_q

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAP_DIR = os.path.join(_ROOT, "tests", "golden", "maps")      # the reference's envs/map PNGs (fixtures)


def synthetic_map(px=110, blocks=True, seed=0):
    """room_10-like occupancy image: 0 = wall, 255 = free, px x px pixels at 0.1 m."""
    img = np.full((px, px), 255, np.uint8)
    img[:5, :] = 0; img[-5:, :] = 0; img[:, :5] = 0; img[:, -5:] = 0
    if blocks:
        rng = np.random.default_rng(seed)
        for _ in range(6):
            r, c = rng.integers(15, px - 25, 2)
            h, w = rng.integers(3, 10, 2)
            img[r:r + h, c:c + w] = 0
        img[px // 2 - 1: px // 2 + 1, 20:45] = 0
    return img


def base_cfg(R=1, P=0, scene="rvoscene", n_obj=4, map_px=110, ped_shape="leg", relation=1, state_dim=3,
             range_total=1000, max_ped=10, control_hz=0.4, robot_shape="circle", image=None, robot_type="diff"):
    """A cfg dict with the keys of /root/reference/envs/cfg/test.yaml that the hot path consumes."""
    rsize = [0, 0, 0.17] if robot_shape == "circle" else [-0.2, 0.2, -0.15, 0.15]
    psize = [0, 0.1, 0.1] if ped_shape == "leg" else [0, 0, 0.17]
    cfg = dict(
        test=False, env_type="robot_nav", robot_type=robot_type, env_num=1, env_id=0, env_name="test", cfg_name="test",
        cfg_type="yaml", control_hz=control_hz, time_max=100, robot_radius=0.17, ped_leg_radius=0.1, ped_safety_space=0.7,
        laser_max=6.0, image_batch=1, image_size=[48, 48], ped_image_size=[48, 48], state_batch=1, state_dim=state_dim,
        state_normalize=False, laser_batch=0, act_dim=2, circle_ranges=[1.8, 2.0], max_ped=max_ped, ped_vec_dim=7,
        ped_image_r=0.3, show_gui=False, sleep_t=0.0, window_height=500, show_image_height=125, is_draw_step=True,
        step_draw=3, use_laser=True, range_total=range_total, view_angle_begin=-1.570795, view_angle_end=1.570795,
        view_min_dist=0.0, view_max_dist=10.0, beep_r=1.0, ped_ca_p=1.0, relation_ped_robo=relation,
        global_map=dict(resolution=0.1, map_file="synthetic", image=image if image is not None else synthetic_map(map_px)),
        view_map=dict(resolution=0.015, width=6, height=6),
        robot=dict(total=R, shape=[robot_shape] * R, size=[list(rsize) for _ in range(R)]),
        object=dict(total=n_obj),
        ped_sim=dict(total=P, type=scene, max_speed=[0.5] * P, shape=[ped_shape] * P, size=[list(psize) for _ in range(P)],
                     go_back="yes"),
        target_min_dist=1.0, node_id=0, wrapper=[])
    return cfg


def make_reset(spec, rng, n_obj=None, lo=2.5, hi=8.5, robots_xy=None, peds_xy=None, ylo=None, yhi=None):
    """Seeded ResetEnv request in the array layout of include/imgenv.h / oracle/ref_driver.cpp."""
    R, P = spec["R"], spec["P"]
    n_obj = spec["max_obstacles"] if n_obj is None else n_obj
    ylo = lo if ylo is None else ylo
    yhi = hi if yhi is None else yhi
    obs = np.zeros((n_obj, 11))
    for k in range(n_obj):
        x, y, yaw = rng.uniform(lo, hi), rng.uniform(ylo, yhi), rng.uniform(-3.14, 3.14)
        if k % 2 == 0:
            obs[k, :5] = [0, 0, 0, 0.3, 0]
        else:
            obs[k, :5] = [1, -0.15, 0.15, -0.15, 0.15]
        obs[k, 5:7] = [x, y]
        obs[k, 7:11] = rpy_to_q(yaw)
    robots = np.zeros((R, 8))
    for j in range(R):
        x, y = (rng.uniform(lo, hi), rng.uniform(ylo, yhi)) if robots_xy is None else robots_xy[j]
        robots[j, :2] = [x, y]
        robots[j, 2:6] = rpy_to_q(rng.uniform(-3.14, 3.14))
        robots[j, 6:8] = [rng.uniform(lo, hi), rng.uniform(ylo, yhi)]
    peds = np.zeros((P, 8)); traj_len = np.zeros(P, np.int32); traj = np.zeros((P, 2, 3))
    for j in range(P):
        x, y = (rng.uniform(lo, hi), rng.uniform(ylo, yhi)) if peds_xy is None else peds_xy[j]
        peds[j, :2] = [x, y]
        peds[j, 2:6] = rpy_to_q(rng.uniform(-3.14, 3.14))
        peds[j, 6:8] = [rng.uniform(lo, hi), rng.uniform(ylo, yhi)]
        traj_len[j] = 2                                    # go_back: yes (reset_helper.py:337-342)
        traj[j, 0] = [peds[j, 6], peds[j, 7], 0]
        traj[j, 1] = [x, y, 0]
    return dict(obs=obs, robots=robots, peds=peds, traj_len=traj_len, traj=traj, ignore_obstacle=int(spec["ignore_obstacle"]))


def random_actions(R, rng, beep=False):
    a = np.zeros((R, 3), np.float32)
    a[:, 0] = rng.uniform(0, 0.6, R)
    a[:, 1] = rng.uniform(-0.9, 0.9, R)
    if beep:
        a[:, 2] = (rng.uniform(0, 1, R) < 0.5).astype(np.float32)
    return a


def make_dataset_reset(spec, rng, T=6, lo=3.0, hi=8.0):
    """ResetEnv request for trajectory-replay pedestrians (PedTrajectoryDatasetWrapper / EnvPos.init_ped_dataset,
    reset_helper.py:417-434): per pedestrian T positions (x, y, yaw) and velocities (vx, vy, 0)."""
    rs = make_reset(spec, rng, lo=lo, hi=hi)
    P = spec["P"]
    traj = np.zeros((P, T, 3)); trajv = np.zeros((P, T, 3))
    for p in range(P):
        pos = np.array([rng.uniform(lo, hi), rng.uniform(lo, hi)])
        for t in range(T):
            v = rng.uniform(-0.8, 0.8, 2)
            traj[p, t] = [pos[0], pos[1], math.atan2(v[1], v[0])]
            trajv[p, t] = [v[0], v[1], 0]
            pos = pos + 0.4 * v
        rs["peds"][p, :2] = traj[p, 0, :2]
        rs["peds"][p, 2:6] = rpy_to_q(traj[p, 0, 2])
    rs["traj"] = traj; rs["traj_v"] = trajv; rs["traj_len"] = np.full(P, T, np.int32)
    return rs


# SURVEY.md section 8(d) synthetic inputs; shapes of BASELINE.json configs[0..4].  `map` names a reference PNG
# (envs/map/*.png, committed as a fixture); when the file is absent the synthetic room of the same size is used.
WORKLOADS = {
    "c1": dict(desc="test.yaml shape: 1 robot, 4 reset objects, no pedestrians, room_10 733^2 grid", R=1, P=0, scene="rvoscene", n_obj=4,
               map="room_10.png", map_px=110, gres=0.1, lo=2.5, hi=8.5, max_ped=10, scenes=8192),
    "c2": dict(desc="8-robot circle crossing (image_circle_fix_8 shape), room_10 733^2 grid", R=8, P=0, scene="rvoscene", n_obj=0,
               map="room_10.png", map_px=110, gres=0.1, lo=3.5, hi=7.5, max_ped=10, scenes=1024),
    "c3": dict(desc="1 robot + 20 ORCA pedestrians (rvoscene), room_10 733^2 grid", R=1, P=20, scene="rvoscene", n_obj=4,
               map="room_10.png", map_px=110, gres=0.1, lo=2.5, hi=8.5, max_ped=20, scenes=8192),
    "c4": dict(desc="200 robots + 200 ervoscene pedestrians + 200 objects on one 7333^2 grid (image_big shape: room_10 at resolution 1)",
               R=200, P=200, scene="ervoscene", n_obj=200, map="room_10.png", map_px=110, gres=1.0, lo=25.0, hi=85.0, max_ped=200,
               scenes=128),
    "c5": dict(desc="64 robots + 64 pedscene (SFM) pedestrians per scene, room_16_empty 1066^2 grid", R=64, P=64, scene="pedscene",
               n_obj=0, map="room_16_empty.png", map_px=160, gres=0.1, lo=2.5, hi=13.5, max_ped=64, scenes=512),
}


def workload_map(w, synthetic=False):
    """-> (uint8 image, description).  synthetic=True: the round-1 room with random blocks (harder: robots sit in / next to walls)."""
    path = os.path.join(MAP_DIR, w.get("map", ""))
    if not synthetic and w.get("map") and os.path.exists(path):
        import cv2
        img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
        if img is not None:
            return img, w["map"]
    return synthetic_map(w["map_px"], blocks=True, seed=7), "synthetic room with blocks"


def make_cfg(w, synthetic_map_=False):
    img, _ = workload_map(w, synthetic_map_)
    cfg = base_cfg(R=w["R"], P=w["P"], scene=w["scene"], n_obj=w["n_obj"], max_ped=w["max_ped"], image=img)
    cfg["global_map"]["resolution"] = w["gres"]
    return cfg


def make_resets(spec, w, n, seed):
    rng = np.random.default_rng(seed)
    return [make_reset(spec, rng, n_obj=w["n_obj"], lo=w["lo"], hi=w["hi"]) for _ in range(n)]
