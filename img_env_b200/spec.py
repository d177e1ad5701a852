"""cfg (the reference's yaml schema, envs/cfg/*.yaml) -> flat arrays for the C ABI.

Mirrors what the reference's Python does before talking to the node:
    ImageEnv._init_static_param / _init_req   /root/reference/envs/env/yaml_env.py:133-209
    EnvPos.init_robot / init_ped / _init_speed_limiter  envs/utils/reset_helper.py:347-412
and what the node does with the map file (GridMap::read_image, grid_map.cpp:28-38).
"""
import math
import os

import numpy as np

SHAPES = {"circle": 0, "rectangle": 1, "leg": 2}
SCENES = {"": 0, "pedscene": 1, "rvoscene": 2, "ervoscene": 3, "dataset": 4}
KTYPES = {"diff": 0, "omni": 1}


def rpy_to_q(yaw):
    """tf.transformations.quaternion_from_euler(0, 0, yaw) (ros_utils.py:21-22): (x, y, z, w)."""
    h = yaw / 2.0
    return (0.0, 0.0, math.sin(h), math.cos(h))


def load_grid(map_image, global_resolution, view_resolution):
    """grid_map.cpp:28-38: grayscale image -> cv::resize (INTER_LINEAR) to the view resolution.
    `map_image` is a file path or an already decoded HxW uint8 array. Returns (grid, raw_h, raw_w)."""
    import cv2
    if isinstance(map_image, str):
        img = cv2.imread(map_image, cv2.IMREAD_GRAYSCALE)
        if img is None:
            raise FileNotFoundError(map_image)
    else:
        img = np.ascontiguousarray(map_image, dtype=np.uint8)
    gr = float(np.float32(global_resolution))      # Env.msg global_resolution is float32
    vr = float(np.float32(view_resolution))        # InitEnv.srv view_resolution is float32
    w = int(img.shape[1] * gr / vr)
    h = int(img.shape[0] * gr / vr)
    grid = cv2.resize(img, (w, h))
    return np.ascontiguousarray(grid), img.shape[0], img.shape[1]


def _limiter(cfg, key, dmin_v, dmax_v):
    out = [0.0] * 9          # comn_pkg/SpeedLimiter defaults: all false / 0
    c = cfg.get(key)
    if c:
        out = [float(bool(c.get("has_velocity_limits", False))), float(bool(c.get("has_acceleration_limits", False))),
               float(bool(c.get("has_jerk_limits", False))), c.get("min_velocity", dmin_v), c.get("max_velocity", dmax_v),
               c.get("min_acceleration", -2), c.get("max_acceleration", 2), c.get("min_jerk", -2), c.get("max_jerk", 2)]
    return [float(x) for x in out]


def build_spec(cfg, map_dir=None, opt_in_beep=False):
    """Returns the dict of arrays both the product (BatchedSim) and the test oracle consume."""
    R = int(cfg["robot"]["total"])
    P = int(cfg["ped_sim"]["total"])
    vm = cfg["view_map"]
    gm = cfg["global_map"]
    image = gm.get("image")
    if image is None:
        path = gm["map_file"]
        if not os.path.isabs(path):
            path = os.path.join(map_dir or ".", path)
        image = path
    grid, raw_h, raw_w = load_grid(image, gm["resolution"], vm["resolution"])
    # InitEnv.srv scalars in field order (yaml_env.py:183-200). beep_r / ped_ca_p are never filled by the
    # reference's Python (always 0.0 on the wire); opt_in_beep=True passes the yaml values instead.
    scalars = [vm["resolution"], vm["width"], vm["height"], cfg["control_hz"], cfg["state_dim"], 0, 0,
               cfg.get("window_height", 500), cfg.get("show_image_height", 125), float(bool(cfg.get("is_draw_step", False))),
               cfg.get("step_draw", 3), float(bool(cfg["use_laser"])), cfg["range_total"], cfg["view_angle_begin"],
               cfg["view_angle_end"], cfg["view_min_dist"], cfg["view_max_dist"],
               cfg.get("beep_r", 0.0) if opt_in_beep else 0.0, cfg.get("ped_ca_p", 0.0) if opt_in_beep else 0.0,
               cfg["relation_ped_robo"]]
    lim_v = _limiter(cfg, "speed_limiter_v", 0, 0.6)
    lim_w = _limiter(cfg, "speed_limiter_w", -0.9, 0.9)
    robot_desc = np.zeros((R, 25))
    size_last = []
    for j in range(R):
        shape = cfg["robot"]["shape"][j]
        size = list(cfg["robot"]["size"][j])
        if shape not in ("circle", "rectangle"):
            raise ValueError("robot shape %r: the node only knows circle/rectangle (agent.cpp:64-77)" % shape)
        sens = cfg["robot"]["sensor_cfgs"][j] if cfg["robot"].get("sensor_cfgs") else [0.0, 0.0]
        robot_desc[j, 0] = SHAPES[shape]
        robot_desc[j, 1:1 + len(size)] = size
        robot_desc[j, 5:7] = sens
        robot_desc[j, 7:16] = lim_v
        robot_desc[j, 16:25] = lim_w
        size_last.append(float(size[-1]))
    ped_desc = np.zeros((P, 8))
    for j in range(P):
        shape = cfg["ped_sim"]["shape"][j]
        size = list(cfg["ped_sim"]["size"][j])
        if shape == "leg":                                         # reset_helper.py:399-403
            size = size + [size[0], -size[1], size[2]]
        ped_desc[j, 0] = SHAPES[shape]
        ped_desc[j, 1:1 + len(size)] = size
        ped_desc[j, 7] = cfg["ped_sim"]["max_speed"][j]
    return dict(R=R, P=P, scalars=np.array(scalars, dtype=np.float64), grid=grid, raw_h=raw_h, raw_w=raw_w,
                global_resolution=float(gm["resolution"]), robot_desc=robot_desc, robot_ktype=cfg["robot_type"],
                ped_desc=ped_desc, scene_type=cfg["ped_sim"]["type"] if P > 0 else "",
                image_size=tuple(cfg["image_size"]), ped_image_size=tuple(cfg["ped_image_size"]), max_ped=int(cfg["max_ped"]),
                ped_vec_dim=int(cfg["ped_vec_dim"]), ped_image_r=float(cfg["ped_image_r"]), laser_max=float(cfg["laser_max"]),
                laser_norm=bool(cfg.get("laser_norm", True)), robot_size_last=size_last,
                max_obstacles=int(cfg.get("object", {}).get("total", 0)), max_traj=int(cfg["ped_sim"].get("max_traj", 2)),
                ignore_obstacle=bool(cfg["ped_sim"].get("ignore_obstacle", False)))
