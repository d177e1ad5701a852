"""TEST INFRASTRUCTURE — not part of the product.

ctypes front-end of ``oracle/_ref/libimgenv_ref.so`` (the UNMODIFIED reference node,
see ref_driver.cpp) plus a restatement of the pure-Python post-processing the
reference applies to the node's reply:

* ``ImageEnv._get_states``          /root/reference/envs/env/yaml_env.py:446-481
* ``ImageEnv._draw_ped_map``        yaml_env.py:392-429
* ``ImageEnv._trans_cv2_sensor_map`` yaml_env.py:431-438 (``cv2.resize`` INTER_CUBIC, pinned to
  OpenCV's own code path with ``cv2.ipp.setUseIPP(False)``)
* ``ImageEnv._norm_lasers``          yaml_env.py:440-444
* ``NearbyPed``                      envs/utils/reset_helper.py:85-99

yaml_env.py itself cannot be imported here (rospy / gym / cv_bridge are absent), so those
functions are restated with the same numpy / python-float arithmetic.  Only tests/,
``__graft_entry__.smoke()`` and bench.py's reference / cpu_baseline legs may import this.
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def ref_lib_path():
    return os.path.join(_HERE, "_ref", "libimgenv_ref.so")


def port_lib_path():
    return os.path.join(_HERE, "liboracle_port.so")


def have_ref():
    return os.path.exists(ref_lib_path())


def have_port():
    return os.path.exists(port_lib_path())


class _Prefixed:
    """lib.ref_xxx / lib.port_xxx -> .xxx"""

    def __init__(self, lib, prefix):
        self._lib, self._prefix = lib, prefix

    def __getattr__(self, name):
        return getattr(self._lib, self._prefix + name)


def _lib(prefix="ref_"):
    if prefix not in _LIBS:
        lib = C.CDLL(ref_lib_path() if prefix == "ref_" else port_lib_path())
        getattr(lib, prefix + "create").restype = C.c_void_p
        for name in ("destroy", "get_states", "get_internal", "set_internal", "rvo_get",
                     "rvo_set", "rvo_get_obstacles", "sfm_get", "sfm_set_pv", "get_map", "get_jerk_limits"):
            try:
                getattr(lib, prefix + name).restype = None
            except AttributeError:      # accessor only one of the two libraries has
                pass
        _LIBS[prefix] = _Prefixed(lib, prefix)
    return _LIBS[prefix]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _dbl(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefEnv:
    """One scene of the reference node (one ROS node == one scene in the reference)."""

    PREFIX = "ref_"

    def __init__(self, spec):
        self.lib = _lib(self.PREFIX)
        self.spec = spec
        self.h = C.c_void_p(self.lib.create())
        self.R, self.P = spec["R"], spec["P"]
        sc = _dbl(spec["scalars"])
        grid = np.ascontiguousarray(spec["grid"], dtype=np.uint8)
        rd = _dbl(spec["robot_desc"]).reshape(self.R, 25)
        pd = _dbl(spec["ped_desc"]).reshape(self.P, 8) if self.P else np.zeros((0, 8))
        rc = self.lib.init(self.h, _p(sc, C.c_double), C.c_double(spec["global_resolution"]),
                               _p(grid, C.c_uint8), grid.shape[0], grid.shape[1], spec["raw_h"], spec["raw_w"],
                               self.R, _p(rd, C.c_double), spec["robot_ktype"].encode(),
                               self.P, _p(pd, C.c_double), spec["scene_type"].encode())
        if rc != 0:
            raise RuntimeError("ref_init failed: %d" % rc)
        self.H, self.W = grid.shape
        self.state_dim = int(spec["scalars"][4])
        self.range_total = int(spec["scalars"][12])

    def __del__(self):
        try:
            self.lib.destroy(self.h)
        except Exception:
            pass

    def reset(self, rs):
        """rs: dict(obs[n,11], robots[R,8], peds[P,8], traj_len[P], traj[P,T,3], ignore_obstacle)"""
        obs = _dbl(rs["obs"]).reshape(-1, 11)
        robots = _dbl(rs["robots"]).reshape(self.R, 8)
        peds = _dbl(rs["peds"]).reshape(self.P, 8) if self.P else np.zeros((0, 8))
        tl = np.ascontiguousarray(rs.get("traj_len", np.zeros(self.P)), dtype=np.int32)
        traj = rs.get("traj")
        flat = []
        for i in range(self.P):
            flat.append(np.asarray(traj[i][: tl[i]], dtype=np.float64).reshape(-1, 3))
        flat = _dbl(np.concatenate(flat, 0)) if flat else np.zeros((0, 3))
        flatv = None
        if rs.get("traj_v") is not None:
            flatv = _dbl(np.concatenate([np.asarray(rs["traj_v"][i][: tl[i]], dtype=np.float64).reshape(-1, 3) for i in range(self.P)], 0))
        rc = self.lib.reset(self.h, obs.shape[0], _p(obs, C.c_double), _p(robots, C.c_double), _p(peds, C.c_double),
                                _p(tl, C.c_int), _p(flat, C.c_double), _p(tl, C.c_int) if flatv is not None else None,
                                _p(flatv, C.c_double) if flatv is not None else None, int(rs.get("ignore_obstacle", 0)))
        if rc != 0:
            raise RuntimeError("ref_reset failed")
        return self.get_states()

    def step(self, actions, alive):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.R, 3)
        al = np.ascontiguousarray(alive, dtype=np.uint8).reshape(self.R)
        rc = self.lib.step(self.h, _p(a, C.c_float), _p(al, C.c_uint8))
        if rc != 0:
            raise RuntimeError("ref_step failed")
        return self.get_states()

    def get_states(self):
        R, P = self.R, self.P
        vs = self.lib.view_size(self.h)
        side = int(round(math.sqrt(vs)))
        nl = self.lib.laser_size(self.h)
        out = dict(view_map=np.zeros((R, vs), np.uint8), state=np.zeros((R, self.state_dim), np.float32),
                   laser=np.zeros((R, nl), np.float32), is_collision=np.zeros(R, np.int8),
                   is_arrive=np.zeros(R, np.uint8), pedinfo=np.zeros((R, P, 5), np.float32))
        self.lib.get_states(self.h, _p(out["view_map"], C.c_uint8), _p(out["state"], C.c_float),
                                _p(out["laser"], C.c_float), _p(out["is_collision"], C.c_int8),
                                _p(out["is_arrive"], C.c_uint8), _p(out["pedinfo"], C.c_float))
        vw = int(float(np.float32(self.spec["scalars"][1])) / float(np.float32(self.spec["scalars"][0])))   # agent.cpp:82
        out["view_map"] = out["view_map"].reshape(R, vs // max(vw, 1), -1) if vs else out["view_map"]
        del side
        return out

    def get_internal(self):
        rb = np.zeros((self.R, 16)); pd = np.zeros((self.P, 20))
        self.lib.get_internal(self.h, _p(rb, C.c_double), _p(pd, C.c_double))
        return rb, pd

    def set_internal(self, rb=None, pd=None):
        rb = _dbl(rb) if rb is not None else None
        pd = _dbl(pd) if pd is not None else None
        self.lib.set_internal(self.h, _p(rb, C.c_double), _p(pd, C.c_double))

    def jerk_limits(self):
        """[R,4] = (lin min_jerk, lin max_jerk, ang min_jerk, ang max_jerk) as the node's robots hold them (min_jerk is
        never assigned by SpeedLimiter(msg), speed_limit.cpp:56-65)."""
        out = np.zeros((self.R, 4))
        self.lib.get_jerk_limits(self.h, _p(out, C.c_double))
        return out

    def rvo_get(self):
        n = self.lib.rvo_num_agents(self.h)
        a = np.zeros((n, 4), np.float32)
        self.lib.rvo_get(self.h, _p(a, C.c_float))
        return a

    def rvo_set(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self.lib.rvo_set(self.h, _p(a, C.c_float))

    def rvo_obstacles(self):
        n = self.lib.rvo_num_obstacles(self.h)
        v = np.zeros((n, 8), np.float32)
        self.lib.rvo_get_obstacles(self.h, _p(v, C.c_float))
        return v

    def sfm_get(self):
        n = self.lib.sfm_num_agents(self.h)
        a = np.zeros((n, 12))
        self.lib.sfm_get(self.h, _p(a, C.c_double))
        return a

    def sfm_set_pv(self, a):
        a = _dbl(a)
        self.lib.sfm_set_pv(self.h, _p(a, C.c_double))

    def sfm_tree(self):
        na = self.lib.sfm_num_agents(self.h) if int(self.spec["scalars"][19]) == 1 else self.P
        nodes = np.zeros((1024, 5)); leaf = np.zeros((na, 4), np.int32); hash_ = np.zeros(na, np.int32)
        n = self.lib.sfm_tree2(self.h, _p(nodes, C.c_double), 1024, _p(leaf, C.c_int), _p(hash_, C.c_int))
        return nodes[:n].copy(), leaf, hash_

    def get_map(self, which):
        m = np.zeros((self.H, self.W), np.uint8)
        self.lib.get_map(self.h, which, _p(m, C.c_uint8))
        return m


class PortEnv(RefEnv):
    """Same interface on top of oracle/liboracle_port.so (our own CPU restatement, oracle/port/)."""
    PREFIX = "port_"

    def sfm_set(self, a):
        a = _dbl(a)
        self.lib.sfm_set(self.h, _p(a, C.c_double))


# ---------------------------------------------------------------------------------------------
# Python post-processing (yaml_env.py:392-481) restated
# ---------------------------------------------------------------------------------------------
def cubic_resize_u8(img, size):
    """cv2.resize(u8, size, INTER_CUBIC) on OpenCV's own (non-IPP) path, yaml_env.py:433-434."""
    import cv2
    if hasattr(cv2, "ipp"):
        cv2.ipp.setUseIPP(False)
    return cv2.resize(np.ascontiguousarray(img, dtype=np.uint8), (size[0], size[1]), interpolation=cv2.INTER_CUBIC)


class PyPost:
    """State carried by the Python ImageEnv between calls + `_get_states`."""

    def __init__(self, spec):
        self.R = spec["R"]
        self.image_size = tuple(spec["image_size"])
        self.ped_image_size = tuple(spec["ped_image_size"])
        self.resolution = 6.0 / self.ped_image_size[0]           # yaml_env.py:161
        self.max_ped = spec["max_ped"]
        self.ped_vec_dim = spec["ped_vec_dim"]
        self.ped_image_r = spec["ped_image_r"]
        self.laser_max = spec["laser_max"]
        self.laser_norm = spec["laser_norm"]
        self.robot_size_last = list(spec["robot_size_last"])    # init_req.env.robots[i].size[-1] (python floats)
        self.min_dist = [float("inf")] * self.R                   # NearbyPed, reset_helper.py:91-92
        self.tmp_distances = None

    def on_reset(self):
        self.tmp_distances = None                                  # yaml_env.py:225

    def _draw_ped_map(self, ped_tmp, pedinfo, robot_index):
        ped_tmp[0] = len(pedinfo)
        H, W = self.ped_image_size
        ped_image = np.zeros([3, H, W], dtype=np.float32)
        d = self.ped_vec_dim
        for j in range(int(ped_tmp[0])):
            px, py, vx, vy, r_ = (float(x) for x in pedinfo[j])   # rospy hands float32 fields over as python floats
            ped_tmp[j * d + 1] = px
            ped_tmp[j * d + 2] = py
            ped_tmp[j * d + 3] = vx
            ped_tmp[j * d + 4] = vy
            ped_r = round(r_, 2)
            ped_tmp[j * d + 5] = ped_r
            ped_tmp[j * d + 6] = ped_r + self.robot_size_last[robot_index]
            ped_tmp[j * d + 7] = math.sqrt(px ** 2 + py ** 2)
            if px > 3 or px < -3 or py > 3 or py < -3:
                continue
            tmx, tmy = -px + 3, -py + 3
            coor_tmx = (tmx - self.ped_image_r) // self.resolution, (tmx + self.ped_image_r) // self.resolution
            coor_tmy = (tmy - self.ped_image_r) // self.resolution, (tmy + self.ped_image_r) // self.resolution
            coor_tmx = list(map(int, coor_tmx))
            coor_tmy = list(map(int, coor_tmy))
            for jj in range(*coor_tmx):
                for kk in range(*coor_tmy):
                    if jj < 0 or jj >= H or kk < 0 or kk >= W:
                        continue
                    dx = (jj + 0.5) * self.resolution - tmx
                    dy = (kk + 0.5) * self.resolution - tmy
                    if dx ** 2 + dy ** 2 < self.ped_image_r ** 2:
                        ped_image[:, jj, kk] = 1.0, vx, vy
        return ped_image

    def get_states(self, st):
        """st: RefEnv.get_states() dict -> dict of the nine ImageState arrays (reference dtypes)."""
        R = self.R
        vec_states, sensor_maps, lasers, ped_infos, ped_maps, distances = [], [], [], [], [], []
        for i in range(R):
            pedinfo = [tuple(x) for x in st["pedinfo"][i]]
            pedinfo.sort(key=lambda x: float(x[0]) ** 2 + float(x[1]) ** 2)   # stable, python-float keys
            ped_tmp = np.zeros([self.max_ped * self.ped_vec_dim + 1], dtype=np.float32)
            ped_image = self._draw_ped_map(ped_tmp, pedinfo, i)
            if len(pedinfo) != 0:
                self.min_dist[i] = ped_tmp[7] - ped_tmp[6]
            ped_infos.append(ped_tmp)
            ped_maps.append(ped_image)
            state = [float(x) for x in st["state"][i]]
            vec_states.append(state)
            sensor_maps.append(cubic_resize_u8(st["view_map"][i], self.image_size).astype("float16") / 255.0)
            lasers.append([float(x) for x in st["laser"][i]])
            distances.append(math.sqrt(state[0] ** 2 + state[1] ** 2))
        step_ds = self.tmp_distances - np.array(distances) if self.tmp_distances is not None else np.zeros_like(distances)
        self.tmp_distances = np.array(distances)
        las = np.array(lasers) / self.laser_max if self.laser_norm else np.array(lasers)
        return dict(vector_states=np.array(vec_states), sensor_maps=np.array(sensor_maps),
                    is_collisions=np.array([int(x) for x in st["is_collision"]]),
                    is_arrives=np.array([bool(x) for x in st["is_arrive"]]),
                    lasers=las, ped_vector_states=np.array(ped_infos), ped_maps=np.array(ped_maps),
                    step_ds=np.asarray(step_ds, dtype=np.float64),
                    ped_min_dists=np.array([float(x) for x in self.min_dist]))
