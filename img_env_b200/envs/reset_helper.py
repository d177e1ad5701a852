"""Host-side episode sampling: start/goal poses, reset objects, pedestrian waypoints.
Follows the reference's EnvPos (envs/utils/reset_helper.py:101-345) decision by decision — same
pose types ('fix', 'rand_angle', 'range', 'range_circle[_fix]', 'range_multi', 'range_view',
'circle_fix'), same rejection tests (free_check_robo_ped d=1.0, free_check_obj, target_min_dist,
50-failure restarts) and the same python `random` draws — but emits the flat arrays of
include/imgenv.h instead of ROS Agent messages."""
import math
import random

import numpy as np

from ..spec import rpy_to_q


def _module_size(size, shape):           # reset_helper.py:167-186
    if shape == "circle":
        return size[2]
    if shape == "rectangle":
        return math.sqrt(size[0] ** 2 + size[2] ** 2)
    if shape == "leg":
        return size[-1] + size[-2]
    if shape == "L":
        return math.sqrt(size[1] ** 2 + size[3] ** 2)
    if shape == "sweep":
        return size[3] + size[1]
    raise ValueError(shape)


def _free_robo_ped(x, y, poses, d=1.0):  # reset_helper.py:35-44
    for pose in poses:
        if pose is None:
            continue
        if math.sqrt((x - pose[0]) * (x - pose[0]) + (y - pose[1]) * (y - pose[1])) <= d:
            return False
    return True


def _free_obj(target, objs):             # reset_helper.py:47-57
    for pose in objs:
        if pose[-1] == 0.0:
            continue
        if math.sqrt((target[0] - pose[0]) ** 2 + (target[1] - pose[1]) ** 2) <= target[-1] + pose[-1]:
            return False
    return True


def _random_pose(x, y, sita):
    return [random.uniform(x[0], x[1]), random.uniform(y[0], y[1]), random.uniform(sita[0], sita[1])]


def _random_view(init_pose, pose_range):  # reset_helper.py:64-82
    tv = [2.5, 4.0, 2.5, 4.0]
    while True:
        p = _random_pose([init_pose[0] - tv[1], init_pose[0] + tv[1]], [init_pose[1] - tv[3], init_pose[1] + tv[3]], [-3.14, 3.14])
        if init_pose[0] - tv[0] <= p[0] <= init_pose[0] + tv[0] and init_pose[1] - tv[2] <= p[1] <= init_pose[1] + tv[2]:
            continue
        if pose_range[0] <= p[0] <= pose_range[1] and pose_range[2] <= p[1] <= pose_range[3]:
            return p


class NearbyPed:                          # reset_helper.py:85-99 (kept for API compatibility; the library owns the values)
    def __init__(self, robots):
        self.min_dist = [float("inf")] * robots

    def set(self, index, value):
        self.min_dist[index] = value

    def get(self):
        return self.min_dist


class EnvPos:
    def __init__(self, cfg):
        self.cfg = cfg

    def reset(self):
        """-> dict(obs[n,11], robots[R,8], peds[P,8], traj_len[P], traj[P,2,3])"""
        obs = self.reset_obs()
        while True:
            out = self._reset_robot_ped()
            if out is not None:
                break
        out["obs"] = obs
        out["ignore_obstacle"] = int(bool(self.cfg["ped_sim"].get("ignore_obstacle", False)))
        return out

    def reset_obs(self):                  # reset_helper.py:122-165
        c = self.cfg.get("object", {"total": 0})
        self.obs_range = []
        rows = []
        for i in range(c["total"]):
            pose_range, size_range = c["poses"][i], c["size_range"][i]
            pose_type, shape = c["poses_type"][i], c["shape"][i]
            if shape == "circle":
                radius = random.uniform(size_range[0], size_range[1])
            else:
                radius = math.sqrt(size_range[0] ** 2 + size_range[2] ** 2)
            if pose_type == "fix":
                self.obs_range.append(list(pose_range) + ([0, radius] if len(pose_range) == 2 else [radius]))
            else:
                if len(pose_range) == 4:
                    p = _random_pose(pose_range[:2], pose_range[2:4], [-3.14, 3.14])
                else:
                    p = _random_pose(pose_range[:2], pose_range[2:4], pose_range[4:6])
                self.obs_range.append(p + [radius])
            o = self.obs_range[i]
            size = [0, 0, o[-1], 0] if shape == "circle" else list(size_range[:4])
            rows.append([0 if shape == "circle" else 1] + size + [o[0], o[1]] + list(rpy_to_q(o[2])))
        return np.array(rows, dtype=np.float64).reshape(-1, 11)

    def _reset_robot_ped(self):           # reset_helper.py:189-345
        cfg = self.cfg
        nr, np_ = cfg["robot"]["total"], cfg["ped_sim"]["total"]
        n = nr + np_
        rb, pd = cfg["robot"], cfg["ped_sim"]
        btype = list(rb["begin_poses_type"][:nr]) + list(pd.get("begin_poses_type", [])[:np_])
        ttype = list(rb["target_poses_type"][:nr]) + list(pd.get("target_poses_type", [])[:np_])
        bpose = list(rb["begin_poses"][:nr]) + list(pd.get("begin_poses", [])[:np_])
        tpose = list(rb["target_poses"][:nr]) + list(pd.get("target_poses", [])[:np_])
        sizes = list(rb["size"][:nr]) + list(pd.get("size", [])[:np_])
        shapes = list(rb["shape"][:nr]) + list(pd.get("shape", [])[:np_])
        msize = [_module_size(sizes[i], shapes[i]) for i in range(n)]
        init, target = [None] * n, [None] * n
        circle_range = random.uniform(cfg["circle_ranges"][0], cfg["circle_ranges"][1])
        for i in range(n):
            if btype[i] == "fix":
                init[i] = bpose[i]
            if ttype[i] == "fix":
                target[i] = tpose[i]
            if btype[i] == "rand_angle":
                t = bpose[i]; init[i] = [t[0], t[1], random.uniform(t[2], t[3])]
            if ttype[i] == "rand_angle":
                t = tpose[i]; target[i] = [t[0], t[1], random.uniform(t[2], t[3])]
        circle_ok = False
        while not circle_ok:
            circle_ok = True
            for i in range(n):
                if init[i] is not None and target[i] is not None:
                    continue
                reset_init = True
                while reset_init:
                    goal_fail = circle_fail = 0
                    if "range" in btype[i]:
                        while reset_init:
                            pr = bpose[i]
                            if "circle" in btype[i]:
                                ang = random.uniform(-3.14, 3.14)
                                if "fix" in btype[i]:
                                    ang = -3.14 + (6.28 / n) * i
                                p = [circle_range * math.cos(ang) + pr[0], circle_range * math.sin(ang) + pr[1], ang + 3.14]
                                p[0] += random.gauss(0, 0.5); p[1] += random.gauss(0, 0.5)
                            else:
                                if "multi" in btype[i]:
                                    pr = pr[random.randint(0, len(pr) - 1)]
                                p = _random_pose(pr[:2], pr[2:4], [-3.14, 3.14] if len(pr) == 4 else pr[4:6])
                            if _free_robo_ped(p[0], p[1], init) and _free_obj([p[0], p[1], msize[i] * 2], self.obs_range):
                                init[i] = p[:]
                                reset_init = False
                                break
                            if "circle" in btype[i]:
                                circle_fail += 1
                                if circle_fail > 50:
                                    circle_ok = False
                                    for j in range(n):
                                        if "circle" in btype[j]:
                                            init[j] = target[j] = None
                    if "circle_fix" in ttype[i] and init[i] is not None:
                        pr, ang = tpose[i], init[i][2]
                        target[i] = [circle_range * math.cos(ang) + pr[0], circle_range * math.sin(ang) + pr[1], ang - 3.14]
                    if "range" in ttype[i]:
                        while True:
                            pr = tpose[i]
                            if "circle" in ttype[i] and init[i] is not None:
                                ang = init[i][2]
                                p = [circle_range * math.cos(ang) + pr[0], circle_range * math.sin(ang) + pr[1], ang - 3.14]
                                p[0] += random.gauss(0, 0.5); p[1] += random.gauss(0, 0.5)
                            if "multi" in ttype[i]:
                                pr = pr[random.randint(0, len(pr) - 1)]
                            if "view" in ttype[i]:
                                if "plus" not in ttype[i]:
                                    p = _random_view(init[i], pr)
                            elif len(pr) == 4:
                                p = _random_pose(pr[:2], pr[2:4], [-3.14, 3.14])
                            elif len(pr) == 6:
                                p = _random_pose(pr[:2], pr[2:4], pr[4:6])
                            if (init[i][0] - p[0]) ** 2 + (init[i][1] - p[1]) ** 2 > cfg["target_min_dist"] ** 2 \
                                    and _free_robo_ped(p[0], p[1], target) and _free_obj([p[0], p[1], msize[i] * 2], self.obs_range):
                                target[i] = p[:]
                                break
                            goal_fail += 1
                            if goal_fail > 50:
                                reset_init = True
                                break
        if any(init[i] is None or target[i] is None for i in range(n)):
            return None
        robots = np.zeros((nr, 8)); peds = np.zeros((np_, 8)); tl = np.zeros(np_, np.int32); traj = np.zeros((np_, 2, 3))
        for i in range(nr):
            robots[i] = [init[i][0], init[i][1], *rpy_to_q(init[i][2]), target[i][0], target[i][1]]
        assert np_ == 0 or pd["go_back"] in ["yes", "no", "random"]
        for k in range(np_):
            i = nr + k
            peds[k] = [init[i][0], init[i][1], *rpy_to_q(init[i][2]), target[i][0], target[i][1]]
            traj[k, 0] = [target[i][0], target[i][1], 0]; tl[k] = 1
            if pd["go_back"] == "yes" or (pd["go_back"] == "random" and random.random() > 0.5):
                traj[k, 1] = [init[i][0], init[i][1], 0]; tl[k] = 2
        self.init_poses, self.target_poses, self.circle_range = init, target, circle_range
        return dict(robots=robots, peds=peds, traj_len=tl, traj=traj)


_MAX_MULTI = 8


def _pose_spec(ptype, pose):
    bits = (1 * (ptype == "fix") | 2 * (ptype == "rand_angle") | 4 * ("range" in ptype) | 8 * ("circle" in ptype) | 16 * ("fix" in ptype)
            | 32 * ("multi" in ptype) | 64 * ("view" in ptype) | 128 * ("plus" in ptype) | 256 * ("circle_fix" in ptype))
    rec = np.zeros(3 + _MAX_MULTI * 6)
    rows = [list(r) for r in pose] if "multi" in ptype else [list(pose)]
    if len(rows) > _MAX_MULTI or any(len(r) > 6 for r in rows) or len({len(r) for r in rows}) != 1:
        raise ValueError("pose spec does not fit the native sampler descriptor: %r" % (pose,))
    rec[0], rec[1], rec[2] = bits, len(rows) if "multi" in ptype else 0, len(rows[0])
    for m, r in enumerate(rows):
        rec[3 + 6 * m: 3 + 6 * m + len(r)] = r
    return rec


def sampler_desc(cfg):
    """Flat float64 descriptor of this cfg's EnvPos for imgenv_sampler_create (include/imgenv.h)."""
    rb, pd = cfg["robot"], cfg["ped_sim"]
    nr, np_ = rb["total"], pd["total"]
    ob = cfg.get("object", {"total": 0})
    go_back = {"yes": 0, "no": 1, "random": 2}[pd.get("go_back", "yes")] if np_ else 0
    cr = cfg.get("circle_ranges", [0.0, 0.0])
    out = [np.array([nr, np_, ob["total"], cr[0], cr[1], cfg.get("target_min_dist", 0.0), go_back, 0.0])]
    for src, n in ((rb, nr), (pd, np_)):
        for i in range(n):
            out.append(np.array([_module_size(src["size"][i], src["shape"][i])]))
            out.append(_pose_spec(src["begin_poses_type"][i], src["begin_poses"][i]))
            out.append(_pose_spec(src["target_poses_type"][i], src["target_poses"][i]))
    for i in range(ob["total"]):
        rec = np.zeros(14)
        pose, size = list(ob["poses"][i]), list(ob["size_range"][i])
        rec[0] = 0 if ob["shape"][i] == "circle" else 1
        rec[1] = 1 if ob["poses_type"][i] == "fix" else 0
        rec[2] = len(pose); rec[3:3 + len(pose)] = pose; rec[9:9 + min(len(size), 4)] = size[:4]
        out.append(rec)
    return np.concatenate(out)
