"""Golden vectors for the host-side episode sampler: runs the REAL reference EnvPos
(/root/reference/envs/utils/reset_helper.py, imported by path with stub ROS message classes) on the
reference's own yaml configs with fixed python `random` seeds and stores the sampled requests.
tests/test_envs_cpu.py replays the same seeds through img_env_b200.envs.reset_helper.EnvPos.
Also writes tests/golden/cfg/*.yaml: the sampler-relevant keys of those configs (input fixtures)."""
import importlib.util
import math
import os
import random
import sys
import types

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
CFGS = ["test", "circle", "random", "10obs_5ped_baseline"]
KEEP = ["robot", "object", "ped_sim", "circle_ranges", "target_min_dist", "env_name"]


def _load_reference_envpos():
    class Point:
        def __init__(self, x=0.0, y=0.0, z=0.0):
            self.x, self.y, self.z = x, y, z

    class Agent:
        def __init__(self):
            self.init_pose = types.SimpleNamespace(position=Point(), orientation=types.SimpleNamespace(x=0, y=0, z=0, w=0))
            self.goal = Point(); self.trajectory = []; self.trajectory_v = []; self.size = []; self.shape = ""

    class SpeedLimiter:
        pass

    ros_utils = types.ModuleType("envs.utils.ros_utils")
    ros_utils.rpy_to_q = lambda rpy: (0.0, 0.0, math.sin(rpy[2] / 2.0), math.cos(rpy[2] / 2.0))   # quaternion_from_euler(0,0,yaw)
    envs = types.ModuleType("envs"); utils = types.ModuleType("envs.utils"); utils.ros_utils = ros_utils
    comn = types.ModuleType("comn_pkg"); msg = types.ModuleType("comn_pkg.msg"); msg.Agent = Agent; msg.SpeedLimiter = SpeedLimiter
    gm = types.ModuleType("geometry_msgs"); gmm = types.ModuleType("geometry_msgs.msg"); gmm.Point = Point
    sys.modules.update({"envs": envs, "envs.utils": utils, "envs.utils.ros_utils": ros_utils, "comn_pkg": comn, "comn_pkg.msg": msg,
                        "geometry_msgs": gm, "geometry_msgs.msg": gmm})
    spec = importlib.util.spec_from_file_location("ref_reset_helper", os.path.join(REF, "envs/utils/reset_helper.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.EnvPos


def main():
    EnvPos = _load_reference_envpos()
    out = {}
    os.makedirs(os.path.join(HERE, "cfg"), exist_ok=True)
    for name in CFGS:
        cfg = yaml.load(open(os.path.join(REF, "envs/cfg", name + ".yaml")), Loader=yaml.FullLoader)
        yaml.safe_dump({k: cfg[k] for k in KEEP if k in cfg}, open(os.path.join(HERE, "cfg", name + ".yaml"), "w"))
        for seed in (0, 1, 2):
            random.seed(seed)
            obs, robots, peds = EnvPos(cfg).reset()
            key = "%s_%d" % (name, seed)
            out[key + "_obs"] = np.array([[o.init_pose.position.x, o.init_pose.position.y, o.init_pose.orientation.z, o.init_pose.orientation.w]
                                          + list(o.size) + [0] * (4 - len(o.size)) for o in obs], dtype=np.float64).reshape(-1, 8)
            out[key + "_robots"] = np.array([[r.init_pose.position.x, r.init_pose.position.y, r.init_pose.orientation.z,
                                              r.init_pose.orientation.w, r.goal.x, r.goal.y] for r in robots], dtype=np.float64).reshape(-1, 6)
            out[key + "_peds"] = np.array([[p.init_pose.position.x, p.init_pose.position.y, p.init_pose.orientation.z, p.init_pose.orientation.w,
                                            p.goal.x, p.goal.y, len(p.trajectory)] + [v for q in p.trajectory for v in (q.x, q.y)]
                                           + [0] * (4 - 2 * len(p.trajectory)) for p in peds], dtype=np.float64).reshape(-1, 11)
    np.savez_compressed(os.path.join(HERE, "reset_helper.npz"), **out)
    print("wrote reset_helper.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
