"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile) + the restated Python post-processing (oracle/pyref.py).

Run in the build container (needs /root/reference for the oracle build):
    python tests/golden/make_golden.py
Each file holds, for one seeded scenario: the reset request, per step the actions / alive mask and
the node's full pre-step internal state (robots, pedestrians, solver), and after every call the nine
ImageState arrays plus the raw 400x400 view_map of every robot.  cv2 version and IPP flag are recorded
because the INTER_CUBIC result depends on them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import base_cfg, build_spec, make_reset, random_actions  # noqa: E402
from scenarios import SCENARIOS  # noqa: E402


def generate(name, sc):
    import cv2
    from oracle.pyref import RefEnv, PyPost
    spec = build_spec(base_cfg(**sc["cfg"]), opt_in_beep=sc.get("opt_in_beep", False))
    R = spec["R"]
    rng = np.random.default_rng(sc["seed"])
    ref = RefEnv(spec); post = PyPost(spec)
    rs = make_reset(spec, rng, lo=sc.get("lo", 2.5), hi=sc.get("hi", 8.5))
    out = {"cv2_version": np.array(cv2.__version__), "ipp": np.array(False)}
    for k, v in rs.items():
        out["reset_" + k] = np.asarray(v)
    st = ref.reset(rs); post.on_reset()
    want = post.get_states(st)
    for k, v in want.items():
        out["r_" + k] = v
    out["r_view_map"] = st["view_map"]
    dones = np.zeros(R, np.int64)
    for t in range(sc["steps"]):
        acts = random_actions(R, rng, beep=sc.get("beep", False))
        alive = (dones == 0).astype(np.uint8)
        rb, pd = ref.get_internal()
        rb = rb.copy(); rb[:, 15] = post.tmp_distances if post.tmp_distances is not None else np.nan
        out["s%d_pre_robot" % t] = rb; out["s%d_pre_ped" % t] = pd
        out["s%d_pre_min_dist" % t] = np.array(post.min_dist, dtype=np.float64)
        if spec["P"]:
            out["s%d_pre_solver" % t] = (ref.rvo_get() if spec["scene_type"] != "pedscene" else ref.sfm_get()).astype(np.float64)
        out["s%d_actions" % t] = acts; out["s%d_alive" % t] = alive
        st = ref.step(acts * alive[:, None], alive)
        want = post.get_states(st)
        for k, v in want.items():
            out["s%d_%s" % (t, k)] = v
        out["s%d_view_map" % t] = st["view_map"]
        rb2, pd2 = ref.get_internal()
        out["s%d_post_robot" % t] = rb2; out["s%d_post_ped" % t] = pd2
        dones = np.clip(np.clip(want["is_collisions"], -1, 1) + want["is_arrives"], 0, 1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "%.1f KB" % (os.path.getsize(os.path.join(HERE, name + ".npz")) / 1024))


if __name__ == "__main__":
    for name, sc in SCENARIOS.items():
        generate(name, sc)
