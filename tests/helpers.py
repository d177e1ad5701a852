"""Shared test helpers: synthetic maps, cfg dicts in the reference's yaml schema, seeded reset
requests, and the comparison rules of SURVEY.md Appendix C."""
import copy
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from img_env_b200.spec import build_spec, rpy_to_q  # noqa: E402


from img_env_b200.scenarios import synthetic_map, base_cfg, make_reset, random_actions, make_dataset_reset  # noqa: E402,F401


def compare_state(got, want, spec, where="", frozen=None):
    """got: dict of numpy arrays from the product for ONE scene ([R,...]); want: PyPost.get_states dict.
    Tolerances: BASELINE.json north_star / SURVEY.md Appendix C."""
    R = spec["R"]
    frozen = np.zeros(R, bool) if frozen is None else frozen
    res = float(np.float32(spec["scalars"][0]))
    lm = spec["laser_max"] if spec["laser_norm"] else 1.0
    msgs = []

    def chk(ok, msg):
        if not ok:
            msgs.append(where + msg)
    chk(np.array_equal(got["is_collisions"].astype(np.int64), want["is_collisions"].astype(np.int64)),
        "is_collisions %s vs %s" % (got["is_collisions"], want["is_collisions"]))
    chk(np.array_equal(got["is_arrives"].astype(bool), want["is_arrives"].astype(bool)), "is_arrives differ")
    sm_g = got["sensor_maps"].view(np.uint16) if got["sensor_maps"].dtype == np.float16 else got["sensor_maps"]
    sm_w = want["sensor_maps"].astype(np.float16).view(np.uint16)
    nbad = int((sm_g != sm_w).sum())
    chk(nbad == 0, "sensor_maps: %d/%d float16 pixels differ" % (nbad, sm_w.size))
    if want["lasers"].ndim == 2 and want["lasers"].shape[1] > 0:      # use_laser=False: the node sends no laser ranges at all
        dl = np.abs(got["lasers"].astype(np.float64) - want["lasers"])
        chk(dl.max(initial=0) <= res / lm * 1.0001 + 1e-6, "lasers: max |d| %.3g (one cell = %.3g)" % (dl.max(initial=0), res / lm))
    chk(np.allclose(got["vector_states"], want["vector_states"], rtol=1e-4, atol=1e-5), "vector_states differ")
    gd = np.sqrt((want["vector_states"][:, :2] ** 2).sum(1))
    chk(np.all(np.abs(got["step_ds"] - want["step_ds"]) <= 1e-4 * np.maximum(gd, 1.0)), "step_ds differ %s vs %s" % (got["step_ds"], want["step_ds"]))
    chk(np.allclose(got["ped_vector_states"], want["ped_vector_states"], rtol=1e-4, atol=1e-5), "ped_vector_states differ")
    chk(np.allclose(got["ped_maps"], want["ped_maps"], rtol=1e-4, atol=1e-5),
        "ped_maps differ in %d cells" % int((np.abs(got["ped_maps"] - want["ped_maps"]) > 1e-4).sum()))
    wm = want["ped_min_dists"]
    chk(np.allclose(got["ped_min_dists"].astype(np.float64), wm, rtol=1e-4, atol=1e-5, equal_nan=True), "ped_min_dists differ")
    return msgs


STATE_KEYS = ["vector_states", "sensor_maps", "is_collisions", "is_arrives", "lasers", "ped_vector_states", "ped_maps", "step_ds",
              "ped_min_dists"]


def load_golden(name):
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    return dict(np.load(path, allow_pickle=False))


def golden_reset_request(g):
    return {k[len("reset_"):]: g[k] for k in g if k.startswith("reset_")}


def golden_state(g, prefix):
    return {k: g["%s_%s" % (prefix, k)] for k in STATE_KEYS}


