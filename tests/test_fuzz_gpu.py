"""GPU: randomized differential test — random configurations (scene type, robot/pedestrian shapes and counts, field of view,
ray count, sensor offset, view size, state_dim, kinematics) x random episodes, CUDA vs the unmodified reference node in
lock-step (tests/test_parity_gpu.py::run_lockstep, same tolerances).  The seeds are fixed: a failure is reproducible."""
import os

import numpy as np
import pytest

from test_parity_gpu import _variant, run_lockstep

pytestmark = pytest.mark.gpu


def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    scene = ["rvoscene", "ervoscene", "pedscene", "rvoscene"][seed % 4]
    R = int(rng.integers(1, 5)); P = int(rng.integers(0, 7))
    kw = dict(R=R, P=P, scene=scene, n_obj=int(rng.integers(0, 5)),
              ped_shape=["leg", "circle"][int(rng.integers(0, 2))], robot_shape=["circle", "rectangle"][int(rng.integers(0, 2))],
              state_dim=int(rng.choice([3, 4, 5])), relation=int(rng.integers(0, 2)), robot_type=["diff", "omni"][int(rng.integers(0, 2))],
              range_total=int(rng.choice([180, 360, 512, 1000])))
    half = float(rng.uniform(0.6, 3.0))
    kw["view_angle_begin"] = -half * float(rng.uniform(0.7, 1.0)); kw["view_angle_end"] = half
    kw["view_min_dist"] = float(rng.choice([0.0, 0.2])); kw["view_max_dist"] = float(rng.choice([2.0, 4.0, 10.0]))
    kw["laser_norm"] = bool(rng.integers(0, 2))
    if rng.integers(0, 3) == 0:
        kw["sensor"] = (float(rng.uniform(-0.1, 0.15)), float(rng.uniform(-0.05, 0.05)))
    if rng.integers(0, 3) == 0:
        kw["view"] = (float(rng.choice([0.015, 0.02, 0.025])), float(rng.choice([4, 6])))
    if rng.integers(0, 4) == 0:
        kw["grey_map"] = True
    if scene == "pedscene":          # keep the reference node inside the region where its quadtree terminates
        kw["R"] = min(R, 3); kw["P"] = min(P, 5)
    return kw


@pytest.mark.parametrize("seed", range(int(os.environ.get("IMGENV_FUZZ_N", "120"))))      # IMGENV_FUZZ_N=1000 for a longer hunt
def test_random_configuration_matches_reference(seed):
    kw = _random_case(seed)
    cfg = _variant(**kw)
    extra = dict(lo=3.0, hi=8.0) if kw["scene"] == "pedscene" else dict(lo=2.5, hi=8.0)
    run_lockstep(cfg, seed=500 + seed, steps=4, **extra)
