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
    if rng.integers(0, 8) == 0:      # (drawn last: the cases of earlier seeds keep their other parameters)
        kw["use_laser"] = False      # the forward rasteriser: view_map_ is the raster, no rays (agent.cpp:439-441)
    return kw


@pytest.mark.parametrize("seed", range(int(os.environ.get("IMGENV_FUZZ_N", "120"))))      # IMGENV_FUZZ_N=1000 for a longer hunt
def test_random_configuration_matches_reference(seed):
    kw = _random_case(seed)
    cfg = _variant(**kw)
    extra = dict(lo=3.0, hi=8.0) if kw["scene"] == "pedscene" else dict(lo=2.5, hi=8.0)
    run_lockstep(cfg, seed=500 + seed, steps=4, **extra)


def _crowded_case(seed):
    """More agents in less space, bigger and smaller bodies, several scenes per handle: overlapping footprints, robots that
    see each other at close range, collision candidates from many records, near-origin (heavy) raster cells."""
    rng = np.random.default_rng(7000 + seed)
    R = int(rng.integers(3, 13)); P = int(rng.integers(2, 17))
    robot_shape = ["circle", "rectangle"][int(rng.integers(0, 2))]
    kw = dict(R=R, P=P, scene=["rvoscene", "ervoscene"][seed % 2], n_obj=int(rng.integers(0, 7)), max_ped=P + int(rng.integers(0, 3)),      # (the reference's _draw_ped_map indexes out of bounds with max_ped < P)
             
              ped_shape=["leg", "circle"][int(rng.integers(0, 2))], robot_shape=robot_shape, relation=int(rng.integers(0, 2)),
              range_total=int(rng.choice([360, 1000])))
    cfg = _variant(**kw)
    if robot_shape == "circle":
        cfg["robot"]["size"] = [[0, 0, float(rng.choice([0.1, 0.17, 0.3, 0.45]))] for _ in range(R)]
    else:
        a, b = float(rng.uniform(0.1, 0.5)), float(rng.uniform(0.1, 0.35))
        cfg["robot"]["size"] = [[-a, a, -b, b] for _ in range(R)]
    pr = float(rng.choice([0.07, 0.1, 0.2]))
    cfg["ped_sim"]["size"] = [([0, pr, pr] if kw["ped_shape"] == "leg" else [0, 0, pr]) for _ in range(P)]
    span = float(rng.choice([1.5, 3.0, 6.0]))
    lo = float(rng.uniform(0.6, 10.4 - span))
    return cfg, dict(lo=lo, hi=lo + span, S=int(rng.integers(1, 4)), beep=bool(seed % 2), opt_in_beep=bool(seed % 2))


@pytest.mark.parametrize("seed", range(int(os.environ.get("IMGENV_FUZZ_CROWDED_N", "40"))))
def test_random_crowded_configuration_matches_reference(seed):
    cfg, extra = _crowded_case(seed)
    run_lockstep(cfg, seed=900 + seed, steps=3, **extra)
