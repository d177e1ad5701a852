"""GPU: size-independent properties at BASELINE.json's full sizes, where the reference node is too
slow (C4: ~3 s per scene-step) or cannot run at all (C5) to serve as a lock-step oracle:

* batch invariance — a scene gives bit-identical State whether it is simulated alone or inside a batch, and whatever
  its scene index (scenes are independent in the reference: one node per scene);
* determinism — the same inputs give the same bits twice (all reductions are order independent);
* footprint records — after every call the per-part cell bitmaps the observation composed are well formed (candidate
  cells are occupied cells, every cell lies inside the map and inside its record's box, no bitmap exceeds its slot);
* range/format checks of the nine State fields.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import bench  # noqa: E402  (workload definitions of the bench = BASELINE.json configs)
from helpers import build_spec, random_actions  # noqa: E402


def _run(spec, resets, actions, steps, scene_ids=None, num_scenes=None):
    import torch
    from img_env_b200.lib import BatchedSim
    n = len(resets)
    S = num_scenes or n
    sim = BatchedSim(spec, num_scenes=S, ped_yaw_mode=1)
    ids = list(range(n)) if scene_ids is None else scene_ids
    if S > n:      # fill the other scenes with different episodes so that cross-talk would show
        other = [i for i in range(S) if i not in ids]
        sim.reset([resets[(k + 1) % n] for k in range(len(other))], scene_ids=other)
    sim.reset(resets, scene_ids=ids)
    assert sim.debug_check_footprints()[3] == 0
    outs = []
    for t in range(steps):
        a = np.zeros((S, spec["R"], 3), np.float32)
        for k, s in enumerate(ids):
            a[s] = actions[t][k]
        for s in range(S):
            if s not in ids:
                a[s] = actions[t][0][::-1]
        sim.step(torch.from_numpy(a).cuda())
        torch.cuda.synchronize()
        outs.append({k: v[ids].cpu().numpy().copy() for k, v in sim.out.items()})
        nrec, nocc, ncand, bad = sim.debug_check_footprints()
        assert bad == 0 and ncand <= nocc, "footprint records broken after step %d: %s" % (t, (nrec, nocc, ncand, bad))
    sim.close()
    return outs


def _same(a, b):
    for t, (x, y) in enumerate(zip(a, b)):
        for k in x:
            assert np.array_equal(x[k], y[k], equal_nan=True), "step %d field %s differs" % (t, k)


@pytest.mark.parametrize("wname,steps", [("c4", 3), ("c5", 3), ("c3", 4)])
def test_batch_invariance_determinism_and_footprint_records(wname, steps):
    w = bench.WORKLOADS[wname]
    spec = build_spec(bench.make_cfg(w))
    resets = bench.make_resets(spec, w, 2, seed=11)
    rng = np.random.default_rng(5)
    actions = [[random_actions(spec["R"], rng) for _ in range(2)] for _ in range(steps)]
    both = _run(spec, resets, actions, steps)
    again = _run(spec, resets, actions, steps)
    _same(both, again)                                             # determinism
    alone = _run(spec, resets[1:], [[a[1]] for a in actions], steps)
    _same([{k: v[1:] for k, v in o.items()} for o in both], alone)    # scene 1 alone == scene 1 in the batch
    moved = _run(spec, resets[1:], [[a[1]] for a in actions], steps, scene_ids=[2], num_scenes=3)
    _same(alone, moved)                                            # ... and at another scene index beside other scenes
    last = both[-1]
    R, img = spec["R"], spec["image_size"][0]
    assert last["sensor_maps"].shape == (2, R, img, img) and last["sensor_maps"].dtype == np.float16
    assert float(last["sensor_maps"].min()) >= 0.0 and float(last["sensor_maps"].max()) <= 1.0
    assert set(np.unique(last["is_collisions"]).tolist()) <= {0, 1, 2, 3}
    assert set(np.unique(last["is_arrives"]).tolist()) <= {0, 1}
    assert np.isfinite(last["lasers"]).all() and last["lasers"].min() >= 0 and last["lasers"].max() <= 1.0 + 1e-6
    assert np.all(last["ped_vector_states"][..., 0] == spec["P"])
    assert set(np.unique(last["ped_maps"][:, :, 0]).tolist()) <= {0.0, 1.0}


@pytest.mark.parametrize("robot_r,ped_shape,ped_r", [(0.17, "leg", 0.1), (0.25, "leg", 0.07), (0.4, "circle", 0.2), (0.12, "circle", 0.3)])
def test_footprint_records_equal_whole_lattice(robot_r, ped_shape, ped_r):
    """k_footprints sets the sure-covered interior of a circle part analytically and evaluates only the rim of its 0.01 m
    lattice (host_tables.h lattice_circle_ring).  debug_check_footprints rebuilds every record from the WHOLE lattice
    (agent.cpp:18-62 point by point) and counts differing words: several radii, random poses incl. the map border."""
    import torch
    from helpers import base_cfg, make_reset
    from img_env_b200.lib import BatchedSim
    cfg = base_cfg(R=6, P=10, scene="rvoscene", n_obj=3, ped_shape=ped_shape, max_ped=10)
    cfg["robot"]["size"] = [[0, 0, robot_r] for _ in range(6)]
    cfg["robot_radius"] = robot_r
    cfg["ped_sim"]["size"] = [([0, ped_r, ped_r] if ped_shape == "leg" else [0, 0, ped_r]) for _ in range(10)]
    spec = build_spec(cfg)
    rng = np.random.default_rng(int(robot_r * 1000))
    S = 6
    sim = BatchedSim(spec, num_scenes=S, ped_yaw_mode=2)
    sim.reset([make_reset(spec, rng, lo=0.2, hi=10.8) for _ in range(S)])
    nrec, nocc, ncand, bad = sim.debug_check_footprints()
    assert nrec > 0 and bad == 0, (nrec, nocc, ncand, bad)
    for t in range(4):
        a = np.stack([random_actions(spec["R"], rng) for _ in range(S)])
        sim.step(torch.from_numpy(a).cuda())
        nrec, nocc, ncand, bad = sim.debug_check_footprints()
        assert bad == 0, "step %d: %d record words differ from the whole-lattice build" % (t, bad)
    sim.close()
