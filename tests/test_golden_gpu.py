"""GPU: the CUDA path replays the committed golden vectors of the UNMODIFIED reference node
(tests/golden/*.npz) — no oracle library needed on the box for this file."""
import numpy as np
import pytest

from helpers import base_cfg, build_spec, compare_state, golden_reset_request, golden_state, load_golden
from scenarios import SCENARIOS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_cuda_matches_reference_golden(name):
    import torch
    from img_env_b200.lib import BatchedSim
    sc = SCENARIOS[name]
    g = load_golden(name)
    spec = build_spec(base_cfg(**sc["cfg"]), opt_in_beep=sc.get("opt_in_beep", False))
    sim = BatchedSim(spec, num_scenes=1, ped_yaw_mode=1)
    out = sim.reset([golden_reset_request(g)])
    torch.cuda.synchronize()
    got = {k: v[0].cpu().numpy() for k, v in out.items()}
    errs = compare_state(got, golden_state(g, "r"), spec, where="reset: ")
    nb = int((sim.debug_view_maps()[0] != g["r_view_map"]).sum())
    if nb:
        errs.append("reset: %d view_map pixels differ" % nb)
    for t in range(sc["steps"]):
        pre = g["s%d_pre_robot" % t]
        sim.set_internal(pre[None], g["s%d_pre_ped" % t][None] if spec["P"] else None,
                         g["s%d_pre_solver" % t][None] if spec["P"] else None)
        acts, alive = g["s%d_actions" % t], g["s%d_alive" % t]
        out = sim.step(torch.from_numpy(acts[None].astype(np.float32)).cuda(), torch.from_numpy(alive[None].astype(np.uint8)).cuda())
        torch.cuda.synchronize()
        got = {k: v[0].cpu().numpy() for k, v in out.items()}
        want = golden_state(g, "s%d" % t)
        if spec["P"] == 0:
            want["ped_min_dists"] = got["ped_min_dists"].astype(np.float64)   # persists (inf) without pedestrians
        errs += compare_state(got, want, spec, where="step %d: " % t)
        vm = sim.debug_view_maps()[0]
        frozen = (pre[:, 12] != 0) | (pre[:, 13] != 0) | (g["s%d_post_robot" % t][:, 13] != 0)   # arriving robots skip their view too
        for j in range(spec["R"]):
            if not frozen[j]:
                nb = int((vm[j] != g["s%d_view_map" % t][j]).sum())
                if nb:
                    errs.append("step %d robot %d: %d view_map pixels differ" % (t, j, nb))
        rb, pd, sv = sim.get_internal()
        if not np.allclose(rb[0][:, :12], g["s%d_post_robot" % t][:, :12], rtol=1e-4, atol=1e-6):
            errs.append("step %d: robot poses differ" % t)
    sim.close()
    assert not errs, "\n".join(errs[:10])
