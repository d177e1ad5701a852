"""Per-robot work counters of k_view for the bench workloads (needs a library built with -DVIEW_STATS=1:
   python -c "from img_env_b200.build import build_variant; build_variant('build_variants/stats.so', ['VIEW_STATS=1'])"
   IMGENV_LIB_PATH=build_variants/stats.so python tools/view_stats.py c4 c1 c5)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from img_env_b200.scenarios import WORKLOADS, make_cfg, make_resets, random_actions
from img_env_b200.spec import build_spec
from img_env_b200.lib import BatchedSim
NAMES = ["robots", "near records", "near words", "static blocks", "candidates", "cells pushed", "dirty outputs", "shadow outputs",
         "ran lattice", "ran edge px", "heavy cells", "any hit", "listed segments"]
for wn in sys.argv[1:] or ["c4"]:
    w = WORKLOADS[wn]; S = min(w["scenes"], 64)
    spec = build_spec(make_cfg(w))
    sim = BatchedSim(spec, S, ped_yaw_mode=1)
    sim.reset(make_resets(spec, w, S, 1))
    rng = np.random.default_rng(0)
    base = sim.debug_view_stats().copy(); base_ph = sim.debug_view_phases().copy()
    for _ in range(5):
        a = np.stack([random_actions(spec["R"], rng) for _ in range(S)])
        sim.step(torch.from_numpy(a).cuda(), torch.ones(S, spec["R"], dtype=torch.uint8, device="cuda"))
    torch.cuda.synchronize()
    st = sim.debug_view_stats() - base
    ph = sim.debug_view_phases() - base_ph
    n = max(int(st[0]), 1)
    print(wn, "robot observations", n, "|", ", ".join("%s %.1f" % (NAMES[k], st[k] / n) for k in range(1, 13)))
    names = ["prologue", "gather", "B (+A)", "heavy + lasers", "D1 segments", "D2 listed", "dirty"]
    tot = max(int(ph[:7].sum()), 1)
    print(wn, "cycles per robot (thread 0, between barriers): %.0f |" % (tot / n), ", ".join("%s %.0f (%.0f%%)" % (names[k], ph[k] / n, 100.0 * ph[k] / tot) for k in range(7)))
    sim.close()
