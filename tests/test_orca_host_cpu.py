"""CPU: the ORCA / ERVO solver the GPU runs (img_env_b200/csrc/orca.cuh), compiled for the host by a small shim
(tests/host/orca_host_harness.cpp), against the UNMODIFIED reference node (oracle/_ref) on crowded scenes: the new velocity
of every pedestrian after one doStep (RVOSimulator.cpp / Agent.cpp:437-1001) within 1e-5.  Crowded scenes reach the corners of
the solver the ordinary cases never touch -- e.g. the relaxed (3-D) program starting from an OBSTACLE line that the disc of
admissible speeds cannot meet (found by tests/test_fuzz_gpu.py, crowded case 607)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, build_spec, make_reset, random_actions


def _corners(obs):
    """The two rotated corners per reset object the RVO ring is built from (get_corners agent.cpp:626-651; imgenv.cu pack_reset_record)."""
    from img_env_b200.lib import load_library
    lib = load_library()
    out = np.zeros((len(obs), 4))
    f32 = lambda v: float(np.float32(v))
    for k, q in enumerate(obs):
        yaw = lib.imgenv_yaw_from_quaternion(C.c_double(q[7]), C.c_double(q[8]), C.c_double(q[9]), C.c_double(q[10]))
        h = yaw * 0.5
        cz, sz = math.cos(h), math.sin(h)
        d = (sz * sz) + (cz * cz); s = 2.0 / d; zs = sz * s; wz = cz * zs; zz = sz * zs
        m00, m01, m10, m11 = 1.0 - zz, 0.0 - wz, wz, 1.0 - zz
        ap = lambda vx, vy: ((m00 * vx + m01 * vy) + q[5], (m10 * vx + m11 * vy) + q[6])
        o = [f32(v) for v in q[1:5]]
        if int(q[0]) == 0:
            a, b = ap(o[0] - o[2], o[1] - o[2]), ap(o[0] + o[2], o[1] + o[2])
        else:
            a, b = ap(o[0], o[2]), ap(o[1], o[3])
        out[k] = [a[0], a[1], b[0], b[1]]
    return out


def _host_tree(corners, n_obj):
    from img_env_b200.lib import load_library
    lib = load_library()
    mv = 16 * max(n_obj, 1) + 16
    verts = np.zeros((mv, 8), np.float32); nodes = np.zeros((mv, 4), np.int32); root = C.c_int32()
    cc = np.ascontiguousarray(corners, dtype=np.float64)
    n = lib.imgenv_host_rvo_tree(cc.ctypes.data_as(C.POINTER(C.c_double)), int(n_obj), mv, C.byref(root),
                                 verts.ctypes.data_as(C.POINTER(C.c_float)), nodes.ctypes.data_as(C.POINTER(C.c_int32)))
    assert n >= 0
    return n, root.value, verts[:n], nodes[:n]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("orca") / "orca_host"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "img_env_b200", "csrc"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "orca_host_harness.cpp")], check=True)
    return str(exe)


def _fmt(a):
    return " ".join("%.9g" % float(x) for x in np.asarray(a).ravel())


@pytest.mark.parametrize("seed", [607, 0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377])
def test_orca_on_the_host_matches_the_reference_node(harness, seed):
    from oracle.pyref import RefEnv, have_ref
    from test_fuzz_gpu import _crowded_case
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    cfg, extra = _crowded_case(seed)
    spec = build_spec(cfg, opt_in_beep=extra["opt_in_beep"])
    R, P = spec["R"], spec["P"]
    ervo = spec["scene_type"] == "ervoscene"
    rng = np.random.default_rng(900 + seed)
    resets = [make_reset(spec, rng, lo=extra["lo"], hi=extra["hi"]) for _ in range(extra["S"])]
    dt = float(np.float32(spec["scalars"][4])) if False else None
    checked = 0
    for rs in resets:
        ref = RefEnv(spec)
        ref.reset(rs)
        n_obj = len(rs["obs"])
        if rs["ignore_obstacle"] or n_obj == 0:
            nv, root, verts, nodes = 0, -1, np.zeros((0, 8), np.float32), np.zeros((0, 4), np.int32)
        else:
            nv, root, verts, nodes = _host_tree(_corners(rs["obs"]), n_obj)
        alive = np.ones(R, np.uint8)
        for t in range(3):
            acts = random_actions(R, rng, beep=extra["beep"])
            pre = ref.rvo_get()
            rb, pd = ref.get_internal()
            goals = np.zeros((P, 3), np.float32)
            for p_ in range(P):      # waypoint cycling (img_env.cpp:306-319, agent.cpp:823-843)
                ti, tl = int(pd[p_, 17]), int(rs["traj_len"][p_])
                if ti < tl and (rs["traj"][p_, ti, 0] - pd[p_, 0]) ** 2 + (rs["traj"][p_, ti, 1] - pd[p_, 1]) ** 2 < 0.04:
                    ti += 1
                goals[p_] = [rs["traj"][p_, ti % tl, 0], rs["traj"][p_, ti % tl, 1], np.float32(cfg["ped_sim"]["max_speed"][p_])]
            beeps = []
            if ervo and extra["opt_in_beep"]:      # img_env.cpp:323-342 with ped_ca_p = 1: every alive robot whose v_y > 0 beeps
                for j in range(R):
                    if alive[j] and acts[j, 2] > 0:
                        beeps.append([np.float32(rb[j, 0]), np.float32(rb[j, 1]), np.float32(cfg["beep_r"])])
            st = ref.step(acts * alive[:, None], alive)
            post = ref.rvo_get()
            step_hz = float(np.float32(cfg["control_hz"]))
            text = "%d %d %.9g %d\n%s\n%d %d\n%s\n%s\n%s\n%d\n%s\n" % (len(pre), P, step_hz, int(ervo), _fmt(pre), nv, root, _fmt(verts), _fmt(nodes),
                                                                     _fmt(goals), len(beeps), _fmt(beeps) if beeps else "")
            r = subprocess.run([harness], input=text, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            got = np.array([[float(x) for x in line.split()] for line in r.stdout.strip().splitlines()], np.float32)
            want = post[:P, 2:4]
            bad = np.where(np.abs(got - want).max(1) > 1e-5)[0]
            assert len(bad) == 0, "seed %d step %d: pedestrians %s: host-compiled solver %s vs reference %s" % (seed, t, bad, got[bad], want[bad])
            checked += P
            alive = (1 - np.clip(np.clip(st["is_collisions"] if "is_collisions" in st else 0, 0, 1), 0, 1)).astype(np.uint8) if False else alive
    assert checked > 0
