"""CPU: the robot kinematics stage the GPU runs (img_env_b200/csrc/kin.cuh: SpeedLimiter::limit speed_limit.cpp:92-173 with its
velocity / acceleration / jerk branches, Agent::cmd agent.cpp:186-283 for diff and omni robots), called on the host
(tests/host/kin_host_harness.cpp) against the UNMODIFIED reference node on random limiter configurations: pose, the two
remembered commands, velocity and is_arrive after every step, to 1e-12 (is_arrives bit-exact)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, base_cfg, build_spec, make_reset, random_actions


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("kin") / "kin_host"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "img_env_b200", "csrc"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "kin_host_harness.cpp")], check=True)
    return str(exe)


def _case(seed):
    rng = np.random.default_rng(3000 + seed)
    cfg = base_cfg(R=4, P=0, n_obj=0, control_hz=float(rng.choice([0.1, 0.25, 0.4])), robot_type=["diff", "omni"][seed % 2])
    def lim(lo, hi, jerk):
        return dict(has_velocity_limits=bool(rng.integers(0, 2)), has_acceleration_limits=bool(rng.integers(0, 2)), has_jerk_limits=jerk,
                    min_velocity=float(rng.uniform(lo, 0.0)), max_velocity=float(rng.uniform(0.1, hi)),
                    min_acceleration=float(rng.uniform(-3, -0.2)), max_acceleration=float(rng.uniform(0.2, 3)),
                    min_jerk=float(rng.uniform(0.3, 4)), max_jerk=float(rng.uniform(0.3, 4)))
    if rng.integers(0, 5):
        cfg["speed_limiter_v"] = lim(-0.2, 0.7, bool(rng.integers(0, 2)))
    if rng.integers(0, 5):
        # (no jerk limit on w: the node's SpeedLimiter(msg) leaves min_jerk unassigned, a negative-going request is clamped to a
        #  denormal and cmd's arc branch divides by it -- the node itself then returns NaN poses)
        cfg["speed_limiter_w"] = lim(-1.0, 1.0, False)
    return cfg, rng


@pytest.mark.parametrize("seed", range(40))
def test_limiter_and_cmd_on_the_host_match_the_reference_node(harness, seed):
    from oracle.pyref import RefEnv, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    cfg, rng = _case(seed)
    spec = build_spec(cfg)
    R = spec["R"]
    ref = RefEnv(spec)
    ref.reset(make_reset(spec, rng, lo=2.0, hi=9.0))
    jl = ref.jerk_limits()            # min_jerk as the node's unassigned member holds it, max_jerk := msg.min_jerk (speed_limit.cpp:56-65)
    f32 = lambda v: float(np.float32(v))
    step_hz = f32(cfg["control_hz"])
    ktype = 0 if cfg["robot_type"] == "diff" else 1
    for t in range(12):
        acts = random_actions(R, rng, beep=(ktype == 1))
        if ktype == 1:
            acts[:, 2] = rng.uniform(-0.4, 0.4, R).astype(np.float32)      # v_y is a real command for omni robots
        rb = ref.get_internal()[0]
        rb[:, 12] = 0; rb[:, 13] = 0                                       # keep every robot alive: the limiter runs every step
        ref.set_internal(rb, None)
        ref.step(acts, np.ones(R))
        post = ref.get_internal()[0]
        for j in range(R):
            d = spec["robot_desc"][j]
            lims = []
            for L, (mn, mx) in ((d[7:16], (jl[j, 0], jl[j, 1])), (d[16:25], (jl[j, 2], jl[j, 3]))):
                lims.append("%d %d %d %.17g %.17g %.17g %.17g %.17g %.17g" % (L[0], L[1], L[2], f32(L[3]), f32(L[4]), f32(L[5]), f32(L[6]), mn, mx))
            text = "%d %.17g 0.05\n%s\n%s\n1\n%s %s\n" % (ktype, step_hz, lims[0], lims[1], " ".join("%.17g" % x for x in rb[j, [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11]]),
                                                       " ".join("%.17g" % float(x) for x in acts[j]))      # (float32 commands, widened exactly)
            r = subprocess.run([harness], input=text, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            got = np.array([float(x) for x in r.stdout.split()])
            want = post[j, [0, 1, 2, 6, 7, 8, 9, 10, 11]]
            assert np.allclose(got[:9], want, rtol=1e-12, atol=1e-12, equal_nan=True), "seed %d step %d robot %d:\n%s\n%s" % (seed, t, j, got[:9], want)
            assert int(got[9]) == int(post[j, 13] != 0), "seed %d step %d robot %d: is_arrive" % (seed, t, j)
