"""CPU tests of the test oracle itself (no GPU): the port (oracle/port, our restatement) is pinned
against the golden vectors generated from the UNMODIFIED reference node (tests/golden/*.npz,
tests/golden/make_golden.py) and, when oracle/_ref is present, against the node step by step."""
import subprocess

import numpy as np
import pytest

from helpers import (ROOT, base_cfg, build_spec, compare_state, golden_reset_request, golden_state, load_golden, make_reset,
                     random_actions)
from scenarios import SCENARIOS


@pytest.fixture(scope="module")
def port_lib():
    r = subprocess.run(["make", "-C", ROOT + "/oracle", "port"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    from oracle import pyref
    return pyref


def _raw_equal(a, b, tag):
    bad = []
    for k in a:
        if k == "view_map" or a[k].dtype.kind != "f":
            n = int((a[k] != b[k]).sum())
        else:
            n = int((~np.isclose(a[k], b[k], rtol=1e-5, atol=1e-6)).sum())
        if n:
            bad.append("%s %s: %d mismatches" % (tag, k, n))
    return bad


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_port_matches_reference_golden(port_lib, name):
    sc = SCENARIOS[name]
    g = load_golden(name)
    spec = build_spec(base_cfg(**sc["cfg"]), opt_in_beep=sc.get("opt_in_beep", False))
    port = port_lib.PortEnv(spec); post = port_lib.PyPost(spec)
    st = port.reset(golden_reset_request(g)); post.on_reset()
    errs = compare_state(post.get_states(st), golden_state(g, "r"), spec, where="reset: ")
    assert int((st["view_map"] != g["r_view_map"]).sum()) == 0
    for t in range(sc["steps"]):
        port.set_internal(g["s%d_pre_robot" % t], g["s%d_pre_ped" % t])
        post.tmp_distances = None if np.isnan(g["s%d_pre_robot" % t][:, 15]).any() else g["s%d_pre_robot" % t][:, 15].copy()
        post.min_dist = list(g["s%d_pre_min_dist" % t])
        if spec["P"]:
            sv = g["s%d_pre_solver" % t]
            if spec["scene_type"] == "pedscene":
                port.sfm_set(sv)
            else:
                port.rvo_set(sv.astype(np.float32))
        acts, alive = g["s%d_actions" % t], g["s%d_alive" % t]
        st = port.step(acts * alive[:, None], alive)
        errs += compare_state(post.get_states(st), golden_state(g, "s%d" % t), spec, where="step %d: " % t)
        nb = int((st["view_map"] != g["s%d_view_map" % t]).sum())
        if nb:
            errs.append("step %d: %d view_map pixels differ" % (t, nb))
        rb, pd = port.get_internal()
        if not np.allclose(rb[:, :14], g["s%d_post_robot" % t][:, :14], rtol=1e-9, atol=1e-12):
            errs.append("step %d: robot internals differ" % t)
    assert not errs, "\n".join(errs[:10])


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_port_matches_reference_live(port_lib, name):
    if not port_lib.have_ref():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    sc = SCENARIOS[name]
    spec = build_spec(base_cfg(**sc["cfg"]), opt_in_beep=sc.get("opt_in_beep", False))
    R = spec["R"]
    rng = np.random.default_rng(sc["seed"] + 1000)
    ref = port_lib.RefEnv(spec); port = port_lib.PortEnv(spec)
    rs = make_reset(spec, rng, lo=sc.get("lo", 2.5), hi=sc.get("hi", 8.5))
    errs = _raw_equal(ref.reset(rs), port.reset(rs), "reset")
    dones = np.zeros(R, np.int64)
    for t in range(5):
        acts = random_actions(R, rng, beep=sc.get("beep", False)); alive = (dones == 0).astype(np.uint8)
        rb, pd = ref.get_internal(); port.set_internal(rb, pd)
        if spec["P"]:
            if spec["scene_type"] == "pedscene":
                port.sfm_set(ref.sfm_get())
            else:
                port.rvo_set(ref.rvo_get())
        a = ref.step(acts * alive[:, None], alive); b = port.step(acts * alive[:, None], alive)
        errs += _raw_equal(a, b, "step %d" % t)
        dones = np.clip(np.clip(a["is_collision"], -1, 1) + a["is_arrive"], 0, 1)
    assert not errs, "\n".join(errs[:10])


def test_cubic_resize_model_matches_cv2(port_lib):
    """The INTER_CUBIC model the CUDA path implements (host tables of the product library)
    reproduces cv2.resize on OpenCV's own (non-IPP) path for 4-level view maps."""
    import ctypes as C
    from img_env_b200.build import build
    from img_env_b200.lib import load_library
    build()
    lib = load_library()
    need = np.zeros(400, np.int16); tap = np.zeros(48 * 4, np.int16); coef = np.zeros(48 * 4, np.int16); ns = C.c_int()
    lib.imgenv_cubic_tables(400, 48, need.ctypes.data_as(C.c_void_p), C.byref(ns), tap.ctypes.data_as(C.c_void_p), coef.ctypes.data_as(C.c_void_p))
    need = need[: ns.value].astype(np.int64); tap = tap.reshape(48, 4).astype(np.int64); coef = coef.reshape(48, 4).astype(np.int64)
    assert ns.value == 144
    rng = np.random.default_rng(0)
    for _ in range(6):
        img = rng.choice(np.array([0, 100, 200, 255], np.uint8), size=(400, 400), p=[0.2, 0.05, 0.35, 0.4])
        img[rng.integers(0, 400):, :] = 255
        src = img[need][:, need].astype(np.int64)                       # [144,144] needed pixels
        hb = np.zeros((144, 48), np.int64)
        for k in range(4):
            hb += src[:, tap[:, k]] * coef[:, k][None, :]
        hb = hb.astype(np.float32)
        scale = np.float32(1.0) / np.float32(2048.0 * 2048.0)
        b = (coef.astype(np.float32) * scale)                           # [48,4]
        # fp32 FMA chain S0*b0 + (S1*b1 + (S2*b2 + S3*b3)): emulate each fma with float64 (exact product, one rounding)
        acc = (hb[tap[:, 3]].astype(np.float64) * b[:, 3][:, None].astype(np.float64)).astype(np.float32)
        for k in (2, 1, 0):
            acc = (hb[tap[:, k]].astype(np.float64) * b[:, k][:, None].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
        out = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
        want = port_lib.cubic_resize_u8(img, (48, 48))
        assert int((out != want).sum()) == 0
