"""GPU parity: the CUDA path (through the C ABI) against the UNMODIFIED reference node
(oracle/_ref, built from /root/reference by oracle/Makefile; the prebuilt .so travels to the GPU
box) + the restated Python post-processing, single-step on identical pre-step state
(SURVEY.md Appendix C)."""
import numpy as np
import pytest

from helpers import base_cfg, build_spec, compare_state, make_dataset_reset, make_reset, random_actions

pytestmark = pytest.mark.gpu


def _to_np(out, s):
    return {k: v[s].detach().cpu().numpy() for k, v in out.items()}


def run_lockstep(cfg, seed, steps, S=1, beep=False, sync=True, opt_in_beep=False, n_obj=None, lo=2.5, hi=8.5,
                 check_view=True, tree_sync="once", ylo=None, yhi=None, map_dir=None, resets=None, action_fn=None):
    import torch
    from img_env_b200.lib import BatchedSim
    from oracle.pyref import RefEnv, PyPost, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libimgenv_ref.so not built")
    spec = build_spec(cfg, map_dir=map_dir, opt_in_beep=opt_in_beep)
    R = spec["R"]
    rng = np.random.default_rng(seed)
    sim = BatchedSim(spec, num_scenes=S, ped_yaw_mode=1)
    refs = [RefEnv(spec) for _ in range(S)]
    posts = [PyPost(spec) for _ in range(S)]
    if resets is not None:
        resets = resets(spec, rng) if callable(resets) else resets
    elif spec["scene_type"] == "dataset":
        resets = [make_dataset_reset(spec, rng, T=spec["max_traj"], lo=lo, hi=hi) for _ in range(S)]
    else:
        resets = [make_reset(spec, rng, n_obj=n_obj, lo=lo, hi=hi, ylo=ylo, yhi=yhi) for _ in range(S)]
    out = sim.reset(resets)
    torch.cuda.synchronize()
    errs = []
    for s in range(S):
        st = refs[s].reset(resets[s]); posts[s].on_reset()
        want = posts[s].get_states(st)
        errs += compare_state(_to_np(out, s), want, spec, where="reset scene %d: " % s)
        if check_view:
            vm = sim.debug_view_maps()
            nb = int((vm[s] != st["view_map"]).sum())
            if nb:
                errs.append("reset scene %d: %d view_map pixels differ" % (s, nb))
    assert not errs, "\n".join(errs[:20])
    sfm = spec["scene_type"] == "pedscene" and spec["P"] > 0
    if sfm:   # the node's quadtree depends on the process-global rand() history: import it once, then only emulate
        for s in range(S):
            sim.sfm_tree_set(*refs[s].sfm_tree(), scene=s)
    dones = np.zeros((S, R), np.int64)
    for t in range(steps):
        acts = np.stack([(action_fn(R, rng, t) if action_fn else random_actions(R, rng, beep=beep)) for _ in range(S)])
        alive = (dones == 0).astype(np.uint8)
        if sync:   # put the product into the node's exact pre-step state
            rbs, pds, svs = [], [], []
            for s in range(S):
                if t == 0 and spec["scene_type"] == "pedscene" and spec["P"]:
                    # libpedsim's socialForce takes sign(theta) of an angle that is pure rounding noise (~1e-16)
                    # while all relative velocities are exactly zero, i.e. only on the very first step of a
                    # process (velocities are never reset, pedscene.h:34-36).  Start from generic velocities.
                    a0 = refs[s].sfm_get()
                    a0[:, 3:5] = rng.uniform(-0.3, 0.3, (a0.shape[0], 2))
                    refs[s].sfm_set_pv(a0)
                rb, pd = refs[s].get_internal()
                rb = rb.copy()
                td = posts[s].tmp_distances
                rb[:, 15] = td if td is not None else np.nan
                rbs.append(rb); pds.append(pd)
                if sim.solver_agents and spec["scene_type"] != "dataset":
                    a = refs[s].rvo_get() if spec["scene_type"] in ("rvoscene", "ervoscene") else refs[s].sfm_get()
                    svs.append(a.astype(np.float64))
            sim.set_internal(np.stack(rbs), np.stack(pds) if spec["P"] else None, np.stack(svs) if svs else None)
            if sfm and tree_sync == "every":
                for s in range(S):
                    sim.sfm_tree_set(*refs[s].sfm_tree(), scene=s)
        out = sim.step(torch.from_numpy(acts).cuda(), torch.from_numpy(alive).cuda())
        torch.cuda.synchronize()
        vm = sim.debug_view_maps() if check_view else None
        for s in range(S):
            frozen_before = np.array([refs[s].get_internal()[0][j, 12] != 0 or refs[s].get_internal()[0][j, 13] != 0 for j in range(R)])
            st = refs[s].step(acts[s] * alive[s][:, None], alive[s])
            want = posts[s].get_states(st)
            errs += compare_state(_to_np(out, s), want, spec, where="step %d scene %d: " % (t, s))
            rb_ref, pd_ref = refs[s].get_internal()
            rb, pd, sv = sim.get_internal()
            if not np.allclose(rb[s][:, :12], rb_ref[:, :12], rtol=1e-4, atol=1e-6):
                errs.append("step %d scene %d: robot internal state differs" % (t, s))
            if sfm:
                vis_ref = refs[s].sfm_get()[: sim.solver_agents, 10]
                if not np.array_equal(sv[s][:, 10], vis_ref):
                    errs.append("step %d scene %d: quadtree membership differs: %s vs %s" % (t, s, sv[s][:, 10], vis_ref))
            if spec["P"] and not np.allclose(pd[s][:, [0, 1, 6, 7, 8, 10, 11, 12, 14, 15, 17]], pd_ref[:, [0, 1, 6, 7, 8, 10, 11, 12, 14, 15, 17]], rtol=1e-4, atol=1e-5):
                errs.append("step %d scene %d: pedestrian state differs\n%s\n%s" % (t, s, pd[s][:, :8], pd_ref[:, :8]))
            if check_view:
                # frozen robots keep a stale view in the node; the debug raster is recomputed -> skip them
                now_frozen = (rb_ref[:, 12] != 0) | (rb_ref[:, 13] != 0)
                for j in range(R):
                    if frozen_before[j] or rb_ref[j, 13] != 0:   # a robot that ARRIVES in this step skips its view too (cmd runs before view)
                        continue
                    nb = int((vm[s, j] != st["view_map"][j]).sum())
                    if nb:
                        errs.append("step %d scene %d robot %d: %d view_map pixels differ" % (t, s, j, nb))
                del now_frozen
            dones[s] = np.clip(np.clip(want["is_collisions"], -1, 1) + want["is_arrives"], 0, 1)
        assert not errs, "\n".join(errs[:20])
    assert sim.debug_counters()[0] == 0, "an ORCA agent saw more obstacle edges than its table holds (the reference keeps all)"
    sim.close()


def test_c1_single_robot_static():
    run_lockstep(base_cfg(R=1, P=0, n_obj=4), seed=1, steps=6)


def test_c2_eight_robots():
    run_lockstep(base_cfg(R=8, P=0, n_obj=0), seed=2, steps=6, lo=4.0, hi=7.0)


def test_c3_orca_20_peds():
    run_lockstep(base_cfg(R=1, P=20, scene="rvoscene", n_obj=4, max_ped=20), seed=3, steps=6)


def test_ervo_beeps():
    run_lockstep(base_cfg(R=3, P=12, scene="ervoscene", n_obj=3, max_ped=12), seed=4, steps=6, beep=True, opt_in_beep=True)


def test_crowded_collisions_multi_scene():
    run_lockstep(base_cfg(R=6, P=10, scene="rvoscene", n_obj=6, max_ped=10), seed=5, steps=8, S=3, lo=4.0, hi=7.0)


def test_circle_peds_rect_robot_state5():
    run_lockstep(base_cfg(R=2, P=5, scene="rvoscene", ped_shape="circle", robot_shape="rectangle", state_dim=5, n_obj=2), seed=6, steps=5)


def test_c5_sfm():
    run_lockstep(base_cfg(R=4, P=6, scene="pedscene", n_obj=3), seed=7, steps=6, lo=3.0, hi=8.0)


def test_sfm_many_agents_in_tree_bounds():
    # y in [10, 20] keeps agents inside the quadtree root box (pedscene.h:19): neighbours stay visible, leaves split
    cfg = base_cfg(R=4, P=12, scene="pedscene", n_obj=2, max_ped=12, map_px=220)
    # (x must stay in [0,10]: the reference itself recurses forever when > 8 agents sit beyond the same side of the root box)
    run_lockstep(cfg, seed=8, steps=10, lo=1.5, hi=8.5, ylo=8.0, yhi=19.0)


def test_sfm_quadtree_splits_64_agents():
    cfg = base_cfg(R=8, P=56, scene="pedscene", n_obj=0, max_ped=56, map_px=220)
    run_lockstep(cfg, seed=10, steps=8, lo=1.0, hi=9.0, ylo=6.0, yhi=19.0, check_view=False)


def test_omni_and_limiters():
    cfg = base_cfg(R=3, P=0, n_obj=3, robot_type="omni", control_hz=0.25)
    cfg["speed_limiter_v"] = dict(has_velocity_limits=True, has_acceleration_limits=True, has_jerk_limits=False, min_velocity=0,
                                  max_velocity=0.5, min_acceleration=-1.6, max_acceleration=1.0, min_jerk=0, max_jerk=0)
    cfg["speed_limiter_w"] = dict(has_velocity_limits=True, has_acceleration_limits=True, has_jerk_limits=False, min_velocity=-0.8,
                                  max_velocity=0.8, min_acceleration=-0.6, max_acceleration=2, min_jerk=0, max_jerk=0)
    run_lockstep(cfg, seed=9, steps=8, beep=True)


def test_dataset_replay_pedestrians():
    cfg = base_cfg(R=2, P=5, scene="dataset", n_obj=2)
    cfg["ped_sim"]["max_traj"] = 5
    run_lockstep(cfg, seed=11, steps=8, lo=3.0, hi=8.0)     # runs past the end of the trajectories (index clamps)


# ---- configuration sweep: every branch of the view / laser code against the reference node -------------------------
def _variant(**kw):
    over = {k: kw.pop(k) for k in list(kw) if k in ("view_angle_begin", "view_angle_end", "view_min_dist", "view_max_dist", "use_laser",
                                                    "laser_norm", "laser_max", "sensor", "view", "grey_map")}
    cfg = base_cfg(**kw)
    for k in ("view_angle_begin", "view_angle_end", "view_min_dist", "view_max_dist", "use_laser", "laser_norm", "laser_max"):
        if k in over:
            cfg[k] = over[k]
    if "sensor" in over:
        cfg["robot"]["sensor_cfgs"] = [list(over["sensor"])] * cfg["robot"]["total"]
    if "view" in over:
        cfg["view_map"] = dict(resolution=over["view"][0], width=over["view"][1], height=over["view"][1])
    if over.get("grey_map"):
        # interpolated / hand-painted greys: 1 and 2 read as pedestrian / robot collisions, < 250 is occupied for the view
        rng = np.random.default_rng(99)
        img = cfg["global_map"]["image"].copy()
        for v in (1, 2, 3, 100, 249, 250, 251):
            r, c = rng.integers(20, 85, 2)
            img[r:r + 6, c:c + 6] = v
        cfg["global_map"]["image"] = img
    return cfg


def test_narrow_fov_with_sensor_offset_and_fewer_rays():
    run_lockstep(_variant(R=3, P=4, scene="rvoscene", n_obj=3, range_total=360, view_angle_begin=-1.0, view_angle_end=1.2, sensor=(0.14, 0.0)),
                 seed=21, steps=5, lo=3.5, hi=7.5)


def test_min_and_max_view_distance():
    run_lockstep(_variant(R=2, P=3, scene="rvoscene", n_obj=3, view_min_dist=0.3, view_max_dist=2.5, range_total=512), seed=22, steps=5, lo=3.5, hi=7.5)


def test_without_lasers_view_map_is_the_raster():
    run_lockstep(_variant(R=3, P=4, scene="rvoscene", n_obj=3, use_laser=False), seed=23, steps=5, lo=3.5, hi=7.5)


def test_unnormalised_lasers_state_dim4():
    run_lockstep(_variant(R=2, P=0, n_obj=4, laser_norm=False, state_dim=4), seed=24, steps=5)


def test_grey_map_values_collide_like_agents():
    run_lockstep(_variant(R=6, P=3, scene="rvoscene", n_obj=2, grey_map=True), seed=25, steps=8, lo=2.0, hi=8.5)


def test_coarser_view_resolution():
    run_lockstep(_variant(R=2, P=3, scene="rvoscene", n_obj=3, view=(0.02, 6)), seed=26, steps=5, lo=3.5, hi=7.5)


def test_wide_fov_two_spans_per_row():
    run_lockstep(_variant(R=2, P=3, scene="rvoscene", n_obj=3, view_angle_begin=-2.6, view_angle_end=2.6, range_total=720), seed=27, steps=5, lo=3.5, hi=7.5)


def test_robot_near_map_border_and_outside_view():
    run_lockstep(_variant(R=4, P=2, scene="rvoscene", n_obj=1), seed=28, steps=6, lo=0.3, hi=2.0)


def test_composited_maps_match_node_rasters():
    """obs_map_ / peds_map_ of the node (img_env.cpp:167-187, 594-618) against the planes the CUDA path keeps instead:
    objects, circle and leg pedestrians incl. the right-leg overwrite quirk, on a map with grey values."""
    import torch
    from img_env_b200.lib import BatchedSim
    from oracle.pyref import RefEnv, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    for ped_shape in ("leg", "circle"):
        cfg = _variant(R=3, P=12, scene="rvoscene", n_obj=6, grey_map=True, ped_shape=ped_shape, max_ped=12)
        spec = build_spec(cfg)
        rng = np.random.default_rng(31)
        sim = BatchedSim(spec, 1, ped_yaw_mode=1); ref = RefEnv(spec)
        rs = make_reset(spec, rng, lo=3.0, hi=6.0)            # crowded: overlapping stamps
        sim.reset([rs]); ref.reset(rs)
        for t in range(3):
            assert np.array_equal(sim.debug_global_map(0, -2), ref.get_map(1)), "obs_map_ differs"
            assert np.array_equal(sim.debug_global_map(0, -1), ref.get_map(2)), "peds_map_ differs (%s, step %d)" % (ped_shape, t)
            acts = random_actions(3, rng)
            rb, pd = ref.get_internal(); rb = rb.copy(); rb[:, 15] = np.nan
            sim.set_internal(rb[None], pd[None], ref.rvo_get().astype(np.float64)[None])
            sim.step(torch.from_numpy(acts[None]).cuda(), torch.ones(1, 3, dtype=torch.uint8, device="cuda"))
            ref.step(acts, np.ones(3))
        sim.close()
