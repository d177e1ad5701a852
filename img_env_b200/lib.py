"""ctypes binding of libimgenv_b200.so (include/imgenv.h) + BatchedSim, a thin torch-facing wrapper.

PyTorch is used only for device memory and streams: the nine ImageState tensors are allocated
here, bound once (imgenv_bind_outputs) and written in place by the CUDA kernels on every
reset/step.  There is no CPU fallback: constructing a BatchedSim without the compiled extension or
without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from .spec import SCENES, KTYPES

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IMGENV_LIB_PATH") or os.path.join(_HERE, "libimgenv_b200.so")    # (override: kernel-variant experiments)
_LIB = None


class ImgenvConfig(C.Structure):
    _fields_ = [("view_resolution", C.c_double), ("view_width", C.c_double), ("view_height", C.c_double),
                ("step_hz", C.c_double), ("state_dim", C.c_int32), ("use_laser", C.c_int32), ("range_total", C.c_int32),
                ("view_angle_begin", C.c_double), ("view_angle_end", C.c_double), ("view_min_dist", C.c_double),
                ("view_max_dist", C.c_double), ("beep_r", C.c_double), ("ped_ca_p", C.c_double),
                ("relation_ped_robo", C.c_int32), ("image_size", C.c_int32), ("ped_image_size", C.c_int32),
                ("max_ped", C.c_int32), ("ped_vec_dim", C.c_int32), ("ped_image_r", C.c_double), ("laser_max", C.c_double),
                ("laser_norm", C.c_int32), ("num_scenes", C.c_int32), ("num_robots", C.c_int32), ("num_peds", C.c_int32),
                ("scene_type", C.c_int32), ("robot_ktype", C.c_int32), ("max_obstacles", C.c_int32), ("max_traj", C.c_int32),
                ("seed", C.c_uint64), ("max_object_radius", C.c_double)]


class ImgenvOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("vector_states", "sensor_maps", "is_collisions", "is_arrives", "lasers",
                                          "ped_vector_states", "ped_maps", "step_ds", "ped_min_dists")]


EXPORTS = ["imgenv_create", "imgenv_destroy", "imgenv_bind_outputs", "imgenv_reset", "imgenv_step", "imgenv_step_host",
           "imgenv_end_episode", "imgenv_get_internal", "imgenv_set_internal", "imgenv_debug_view_maps", "imgenv_debug_view_maps2", "imgenv_debug_global_map",
           "imgenv_solver_agents", "imgenv_view_dims", "imgenv_launches_per_step",
           "imgenv_algorithmic_bytes_per_robot_step", "imgenv_last_error", "imgenv_version",
           "imgenv_sampler_create", "imgenv_sampler_destroy", "imgenv_sampler_seed", "imgenv_sampler_sample", "imgenv_sampler_draw",
           "imgenv_reset_sampled", "imgenv_debug_check_footprints",
           "imgenv_record_enable", "imgenv_record_fetch", "imgenv_debug_set_min_jerk", "imgenv_set_ped_yaw_mode", "imgenv_debug_counters", "imgenv_debug_view_stats", "imgenv_debug_view_phases", "imgenv_debug_rvo_tree", "imgenv_host_rvo_tree", "imgenv_autoreset_enable", "imgenv_autoreset_refill",
           "imgenv_reset_masked"]


def load_library(path=None):
    """Loads the compiled extension; raises if it has not been built (no silent fallback)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("libimgenv_b200.so is not built (run `python -m img_env_b200.build` or __graft_entry__.build()); "
                           "img_env_b200 has no CPU fallback")
    lib = C.CDLL(p)
    lib.imgenv_last_error.restype = C.c_char_p
    lib.imgenv_version.restype = C.c_char_p
    lib.imgenv_algorithmic_bytes_per_robot_step.restype = C.c_int64
    lib.imgenv_algorithmic_bytes_per_robot_step.argtypes = [C.c_void_p]
    lib.imgenv_yaw_from_quaternion.restype = C.c_double
    lib.imgenv_yaw_from_quaternion.argtypes = [C.c_double] * 4
    if path is None:
        _LIB = lib
    return lib


def _ptr(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class NativeSampler:
    """EnvPos.reset (reset_helper.py:115-345) in native code; host-only (usable without a GPU).
    desc comes from img_env_b200.envs.reset_helper.sampler_desc(cfg)."""

    def __init__(self, desc, num_scenes=1, seed=0, max_obs=None, max_traj=2):
        self.lib = load_library()
        d = np.ascontiguousarray(desc, dtype=np.float64)
        self.R, self.P, self.n_obj = int(d[0]), int(d[1]), int(d[2])
        self.max_obs = int(max_obs if max_obs is not None else self.n_obj); self.max_traj = int(max_traj)
        self.num_scenes = int(num_scenes)
        h = C.c_void_p()
        rc = self.lib.imgenv_sampler_create(_ptr(d), C.c_int64(d.size), self.num_scenes, C.c_uint64(int(seed)), C.byref(h))
        if rc != 0:
            raise ValueError(self.lib.imgenv_last_error().decode())
        self.h = h

    def seed(self, scene, seed):
        if self.lib.imgenv_sampler_seed(self.h, int(scene), C.c_uint64(int(seed))) != 0:
            raise ValueError(self.lib.imgenv_last_error().decode())

    def draw(self, scene, kind, a=0.0, b=0.0):
        out = C.c_double()
        if self.lib.imgenv_sampler_draw(self.h, int(scene), {"random": 0, "uniform": 1, "gauss": 2, "randint": 3}[kind],
                                        C.c_double(a), C.c_double(b), C.byref(out)) != 0:
            raise ValueError(self.lib.imgenv_last_error().decode())
        return out.value

    def sample(self, scene_ids=None):
        """-> dict of arrays shaped like BatchedSim.reset's packed arguments (n_obs, obs, robots, peds, traj_len, traj)."""
        ids = np.ascontiguousarray(scene_ids if scene_ids is not None else np.arange(self.num_scenes), dtype=np.int32)
        n = ids.size
        mo = max(self.max_obs, 1)
        out = dict(n_obs=np.zeros(n, np.int32), obs=np.zeros((n, mo, 11)), robots=np.zeros((n, self.R, 8)), peds=np.zeros((n, self.P, 8)),
                   traj_len=np.zeros((n, self.P), np.int32), traj=np.zeros((n, self.P, self.max_traj, 3)))
        rc = self.lib.imgenv_sampler_sample(self.h, n, _ptr(ids, C.c_int32), self.max_obs, self.max_traj, _ptr(out["n_obs"], C.c_int32),
                                            _ptr(out["obs"]), _ptr(out["robots"]), _ptr(out["peds"]), _ptr(out["traj_len"], C.c_int32), _ptr(out["traj"]))
        if rc != 0:
            raise ValueError(self.lib.imgenv_last_error().decode())
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.imgenv_sampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchedSim:
    """S scenes of one configuration on one GPU. See include/imgenv.h for the contract."""

    def __init__(self, spec, num_scenes=1, device=0, seed=0, ped_yaw_mode=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("img_env_b200 needs a CUDA device (B200 / sm_100a); there is no CPU path")
        self.torch = torch
        self.lib = load_library()
        self.spec = spec
        self.S, self.R, self.P = int(num_scenes), spec["R"], spec["P"]
        self.device = torch.device("cuda", device)
        sc = spec["scalars"]
        cfg = ImgenvConfig(view_resolution=sc[0], view_width=sc[1], view_height=sc[2], step_hz=sc[3], state_dim=int(sc[4]),
                           use_laser=int(sc[11]), range_total=int(sc[12]), view_angle_begin=sc[13], view_angle_end=sc[14],
                           view_min_dist=sc[15], view_max_dist=sc[16], beep_r=sc[17], ped_ca_p=sc[18],
                           relation_ped_robo=int(sc[19]), image_size=spec["image_size"][0],
                           ped_image_size=spec["ped_image_size"][0], max_ped=spec["max_ped"], ped_vec_dim=spec["ped_vec_dim"],
                           ped_image_r=spec["ped_image_r"], laser_max=spec["laser_max"], laser_norm=int(spec["laser_norm"]),
                           num_scenes=self.S, num_robots=self.R, num_peds=self.P, scene_type=SCENES[spec["scene_type"]],
                           robot_ktype=KTYPES[spec["robot_ktype"]], max_obstacles=spec["max_obstacles"],
                           max_traj=spec["max_traj"], seed=seed, max_object_radius=float(spec.get("max_object_radius", 0.0)))
        self.cfg = cfg
        grid = np.ascontiguousarray(spec["grid"], dtype=np.uint8)
        rd = np.ascontiguousarray(spec["robot_desc"], dtype=np.float64)
        pd = np.ascontiguousarray(spec["ped_desc"], dtype=np.float64) if self.P else np.zeros((1, 8))
        sl = np.ascontiguousarray(spec["robot_size_last"], dtype=np.float64)
        h = C.c_void_p()
        rc = self.lib.imgenv_create(C.byref(cfg), _ptr(grid, C.c_uint8), grid.shape[0], grid.shape[1], _ptr(rd), _ptr(pd),
                                    _ptr(sl), device, C.byref(h))
        self._check(rc)
        self.h = h
        self.lib.imgenv_set_ped_yaw_mode(self.h, int(ped_yaw_mode))
        S, R = self.S, self.R
        img = spec["image_size"][0]
        self.state_dim, self.range_total = int(sc[4]), int(sc[12])
        self.pvs_len = 1 + spec["ped_vec_dim"] * spec["max_ped"]
        dev = self.device
        self.out = dict(
            vector_states=torch.zeros(S, R, self.state_dim, dtype=torch.float32, device=dev),
            sensor_maps=torch.zeros(S, R, img, img, dtype=torch.float16, device=dev),
            is_collisions=torch.zeros(S, R, dtype=torch.int8, device=dev),
            is_arrives=torch.zeros(S, R, dtype=torch.uint8, device=dev),
            lasers=torch.zeros(S, R, self.range_total, dtype=torch.float32, device=dev),
            ped_vector_states=torch.zeros(S, R, self.pvs_len, dtype=torch.float32, device=dev),
            ped_maps=torch.zeros(S, R, 3, img, img, dtype=torch.float32, device=dev),
            step_ds=torch.zeros(S, R, dtype=torch.float32, device=dev),
            ped_min_dists=torch.full((S, R), float("inf"), dtype=torch.float32, device=dev))
        o = ImgenvOutputs(**{k: v.data_ptr() for k, v in self.out.items()})
        self._check(self.lib.imgenv_bind_outputs(self.h, C.byref(o)))
        self.solver_agents = self.lib.imgenv_solver_agents(self.h)
        self.launches_per_step = self.lib.imgenv_launches_per_step(self.h)
        self.bytes_per_robot_step = int(self.lib.imgenv_algorithmic_bytes_per_robot_step(self.h))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("imgenv: " + self.lib.imgenv_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.imgenv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, resets, scene_ids=None):
        """resets: list (one per listed scene) of dicts obs[n,11], robots[R,8], peds[P,8], traj_len[P], traj[P,T,3]."""
        n = len(resets)
        ids = np.ascontiguousarray(scene_ids if scene_ids is not None else np.arange(n), dtype=np.int32)
        mo, mt = max(self.spec["max_obstacles"], 1), max(self.spec["max_traj"], 1)
        n_obs = np.zeros(n, np.int32)
        obs = np.zeros((n, mo, 11)); robots = np.zeros((n, self.R, 8)); peds = np.zeros((n, max(self.P, 1), 8))
        tl = np.zeros((n, max(self.P, 1)), np.int32); traj = np.zeros((n, max(self.P, 1), mt, 3))
        trajv = np.zeros((n, max(self.P, 1), mt, 3)) if self.spec["scene_type"] == "dataset" else None
        ign = 0
        for i, rs in enumerate(resets):
            o = np.asarray(rs["obs"], dtype=np.float64).reshape(-1, 11)
            n_obs[i] = o.shape[0]
            obs[i, : o.shape[0]] = o
            robots[i] = np.asarray(rs["robots"], dtype=np.float64).reshape(self.R, 8)
            if self.P:
                peds[i] = np.asarray(rs["peds"], dtype=np.float64).reshape(self.P, 8)
                tl[i] = np.asarray(rs["traj_len"], dtype=np.int32)
                for p in range(self.P):
                    traj[i, p, : tl[i, p]] = np.asarray(rs["traj"][p], dtype=np.float64).reshape(-1, 3)[: tl[i, p]]
                    if trajv is not None:
                        trajv[i, p, : tl[i, p]] = np.asarray(rs["traj_v"][p], dtype=np.float64).reshape(-1, 3)[: tl[i, p]]
            ign = int(rs.get("ignore_obstacle", 0))
        rc = self.lib.imgenv_reset(self.h, n, _ptr(ids, C.c_int32), _ptr(n_obs, C.c_int32), _ptr(obs), _ptr(robots), _ptr(peds),
                                   _ptr(tl, C.c_int32), _ptr(traj), _ptr(trajv) if trajv is not None else None, ign, self._stream())
        self._check(rc)
        return self.out

    def reset_sampled(self, sampler, scene_ids=None, ignore_obstacle=0):
        """EnvPos.reset + reset service for the listed scenes without leaving native code."""
        ids = np.ascontiguousarray(scene_ids if scene_ids is not None else np.arange(self.S), dtype=np.int32)
        self._check(self.lib.imgenv_reset_sampled(self.h, sampler.h, ids.size, _ptr(ids, C.c_int32), int(ignore_obstacle), self._stream()))
        return self.out

    def autoreset_enable(self, sampler, depth=4, ignore_obstacle=0):
        """Episode queues for device-side auto-reset: `depth` pre-sampled episodes per scene (see include/imgenv.h)."""
        self._check(self.lib.imgenv_autoreset_enable(self.h, sampler.h, int(depth), int(ignore_obstacle), self._stream()))
        self._autoreset_sampler = sampler          # keep it alive: the library samples from it

    def autoreset_refill(self):
        self._check(self.lib.imgenv_autoreset_refill(self.h, self._stream()))

    def reset_masked(self, mask, refill=True):
        """Resets the scenes with a non-zero mask byte (uint8 / bool CUDA tensor [S]) from their episode queues; no host sync."""
        torch = self.torch
        m = mask if mask.dtype == torch.uint8 else mask.to(torch.uint8)
        assert m.is_cuda and m.is_contiguous() and m.numel() == self.S
        self._check(self.lib.imgenv_reset_masked(self.h, C.c_void_p(m.data_ptr()), int(bool(refill)), self._stream()))
        return self.out

    def record_enable(self, max_steps):
        """Episode record (EpRes.msg): keep the poses / speeds of the next `max_steps` steps after every reset; 0 disables."""
        self._check(self.lib.imgenv_record_enable(self.h, int(max_steps)))
        self._rec_T = int(max_steps)

    def record_fetch(self, scene):
        """-> dict(robots[T,R,6] = x, y, yaw, v, w, alive; peds[T,P,5] = x, y, yaw, vx, vy) of `scene` since its last reset."""
        T = getattr(self, "_rec_T", 0)
        rb = np.zeros((max(T, 1), self.R, 6)); pd = np.zeros((max(T, 1), max(self.P, 1), 5))
        n = C.c_int32(0)
        self._check(self.lib.imgenv_record_fetch(self.h, int(scene), C.byref(n), _ptr(rb), _ptr(pd), self._stream()))
        return dict(robots=rb[: n.value], peds=pd[: n.value, : self.P])

    def debug_check_footprints(self):
        """-> (non-empty footprint records, occupied cells, candidate cells, violations of the record invariants)."""
        out = np.zeros(4, np.int64)
        self._check(self.lib.imgenv_debug_check_footprints(self.h, _ptr(out, C.c_int64), self._stream()))
        return tuple(int(x) for x in out)

    def step(self, actions, alive=None):
        """actions: float32 CUDA tensor [S,R,3] (v, w, beep); alive: uint8 CUDA tensor [S,R] or None."""
        torch = self.torch
        assert actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()
        assert tuple(actions.shape) == (self.S, self.R, 3)
        ap = None
        if alive is not None:
            assert alive.is_cuda and alive.dtype == torch.uint8 and alive.is_contiguous()
            ap = C.c_void_p(alive.data_ptr())
        self._check(self.lib.imgenv_step(self.h, C.c_void_p(actions.data_ptr()), ap, self._stream()))
        return self.out

    def step_host(self, actions_np, alive_np=None):
        """Host-buffer entry point (the call the reference-facing plugin makes): numpy / pinned arrays."""
        a = np.ascontiguousarray(actions_np, dtype=np.float32)
        al = np.ascontiguousarray(alive_np, dtype=np.uint8) if alive_np is not None else None
        self._check(self.lib.imgenv_step_host(self.h, _ptr(a, C.c_float), _ptr(al, C.c_uint8) if al is not None else None,
                                              self._stream()))
        return self.out

    def step_host_ptr(self, actions_ptr, alive_ptr=None):
        self._check(self.lib.imgenv_step_host(self.h, C.c_void_p(actions_ptr), C.c_void_p(alive_ptr) if alive_ptr else None,
                                              self._stream()))
        return self.out

    def debug_counters(self):
        """-> (ORCA obstacle-neighbour table overflows, 0, 0, 0) since creation."""
        out = np.zeros(4, np.int64)
        self._check(self.lib.imgenv_debug_counters(self.h, _ptr(out, C.c_int64), self._stream()))
        return tuple(int(x) for x in out)

    def debug_view_stats(self):
        """Work counters of the observation kernel (instrumented builds only, see include/imgenv.h)."""
        out = np.zeros(16, np.int64)
        self._check(self.lib.imgenv_debug_view_stats(self.h, _ptr(out, C.c_int64), self._stream()))
        return out

    def debug_view_phases(self):
        """Cycles per phase of the observation CTAs (instrumented builds only, see include/imgenv.h)."""
        out = np.zeros(8, np.int64)
        self._check(self.lib.imgenv_debug_view_phases(self.h, _ptr(out, C.c_int64), self._stream()))
        return out

    def debug_rvo_tree(self, scene=0):
        """-> dict(n, root, verts[n,8], nodes[n,4], corners[max_obs,4]) of the device-built RVO obstacle set of `scene`."""
        mv = 16 * max(self.spec["max_obstacles"], 1) + 16
        verts = np.zeros((mv, 8), np.float32); nodes = np.zeros((mv, 4), np.int32); corners = np.zeros((max(self.spec["max_obstacles"], 1), 4))
        n, root = C.c_int32(), C.c_int32()
        self._check(self.lib.imgenv_debug_rvo_tree(self.h, int(scene), C.byref(n), C.byref(root), _ptr(verts, C.c_float), _ptr(nodes, C.c_int32),
                                                   _ptr(corners), self._stream()))
        return dict(n=n.value, root=root.value, verts=verts[: n.value], nodes=nodes[: n.value], corners=corners)

    def host_rvo_tree(self, corners, n_obj):
        """The host restatement of the same construction on `corners[:n_obj]` -> dict(n, root, verts, nodes)."""
        mv = 16 * max(self.spec["max_obstacles"], 1) + 16
        verts = np.zeros((mv, 8), np.float32); nodes = np.zeros((mv, 4), np.int32)
        root = C.c_int32()
        cc = np.ascontiguousarray(corners, dtype=np.float64)
        n = self.lib.imgenv_host_rvo_tree(_ptr(cc), int(n_obj), mv, C.byref(root), _ptr(verts, C.c_float), _ptr(nodes, C.c_int32))
        return dict(n=n, root=root.value, verts=verts[: max(n, 0)], nodes=nodes[: max(n, 0)])

    def debug_set_min_jerk(self, min_jerk):
        """min_jerk[R,2] (linear, angular): the value the node's unassigned SpeedLimiter::min_jerk holds (tests)."""
        a = np.ascontiguousarray(min_jerk, dtype=np.float64).reshape(self.R, 2)
        self._check(self.lib.imgenv_debug_set_min_jerk(self.h, _ptr(a)))

    def revive(self):
        self._check(self.lib.imgenv_revive(self.h, self._stream()))

    def profile_begin(self, max_steps):
        self._check(self.lib.imgenv_profile_begin(self.h, int(max_steps)))

    def profile_end(self):
        ms = (C.c_float * 4)()
        n = self.lib.imgenv_profile_end(self.h, ms)
        return n, dict(k_dynamics=ms[0], k_footprints=ms[1], k_view=ms[2])

    def get_internal(self):
        rb = np.zeros((self.S, self.R, 16)); pd = np.zeros((self.S, max(self.P, 1), 20))
        na = self.solver_agents
        width = 12 if self.spec["scene_type"] == "pedscene" else 4
        sv = np.zeros((self.S, max(na, 1), width))
        self._check(self.lib.imgenv_get_internal(self.h, _ptr(rb), _ptr(pd) if self.P else None, _ptr(sv) if na else None))
        return rb, pd[:, : self.P], sv[:, :na]

    def set_internal(self, rb=None, pd=None, sv=None):
        rb = np.ascontiguousarray(rb, dtype=np.float64) if rb is not None else None
        pd = np.ascontiguousarray(pd, dtype=np.float64) if (pd is not None and self.P) else None
        sv = np.ascontiguousarray(sv, dtype=np.float64) if (sv is not None and self.solver_agents) else None
        self._check(self.lib.imgenv_set_internal(self.h, _ptr(rb), _ptr(pd), _ptr(sv)))

    def sfm_tree_get(self, scene=0):
        na = self.solver_agents
        nodes = np.zeros((1024, 5)); leaf = np.zeros((na, 4), np.int32); hash_ = np.zeros(na, np.int32); n = C.c_int32()
        self._check(self.lib.imgenv_sfm_tree_get(self.h, int(scene), C.byref(n), _ptr(nodes), _ptr(leaf, C.c_int32), _ptr(hash_, C.c_int32)))
        return nodes[: n.value].copy(), leaf, hash_

    def sfm_tree_set(self, nodes, leaf, hash_, scene=0):
        nodes = np.ascontiguousarray(nodes, dtype=np.float64); leaf = np.ascontiguousarray(leaf, dtype=np.int32)
        hash_ = np.ascontiguousarray(hash_, dtype=np.int32)
        self._check(self.lib.imgenv_sfm_tree_set(self.h, int(scene), nodes.shape[0], _ptr(nodes), _ptr(leaf, C.c_int32), _ptr(hash_, C.c_int32)))

    def debug_global_map(self, scene=0, robot=-1):
        """robot >= 0: that robot's global_map_; -1: peds_map_; -2: obs_map_ (u8 [H, W])."""
        H, W = self.spec["grid"].shape
        out = np.zeros((H, W), np.uint8)
        self._check(self.lib.imgenv_debug_global_map(self.h, int(scene), int(robot), _ptr(out, C.c_uint8), self._stream()))
        return out

    def debug_stats(self):
        out = np.zeros((self.S, self.R, 4), np.int32)
        self._check(self.lib.imgenv_debug_view_maps2(self.h, None, _ptr(out, C.c_int32), self._stream()))
        return out

    def debug_view_maps(self):
        vh, vw = C.c_int32(), C.c_int32()
        self.lib.imgenv_view_dims(self.h, C.byref(vh), C.byref(vw))
        out = np.zeros((self.S, self.R, vh.value, vw.value), np.uint8)
        self._check(self.lib.imgenv_debug_view_maps(self.h, _ptr(out, C.c_uint8), self._stream()))
        return out
