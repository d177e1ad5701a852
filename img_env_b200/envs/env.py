"""ImageEnv: the reference's Gym-style environment (envs/env/yaml_env.py:52-390) on top of the CUDA
library. Same constructor argument (the yaml cfg dict), same reset()/step(actions)/end_ep() surface
and attributes; `env_num` scenes of the cfg are simulated by ONE env object on one GPU (the reference
needs one ROS node + one ImageEnv per scene), so State arrays carry S*R robots, scene-major."""
import numpy as np

import random

from ..lib import BatchedSim, NativeSampler
from ..spec import build_spec
from .action import ContinuousAction
from .reset_helper import EnvPos, NearbyPed, sampler_desc
from .state import ImageState


class ImageEnv:
    def __init__(self, cfg: dict, num_scenes=None, device=0, numpy_state=False, map_dir=None, copy_state=None):
        """copy_state (default True, or cfg['copy_state']): every reset()/step() returns FRESH tensors, like the reference's
        fresh numpy arrays, so a State kept by the caller (replay buffer: obs vs next_obs) does not change under it.
        copy_state=False returns views of the library's persistent output buffers, which the next reset()/step() overwrites
        in place -- zero-copy for consumers that use the State before stepping again."""
        import torch
        self.copy_state = bool(cfg.get("copy_state", True) if copy_state is None else copy_state)
        self.torch = torch
        self.cfg = cfg
        self._init_static_param(cfg)
        self.num_scenes = int(num_scenes if num_scenes is not None else cfg.get("batch_scenes", 1))
        self.numpy_state = numpy_state
        self.env_pose = [EnvPos(cfg) for _ in range(self.num_scenes)]
        self.nearby_ped = NearbyPed(self.robot_total * self.num_scenes)
        self.spec = build_spec(cfg, map_dir=map_dir, opt_in_beep=bool(cfg.get("opt_in_beep", False)))
        self.sim = BatchedSim(self.spec, num_scenes=self.num_scenes, device=device, seed=int(cfg.get("seed", 0)),
                              ped_yaw_mode=int(cfg.get("ped_yaw_mode", 1)))
        # Episode sampling: the native EnvPos port (one CPython-compatible generator per scene, seeded from cfg['seed'] or
        # from python's `random`, so random.seed() in the training script still fixes the episodes); cfg['native_sampler']
        # = False keeps the per-scene python EnvPos objects drawing from the global `random`, exactly like the reference.
        self.sampler = None
        if cfg.get("native_sampler", True) and self.spec["scene_type"] != "dataset":
            self.sampler = NativeSampler(sampler_desc(cfg), num_scenes=self.num_scenes, seed=int(cfg.get("sampler_seed", random.getrandbits(62))),
                                         max_obs=self.spec["max_obstacles"], max_traj=self.spec["max_traj"])
        self._ignore_obstacle = int(bool(cfg["ped_sim"].get("ignore_obstacle", False)))
        # Device-side auto-reset (default with the native sampler): every scene holds a queue of pre-sampled episodes in device
        # memory, so reset(scene_mask=<device bool tensor>) -- what NeverStopWrapper calls every step -- costs no host sync.
        self.masked_reset = self.sampler is not None and not numpy_state and bool(cfg.get("device_autoreset", True))
        if self.masked_reset:
            self.sim.autoreset_enable(self.sampler, depth=int(cfg.get("autoreset_depth", 4)), ignore_obstacle=self._ignore_obstacle)
        self._record_steps = int(cfg.get("record_steps", 0))
        if self._record_steps:
            self.sim.record_enable(self._record_steps)
        self.dones = None
        self._act = torch.zeros(self.num_scenes, self.robot_total, 3, dtype=torch.float32, device=self.sim.device)

    def _init_static_param(self, cfg):     # yaml_env.py:133-181 (only the keys the hot path consumes)
        self.test = cfg.get("test", False)
        self.env_name = cfg.get("env_name", "img_env")
        self.env_type = cfg.get("env_type", "robot_nav")
        self.robot_type = cfg["robot_type"]
        self.image_size = tuple(cfg["image_size"]); self.ped_image_size = tuple(cfg["ped_image_size"])
        self.state_dim = cfg["state_dim"]; self.laser_max = cfg["laser_max"]; self.control_hz = cfg["control_hz"]
        self.robot_total = cfg["robot"]["total"]; self.ped_total = cfg["ped_sim"]["total"]
        self.max_ped = cfg["max_ped"]; self.ped_vec_dim = cfg["ped_vec_dim"]; self.ped_image_r = cfg["ped_image_r"]
        self.laser_norm = cfg.get("laser_norm", True)
        self.node_id = str(cfg.get("node_id", 0))

    def __len__(self):
        return self.num_scenes * self.robot_total

    def _state(self):
        o = self.sim.out
        n = len(self)
        flat = {k: v.reshape((n,) + tuple(v.shape[2:])) for k, v in o.items()}
        if self.numpy_state:                # reference dtypes (yaml_env.py:472-481)
            self.torch.cuda.synchronize()
            f = {k: v.cpu().numpy() for k, v in flat.items()}
            return ImageState(f["vector_states"].astype(np.float64), f["sensor_maps"], f["is_collisions"].astype(np.int64),
                              f["is_arrives"].astype(bool), f["lasers"].astype(np.float64), f["ped_vector_states"], f["ped_maps"],
                              f["step_ds"].astype(np.float64), f["ped_min_dists"].astype(np.float64))
        if self.copy_state:
            flat = {k: v.clone() for k, v in flat.items()}
        return ImageState(**flat)

    def _dataset_resets(self, ids, datas):
        """EnvPos.init_ped_dataset (reset_helper.py:417-434) on top of a sampled episode: per pedestrian T rows of
        (x, y, yaw, vx, vy) replace its start pose and trajectory (yaml_env.py:245-247)."""
        if datas is None:
            raise ValueError("ped_sim.type 'dataset' replays recorded trajectories: pass reset(cur_ped_pos_v_datas=array[P,T,5]) "
                             "(or [n_scenes,P,T,5]) as the reference's PedTrajectoryDatasetWrapper does (yaml_env.py:245-247)")
        from ..spec import rpy_to_q
        d = np.asarray(datas, dtype=np.float64)
        if d.ndim == 3:
            d = np.broadcast_to(d, (len(ids),) + d.shape)
        P, T = self.ped_total, d.shape[2]
        if d.shape[0] != len(ids) or d.shape[1] != P or d.shape[3] != 5 or T > self.spec["max_traj"]:
            raise ValueError("cur_ped_pos_v_datas must be [P=%d, T<=%d, 5] per scene, got %s" % (P, self.spec["max_traj"], d.shape))
        out = []
        for k, s in enumerate(ids):
            rs = self.env_pose[s].reset()
            rs["peds"] = np.array(rs["peds"], dtype=np.float64).reshape(P, 8)
            traj = np.zeros((P, T, 3)); trajv = np.zeros((P, T, 3))
            for p in range(P):
                traj[p] = d[k, p, :, :3]
                trajv[p, :, :2] = d[k, p, :, 3:5]
                rs["peds"][p, :2] = d[k, p, 0, :2]
                rs["peds"][p, 2:6] = rpy_to_q(d[k, p, 0, 2])
            rs["traj"], rs["traj_v"], rs["traj_len"] = traj, trajv, np.full(P, T, np.int32)
            out.append(rs)
        return out

    def reset(self, scene_ids=None, scene_mask=None, **kwargs):
        """scene_ids: list of scenes to reset (None = all).  scene_mask: bool / uint8 DEVICE tensor [num_scenes] instead of
        scene_ids (device-side auto-reset: no host synchronisation; needs the native sampler)."""
        torch = self.torch
        if self.masked_reset:
            # every sampled reset goes through the scenes' episode queues, so that a scene sees the same episode sequence
            # whether it is reset by list or by mask
            everything = scene_mask is None and scene_ids is None
            if scene_mask is None:
                scene_mask = torch.ones(self.num_scenes, dtype=torch.uint8, device=self.sim.device)
                if scene_ids is not None:
                    scene_mask.zero_(); scene_mask[torch.as_tensor(list(scene_ids), dtype=torch.long, device=self.sim.device)] = 1
            self.sim.reset_masked(scene_mask, refill=not getattr(self, "_capturing", False))
            state = self._state()
            if self.dones is None or everything:
                self.dones = torch.zeros(len(self), dtype=torch.int64, device=self.sim.device)
            return state
        if scene_mask is not None:
            raise ValueError("reset(scene_mask=...) needs the device-side auto-reset (native sampler, torch state)")
        ids = list(range(self.num_scenes)) if scene_ids is None else list(scene_ids)
        if self.spec["scene_type"] == "dataset":
            self.sim.reset(self._dataset_resets(ids, kwargs.get("cur_ped_pos_v_datas")), scene_ids=ids)
        elif self.sampler is not None:
            self.sim.reset_sampled(self.sampler, ids, self._ignore_obstacle)
        else:
            self.sim.reset([self.env_pose[s].reset() for s in ids], scene_ids=ids)
        state = self._state()
        if scene_ids is None or self.dones is None:
            self.dones = self.torch.zeros(len(self), dtype=self.torch.int64, device=self.sim.device)
        return state

    def step(self, actions):
        """actions: list of ContinuousAction (len S*R) or a float tensor/array [S*R, 2|3] of (v, w[, beep])."""
        torch = self.torch
        if isinstance(actions, (list, tuple)) and len(actions) and isinstance(actions[0], ContinuousAction):
            a = np.array([[x.v, x.w, x.beep] for x in actions], dtype=np.float32)
            self._act.copy_(torch.from_numpy(a).view_as(self._act), non_blocking=True)
        else:
            t = torch.as_tensor(actions, dtype=torch.float32, device=self.sim.device).reshape(self.num_scenes, self.robot_total, -1)
            self._act.zero_(); self._act[..., : t.shape[-1]] = t
        self.sim.step(self._act, None)     # alive = library dones (yaml_env.py:319-331)
        state = self._state()
        if self.numpy_state:
            rewards = state.is_arrives.astype(np.int64) - state.is_collisions
            dones = np.clip(np.clip(state.is_collisions, -1, 1) + state.is_arrives, 0, 1)
            self.dones = dones
            return state, rewards, dones.copy(), {"dones_info": np.zeros_like(dones)}
        coll = state.is_collisions.to(torch.int64); arr = state.is_arrives.to(torch.int64)
        rewards = arr - coll                                            # yaml_env.py:373
        self.dones = (coll.clamp(-1, 1) + arr).clamp(0, 1)               # yaml_env.py:374-376
        return state, rewards, self.dones.clone(), {"dones_info": torch.zeros_like(self.dones)}

    def end_ep(self, robot_res=None, scene=0):
        """yaml_env.py:379-390 / img_env.cpp:527-545.  The node publishes an EpRes message (poses and speeds of the episode);
        with cfg['record_steps'] > 0 the same record is returned for `scene` (None -> every scene) instead of True."""
        if not self._record_steps:
            return True
        scenes = range(self.num_scenes) if scene is None else [scene]
        out = []
        for sc in scenes:
            rec = self.sim.record_fetch(sc)
            rec["result"] = list(robot_res) if robot_res is not None else None
            rec["resolution"], rec["step_hz"], rec["env_name"] = self.spec["scalars"][0], self.control_hz, self.env_name
            out.append(rec)
        return out if scene is None else out[0]

    def close(self):
        self.sim.close()
