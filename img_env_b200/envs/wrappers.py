"""The reference's wrapper stack (envs/wrapper/base.py, filter_states.py), vectorised.

Same class names, constructor signature `(env, cfg)`, registry `wrapper_dict` and per-robot semantics
as the reference, but every wrapper works on whole arrays (torch CUDA tensors straight from the
library, or numpy arrays) instead of per-robot Python loops, and understands an env that batches
several scenes: arrays have N = S*R rows, scene-major, and `reset(scene_ids=[...])` re-initialises
only the listed scenes (NeverStopWrapper auto-resets a scene once all of its robots are done).
Wrappers that only exist for ROS bags / real robots / evaluation harnesses are out of scope
 and resolve to a pass-through.
"""
import math
import time

import numpy as np

from .action import ContinuousAction, DiscreteActions


# --------------------------------------------------------------------------------------------
# tiny array shim: the same code runs on numpy arrays and torch tensors
# --------------------------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _where(c, a, b):
    if _is_torch(c):
        import torch
        a = a if _is_torch(a) else torch.as_tensor(a, device=c.device)
        b = b if _is_torch(b) else torch.as_tensor(b, device=c.device)
        return torch.where(c, a, b)
    return np.where(c, a, b)


def _zeros_like(x):
    if _is_torch(x):
        import torch
        return torch.zeros_like(x)
    return np.zeros_like(x)


def _cat(xs, axis):
    if _is_torch(xs[0]):
        import torch
        return torch.cat(xs, dim=axis)
    return np.concatenate(xs, axis=axis)


def _f64(x):
    return x.double() if _is_torch(x) else np.asarray(x, dtype=np.float64)


def _i64(x):
    return x.long() if _is_torch(x) else np.asarray(x).astype(np.int64)


def _sqrt(x):
    return x.sqrt() if _is_torch(x) else np.sqrt(x)


class Wrapper:
    """Minimal stand-in for gym.Wrapper (gym is not a dependency)."""

    def __init__(self, env, cfg=None):
        self.env, self.cfg = env, cfg

    def __getattr__(self, name):
        return getattr(self.env, name)

    def __len__(self):
        return len(self.env)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)

    # State carried between steps is normally REBOUND to fresh arrays (a caller may keep what a step returned).  Under CUDA-graph
    # capture (GraphedStep) it has to live at fixed addresses instead: _keep() then updates the old tensor in place.
    _inplace = False

    def _keep(self, old, new):
        if self._inplace and old is not None and _is_torch(old) and old.shape == new.shape and old.dtype == new.dtype:
            old.copy_(new)
            return old
        return new

    # helpers for batched scenes
    def _robots_per_scene(self):
        return int(self.cfg["robot"]["total"])

    def _row_mask(self, like, kwargs):
        """bool mask (same array family as `like`) of the robots whose scene is being reset; None = every row.
        NeverStopWrapper hands its device-side mask down as `row_mask` so that a partial reset costs no host sync."""
        m = kwargs.get("row_mask")
        if m is None:
            ids = kwargs.get("scene_ids")
            if ids is None:
                return None
            m = np.zeros(like.shape[0], bool)
            m.reshape(-1, self._robots_per_scene())[list(ids)] = True
        if _is_torch(like) and not _is_torch(m):
            import torch
            m = torch.as_tensor(m, device=like.device)
        elif not _is_torch(like) and _is_torch(m):
            m = m.cpu().numpy()
        return m


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        state, reward, done, info = self.env.step(action)
        return self.observation(state), reward, done, info

    def observation(self, state):
        return state


class StatePedVectorWrapper(ObservationWrapper):
    """base.py:19-34: (x - avg) / std on the first ped_tmp[0] pedestrians' 7-vectors."""
    avg = [0.0, 0.0, 0.0, 0.0, 0.25, 0.25, 0.0]
    std = [6.0, 6.0, 0.6, 0.9, 0.50, 0.5, 6.0]

    def observation(self, state):
        # out of place: in torch mode ped_vector_states may be a view of the library's output buffer, and rows of scenes that
        # a partial reset did not touch would otherwise be normalised a second time
        p = state.ped_vector_states
        p = p.clone() if _is_torch(p) else np.array(p, copy=True)
        n = p.shape[0]
        k = (p.shape[1] - 1) // 7
        body = p[:, 1:1 + 7 * k].reshape(n, k, 7)
        if _is_torch(p):
            import torch
            avg = torch.tensor(self.avg, dtype=torch.float64, device=p.device); std = torch.tensor(self.std, dtype=torch.float64, device=p.device)
            idx = torch.arange(k, device=p.device)[None, :] < p[:, :1]
            norm = ((body.double() - avg) / std).to(p.dtype)
            p[:, 1:1 + 7 * k] = torch.where(idx[..., None], norm, body).reshape(n, 7 * k)
        else:
            idx = np.arange(k)[None, :] < p[:, :1]
            norm = ((body.astype(np.float64) - np.array(self.avg)) / np.array(self.std)).astype(p.dtype)
            p[:, 1:1 + 7 * k] = np.where(idx[..., None], norm, body).reshape(n, 7 * k)
        state.ped_vector_states = p
        return state


class VelActionWrapper(Wrapper):
    """base.py:37-66. Discrete: index -> (v, w, beep) table; continuous: per-component clip to cfg ranges."""

    def __init__(self, env, cfg):
        super().__init__(env, cfg)
        self.discrete = bool(cfg["discrete_action"])
        if self.discrete:
            self.actions = DiscreteActions(cfg["discrete_actions"])
            self.table = np.array([a.reverse() for a in self.actions.actions], dtype=np.float32)
        else:
            self.clip_range = cfg["continuous_actions"]

    def _action_torch(self, actions):
        import torch
        actions = actions.detach()
        if self.discrete and actions.ndim == 1:
            if getattr(self, "_table_t", None) is None or self._table_t.device != actions.device:
                self._table_t = torch.as_tensor(self.table, device=actions.device)
            return self._table_t[actions.long()]
        out = torch.zeros((actions.shape[0], 3), dtype=torch.float32, device=actions.device)
        if self.discrete:
            out[:, : actions.shape[1]] = actions
        else:
            for i in range(actions.shape[1]):
                out[:, i] = actions[:, i].clamp(self.clip_range[i][0], self.clip_range[i][1])
        return out

    def action(self, actions):
        """-> float32 array [N, 3] of (v, w, beep); torch actions stay on their device (no host round trip)"""
        if _is_torch(actions):
            return self._action_torch(actions)
        actions = np.asarray(actions)
        if self.discrete:
            if actions.ndim == 1:
                return self.table[actions.astype(np.int64)]
            out = np.zeros((actions.shape[0], 3), np.float32); out[:, : actions.shape[1]] = actions
            return out
        out = np.zeros((actions.shape[0], 3), np.float32)
        for i in range(actions.shape[1]):
            out[:, i] = np.clip(actions[:, i], self.clip_range[i][0], self.clip_range[i][1])
        return out

    def step(self, action):
        a = self.action(action)
        state, reward, done, info = self.env.step(a)
        info["speeds"] = _f64(a[:, :2])
        return state, reward, done, info

    def reverse_action(self, actions):
        return actions


class MultiRobotCleanWrapper(Wrapper):
    """base.py:69-95: a robot that finished keeps stepping but its reward / speeds are masked afterwards."""

    def __init__(self, env, cfg):
        super().__init__(env, cfg)
        self.is_clean = None

    def step(self, action):
        state, reward, done, info = self.env.step(action)
        if self.is_clean is None:
            self.is_clean = _zeros_like(done) == 0
        clean = self.is_clean          # never mutated in place: every update below rebinds self.is_clean to a new array
        info["is_clean"] = clean
        reward = _where(clean, reward, _zeros_like(reward))
        if "speeds" in info:
            sp = info["speeds"]
            if _is_torch(sp) != _is_torch(clean):       # actions and state live in different array families
                sp = sp.cpu().numpy() if _is_torch(sp) else sp
                c = clean.cpu().numpy() if _is_torch(clean) else clean
                info["speeds"] = np.where(c[:, None], sp, 0.0)
            else:
                if _is_torch(sp) and sp.device != clean.device:
                    sp = sp.to(clean.device)
                info["speeds"] = _where(clean[:, None], sp, _zeros_like(sp))
        self.is_clean = self._keep(clean, _where(done > 0, _zeros_like(clean), clean))
        return state, reward, done, info

    def reset(self, **kwargs):
        state = self.env.reset(**kwargs)
        if self.is_clean is not None:
            m = self._row_mask(self.is_clean, kwargs)
            self.is_clean = self._keep(self.is_clean, (self.is_clean | m) if m is not None else (_zeros_like(self.is_clean) == 0))
        return state


class StateBatchWrapper(Wrapper):
    """base.py:97-150: frame stacking of sensor_maps / vector_states / lasers (zeros until the queue is full)."""

    def __init__(self, env, cfg):
        super().__init__(env, cfg)
        self.depth = {"sensor_maps": cfg["image_batch"] if cfg["image_batch"] > 0 else None,
                      "vector_states": cfg["state_batch"] if cfg["state_batch"] > 0 else None,
                      "lasers": max(cfg["laser_batch"], 1) if cfg["laser_batch"] >= 0 else None}
        self.q = {}

    def _concate(self, name, t, reset_mask=None):
        """push `t`; with a reset_mask only the masked rows restart (cleared, then pushed) and the others keep their queue."""
        k = self.depth[name]
        if k is None:
            return t
        t1 = t[:, None]
        if name not in self.q:
            self.q[name] = _cat([_zeros_like(t1)] * k, 1)
        q = self.q[name]
        if reset_mask is None:
            q = _cat([q[:, 1:], t1], 1)
        else:
            m = reset_mask.reshape((-1,) + (1,) * (q.ndim - 1))
            q = _where(m, _cat([_zeros_like(q[:, 1:]), t1], 1), q)
        old = self.q[name]
        self.q[name] = self._keep(old, q)          # normally a fresh array every call (cat / where): the caller may keep the returned one
        return self.q[name].clone() if self.q[name] is old and self._inplace else q

    def batch_state(self, state, reset_mask=None):
        state.sensor_maps = self._concate("sensor_maps", state.sensor_maps, reset_mask)
        tmp = self._concate("vector_states", state.vector_states, reset_mask)
        if self.depth["vector_states"] is not None:
            tmp = tmp.reshape(tmp.shape[0], tmp.shape[1] * tmp.shape[2])
        state.vector_states = tmp
        state.lasers = self._concate("lasers", state.lasers, reset_mask)
        return state

    def step(self, action):
        state, reward, done, info = self.env.step(action)
        return self.batch_state(state), reward, done, info

    def reset(self, **kwargs):
        state = self.env.reset(**kwargs)
        m = self._row_mask(state.sensor_maps, kwargs)
        if m is None:
            self.q = {}
        # partial reset: only the listed scenes restart their queues; the other rows keep showing their current stack
        return self.batch_state(state, reset_mask=m)


class SensorsPaperRewardWrapper(Wrapper):
    """base.py:152-190 (Sensors-20 reward), one expression over all robots."""

    def __init__(self, env, cfg):
        super().__init__(env, cfg)
        self.ped_safety_space = cfg["ped_safety_space"]

    def reward(self, reward, states):
        md = _f64(states.ped_min_dists)
        vs = _f64(states.vector_states)
        coll = _i64(states.is_collisions) > 0
        arr = _i64(states.is_arrives) > 0
        step_d = _f64(states.step_ds)
        zero = _zeros_like(md)
        collision_reward = _where(md <= self.ped_safety_space, -50 * (self.ped_safety_space - md), zero)
        collision_reward = _where(coll, zero - 500.0, collision_reward)
        d = _sqrt(vs[:, 0] ** 2 + vs[:, 1] ** 2)
        reached = (d < 0.3) | arr
        reach_reward = _where(~coll & reached, zero + 500.0, zero)
        moving = ~coll & ~reached
        distance_reward = _where(moving, step_d * 200, zero)
        step_reward = _where(moving, zero - 5.0, zero)
        return collision_reward + reach_reward + step_reward + distance_reward

    def step(self, action):
        states, reward, done, info = self.env.step(action)
        return states, self.reward(reward, states), done, info


class NeverStopWrapper(Wrapper):
    """base.py:193-211: reset as soon as every robot (of a scene) is done; must be the outermost wrapper."""

    def step(self, action):
        states, reward, done, info = self.env.step(action)
        ad = info["all_down"]
        r = self._robots_per_scene()
        per_scene = ad.reshape(-1, r)[:, 0]
        if _is_torch(per_scene) and getattr(self.env, "masked_reset", False):
            # device-side auto-reset: the scenes to restart are selected by a device mask, every step, without reading it
            # (an all-false mask is a few empty launches); the stack below restarts its per-row state from the same mask
            states = self.env.reset(scene_mask=per_scene, row_mask=ad)
            return states, reward, done, info
        if _is_torch(per_scene):        # one small read per step; the scene list is only fetched when something ended
            scenes = per_scene.nonzero().flatten().tolist() if bool(per_scene.any()) else []
        else:
            scenes = np.flatnonzero(np.asarray(per_scene)).tolist()
        if scenes:
            if len(scenes) == per_scene.shape[0]:
                states = self.env.reset(**{k: v for k, v in info.items() if k == "dones_info"})
            else:
                states = self.env.reset(scene_ids=scenes, row_mask=info["all_down"])
        return states, reward, done, info


class TimeLimitWrapper(Wrapper):
    """base.py:214-230: done / dones_info=10 once a robot's episode ran longer than cfg time_max."""

    def __init__(self, env, cfg):
        super().__init__(env, cfg)
        self._max_episode_steps = cfg["time_max"]
        self._elapsed_steps = None

    def step(self, ac):
        observation, reward, done, info = self.env.step(ac)
        if self._elapsed_steps is None:
            self._elapsed_steps = _i64(_zeros_like(done))
        self._elapsed_steps = self._keep(self._elapsed_steps, self._elapsed_steps + 1)
        over = self._elapsed_steps > self._max_episode_steps
        done = _where(over, _zeros_like(done) + 1, done)
        info["dones_info"] = _where(over, _zeros_like(info["dones_info"]) + 10, info["dones_info"])
        return observation, reward, done, info

    def reset(self, **kwargs):
        if self._elapsed_steps is not None:
            m = self._row_mask(self._elapsed_steps, kwargs)
            self._elapsed_steps = self._keep(self._elapsed_steps, _zeros_like(self._elapsed_steps) if m is None else
                                             _where(m, _zeros_like(self._elapsed_steps), self._elapsed_steps))
        return self.env.reset(**kwargs)


class InfoLogWrapper(Wrapper):
    """base.py:233-254."""

    def __init__(self, env, cfg):
        super().__init__(env, cfg)
        self.robot_total = cfg["robot"]["total"]
        self.ped = cfg["ped_sim"]["total"] > 0 and cfg["env_type"] == "robot_nav"

    def step(self, action):
        states, reward, done, info = self.env.step(action)
        coll, arr = _i64(states.is_collisions), _i64(states.is_arrives)
        info["arrive"] = states.is_arrives
        info["collision"] = states.is_collisions
        di = _where(coll > 0, coll, _i64(info["dones_info"]))
        info["dones_info"] = _where(arr == 1, _zeros_like(di) + 5, di)
        r = self.robot_total
        down = (done > 0).reshape(-1, r)
        if _is_torch(down):
            per_scene = down.sum(1) == r
            info["all_down"] = per_scene[:, None].expand(-1, r).reshape(-1)
        else:
            per_scene = down.sum(1) == r
            info["all_down"] = np.repeat(per_scene, r)
        if self.ped:
            info["bool_get_close_to_human"] = _where(states.ped_min_dists < 1, _zeros_like(coll) + 1, _zeros_like(coll))
        return states, reward, done, info


class ObsStateTmp(ObservationWrapper):
    """filter_states.py:6-12"""

    def observation(self, states):
        return [states.sensor_maps, states.vector_states, states.ped_maps]


class ObsLaserStateTmp(ObservationWrapper):
    """filter_states.py:15-20"""

    def observation(self, states):
        return [states.lasers, states.vector_states, states.ped_maps]


class TimeControlWrapper(Wrapper):
    """base.py:301-312: wall-clock pacing to control_hz."""

    def step(self, action):
        start = time.time()
        out = self.env.step(action)
        while time.time() - start < self.cfg["control_hz"]:
            time.sleep(0.02)
        return out


class _PassThrough(Wrapper):
    """ROS-bag recording, real-robot and evaluation harness wrappers: out of scope, kept loadable."""


wrapper_dict = {
    "StatePedVectorWrapper": StatePedVectorWrapper,
    "VelActionWrapper": VelActionWrapper,
    "StateBatchWrapper": StateBatchWrapper,
    "SensorsPaperRewardWrapper": SensorsPaperRewardWrapper,
    "NeverStopWrapper": NeverStopWrapper,
    "ObsStateTmp": ObsStateTmp,
    "TimeLimitWrapper": TimeLimitWrapper,
    "MultiRobotCleanWrapper": MultiRobotCleanWrapper,
    "InfoLogWrapper": InfoLogWrapper,
    "ObsLaserStateTmp": ObsLaserStateTmp,
    "TimeControlWrapper": TimeControlWrapper,
    "BagRecordWrapper": _PassThrough,
    "TestEpisodeWrapper": _PassThrough,
    "BarnDataSetWrapper": _PassThrough,
    "RealTestRecoderWrapper": _PassThrough,
    "PedTrajectoryDatasetWrapper": _PassThrough,
}


class GraphedStep:
    """One env.step() of a wrapper stack captured into a CUDA graph and replayed: the ~100 small launches of the simulator and of
    the vectorised wrappers cost one graph launch.  Needs the device-side auto-reset (ImageEnv.masked_reset: no host decision
    inside a step) and torch actions of a fixed shape.  Everything step() returns lives in static buffers that the next replay
    overwrites.  Between replays the episode queues are topped up (host sampler, non-blocking).

        env = make_env(cfg, num_scenes=S); env.reset()
        fast = GraphedStep(env, example_actions)
        obs, reward, done, info = fast.step(actions)
    """

    def __init__(self, env, example_actions, warmup=3):
        import torch
        self.torch = torch
        self.env = env
        base = env
        while isinstance(base, Wrapper):
            base._inplace = True
            base = base.env
        self.base = base
        if not getattr(base, "masked_reset", False):
            raise RuntimeError("GraphedStep needs the device-side auto-reset (native sampler, torch state)")
        if base.copy_state is False:
            raise RuntimeError("GraphedStep needs copy_state=True (captured steps must not alias the library's output buffers)")
        self.actions = example_actions.detach().clone().to(base.sim.device)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):              # lazily created wrapper state comes into being before the capture
            for _ in range(warmup):
                env.step(self.actions)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        base.sim.autoreset_refill()
        base._capturing = True
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = env.step(self.actions)
        base._capturing = False

    def step(self, actions):
        self.actions.copy_(actions, non_blocking=True)
        self.base.sim.autoreset_refill()           # host sampler tops the episode queues up; never blocks on the device
        self.graph.replay()
        return self.out
