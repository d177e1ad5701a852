"""Wrapper registry. The reference's wrapper stack (envs/wrapper/base.py) is a "next" row of the scope
table; round 1 ships the pass-through registry so cfg `wrapper:` lists load, plus the
two wrappers that only touch actions/time."""


class _Wrapper:
    def __init__(self, env, cfg):
        self.env, self.cfg = env, cfg

    def __getattr__(self, name):
        return getattr(self.env, name)

    def __len__(self):
        return len(self.env)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)


class VelActionWrapper(_Wrapper):
    """envs/wrapper/base.py:37-59: discrete index -> (v, w[, beep]) table lookup, or pass continuous through."""

    def __init__(self, env, cfg):
        super().__init__(env, cfg)
        import torch
        self.discrete = bool(cfg.get("discrete_action", False))
        if self.discrete:
            tab = [list(a) + [0] * (3 - len(a)) for a in cfg["discrete_actions"]]
            self.table = torch.tensor(tab, dtype=torch.float32, device=env.sim.device)

    def step(self, action):
        import torch
        if self.discrete:
            idx = torch.as_tensor(action, device=self.env.sim.device).long().reshape(-1)
            return self.env.step(self.table[idx])
        return self.env.step(action)


class TimeLimitWrapper(_Wrapper):
    """envs/wrapper/base.py:62-79: dones |= step >= time_max (reported in info['dones_info'] as 10)."""

    def __init__(self, env, cfg):
        super().__init__(env, cfg)
        self.time_max = cfg.get("time_max", 100)
        self.t = 0

    def reset(self, **kw):
        self.t = 0
        return self.env.reset(**kw)

    def step(self, action):
        state, reward, done, info = self.env.step(action)
        self.t += 1
        if self.t >= self.time_max:
            info["dones_info"] = info["dones_info"] + (done == 0) * 10
            done = done * 0 + 1
        return state, reward, done, info


wrapper_dict = {"VelActionWrapper": VelActionWrapper, "TimeLimitWrapper": TimeLimitWrapper}
