"""A seeded stand-in for ImageEnv used to compare wrapper stacks without a GPU."""
import numpy as np


def gen_state(rng, n, max_ped=3, rays=16):
    pv = np.zeros((n, 1 + 7 * max_ped), np.float32)
    pv[:, 0] = rng.integers(0, max_ped + 1, n)
    pv[:, 1:] = rng.normal(0, 2, (n, 7 * max_ped)).astype(np.float32)
    return dict(vector_states=rng.normal(0, 2, (n, 3)), sensor_maps=rng.random((n, 8, 8)).astype(np.float16),
                is_collisions=(rng.random(n) < 0.12).astype(np.int64) * rng.integers(1, 4, n), is_arrives=rng.random(n) < 0.1,
                lasers=rng.random((n, rays)), ped_vector_states=pv, ped_maps=rng.random((n, 3, 8, 8)).astype(np.float32),
                step_ds=rng.normal(0, 0.1, n), ped_min_dists=rng.random(n) * 2)


class FakeEnv:
    def __init__(self, state_cls, n, seed, list_actions):
        self.state_cls, self.n, self.rng, self.list_actions = state_cls, n, np.random.default_rng(seed), list_actions
        self.seen_actions = []

    def __len__(self):
        return self.n

    def _state(self):
        g = gen_state(self.rng, self.n)
        return self.state_cls(g["vector_states"], g["sensor_maps"], g["is_collisions"], g["is_arrives"], g["lasers"],
                              g["ped_vector_states"], g["ped_maps"], g["step_ds"], g["ped_min_dists"])

    def reset(self, **kwargs):
        return self._state()

    def step(self, actions):
        if self.list_actions:
            self.seen_actions.append(np.array([a.reverse() for a in actions], dtype=np.float32))
        else:
            self.seen_actions.append(np.asarray(actions, dtype=np.float32).copy())
        s = self._state()
        rewards = s.is_arrives.astype(np.int64) - s.is_collisions
        dones = np.clip(np.clip(s.is_collisions, -1, 1) + s.is_arrives, 0, 1)
        return s, rewards, dones.copy(), {"dones_info": np.zeros_like(dones)}
