"""Action containers with the reference's public surface (envs/action/action.py): `ContinuousAction(v, w, beep)` with
`.v/.w/.beep/.reverse()` and `DiscreteActions(table)` indexable by the discrete action id.  Internally a discrete table is
also kept as one float32 array [n, 3] so that VelActionWrapper can decode a whole batch of ids with a single gather."""
from dataclasses import dataclass

import numpy as np


class Action:
    """Marker base class (kept for isinstance checks written against the reference)."""


@dataclass
class ContinuousAction(Action):
    v: float                 # forward speed (m/s)
    w: float                 # yaw rate (rad/s), or lateral speed for omni robots
    beep: float = 0          # > 0: the robot beeps (emotional ORCA pedestrians react)

    def reverse(self):
        """The request triple (v, w, v_y) in the order the step service takes it."""
        return [self.v, self.w, self.beep]


class DiscreteActions:
    """Lookup table id -> ContinuousAction built from rows (v, w) or (v, w, beep); v must not be negative."""

    def __init__(self, table):
        rows = [tuple(r) for r in table]
        bad = [r for r in rows if len(r) not in (2, 3) or r[0] < 0]
        assert not bad, "discrete actions are (v >= 0, w[, beep]) rows: %r" % (bad[:1],)
        self.array = np.array([r if len(r) == 3 else r + (0,) for r in rows], dtype=np.float32).reshape(-1, 3)
        self.actions = [ContinuousAction(*(r if len(r) == 3 else r + (0,))) for r in rows]

    def __len__(self):
        return len(self.actions)

    def __getitem__(self, index):
        return self.actions[index]

    def __iter__(self):
        return iter(self.actions)
