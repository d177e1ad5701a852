"""img_env_b200 — B200-native implementation of the img_env per-step simulation hot path.

Public surface (mirrors /root/reference/envs):
    img_env_b200.envs.make_env / ImageEnv / ImageState / ContinuousAction   (Gym-style API)
    img_env_b200.lib.BatchedSim                                              (thin wrapper of the C ABI)
"""
from .spec import build_spec, load_grid, rpy_to_q   # noqa: F401
from .lib import BatchedSim, load_library          # noqa: F401

__version__ = "0.1"
