"""Gym-style API of the reference (`envs/__init__.py:9-33`): read_yaml, make_env."""
import yaml

from .action import Action, ContinuousAction, DiscreteActions   # noqa: F401
from .env import ImageEnv                                        # noqa: F401
from .reset_helper import EnvPos, NearbyPed                      # noqa: F401
from .state import ImageState                                    # noqa: F401


from .wrappers import GraphedStep                                # noqa: F401,E402


def read_yaml(file: str) -> dict:
    with open(file, "r", encoding="utf-8") as f:
        return yaml.load(f.read(), Loader=yaml.FullLoader)


def make_env(cfg, **kwargs):
    if isinstance(cfg, str):
        cfg = read_yaml(cfg)
    if cfg.get("env_type", "robot_nav") != "robot_nav":
        raise ValueError("only env_type 'robot_nav' (ImageEnv) is implemented; gazebo/real envs are out of scope")
    from .wrappers import wrapper_dict
    unknown = [name for name in cfg.get("wrapper", []) if name not in wrapper_dict]
    if unknown:      # the reference raises KeyError (envs/__init__.py:29); a typo must not silently change reward / done semantics
        raise ValueError("unknown wrapper(s) %s; known: %s" % (unknown, sorted(wrapper_dict)))
    env = ImageEnv(cfg, **kwargs)
    for name in cfg.get("wrapper", []):
        env = wrapper_dict[name](env, cfg)
    cfg["node_id"] = cfg.get("node_id", 0) + 1
    return env
