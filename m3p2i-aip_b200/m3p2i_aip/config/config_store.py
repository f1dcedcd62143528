"""Structured configuration schema of the planner stack.

`ExampleConfig` carries the attribute surface that MPPI / M3P2I (cfg.env_type, cfg.multi_modal, cfg.mppi.*), the
Objective (cfg.kp_suction, cfg.pre_height_diff, cfg.mppi.num_samples) and the scripts (cfg.isaacgym, cfg.task, cfg.goal,
cfg.cube_on_shelf, cfg.suction_active) read; same field names, types and defaults as the schema the reference registers
(config/config_store.py:7-23), so YAML trees written for it validate here. The class is built from a table, and hydra
is optional: without hydra-core any attribute-style object (e.g. m3p2i_b200.scene.make_cfg) serves as cfg.
"""
import dataclasses
from typing import List

from m3p2i_aip.planners.motion_planner.mppi import MPPIConfig
from m3p2i_aip.utils.isaacgym_utils.isaacgym_wrapper import IsaacGymConfig

_REQUIRED = dataclasses.MISSING
# (attribute, type, default) -- _REQUIRED: must come from the YAML / caller
_SCHEMA = (
    ("render", bool, _REQUIRED), ("n_steps", int, _REQUIRED),
    ("mppi", MPPIConfig, _REQUIRED), ("isaacgym", IsaacGymConfig, _REQUIRED),
    ("env_type", str, _REQUIRED), ("task", str, _REQUIRED), ("goal", List[float], _REQUIRED), ("nx", int, _REQUIRED),
    ("actors", List[str], _REQUIRED), ("initial_actor_positions", List[List[float]], _REQUIRED),
    ("kp_suction", int, 0),            # gain of the suction force pair (skill_utils.py:86-90)
    ("suction_active", bool, False),   # pull preference of the last tick (sim.py:47-50)
    ("multi_modal", bool, False),
    ("pre_height_diff", float, 0.0),   # height of the pre-grasp / pre-place pose above the cube
    ("cube_on_shelf", bool, False),
)

ExampleConfig = dataclasses.make_dataclass(
    "ExampleConfig",
    [(n, t) if d is _REQUIRED else (n, t, dataclasses.field(default=d)) for n, t, d in _SCHEMA])
ExampleConfig.__module__ = __name__

_HYDRA_NODES = (("config_point", None, ExampleConfig), ("config_panda", None, ExampleConfig),
                ("base_mppi", "mppi", MPPIConfig), ("base_isaacgym", "isaacgym", IsaacGymConfig))


def register_with_hydra():
    """Make the schema known to hydra's ConfigStore under the names the reference's YAML `defaults:` lists use.
    Returns the store, or None when hydra-core is not installed."""
    try:
        from hydra.core.config_store import ConfigStore
    except ImportError:
        return None
    store = ConfigStore.instance()
    for name, group, node in _HYDRA_NODES:
        store.store(name=name, node=node) if group is None else store.store(group=group, name=name, node=node)
    return store


cs = register_with_hydra()
