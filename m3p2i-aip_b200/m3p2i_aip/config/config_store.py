"""Structured configuration (reference: config/config_store.py:7-29). hydra is optional: the dataclasses are plain
Python; when hydra-core is installed the same ConfigStore registrations as the reference are made."""
from dataclasses import dataclass
from typing import List

from m3p2i_aip.planners.motion_planner.mppi import MPPIConfig
from m3p2i_aip.utils.isaacgym_utils.isaacgym_wrapper import IsaacGymConfig


@dataclass
class ExampleConfig:
    render: bool
    n_steps: int
    mppi: MPPIConfig
    isaacgym: IsaacGymConfig
    env_type: str
    task: str
    goal: List[float]
    nx: int
    actors: List[str]
    initial_actor_positions: List[List[float]]
    kp_suction: int = 0
    suction_active: bool = False
    multi_modal: bool = False
    pre_height_diff: float = 0.
    cube_on_shelf: bool = False


try:  # pragma: no cover - hydra is not part of this image
    from hydra.core.config_store import ConfigStore
    cs = ConfigStore.instance()
    cs.store(name="config_point", node=ExampleConfig)
    cs.store(name="config_panda", node=ExampleConfig)
    cs.store(group="mppi", name="base_mppi", node=MPPIConfig)
    cs.store(group="isaacgym", name="base_isaacgym", node=IsaacGymConfig)
except ImportError:
    cs = None
