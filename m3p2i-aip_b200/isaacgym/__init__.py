"""Inert stand-in for NVIDIA IsaacGym so that `from isaacgym import gymtorch` / `gymapi` at the top of the
reference's scripts (scripts/reactive_tamp.py:1, scripts/sim.py:1) imports. Nothing here simulates anything: the
rollout environments live in libm3p2i_b200.so behind m3p2i_aip.utils.isaacgym_utils.isaacgym_wrapper."""
from . import gymapi, gymtorch  # noqa: F401
