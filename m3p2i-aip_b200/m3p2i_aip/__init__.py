"""Drop-in `m3p2i_aip` package: the modules on the MPPI hot path are provided here (B200-native); every other
module of the reference (task planner, active-inference agent, plotting, ...) is out of scope and, when a reference
checkout is available, is imported from it unchanged.

Set M3P2I_REFERENCE_SRC to the reference's `src` directory (the one that contains `m3p2i_aip/`). The sub-packages
of this package then extend their __path__ with the reference's directories, so that e.g.
`m3p2i_aip.planners.task_planner.task_planner` resolves to the reference file while
`m3p2i_aip.planners.motion_planner.m3p2i` resolves to this repository.
"""
import os


def _reference_dir(*parts):
    src = os.environ.get("M3P2I_REFERENCE_SRC")
    if not src:
        return None
    d = os.path.join(src, "m3p2i_aip", *parts)
    return d if os.path.isdir(d) else None


def extend_path(path_list, *parts):
    """Append the reference's directory for sub-package `parts` to a package __path__ (our modules stay first)."""
    d = _reference_dir(*parts)
    if d and d not in path_list:
        path_list.append(d)
    return path_list


extend_path(__path__)
