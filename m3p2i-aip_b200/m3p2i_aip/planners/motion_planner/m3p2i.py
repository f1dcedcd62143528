"""M3P2I: the multi-modal MPPI planner (reference: planners/motion_planner/m3p2i.py:5-92).

Two half-K modes with their own means and best trajectories, three weight sets and the on-the-fly beta search
(m3p2i.py:24-92) are evaluated by the native softmin kernels; this class keeps the reference's public methods.
"""
import m3p2i_aip.planners.motion_planner.mppi as mppi


class M3P2I(mppi.MPPI):
    def __init__(self, cfg, dynamics=None, running_cost=None):
        super().__init__(cfg, dynamics, running_cost)
        self.suction_active = getattr(cfg, "suction_active", False)

    def update_gripper_command(self, task):
        if task in ["reach", "place"]:
            self.gripper_command = "open"
        elif task == "pick":
            self.gripper_command = "close"

    def get_pull_preference(self):
        if self.multi_modal:
            if self._info is None:   # before the first command() the reference's weights are all zero
                return 0
            return int(self._info.weight_pull > self._info.weight_push)
        return self.suction_active
