"""isaacgym.gymapi stand-in: only the names the reference touches at import time."""


class SimParams:
    pass


class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z


class Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = x, y, z, w


def acquire_gym():
    raise RuntimeError("IsaacGym is not part of this build: use m3p2i_aip.utils.isaacgym_utils.isaacgym_wrapper")
