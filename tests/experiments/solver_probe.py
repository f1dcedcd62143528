"""Three physics probes of the Panda integrator on the CPU oracle (K=1 "real world" env through the sim facade):
resting creep of cubeA over 400 ticks, squeeze-and-lift (finger creep, tangential table load, cube creep, lift), and a
closed gripper pressed down onto cubeA (does the cube tunnel through the table?).
    python tests/experiments/solver_probe.py [link_sweeps]
Used at the end of round 1 to evaluate the pass order in fixed_boxes_last.patch (see README.md here)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "m3p2i-aip_b200"), os.path.join(ROOT, "oracle"), ROOT]
import oracle_py as O  # noqa: E402
from m3p2i_b200 import scene as S  # noqa: E402
from m3p2i_aip.utils.isaacgym_utils import isaacgym_wrapper as wrapper  # noqa: E402

if len(sys.argv) > 1:
    S.PANDA_SCENE_OVERRIDES["link_sweeps"] = int(sys.argv[1])
S.PANDA_SCENE_OVERRIDES["report_cube_contacts"] = 1   # the physical table force
GRASP_Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333, 0.04, 0.04]
cfg = S.make_cfg("panda_env", "pick", None, 1, 16)


def world():
    real = wrapper.IsaacGymWrapper(cfg.isaacgym, "panda_env", num_envs=1, device="cpu", backend_factory=O.Oracle.for_sim)
    for _ in range(30):
        real.step()
    return real


def cube(real):
    return real.get_actor_link_by_name("cubeA", "box")[0, :3].clone()


real = world()
c0 = cube(real)
for _ in range(400):
    real.step()
print(f"creep over 400 ticks: {float(torch.linalg.norm(cube(real) - c0)):.6f} m")

real = world()
real._dof_state[0, 0::2] = torch.tensor(GRASP_Q)
real._dof_state[0, 1::2] = 0
real.set_dof_state_tensor(real._dof_state)
F, FT, C = [], [], []
for i in range(70):
    a = torch.zeros(1, 9)
    a[0, 7:] = -1.5
    if i >= 40:
        a[0, 1], a[0, 3] = -0.5, 0.5
    real.set_dof_velocity_target_tensor(a)
    real.step()
    F.append(real._dof_state[0, [14, 16]].clone().numpy())
    FT.append(real.get_actor_contact_forces_by_name("table", "box")[0].clone().numpy())
    C.append(cube(real).numpy())
F, FT, C = np.array(F), np.array(FT), np.array(C)
hold = slice(10, 40)
print(f"squeeze: finger min {F[hold].min():.4f} m, finger creep {np.abs(F[39] - F[10]).max():.5f} m, "
      f"table |Fx|+|Fy| max {np.abs(FT[hold, :2]).sum(1).max():.3f} N, cube creep {np.linalg.norm(C[39] - C[10]):.5f} m, "
      f"lift {C[-1, 2] - C[39, 2]:.3f} m")

real = world()
q = list(GRASP_Q)
q[7] = q[8] = 0.0
real._dof_state[0, 0::2] = torch.tensor(q)
real._dof_state[0, 1::2] = 0
real._dof_state[0, 2] -= 0.12          # start with the closed gripper 6 cm above the cube
real.set_dof_state_tensor(real._dof_state)
zmin = 9.0
for i in range(120):
    a = torch.zeros(1, 9)
    a[0, 1] = 0.6
    a[0, 7:] = -1.5
    real.set_dof_velocity_target_tensor(a)
    real.step()
    zmin = min(zmin, float(cube(real)[2]))
print(f"press: lowest cube centre {zmin:.4f} m (resting 1.0497, table top 1.025)")
