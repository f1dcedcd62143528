"""Closed-loop reactive pick on the native backend: reach -> pick (-> place), task switching by the thresholds of
PLANNER_AIF_PANDA (task_planner.py:57-75). Prints the trace; exit code 0 if the cube ends within 5 cm of the goal.
usage: [PICK_BACKEND=oracle] [PICK_REAL=oracle|native] [PICK_SAMPLING=halton|philox|philox-spline] python tests/experiments/pick_episode.py [K] [H] [ticks]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "m3p2i-aip_b200"), os.path.join(ROOT, "tests"), ROOT]
from m3p2i_b200 import scene as S
from m3p2i_aip.planners.motion_planner import m3p2i
from m3p2i_aip.planners.motion_planner.cost_functions import Objective
from m3p2i_aip.utils.isaacgym_utils import isaacgym_wrapper as wrapper

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16
TICKS = int(sys.argv[3]) if len(sys.argv) > 3 else 900
factory = None
if os.environ.get("PICK_BACKEND") == "oracle" or os.environ.get("PICK_REAL") == "oracle":
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    O.set_threads(os.cpu_count())
    factory = O.Oracle.for_sim if os.environ.get("PICK_BACKEND") == "oracle" else None
# PICK_REAL=oracle|native: backend of the K=1 "real world" env only (default: the planner's)
real_factory = {"oracle": lambda: O.Oracle.for_sim, "native": lambda: None}.get(os.environ.get("PICK_REAL"), lambda: factory)()


class Tamp:
    def __init__(self, cfg):
        self.sim = wrapper.IsaacGymWrapper(cfg.isaacgym, cfg.env_type, num_envs=cfg.mppi.num_samples, device="cpu", backend_factory=factory)
        self.objective = Objective(cfg)
        self.mp = m3p2i.M3P2I(cfg, dynamics=self.dynamics, running_cost=self.running_cost)
    def dynamics(self, _, u, t=None): raise AssertionError
    def running_cost(self, _): raise AssertionError
    def run_tamp(self, dof, root, task, goal):
        self.sim._dof_state[:] = dof; self.sim._root_state[:] = root
        self.sim.set_dof_state_tensor(self.sim._dof_state); self.sim.set_actor_root_state_tensor(self.sim._root_state)
        self.mp.update_gripper_command(task); self.objective.update_objective(task, goal)
        return self.mp.command(self.sim._dof_state[0])[0]


cfg = S.make_cfg("panda_env", "reach", None, K, H)
cfg.mppi.sampling_method = os.environ.get("PICK_SAMPLING", "philox" if factory is None else "halton")
tamp = Tamp(cfg)
real = wrapper.IsaacGymWrapper(cfg.isaacgym, "panda_env", num_envs=1, device="cpu", backend_factory=real_factory)
for _ in range(30):
    real.step()
task, goal = "reach", torch.zeros(7)
thr = cfg.pre_height_diff + 0.005
for i in range(TICKS):
    a = tamp.run_tamp(real._dof_state.clone(), real._root_state.clone(), task, goal)
    real.set_dof_velocity_target_tensor(a.view(1, -1)); real.step()
    lf = real.get_actor_link_by_name("panda", "panda_leftfinger")[0, :3]; rf = real.get_actor_link_by_name("panda", "panda_rightfinger")[0, :3]
    ee = 0.5 * (lf + rf)
    cube = real.get_actor_link_by_name("cubeA", "box")[0, :7].clone(); cb = real.get_actor_link_by_name("cubeB", "box")[0, :7].clone()
    d = float(torch.linalg.norm(ee - cube[:3]))
    if task == "reach" and d < thr:
        task = "pick"; goal = cb.clone(); goal[2] += thr
        print(f"tick {i}: reach done (ee-cube {d:.3f}) -> pick, goal {goal[:3].tolist()}", flush=True)
    gd = float(torch.linalg.norm(goal[:3] - cube[:3])) if task != "reach" else float("nan")
    if task == "pick" and float(torch.linalg.norm(goal[:2] - cube[:2])) < 0.03:
        task = "place"; print(f"tick {i}: pick done (cube-goal {gd:.3f}) -> place", flush=True)
    if i % 20 == 0:
        print(f"{i:4d} {task:5s} ee-cube {d:.3f} cube {[round(float(x), 3) for x in cube[:3]]} fingers {[round(float(x), 3) for x in real._dof_state[0, [14, 16]]]} goal-dist {gd:.3f}", flush=True)
    if task == "place" and float(real._dof_state[0, 14]) > 0.035:
        break
final = float(torch.linalg.norm(goal[:3] - cube[:3])) if task != "reach" else 9.9
print(f"final task {task}, cube-goal distance {final:.3f}")
sys.exit(0 if final < 0.05 else 1)
