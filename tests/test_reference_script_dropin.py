"""scripts/reactive_tamp.py of the reference, loaded UNCHANGED from the reference checkout, runs on this package.

Only in the build container (needs /root/reference; skipped elsewhere). hydra and zerorpc are third-party packages
that are not installed here: they are replaced by import stubs (the REACTIVE_TAMP class never calls them; only the
`run_reactive_tamp` entry point does). Everything on the hot path resolves to this repo, the task planner resolves
to the reference's own files through M3P2I_REFERENCE_SRC. The CPU variant swaps the native backend for the oracle
(no GPU in the build container); the -m gpu variant would use the native backend but also needs the checkout.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not available")


@pytest.fixture()
def reactive_tamp(monkeypatch):
    monkeypatch.setenv("M3P2I_REFERENCE_SRC", os.path.join(REF, "src"))
    saved = {k: v for k, v in sys.modules.items() if k == "m3p2i_aip" or k.startswith("m3p2i_aip.")}
    for k in saved:
        del sys.modules[k]
    hydra = types.ModuleType("hydra")
    hydra.main = lambda **kw: (lambda f: f)
    monkeypatch.setitem(sys.modules, "hydra", hydra)
    monkeypatch.setitem(sys.modules, "zerorpc", types.ModuleType("zerorpc"))
    import oracle_py as O
    from m3p2i_b200 import native
    monkeypatch.setattr(native.NativePlanner, "for_sim", classmethod(lambda cls, sim, cfg=None, **kw: O.Oracle.for_sim(sim, cfg, **kw)))
    spec = importlib.util.spec_from_file_location("ref_reactive_tamp", os.path.join(REF, "scripts", "reactive_tamp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    yield mod
    for k in [k for k in sys.modules if k == "m3p2i_aip" or k.startswith("m3p2i_aip.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_module_resolution(reactive_tamp):
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert reactive_tamp.m3p2i.__file__.startswith(here)            # hot path: this repo
    assert reactive_tamp.Objective.__module__ == "m3p2i_aip.planners.motion_planner.cost_functions"
    assert sys.modules[reactive_tamp.Objective.__module__].__file__.startswith(here)
    assert reactive_tamp.wrapper.__file__.startswith(here)
    assert reactive_tamp.task_planner.__file__.startswith(REF)      # out of scope: the reference's own file


@pytest.mark.parametrize("task,goal,mm", [("navigation", [-3.0, 3.0], False), ("push_pull", [-3.75, -3.75], True)])
def test_reactive_tamp_point_env_unchanged(reactive_tamp, task, goal, mm):
    from m3p2i_b200 import scene as S
    from m3p2i_aip.utils.data_transfer import bytes_to_torch, torch_to_bytes
    cfg = S.make_cfg("point_env", task, goal, 64, 12, multi_modal=mm)
    tamp = reactive_tamp.REACTIVE_TAMP(cfg)
    tamp.motion_planner.delta = torch.from_numpy(np.random.default_rng(0).standard_normal((64, 12, 2)).astype(np.float32))
    actors = S.default_actors("point_env")
    dof = torch.from_numpy(S.initial_dof_state(actors)).view(1, -1)
    root = torch.from_numpy(S.initial_root_state(actors)).view(1, -1, 13)
    for _ in range(3):
        action = bytes_to_torch(tamp.run_tamp(torch_to_bytes(dof), torch_to_bytes(root)))
        assert action.shape == (2,) and torch.isfinite(action).all()
    assert bytes_to_torch(tamp.get_trajs()).shape == (20, 12, 2)
    assert bytes_to_torch(tamp.get_suction()) in (0, 1, True, False)
    # first tick of navigation heads towards the goal
    if task == "navigation":
        assert action[0] < 0 and action[1] > 0


def test_reactive_tamp_panda_unchanged(reactive_tamp, capsys):
    """config_panda: the reference's active-inference task planner picks the task; the planner tick runs natively."""
    from m3p2i_b200 import scene as S
    from m3p2i_aip.utils.data_transfer import bytes_to_torch, torch_to_bytes
    cfg = S.make_cfg("panda_env", "reactive_pick", None, 32, 12)
    tamp = reactive_tamp.REACTIVE_TAMP(cfg)
    tamp.motion_planner.delta = torch.from_numpy(np.random.default_rng(0).standard_normal((32, 12, 9)).astype(np.float32))
    actors = S.default_actors("panda_env")
    dof = torch.from_numpy(S.initial_dof_state(actors)).view(1, -1)
    root = torch.from_numpy(S.initial_root_state(actors)).view(1, -1, 13)
    for _ in range(2):
        action = bytes_to_torch(tamp.run_tamp(torch_to_bytes(dof), torch_to_bytes(root)))
    assert action.shape == (9,) and torch.isfinite(action).all()
    assert tamp.task_planner.task == "reach"          # cube far from the gripper: the AIF planner asks for "reach"
    assert (action[7:] > 1.4).all()                   # gripper_command "open" (m3p2i.py:10-12, mppi.py:412-414)


def test_host_skill_utils_match_reference():
    """The host-side quaternion costs of this repo against the reference's own functions on random quaternions."""
    spec = importlib.util.spec_from_file_location("ref_skill_utils", os.path.join(REF, "src", "m3p2i_aip", "utils", "skill_utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from m3p2i_aip.utils import skill_utils as ours
    g = torch.Generator().manual_seed(0)
    a = torch.nn.functional.normalize(torch.randn(64, 4, generator=g), dim=1)
    b = torch.nn.functional.normalize(torch.randn(64, 4, generator=g), dim=1)
    assert torch.allclose(ours.quaternion_rotation_matrix(a), ref.quaternion_rotation_matrix(a), atol=1e-6)
    assert torch.allclose(ours.get_general_ori_cube2goal(a, b), ref.get_general_ori_cube2goal(a, b), atol=1e-5)
    for tilt in (0, 0.5):
        assert torch.allclose(ours.get_general_ori_ee2cube(a, b, tilt), ref.get_general_ori_ee2cube(a, b, tilt), atol=1e-5)


class _EpisodeOver(Exception):
    pass


def test_sim_and_reactive_tamp_loop_unchanged(reactive_tamp, monkeypatch, capsys):
    """The reference's two-process loop, both scripts UNCHANGED: scripts/sim.py (the "real world", sim.py:19-58) drives
    scripts/reactive_tamp.py's REACTIVE_TAMP through the reference's RPC method names (run_tamp / get_suction; byte
    frames of data_transfer.py) -- the zerorpc socket is replaced by an in-process client, everything else is the
    reference's own control flow: update_dyn_obs, play_with_cube, set_dof_velocity_target_tensor,
    check_and_apply_suction, step, time_tracking. The K=1 env behind sim.py is this repo's integrator."""
    from m3p2i_b200 import scene as S
    cfg = S.make_cfg("point_env", "navigation", [1.0, -0.8], 64, 12)
    cfg.mppi.sampling_method = "halton"
    tamp = reactive_tamp.REACTIVE_TAMP(cfg)
    ticks = {"n": 0}

    class Client:
        def connect(self, addr):
            assert addr.startswith("tcp://")

        def run_tamp(self, dof_bytes, root_bytes):
            if ticks["n"] >= 60:
                raise _EpisodeOver
            ticks["n"] += 1
            return tamp.run_tamp(dof_bytes, root_bytes)

        def get_suction(self):
            return tamp.get_suction()

        def get_trajs(self):
            return tamp.get_trajs()

    sys.modules["zerorpc"].Client = Client
    spec = importlib.util.spec_from_file_location("ref_sim", os.path.join(REF, "scripts", "sim.py"))
    sim_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim_mod)
    assert sim_mod.wrapper.__file__.startswith(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    monkeypatch.setattr(sim_mod, "time_tracking", lambda t, cfg: t)    # no real-time pacing in a test
    created = []
    real_cls = sim_mod.wrapper.IsaacGymWrapper

    def make(*a, **k):
        import oracle_py as O
        created.append(real_cls(*a, backend_factory=O.Oracle.for_sim, **k))
        return created[-1]
    monkeypatch.setattr(sim_mod.wrapper, "IsaacGymWrapper", make)
    with pytest.raises(_EpisodeOver):
        sim_mod.run_sim(cfg)
    assert ticks["n"] == 60
    robot = created[0].robot_pos[0]
    d0 = float(np.linalg.norm(np.array([1.0, -0.8])))
    assert float(torch.linalg.norm(robot - torch.tensor([1.0, -0.8]))) < 0.5 * d0, robot.tolist()


def test_small_host_helpers_match_reference():
    """scale_ctrl / cost_to_go (public helpers of utils/mppi_utils.py) against the reference's own functions."""
    sys.modules.setdefault("ghalton", types.ModuleType("ghalton"))
    spec = importlib.util.spec_from_file_location("ref_mppi_utils", os.path.join(REF, "src", "m3p2i_aip", "utils", "mppi_utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from m3p2i_aip.utils import mppi_utils as ours
    g = torch.Generator().manual_seed(1)
    c = torch.rand(16, 12, generator=g) * 5
    gam = 0.95 ** torch.arange(12, dtype=torch.float32)
    assert torch.allclose(ours.cost_to_go(c.clone(), gam), ref.cost_to_go(c.clone(), gam), atol=1e-5)
    u = torch.randn(8, 5, 2, generator=g) * 3
    lo, hi = torch.tensor([-1.0, -2.0]), torch.tensor([1.5, 2.0])
    for fn in ("clamp", "clamp_rescale", "tanh", "identity"):
        assert torch.allclose(ours.scale_ctrl(u.clone(), lo, hi, fn), ref.scale_ctrl(u.clone(), lo, hi, fn), atol=1e-6), fn
