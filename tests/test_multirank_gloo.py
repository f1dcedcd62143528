"""N>1 host logic on CPU: two processes (gloo), each owning half of the K samples, must reproduce the unsharded
tick. The shards here are oracle-backed (no GPU in this test); the same ShardedPlanner drives native shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_py as O
from helpers import make_backend
from m3p2i_b200 import _abi as A
from m3p2i_b200 import scene as S
from m3p2i_b200.sharded import ShardedPlanner, shard_bounds


def _inputs(env, task, mm):
    actors = S.default_actors(env)
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors, mm and env == "panda_env").copy()
    if env == "point_env":
        dof[0], dof[2] = 0.3, 2.4
        goal = np.array([-3.75, -3.75], np.float32)
    else:
        goal = np.zeros(7, np.float32)
    grip = {"reach": "open", "pick": "close"}.get(task)
    return dof, root, goal, grip


def _worker(rank, world, port, env, task, mm, K, T, noise_mode, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O.set_threads(1)
    cfg = S.make_cfg(env, task, None, K, T, multi_modal=mm, cube_on_shelf=mm and env == "panda_env")
    Kl, off = shard_bounds(K, rank, world)
    b = make_backend(O.Oracle, cfg, noise_mode=noise_mode, seed=5, K_local=Kl, offset=off)
    dof, root, goal, grip = _inputs(env, task, mm)
    if noise_mode == A.NOISE_TABLE:
        delta = np.random.default_rng(1).standard_normal((K, T, b.nu)).astype(np.float32)
        b.set_noise_table(delta[off:off + Kl])
        b.set_noise_row0(delta[0])
    sp = ShardedPlanner(b)
    res = []
    for _ in range(3):
        b.set_state(dof, root)
        b.set_objective(task, goal, grip)
        a, c, info = sp.command()
        res.append((np.array(a), np.array(c)))
    if rank == 0:
        np.savez(out, **{f"a{i}": r[0] for i, r in enumerate(res)}, **{f"c{i}": r[1] for i, r in enumerate(res)})
    dist.destroy_process_group()


@pytest.mark.parametrize("env,task,mm,noise", [("point_env", "push_pull", True, A.NOISE_PHILOX),
                                               ("panda_env", "reach", False, A.NOISE_TABLE),
                                               ("panda_env", "reach", True, A.NOISE_PHILOX)])
def test_two_rank_gloo_matches_single(tmp_path, env, task, mm, noise):
    K, T, world = 64, 12, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, port, env, task, mm, K, T, noise, out), nprocs=world, join=True)
    got = np.load(out)
    cfg = S.make_cfg(env, task, None, K, T, multi_modal=mm, cube_on_shelf=mm and env == "panda_env")
    o = make_backend(O.Oracle, cfg, noise_mode=noise, seed=5)
    dof, root, goal, grip = _inputs(env, task, mm)
    if noise == A.NOISE_TABLE:
        o.set_noise_table(np.random.default_rng(1).standard_normal((K, T, o.nu)).astype(np.float32))
    for i in range(3):
        o.set_state(dof, root)
        o.set_objective(task, goal, grip)
        a, c, _ = o.command()
        np.testing.assert_allclose(got[f"a{i}"], a, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got[f"c{i}"], c[: K // world], rtol=1e-5, atol=1e-4)
