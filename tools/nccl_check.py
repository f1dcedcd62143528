"""Run under torchrun with N ranks: K sharded over N GPUs must reproduce the unsharded command, with either exchange:
    EXCHANGE=nccl  m3p2i_comm_init: NCCL all-gather of the discounted costs + all-reduce of the packed partial sums
    EXCHANGE=peer  m3p2i_peer_export / m3p2i_peer_attach: stores into peer HBM (cudaIpc over NVLink) issued by the
                   rollout and weighted-sum kernels themselves, no collective call (default)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "m3p2i-aip_b200"))
sys.path.insert(0, ROOT)
from m3p2i_b200 import _abi as A, native, scene as S, sharded  # noqa: E402
import bench  # noqa: E402


def make(task, mm, K, T, Kl, off, dev):
    cfg = S.make_cfg("panda_env", task, None, K, T, multi_modal=mm, cube_on_shelf=mm)
    c = S.build_config(cfg, num_samples_local=Kl, sample_offset=off, noise_mode=A.NOISE_PHILOX, seed=3)
    p = native.NativePlanner(c, S.build_panda_scene(), device=dev)
    p.set_filter_matrix(S.savgol_matrix(T))
    actors = S.default_actors("panda_env")
    dof, root = S.initial_dof_state(actors), S.initial_root_state(actors, mm)
    _, _, goal = bench.scene_inputs()
    p.set_state(dof, root)
    p.set_objective(task, goal if task == "pick" else np.zeros(7, np.float32), "close" if task == "pick" else "open")
    return p


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ok = True
    exchange = os.environ.get("EXCHANGE", "peer")
    for task, mm in (("pick", False), ("reach", True)):
        K, T = 1024 * world, 16
        Kl = K // world
        p = make(task, mm, K, T, Kl, rank * Kl, lr)
        if exchange == "peer":
            sharded.attach_peers(p)
        else:
            uid = [native.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            p.comm_init(rank, world, uid[0])
        outs = []
        for _ in range(3):
            a, c, info = p.command()
            outs.append((a.copy(), c.copy()))
        if rank == 0:
            full = make(task, mm, K, T, K, 0, lr)
            for i in range(3):
                a, c, _ = full.command()
                da = np.abs(outs[i][0] - a).max()
                dc = np.abs(outs[i][1] - c[:Kl]).max()
                good = da < 1e-5 and dc < 1e-3
                ok &= good
                print(f"{task} mm={mm} tick {i}: max|action diff|={da:.2e} max|cost_total diff|={dc:.2e} {'OK' if good else 'MISMATCH'}",
                      flush=True)
            full.close()
        p.close()
        dist.barrier()
    if rank == 0:
        print(f"sharded command over {world} GPUs, exchange={exchange}:", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
