#!/usr/bin/env python
"""bench.py -- sample-steps/s of one MPPI command() (the hot path of BASELINE.json) on N B200s.

A step is one planner tick: noise -> perturbation -> K x H rollout (dynamics + cost) -> softmin -> mean update.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c2|c3|c4|c4_grasp|c5] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs; K is per GPU, sharded commands run K_global = K * N):
    c1        navigation, point robot, K=200 H=12
    c2        push, point robot + block, K=1024 H=20
    c3        push_pull multi_modal, 2 x 2048, H=20
    c4        config_panda reactive pick, K=4096 H=32, arm at its initial pose        <- default at N = 1 (the headline)
    c4_grasp  the same command with the fingers closed around cubeA (every rollout is contact-rich)
    c5        config_panda multi_modal=True cube_on_shelf=True (reach), K=4096 per GPU, H=32: K_global = 32768 at N = 8
                                                                                        <- default at N > 1
Prints ONE JSON line on rank 0. `--impl reference` times the reference's own CPU implementation of the same command:
the UNMODIFIED reference Python (baseline/_ref: M3P2I.command + Objective.compute_cost, device='cpu') over the sim
facade, with the oracle's C integrator standing in for IsaacGym/PhysX (not installable), on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "m3p2i-aip_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GRASP_Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333, 0.027, 0.027]
CONFIGS = {
    "c1": dict(env="point_env", task="navigation", goal=[-3.0, 3.0], K=200, T=12, mm=False,
               workload="navigation task, point-robot 2D, K=200 H=12 single-mode (BASELINE configs[0])"),
    "c2": dict(env="point_env", task="push", goal=[-1.0, -1.0], K=1024, T=20, mm=False, robot=[0.2, 2.45],
               workload="push task, point-robot with box contact, K=1024 H=20 (BASELINE configs[1])"),
    "c3": dict(env="point_env", task="push_pull", goal=[-3.75, -3.75], K=4096, T=20, mm=True, robot=[0.3, 2.4],
               workload="push_pull multi_modal=True, 2 modes x K=2048 H=20 (BASELINE configs[2])"),
    "c4": dict(env="panda_env", task="pick", K=4096, T=32, mm=False, grip="close",
               workload="config_panda reactive pick, 7-DoF Panda + cube, K=4096 per GPU, H=32 (BASELINE configs[3])"),
    "c4_grasp": dict(env="panda_env", task="pick", K=4096, T=32, mm=False, grip="close", q=GRASP_Q,
                     workload="config_panda reactive pick, K=4096 per GPU, H=32, fingers closed around cubeA (contact-rich "
                              "state of configs[3])"),
    "c5": dict(env="panda_env", task="reach", K=4096, T=32, mm=True, shelf=True, grip="open",
               workload="config_panda multi_modal=True cube_on_shelf=True (reach), K=4096 per GPU, H=32 "
                        "(BASELINE configs[4]: K_global=32768 over 8 GPUs)"),
}
# algorithmic HBM bytes per sample-step (DESIGN.md "Algorithmic bytes"): the rollout kernel writes the action row
# (nu f32), the float4 state row and the cost; the weighted-sum pass re-reads the action row.
B_ROLLOUT = {"point_env": 4 * 2 + 16 + 4, "panda_env": 4 * 9 + 16 + 4}
B_PATH = {"point_env": 4 * 2 + 16 + 4 + 4 * 2, "panda_env": 4 * 9 + 16 + 4 + 4 * 9}


def scene_inputs(conf=None):
    """(dof_state, root_state, goal) of a workload: the reference's initial scene, cubes at rest."""
    from m3p2i_b200 import scene as S
    conf = conf or CONFIGS["c4"]
    actors = S.default_actors(conf["env"])
    shelf = bool(conf.get("shelf"))
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors, shelf).copy()
    if conf["env"] == "point_env":
        if conf.get("robot"):
            dof[0], dof[2] = conf["robot"]
        return dof, root, np.asarray(conf["goal"], np.float32)
    # cubes resting on the table, as in a running episode
    if not shelf:
        root[S.actor_index(actors, "cubeA"), 2] -= 0.0095
    root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
    if conf.get("q"):
        dof[0::2] = conf["q"]
    cb = root[S.actor_index(actors, "cubeB")]
    goal = np.concatenate([cb[:3] + np.array([0, 0, 0.055], np.float32), cb[3:7]]).astype(np.float32) \
        if conf["task"] == "pick" else np.zeros(7, np.float32)
    return dof, root, goal


class stdout_to_stderr:
    """NCCL prints its version banner on fd 1 during communicator creation; the driver wants exactly one JSON line
    on stdout, so fd 1 is pointed at stderr while communicators are being built."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons while the benchmark runs (B200_PROFILING.md): NVML polled every 50 ms in this
    process (nvidia_ml_py), or nvidia-smi every 200 ms when NVML cannot be loaded. Started BEFORE the warm-up so that
    the thread is in steady state during the timed region; the coarse period keeps it from competing with the
    launching thread for the GIL."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]   # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        except Exception:
            self.nvml = self.handle = None

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    sm = int(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                    mask = int(self.reasons_fn(self.handle))
                    self.rows.append([time.perf_counter(), str(sm), str(self.max_sm)] +
                                     ["Active" if mask & b else "Not Active" for b in self.BITS])
                    self.stop_flag.wait(0.05)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([time.perf_counter()] + [x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self, t0=None, t1=None):
        """clocks over [t0, t1] (perf_counter; the timed regions), all samples if that window caught none"""
        self.stop_flag.set()
        self.join(timeout=6)
        rows = [r for r in self.rows if t0 is None or t0 <= r[0] <= t1] or self.rows
        sm = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        mx = [int(r[2]) for r in rows if len(r) > 2 and r[2].isdigit()]
        reasons = sorted({self.NAMES[i] for r in rows for i in range(4) if len(r) > 3 + i and r[3 + i] == "Active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


NCU_SUMMARY = {"c4": "r02_ncu_rollout_far_c4_summary.csv", "c5": "r02_ncu_rollout_far_c5_summary.csv",
               "c4_grasp": "r02_ncu_rollout_team_grasp_summary.csv"}


def ncu_numbers(name):
    """(dram bytes per launch, issue-slot utilisation in percent) of the rollout kernel from the committed
    ncu --set full summary of this workload (profiles/), or (None, None)."""
    path = os.path.join(ROOT, "profiles", NCU_SUMMARY.get(name, "-"))
    if not os.path.exists(path):
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, issue = 0.0, None
    for line in open(path):
        parts = line.strip().split(",")
        if len(parts) >= 4 and parts[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(parts[-1]) * scale.get(parts[-2], 1.0)
        if len(parts) >= 4 and parts[-3] == "smsp__issue_active.avg.pct_of_peak_sustained_active":
            issue = float(parts[-1])
    return (tot or None), issue


def cpu_port_rate(conf, threads, seconds=5.0):
    """The C port of the same path (oracle/, OpenMP over samples), timed for about `seconds`."""
    import oracle_py as O
    from m3p2i_b200 import _abi as A
    from m3p2i_b200 import scene as S
    O.set_threads(threads)
    K, T = conf["K"], conf["T"]
    dof, root, goal = scene_inputs(conf)
    cfg = S.make_cfg(conf["env"], conf["task"], goal.tolist(), K, T, multi_modal=conf["mm"], cube_on_shelf=bool(conf.get("shelf")))
    scene = S.build_point_scene() if conf["env"] == "point_env" else S.build_panda_scene()
    o = O.Oracle(S.build_config(cfg, noise_mode=A.NOISE_PHILOX, seed=0), scene)
    o.set_filter_matrix(S.savgol_matrix(T))
    o.set_objective(conf["task"], goal, conf.get("grip"))
    times, t_end, n = [], time.perf_counter() + seconds, 0
    while n < 3 or time.perf_counter() < t_end:
        o.set_state(dof, root)
        t0 = time.perf_counter()
        o.command()
        if n >= 1:
            times.append(time.perf_counter() - t0)
        n += 1
        if len(times) >= 200:
            break
    o.close()
    return K * T / float(np.median(times)), len(times), float(np.median(times))


def cpu_reference_python(conf, K, threads, steps, warmup):
    """Per-command seconds of the unmodified reference Python over the facade (oracle/reference_rig.py)."""
    import reference_rig as R
    dof, root, goal = scene_inputs(conf)
    return R.time_reference(conf["env"], conf["task"], goal.tolist(), K, conf["T"], dof, root, multi_modal=conf["mm"],
                            cube_on_shelf=bool(conf.get("shelf")), steps=steps, warmup=warmup, threads=threads)


REF_KIND = "reference-python+port-dynamics"
REF_NOTE = ("unmodified reference M3P2I.command + Objective.compute_cost (baseline/_ref, torch CPU tensors, all host cores) "
            "driven as scripts/reactive_tamp.py does, over the sim facade with the oracle's C integrator standing in for "
            "IsaacGym/PhysX (closed binary, not installable); noise table injected (the first-call scipy spline loop is not timed)")


def run_reference(args, rank, world, name, conf):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    T = conf["T"]
    Kg = conf["K"] * max(args.gpus, 1)
    line = {"impl": "reference", "metric": "sample-steps/sec (K x H per command)", "unit": "sample-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "gpu_launches": 0}
    try:
        import reference_rig as R
        have_ref = R.reference_dir() is not None
    except Exception:
        have_ref = False
    if have_ref:
        # full K_global when the whole run fits in a few minutes, else one GPU's shard per step (the reference's cost
        # is linear in K): decided from one untimed command
        K_run, scope = Kg, "full K_global"
        t_probe = cpu_reference_python(conf, K_run, cores, 1, 0)[0][0]
        if t_probe * (args.steps + args.warmup) > 150.0 and Kg > conf["K"]:
            K_run, scope = conf["K"], "per-shard: one GPU's K per step (the reference's cost is linear in K)"
        times, src = cpu_reference_python(conf, K_run, cores, args.steps, args.warmup)
        dt = float(np.mean(times))
        value = K_run * T / dt
        kind, sample = REF_KIND, f"{args.steps} commands of K={K_run}, H={T} ({scope}); reference from {os.path.relpath(src, ROOT)}"
        note = REF_NOTE
        med = float(np.median(times))
    else:
        v, n, med = cpu_port_rate(conf, cores, seconds=max(5.0, 0.05 * args.steps))
        dt, value, K_run, scope = med, v, conf["K"], "per-shard"
        kind, sample = "port", f"{n} commands of K={conf['K']}, H={T}; baseline/_ref missing: C port of the path (oracle/, OpenMP)"
        note = "baseline/_ref is absent (pip install --no-deps --target baseline/_ref <reference>): timed the C port instead"
    line.update({"value": value, "ms_per_step": dt * 1e3, "ms_per_step_median": med * 1e3,
                 "config": {"workload": conf["workload"] + ("" if scope == "full K_global" else f" [{scope}]"), "name": name,
                            "K_global": Kg, "K_timed": K_run, "H": T, "note": note},
                 "cpu_baseline": {"value": value, "unit": "sample-steps/s", "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def build_planner(conf, world, rank, local_rank, exchange_arg):
    import torch.distributed as dist
    from m3p2i_b200 import _abi as A
    from m3p2i_b200 import native
    from m3p2i_b200 import scene as S
    K, T = conf["K"], conf["T"]
    Kg = K * world
    dof, root, goal = scene_inputs(conf)
    cfg = S.make_cfg(conf["env"], conf["task"], goal.tolist(), Kg, T, multi_modal=conf["mm"], cube_on_shelf=bool(conf.get("shelf")))
    cfg.mppi.sampling_method = "philox"
    c = S.build_config(cfg, num_samples_local=K, sample_offset=rank * K, noise_mode=A.NOISE_PHILOX, seed=0)
    scene = S.build_point_scene() if conf["env"] == "point_env" else S.build_panda_scene()
    planner = native.NativePlanner(c, scene, device=local_rank)
    planner.set_filter_matrix(S.savgol_matrix(T))
    exchange = "none"
    if world > 1:
        exchange = exchange_arg
        with stdout_to_stderr():
            if exchange == "peer":
                from m3p2i_b200 import sharded
                try:
                    sharded.attach_peers(planner)   # raises on every rank if any rank could not map its peers
                except RuntimeError as exc:
                    print(f"[rank {rank}] {exc}; using NCCL", file=sys.stderr)
                    exchange = "nccl (peer memory unavailable)"
            if exchange != "peer":
                uid = [native.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                planner.comm_init(rank, world, uid[0])
    planner.set_objective(conf["task"], goal, conf.get("grip"))
    planner.set_state(dof, root)
    return planner, exchange, dof, root


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="default: c4 on one GPU, c5 on several")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: exchange fused into the kernels over NVLink peer memory (default) or two NCCL collectives")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_for_default = max(world, args.gpus)
    name = args.config or ("c4" if n_for_default == 1 else "c5")
    conf = CONFIGS[name]
    if args.impl == "reference":
        run_reference(args, rank, world, name, conf)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
        args.gpus = world

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()

    K, T, env = conf["K"], conf["T"], conf["env"]
    Kg = K * world
    planner, exchange, dof, root = build_planner(conf, world, rank, local_rank, args.exchange)
    # a stream torch owns, so that torch.cuda.Event brackets exactly the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    planner.set_stream(stream.cuda_stream)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing: inputs already in HBM, CUDA events on the launching stream
    for _ in range(args.warmup):
        planner.command_resident()
    barrier()
    t_timed0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 1)  # > L2 (126 MB), outside the timed events
        ev[i][0].record(stream)
        planner.command_resident()
        ev[i][1].record(stream)
    barrier()
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)  # max over ranks, per step
    ms_per_step = float(step_ms.mean())
    ms_median = float(step_ms.median())
    # rollout-kernel duration and the waits of the peer exchange (CUDA events / globaltimer inside the library, same
    # stream), L2 flushed before each launch
    roll_ms, wait_j, wait_p = [], [], []
    for i in range(min(args.steps, 50)):
        if flush is not None:
            flush.fill_(i & 1)
        if world > 1:
            dist.barrier()   # every rank enters the command together: what is left in peer_wait_ms is rollout skew + exchange
        info = planner.command_resident(sync=True)
        roll_ms.append(info.rollout_ms)
        wait_j.append(info.peer_wait_ms[0])
        wait_p.append(info.peer_wait_ms[1])
    last = planner.command_resident(sync=True)
    launches_per_step, lanes = last.launches, int(last.rollout_lanes)
    near = int(last.near_samples)   # samples the far-field kernel left to the full rollout kernel (-1: it did not run)
    beta_iters = int(last.beta_iters)
    barrier()

    # ---------------- end to end through the public API with host buffers (H2D state in, D2H action out)
    n_e2e = args.steps
    for _ in range(3):
        planner.set_state(dof, root)
        planner.command(want_cost=False, want_info=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        planner.set_state(dof, root)                      # pinned staging + H2D inside
        act, _, _ = planner.command(want_cost=False, want_info=False)   # D2H of the action inside, synchronous
    barrier()
    t_timed1 = time.perf_counter()
    e2e_s = torch.tensor([(t_timed1 - t0) / n_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s)
    waits = torch.tensor([float(np.mean(wait_j)), float(np.mean(wait_p)), float(np.mean(roll_ms))], dtype=torch.float64, device="cuda")
    waits_all = [torch.zeros_like(waits) for _ in range(world)]
    if world > 1:
        dist.all_gather(waits_all, waits)
    else:
        waits_all = [waits]
    # N > 1: the single-GPU command of the SAME workload (one shard's K, unsharded) on rank 0's GPU, so that the weak
    # scaling of this workload can be read from this line alone (the default N = 1 line is another workload: c4)
    n1_same = None
    if world > 1:
        if rank == 0:
            solo, _, _, _ = build_planner(conf, 1, 0, local_rank, "none")
            solo.set_stream(stream.cuda_stream)
            for _ in range(args.warmup):
                solo.command_resident()
            torch.cuda.synchronize()
            n_solo = min(args.steps, 50)
            ev1 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_solo)]
            for i in range(n_solo):
                if flush is not None:
                    flush.fill_(i & 1)
                ev1[i][0].record(stream)
                solo.command_resident()
                ev1[i][1].record(stream)
            torch.cuda.synchronize()
            ms1 = float(np.mean([a.elapsed_time(b) for a, b in ev1]))
            n1_same = {"ms_per_step": ms1, "value": K * T / (ms1 * 1e-3), "steps": n_solo,
                       "note": "same workload, one GPU, K = K_per_gpu unsharded, measured on rank 0 after the sharded run"}
            solo.close()
        barrier()
    # N > 1 with the default workload (c5): the driver's N = 1 line is the N = 1 default (c4, pick). Time that workload
    # sharded over these N GPUs too, so that a cross-N ratio of like with like can be read from the lines the driver has
    n1_default = None
    if world > 1 and args.config is None and name != "c4":
        conf4 = CONFIGS["c4"]
        p4, ex4, _, _ = build_planner(conf4, world, rank, local_rank, args.exchange)
        p4.set_stream(stream.cuda_stream)
        for _ in range(args.warmup):
            p4.command_resident()
        barrier()
        n4 = min(args.steps, 50)
        ev4 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n4)]
        for i in range(n4):
            if flush is not None:
                flush.fill_(i & 1)
            ev4[i][0].record(stream)
            p4.command_resident()
            ev4[i][1].record(stream)
        barrier()
        ms4 = torch.tensor([a.elapsed_time(b) for a, b in ev4], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms4, op=dist.ReduceOp.MAX)
        m4 = float(ms4.mean())
        n1_default = {"name": "c4", "workload": conf4["workload"], "K_global": conf4["K"] * world, "ms_per_step": m4,
                      "value": conf4["K"] * world * conf4["T"] / (m4 * 1e-3), "steps": n4, "exchange": ex4,
                      "note": "the workload of the default N = 1 line (bench.py --gpus 1), sharded over these GPUs the same "
                              "way and timed the same way (device events, max over ranks, L2 flushed)"}
        p4.close()
        barrier()
    # N = 1 default line: the same command in the CONTACT-RICH state of the same config (fingers closed around cubeA: every
    # rollout is on the near list and runs the full contact solver) -- the state the planner lives in during pick --, timed
    # the same way, so that the headline (arm at its initial pose: all rollouts contact-free) is not read on its own
    contact_rich = None
    if world == 1 and args.config is None and name == "c4":
        try:
            pg, _, _, _ = build_planner(CONFIGS["c4_grasp"], 1, 0, local_rank, "none")
            pg.set_stream(stream.cuda_stream)
            for _ in range(3):
                pg.command_resident()
            torch.cuda.synchronize()
            n_g = min(args.steps, 20)
            evg = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_g)]
            for i in range(n_g):
                if flush is not None:
                    flush.fill_(i & 1)
                evg[i][0].record(stream)
                pg.command_resident()
                evg[i][1].record(stream)
            torch.cuda.synchronize()
            msg = float(np.mean([a.elapsed_time(b) for a, b in evg]))
            ig = pg.command_resident(sync=True)
            contact_rich = {"name": "c4_grasp", "workload": CONFIGS["c4_grasp"]["workload"], "ms_per_step": msg,
                            "value": K * T / (msg * 1e-3), "steps": n_g, "near_samples": int(ig.near_samples),
                            "note": "same K, H and timing as the headline; not part of `value`"}
            pg.close()
        except Exception as exc:   # never let the extra leg take the headline line down
            contact_rich = {"error": repr(exc)}
    clocks = sampler.summary(t_timed0, t_timed1)
    nf = 22 if env == "point_env" else 53
    nu = 2 if env == "point_env" else 9
    h2d = 4 * nf
    d2h = 4 * (2 * T * nu) + C_sizeof_info()

    if rank == 0:
        peak, peak_src = peaks()
        r_ms = float(np.mean(roll_ms))
        b_roll = B_ROLLOUT[env]
        achieved = b_roll * K * T / (r_ms * 1e-3) / 1e9
        traffic, issue = ncu_numbers(name)
        kernel = (f"k_rollout_team (panda_env, {lanes} lanes per sample)" if lanes > 1 else f"k_rollout<{env}> (thread per sample)")
        if near >= 0:
            kernel = (f"k_rollout_far (panda_env, far-field samples, 16 lanes per sample) + {kernel} over the near list "
                      f"({near} of {K} samples of this command)")
        line = {
            "metric": "sample-steps/sec (K x H per command)", "value": Kg * T / (ms_per_step * 1e-3),
            "unit": "sample-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "ms_per_step_median": ms_median, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": conf["workload"], "name": name, "K_global": Kg, "K_per_gpu": K, "H": T,
                       "multi_modal": bool(conf["mm"]), "noise": "philox4x32-10 in-kernel",
                       "dt": 0.05 if env == "point_env" else 0.01, "substeps": 2, "solver_passes": 2, "link_sweeps": 2,
                       "exchange": {"none": "single rank", "peer": "stores into peer HBM over NVLink from inside the rollout / "
                                    "weighted-sum kernels (no collective call)"}.get(exchange, exchange),
                       "l2": "not flushed" if flush is None else "flushed between steps (256 MiB fill outside the timed events)",
                       "timing": "CUDA events per step on the launching stream, max over ranks per step; value from the mean over steps"},
            "e2e": {"value": Kg * T / e2e_s, "unit": "sample-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "api": "NativePlanner.set_state + command (m3p2i_set_state / m3p2i_command), host arrays in and out; the start state rides in the kernel parameters of the rollout launch, the action and info rows are written by the update kernel into mapped pinned host memory (no separate copies), the call returns after a stream synchronize"},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "beta_search_iterations": beta_iters,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_unit": f"bytes per launch (ncu --set full, profiles/{NCU_SUMMARY.get(name, 'none for this workload')})",
                         "algorithmic_bytes_per_launch": b_roll * K * T, "peak_source": peak_src,
                         "bytes_per_sample_step": b_roll, "kernel_ms": r_ms, "path_bytes_per_sample_step": B_PATH[env],
                         "issue_frac": None if issue is None else issue / 100.0,
                         "issue_frac_source": "smsp__issue_active.avg.pct_of_peak_sustained_active of the same ncu capture: the "
                                              "roofline that binds this kernel (instruction issue / dependent-instruction latency)",
                         "note": "issue/latency-bound, not HBM-bound: the outputs stay in the 126 MB L2 (DRAM traffic < "
                                 "algorithmic bytes); see DESIGN.md 5"},
        }
        if near >= 0:
            line["far_field"] = {"near_samples": near, "far_samples": K - near,
                                 "note": "rank 0, last command: samples whose gripper never comes within contact range of a cube, the "
                                         "table or the shelf are rolled out by k_rollout_far (joints + geometry tests + cost only; exact: "
                                         "the full path would skip the same work step by step); the others by the full rollout kernel "
                                         "from their hand-over boundary. bench.py --config c4_grasp is the all-near case"}
        if contact_rich is not None:
            line["contact_rich_state"] = contact_rich
        if n1_default is not None:
            line["n1_default_workload_at_this_n"] = n1_default
        if n1_same is not None:
            line["single_gpu_same_workload"] = n1_same
            line["weak_scaling_efficiency_same_workload"] = (line["value"] / world) / n1_same["value"]
        if world > 1:
            w = torch.stack(waits_all).cpu().numpy()
            line["peer_wait_ms"] = {"wait_costs_mean_over_ranks": float(w[:, 0].mean()), "wait_costs_max_rank": float(w[:, 0].max()),
                                    "wait_partials_mean_over_ranks": float(w[:, 1].mean()), "wait_partials_max_rank": float(w[:, 1].max()),
                                    "rollout_ms_per_rank": [float(x) for x in w[:, 2]],
                                    "note": "device time (globaltimer) each rank spends at the two hand-overs of the exchange, "
                                            "commands entered together after a host barrier: the wait for the costs is the skew of "
                                            "the rollout kernels (different samples, different contact counts) plus the NVLink stores"}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            try:
                import reference_rig as R
                have_ref = R.reference_dir() is not None
            except Exception:
                have_ref = False
            port_v, port_n, port_med = cpu_port_rate(conf, cores, 4.0)
            if have_ref:
                t1 = cpu_reference_python(conf, K, cores, 1, 0)[0][0]
                n = int(max(2, min(40, args.cpu_seconds / max(t1, 1e-3))))
                times, src = cpu_reference_python(conf, K, cores, n, 1)
                v = K * T / float(np.median(times))
                line["cpu_baseline"] = {"value": v, "unit": "sample-steps/s", "cores": cores, "kind": REF_KIND,
                                        "sample": f"{n} commands of K={K}, H={T} (median {np.median(times) * 1e3:.1f} ms); " + REF_NOTE,
                                        "port": {"value": port_v, "kind": "port", "sample": f"{port_n} commands, median {port_med * 1e3:.1f} ms: "
                                                 "the whole path in C (oracle/, OpenMP over samples)"}}
            else:
                line["cpu_baseline"] = {"value": port_v, "unit": "sample-steps/s", "cores": cores, "kind": "port",
                                        "sample": f"{port_n} commands of K={K}, H={T} (median {port_med * 1e3:.1f} ms), oracle/ C port, "
                                                  "OpenMP over samples (baseline/_ref missing)"}
        print(json.dumps(line), flush=True)
    planner.close()
    if world > 1:
        dist.destroy_process_group()


def C_sizeof_info():
    import ctypes
    from m3p2i_b200 import _abi as A
    return ctypes.sizeof(A.CommandInfo)


if __name__ == "__main__":
    main()
