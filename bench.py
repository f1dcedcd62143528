#!/usr/bin/env python
"""bench.py -- sample-steps/s of one MPPI command() (the hot path of BASELINE.json) on N B200s.

A step is one planner tick: noise -> perturbation -> K x H rollout (dynamics + cost) -> softmin -> mean update.
Workload: config_panda reactive pick, 7-DoF Panda + cube, K=4096 per GPU, H=32 (BASELINE.json configs[3]; at N>1
K is sharded, K_global = 4096*N, which at N=8 is configs[4]'s K=32768).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (see the keys below). `--impl reference` times the CPU oracle port of the same path
on the host cores (the reference's own rollout needs IsaacGym, which cannot be installed: DESIGN.md).
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

K_PER_GPU = 4096
HORIZON = 32
WORKLOAD = "config_panda reactive pick, 7-DoF Panda + cube, K=4096 per GPU, H=32"
# algorithmic HBM bytes per sample-step (DESIGN.md "Algorithmic bytes"): the rollout kernel writes the action row
# (9 f32), the float4 state row and the cost; the weighted-sum pass re-reads the action row.
B_ROLLOUT = 4 * 9 + 16 + 4          # 56 B, fused rollout kernel (Philox noise: no table read)
B_PATH = B_ROLLOUT + 4 * 9          # 92 B, whole command


def scene_inputs():
    from m3p2i_b200 import scene as S
    actors = S.default_actors("panda_env")
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors).copy()
    # cubes resting on the table, as in a running episode
    root[S.actor_index(actors, "cubeA"), 2] -= 0.0095
    root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
    cb = root[S.actor_index(actors, "cubeB")]
    goal = np.concatenate([cb[:3] + np.array([0, 0, 0.055], np.float32), cb[3:7]]).astype(np.float32)
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
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md): NVML polled every 5 ms in this
    process (nvidia_ml_py), or nvidia-smi every 100 ms when NVML cannot be loaded."""
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
                    self.rows.append([str(sm), str(self.max_sm)] + ["Active" if mask & b else "Not Active" for b in self.BITS])
                    self.stop_flag.wait(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i] == "Active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one rollout launch from the committed ncu --set full capture
    (profiles/), in bytes; None if the summary is missing."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_rollout_team_summary.csv")
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(path):
        parts = line.strip().split(",")
        if len(parts) >= 4 and parts[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(parts[-1]) * scale.get(parts[-2], 1.0)
    return tot or None


def cpu_oracle_rate(threads, seconds=12.0, K=K_PER_GPU, T=HORIZON, min_steps=2, warmup=1):
    """The CPU port of the same path (oracle/), OpenMP over samples, timed for about `seconds`."""
    import oracle_py as O
    from m3p2i_b200 import _abi as A
    from m3p2i_b200 import scene as S
    O.set_threads(threads)
    cfg = S.make_cfg("panda_env", "pick", None, K, T)
    o = O.Oracle(S.build_config(cfg, noise_mode=A.NOISE_PHILOX, seed=0), S.build_panda_scene())
    o.set_filter_matrix(S.savgol_matrix(T))
    dof, root, goal = scene_inputs()
    o.set_objective("pick", goal, "close")
    times = []
    t_end = time.perf_counter() + seconds
    n = 0
    while n < warmup + min_steps or time.perf_counter() < t_end:
        o.set_state(dof, root)
        t0 = time.perf_counter()
        o.command()
        dt = time.perf_counter() - t0
        if n >= warmup:
            times.append(dt)
        n += 1
        if len(times) >= 200:
            break
    o.close()
    return K * T / float(np.median(times)), len(times), float(np.median(times))


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    import oracle_py as O
    from m3p2i_b200 import _abi as A
    from m3p2i_b200 import scene as S
    O.set_threads(cores)
    Kg = K_PER_GPU * args.gpus
    # bounded sample: the full K of one GPU's shard per step (the port's cost is linear in K)
    Ks = K_PER_GPU
    cfg = S.make_cfg("panda_env", "pick", None, Ks, HORIZON)
    o = O.Oracle(S.build_config(cfg, noise_mode=A.NOISE_PHILOX, seed=0), S.build_panda_scene())
    o.set_filter_matrix(S.savgol_matrix(HORIZON))
    dof, root, goal = scene_inputs()
    o.set_objective("pick", goal, "close")
    for _ in range(args.warmup):
        o.set_state(dof, root)
        o.command()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.set_state(dof, root)
        o.command()
    dt = (time.perf_counter() - t0) / args.steps
    value = Ks * HORIZON / dt
    line = {"impl": "reference", "metric": "sample-steps/sec (K x H per command)", "value": value, "unit": "sample-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "K_global": Kg, "H": HORIZON, "noise": "philox4x32-10",
                       "note": "CPU port of the reference path (oracle/, OpenMP over samples); the reference's own rollout "
                               "is IsaacGym/PhysX which is not installable; throughput of the port is independent of K"},
            "cpu_baseline": {"value": value, "unit": "sample-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} commands of K={Ks}, H={HORIZON} (one GPU's shard)"},
            "e2e": {"value": value, "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
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
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
        args.gpus = world

    import torch
    import torch.distributed as dist
    from m3p2i_b200 import _abi as A
    from m3p2i_b200 import native
    from m3p2i_b200 import scene as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()

    Kg = K_PER_GPU * world
    cfg = S.make_cfg("panda_env", "pick", None, Kg, HORIZON)
    cfg.mppi.sampling_method = "philox"
    c = S.build_config(cfg, num_samples_local=K_PER_GPU, sample_offset=rank * K_PER_GPU, noise_mode=A.NOISE_PHILOX, seed=0)
    planner = native.NativePlanner(c, S.build_panda_scene(), device=local_rank)
    planner.set_filter_matrix(S.savgol_matrix(HORIZON))
    exchange = "none"
    if world > 1:
        exchange = args.exchange
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
    dof, root, goal = scene_inputs()
    planner.set_objective("pick", goal, "close")
    planner.set_state(dof, root)
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
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    roll_ms = []
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
    # rollout-kernel duration (CUDA events inside the library, same stream), L2 flushed before each launch
    for i in range(min(args.steps, 50)):
        if flush is not None:
            flush.fill_(i & 1)
        info = planner.command_resident(sync=True)
        roll_ms.append(info.rollout_ms)
    last = planner.command_resident(sync=True)
    launches_per_step, lanes = last.launches, int(last.rollout_lanes)
    barrier()

    # ---------------- end to end through the public API with host buffers (H2D state in, D2H action out)
    n_e2e = args.steps
    for _ in range(3):
        planner.set_state(dof, root)
        planner.command(want_cost=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        planner.set_state(dof, root)                      # pinned staging + H2D inside
        act, _, _ = planner.command(want_cost=False)      # D2H of the action inside, synchronous
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s)
    clocks = sampler.summary()
    h2d = 4 * 53
    d2h = 4 * (2 * HORIZON * 9) + C_sizeof_info()

    if rank == 0:
        peak, peak_src = peaks()
        r_ms = float(np.mean(roll_ms))
        achieved = B_ROLLOUT * K_PER_GPU * HORIZON / (r_ms * 1e-3) / 1e9
        line = {
            "metric": "sample-steps/sec (K x H per command)", "value": Kg * HORIZON / (ms_per_step * 1e-3),
            "unit": "sample-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "K_global": Kg, "H": HORIZON, "noise": "philox4x32-10 in-kernel",
                       "dt": 0.01, "substeps": 2, "solver_passes": 2, "link_sweeps": 4,
                       "exchange": {"none": "single rank", "peer": "stores into peer HBM over NVLink from inside the rollout / "
                                    "weighted-sum kernels (no collective call)"}.get(exchange, exchange),
                       "l2": "not flushed" if flush is None else "flushed between steps (256 MiB fill outside the timed events)",
                       "timing": "CUDA events per step on the launching stream, max over ranks, mean over steps"},
            "e2e": {"value": Kg * HORIZON / e2e_s, "unit": "sample-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "api": "NativePlanner.set_state + command (m3p2i_set_state / m3p2i_command), host arrays"},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": (f"k_rollout_team (panda_env, {lanes} lanes per sample)" if lanes > 1 else "k_rollout<panda_env> (thread per sample)"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "traffic_unit": "bytes per launch (ncu --set full, profiles/r01_ncu_rollout_team_summary.csv)",
                         "algorithmic_bytes_per_launch": B_ROLLOUT * K_PER_GPU * HORIZON, "peak_source": peak_src,
                         "bytes_per_sample_step": B_ROLLOUT, "kernel_ms": r_ms,
                         "path_bytes_per_sample_step": B_PATH,
                         "note": "issue/latency-bound, not HBM-bound: the 7.3 MB of outputs stay in the 126 MB L2 (DRAM traffic < algorithmic bytes); see DESIGN.md 5"},
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            v, n, med = cpu_oracle_rate(cores, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "sample-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} commands of K={K_PER_GPU}, H={HORIZON} (median {med * 1e3:.1f} ms), "
                                              "oracle/ C port, OpenMP over samples"}
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
