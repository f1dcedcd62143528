"""bench.py's reference arm (the unmodified reference Python from baseline/_ref -- or /root/reference/src in the build
container -- over the sim facade with the oracle's integrator, timed on the host cores) runs without a GPU: check the
JSON contract of the line the driver parses. The native arm needs a GPU and is exercised on the box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sample-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("sample-steps/sec") and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert abs(d["value"] - 4096 * 32 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["config"]["workload"].startswith("config_panda reactive pick") and d["config"]["K_global"] == 4096
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference-python+port-dynamics", "port")
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    if os.path.exists(os.path.join(ROOT, "baseline", "_ref")) or os.path.exists("/root/reference/src"):
        assert cb["kind"] == "reference-python+port-dynamics"   # the reference's own code is the baseline of record
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["dtype"] == "f32" and d["scaling"] == "weak" and d["vs_baseline"] is None


def test_reference_arm_c5_is_the_multi_modal_shelf_reach():
    """At N > 1 the default workload is BASELINE configs[4] (multi-modal reach, cube on the shelf)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.strip()][-1])
    assert d["config"]["name"] == "c5" and "multi_modal=True cube_on_shelf=True" in d["config"]["workload"]
    assert d["config"]["K_global"] == 8192 and d["n_gpus"] == 2


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
