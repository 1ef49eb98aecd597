"""bench.py's reference arm (the part of the benchmark contract that runs without a GPU): one JSON line with the keys the driver
reads, the same `config` the CUDA arm emits for the same flags, and no work on ranks other than 0."""
import json
import os
import subprocess
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_reference(extra=(), env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--frame-size", "640x400", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, **(env or {})})
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_prints_the_contract_line():
    lines = run_reference()
    d = json.loads(lines[-1])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["unit"] == "frames/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["gpu_launches"] == 0 and d["data"] == "synthetic" and d["dtype"] == "u8/int32/f32"
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["config"]["frame_size"] == [640, 400] and "workload" in d["config"] and "model" not in d["config"]

    # the CUDA arm builds its `config` with the same function from the same flags: identical keys AND values
    import bench
    args = SimpleNamespace(size=(640, 400), config=2, frame_size="640x400", batch=d["config"]["frames_in_ring_per_gpu"], k2=0.0, tilt=0.0)
    lp, _ = bench.build_workload(640, 400, 1)
    assert bench.workload_config(args, lp, 1) == d["config"]


def test_reference_arm_on_other_ranks_does_nothing():
    lines = run_reference(env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert lines == [] or all(not ln.startswith("{") for ln in lines)
