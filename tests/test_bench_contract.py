"""CPU: the reference arm of bench.py (the reference's own CPU SpMV, oracle/_ref or the C restatement) prints exactly one
JSON line on stdout with the keys the driver reads, on the same `config.workload` string the GPU arm prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "8"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GOPS" and d["higher_is_better"] is True and d["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["e2e"] == {"value": d["value"], "unit": "GOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    # the same workload string as the GPU arm builds (bench.workload_name)
    sys.path.insert(0, ROOT)
    import bench
    bench.WORKLOAD = "c2"
    assert d["config"]["workload"].startswith(bench.WORKLOADS["c2"][0] % ())
    assert d["config"]["spmv_per_step"] == 8
