"""bench.py contract, CPU side: the reference arm (`--impl reference` = the fp32 oracle port timed on the host cores)
prints ONE JSON line with the keys the driver reads, on the same workload description as the CUDA arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config("vit_b", 32, 1, "box")   # identical to the CUDA arm's config object
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]


def test_cuda_arm_and_reference_arm_describe_the_same_workload():
    sys.path.insert(0, ROOT)
    import bench

    a = bench.workload_config("vit_b", 32, 1)
    b = bench.workload_config("vit_b", 32, 8)
    assert a["workload"] == b["workload"] and a["images_per_gpu_per_step"] == 32
    assert a["global_batch"] == 32 and b["global_batch"] == 256   # weak scaling: per-GPU work is fixed
    assert a["collective"] == "none" and "all_gather" in b["collective"]
