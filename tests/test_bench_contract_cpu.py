"""bench.py contract checks that need no GPU: the reference arm (CPU oracle of the path) prints ONE JSON line with the
keys the driver reads, and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, cwd=ROOT,
                          env=env, timeout=600)


def test_reference_arm_prints_one_contract_line():
    res = _run("--impl", "reference", "--workload", "cora", "--steps", "2", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "gconv_adapter_fwd_bwd_edges_per_sec" and j["unit"] == "edges/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 2 and j["value"] > 0
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert "cora" in j["config"]["workload"] and j["vs_baseline"] is None and j["dtype"] == "f32"


def test_product_arm_has_no_cpu_fallback():
    res = _run("--workload", "cora", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-e2e")
    assert res.returncode != 0
    assert "CUDA" in (res.stderr + res.stdout)
