"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU forward from oracle/_ref, else the oracle port) prints exactly ONE JSON line on stdout
with the keys the driver reads; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_contract_keys():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-budget", "1", "--height", "64", "--width", "64")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MP/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("megapixels/sec") and d["value"] > 0 and d["vs_baseline"] is None
    from oracle.build_ref import ref_available
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_available() else "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box the product arm is exercised by the driver itself
    out = _run("--steps", "1", "--warmup", "3")
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
    assert out.stdout.strip() == ""  # nothing that could be mistaken for a measurement
