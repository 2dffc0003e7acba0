"""bench.py's reference arm runs without a GPU: check the keys of its JSON line (the contract the
driver relies on) on a small sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line(kit):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--genome-mb", "0.8", "--chunk-mb", "0.4", "--cpu-threads", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-1500:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "classified k-mers/sec" and line["unit"] == "k-mers/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["config"]["kmers"] > 0
    cb = line["cpu_baseline"]
    assert cb["same_read_set_as_gpu_arm"] is True and cb["full_set_run"]["kmers"] == line["config"]["kmers"]
    if cb["kind"] == "reference":
        assert cb["t1"]["kmers_per_user_s"] > 0          # the per-core figure BASELINE.md section 4 asks for


def test_reference_arm_shrinks_to_its_budget(kit):
    """A budget that K runs of the full set cannot meet: the steps use a prefix of the chromosomes and say so."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "40", "--warmup", "0",
                        "--genome-mb", "1.2", "--chunk-mb", "0.2", "--cpu-threads", "2", "--ref-budget-s", "14"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-1500:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    cb = line["cpu_baseline"]
    if cb["kind"] == "reference" and not cb["same_read_set_as_gpu_arm"]:
        assert "chromosomes per step" in cb["sample"] and line["steps"] == 40


def test_product_arm_needs_a_gpu():
    """No CUDA device here: the product arm must stop with a message, not fall back on anything."""
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--genome-mb", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
    assert p.returncode != 0 and "no CPU path" in (p.stderr + p.stdout)
