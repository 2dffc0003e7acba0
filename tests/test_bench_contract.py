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


def test_no_collective_inside_a_rank_guard():
    """Every collective of bench.py must be reached by EVERY rank: an all-reduce inside the `if rank == 0:` block
    that prints the JSON line once hung an 8-GPU run until NCCL's 10-minute watchdog (round 2).  Static check: no call
    of the collective helpers / torch.distributed collectives lexically inside an `if` whose test mentions `rank`."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    collectives = {"sum_over_ranks", "max_over_ranks", "min_over_ranks", "reduce_ranks", "barrier", "all_reduce",
                   "all_gather_object", "all_gather", "broadcast", "broadcast_object_list"}
    bad = []

    class V(ast.NodeVisitor):
        def __init__(self):
            self.guard = 0

        def visit_If(self, node):
            names = {n.id for n in ast.walk(node.test) if isinstance(n, ast.Name)}
            g = "rank" in names or "local" in names
            self.visit(node.test)
            self.guard += g
            for b in node.body:
                self.visit(b)
            self.guard -= g
            for b in node.orelse:          # `else` of a rank test is a rank guard too
                self.guard += g
                self.visit(b)
                self.guard -= g

        def visit_Call(self, node):
            f = node.func
            name = f.id if isinstance(f, ast.Name) else (f.attr if isinstance(f, ast.Attribute) else "")
            if self.guard and name in collectives:
                bad.append((name, node.lineno))
            self.generic_visit(node)

    V().visit(tree)
    assert not bad, "collectives under a rank guard: %s" % bad
