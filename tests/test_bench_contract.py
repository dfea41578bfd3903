"""bench.py's output contract, checked on the CPU arm (the reference arm runs the oracle: no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(oracle, built_libs):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]  # libraries that print on stdout must not reach it (bench.py keeps fd 1 for the line)
    line = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["impl"] == "reference" and line["metric"] in base["metric"] and line["unit"] == "ms/frame"
    assert line["higher_is_better"] is False and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and line["ms_per_step"] == line["value"]
    assert line["config"]["workload"].startswith("synthetic 6M") and line["config"]["n_gaussians"] == 6_000_000
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "ms/frame", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["pairs"] > 10_000_000


def test_reference_arm_uses_all_cores_under_torchrun_and_none_of_the_product(oracle, built_libs):
    """torchrun exports OMP_NUM_THREADS=1 to its workers (round 1: the N >= 2 reference lines ran on one core), and the arm
    must not load the product's libraries: its camera comes from oracle/_ref, its scene from the pure-numpy generators."""
    code = (
        "import os, sys, json; sys.path.insert(0, %r); import bench\n"
        "from oracle import oracle as O\n"
        "cores = O.use_all_cores(); ubo, src = bench.garden_ubo(); from torpedo_b200 import scenes\n"
        "maps = open('/proc/self/maps').read()\n"
        "print(json.dumps({'cores': cores, 'affinity': len(os.sched_getaffinity(0)), 'engine': 'torpedo_b200.engine' in sys.modules,\n"
        "                  'tpdcu': 'libtpdcu' in maps or 'libtpdhost' in maps, 'oracle': 'libtpd_oracle' in maps, 'src': src}))\n" % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    assert got["cores"] == got["affinity"] and got["oracle"]
    assert not got["engine"] and not got["tpdcu"], got
