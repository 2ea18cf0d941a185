"""bench.py --impl reference on the CPU (no GPU needed): the arm the driver runs beside the GPU arm.  It must print
one JSON line with the contract's keys, run the CPU port on every host thread whatever OMP_NUM_THREADS torchrun
exports, load nothing of the product, and -- under torchrun -- run on rank 0 only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra, *args):
    env = dict(os.environ, SDFGPU_BENCH_REF_SECONDS="2", **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True,
                          text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_line():
    r = run({"OMP_NUM_THREADS": "1"}, "--steps", "3", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sdf_samples_per_sec" and d["unit"] == "samples/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 1e6 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["grid"] == [512, 512, 512] and d["config"]["frame"] == [1920, 1080]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == d["detail"]["omp_threads"]
    assert cb["cores"] == (os.cpu_count() or 1)            # torchrun's OMP_NUM_THREADS=1 is not obeyed
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_under_torchrun_runs_on_rank_0_only():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0 and r.stdout.strip() == "", r.stdout + r.stderr


def test_reference_arm_does_not_load_the_product():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1'];\n"
            "runpy.run_path('bench.py', run_name='__main__');\n"
            "import re; maps = open('/proc/self/maps').read();\n"
            "assert 'liboracle' in maps and 'libsdfgpu' not in maps and 'sdf_viewer_b200' not in sys.modules, 'product loaded'\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, SDFGPU_BENCH_REF_SECONDS="1"))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
