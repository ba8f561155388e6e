"""bench.py contract checks that need no GPU: the reference arm (CPU oracle) prints ONE JSON line with the keys the
driver reads, on the same metric / config naming as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0", *args], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys():
    sys.path.insert(0, ROOT)
    import bench
    j = run_reference()
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["metric"] == bench.metric_name("c1") and j["config"] == bench.config_of("c1", 1)
    assert j["value"] > 0 and j["n_gpus"] == 1 and j["vs_baseline"] is None
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["single_thread_value"] > 0
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_ignores_a_single_thread_openmp_default_and_non_zero_ranks_stay_silent():
    # torchrun exports OMP_NUM_THREADS=1; the arm must still describe the cores it could use, and only rank 0 prints
    j = run_reference({"OMP_NUM_THREADS": "1"})
    assert j["cpu_baseline"]["cores"] >= 1
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env,
                         timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
