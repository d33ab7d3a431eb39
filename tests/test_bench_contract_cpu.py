"""The bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`
runs the CPU restatement of the reference loops) prints ONE JSON line with the keys the driver
reads, on the same metric / unit / config as our arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = run_bench("--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1", "--cpu-budget", "0.2")
    assert d["impl"] == "reference" and d["unit"] == "DOFs/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--gpus", "2",
                          "--steps", "1", "--warmup", "1", "--cpu-budget", "0.1"], capture_output=True, text=True, timeout=300,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_profiles_traffic_covers_every_bench_config():
    """bench.py fills roofline.traffic from profiles/traffic.json: the headline workload and every sub-config must
    have the kernels it reports, as positive byte counts with the ncu source named"""
    sys.path.insert(0, ROOT)
    import bench
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for name in (bench.HEADLINE,) + tuple(bench.SUB_CONFIGS):
        assert name in t, name
        for kernel in ("evaluate", "interpolate"):
            assert t[name][kernel] > 0 and bench.traffic_from_profiles(kernel, name) == t[name][kernel], (name, kernel)
        assert "ncu" in t[name]["source"]
    assert t["c3"]["indices"] > 0 and t["c3"]["indices_multi"] > 0 and t["c4"]["evaluate_values_only"] > 0
