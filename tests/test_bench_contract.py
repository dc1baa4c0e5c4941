"""bench.py contract on a GPU-less box: the reference arm (`--impl reference`) runs the CPU oracle port on a bounded
sample and prints ONE JSON line with the keys the driver reads; under torchrun only rank 0 prints.  (The GPU arm is
exercised on the GPU box; here only what needs no GPU.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--no-train"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("720p output frames/s")
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_committed_bench_lines_keep_the_contract():
    """The final bench lines under profiles/ are what DESIGN.md and README.md quote: each must be one JSON object with the
    contract's keys, a roofline measured against a stated peak, launches counted, clocks sampled and no thermal / hardware
    slowdown reason."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_v10*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r01_bench_v9.json")) + glob.glob(os.path.join(ROOT, "profiles", "r01_bench_v8.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r02_bench_v*.json")))
    assert paths
    for path in paths:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        if d.get("impl") == "reference":
            assert d["gpu_launches"] == 0 and d["cpu_baseline"]["kind"] == "port"
            continue
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (path, k)
        assert d["gpu_launches"] > 0 and d["dtype"] == "bf16" and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] < d["value"] * 1.02
        r = d["roofline"]
        assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.3 < r["frac"] < 1.0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] == 1 and d.get("cpu_baseline"):      # (the same-box companion of the 2-GPU run skips the CPU leg)
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "glue" in d
