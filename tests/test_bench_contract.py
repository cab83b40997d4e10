"""bench.py's driver-facing contract, the parts that need no GPU: the reference arm prints ONE JSON line with the keys
the driver reads, and the product arm fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          cwd=ROOT, timeout=600)


def test_reference_arm_line():
    # a reduced sample (PBX_BENCH_REF_N) keeps this a matter of seconds; the driver's run uses the default sample
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", env={"PBX_BENCH_REF_N": "64"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "GDoF/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert "512^3" in d["metric"] and d["config"]["grid"] == [512, 512, 512]


def test_grid_flag_spelling():
    """`--grid` is the spelling that survives torchrun's own option parser (`--n` is an ambiguous abbreviation there)"""
    r = _run("--help")
    assert r.returncode == 0 and "--grid" in r.stdout and "--workload" in r.stdout
