"""CPU: the reference arm of bench.py (`--impl reference`: the oracle timed on the host cores) prints ONE JSON line with the
contract's keys; under a multi-rank launch only rank 0 works.  Runs the 224^2 workload (seconds on a few cores)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, extra_args=()):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "224", "--steps", "1",
                          "--warmup", "0", "--gpus", env_extra.get("WORLD_SIZE", "1"), *extra_args],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run({})
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("image-pairs/sec (fwd+bwd) DUSt3R ViT-L/16")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    # "reference": the unmodified reference package (baseline/_ref or /root/reference) ran; "port": the oracle stood in
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "224x224" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    lines = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29555"})
    assert lines == []


def test_reference_arm_never_loads_the_product_library():
    """the reference arm's process must not import `uniception_b200` (which loads libuc_b200.so); bench.py asserts
    it before printing, and the oracle-port fallback (no reference package on the path) obeys the same rule."""
    env = {"UC_REFERENCE_ROOT": "/nonexistent"}
    code = ("import os, sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'linear224', '--steps', '1', '--warmup', '0'];"
            "import bench; bench._reference_root = lambda: None; bench.main();"
            "assert 'uniception_b200' not in sys.modules; assert not any('libuc_b200' in l for l in open('/proc/self/maps'))")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][-1])
    assert d["cpu_baseline"]["kind"] == "port"
