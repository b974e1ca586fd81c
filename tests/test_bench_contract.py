"""bench.py's contract on the CPU side (no GPU): the algorithmic-work figures the roofline is computed from, and the
`--impl reference` arm -- the reference's own modules (or the oracle port when they are not staged) on the host cores -- printing
ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_algorithmic_flops_match_baseline_md():
    """BASELINE.md section 4 / SURVEY.md section 8(d): 2 x MAC of the Linear layers per sample."""
    import bench
    assert bench.flop_per_sample(10, 128, bwd=False) == 131584
    assert bench.flop_per_sample(10, 128, bwd=True) == 362496
    assert bench.flop_per_sample(10, 256, bwd=False) == 459776
    # by hand: layers 63->128, 128->128, (128+63)->128, 128->128, heads 128->4
    macs = 63 * 128 + 128 * 128 + 191 * 128 + 128 * 128 + 128 * 4
    assert bench.flop_per_sample(10, 128, bwd=False) == 2 * macs


def test_peaks_come_from_the_measured_file_or_the_stated_fallback():
    import bench
    pk = bench.peaks()
    assert pk["source"] in ("measured", "fallback") and pk["tf_burst"] >= pk["tf_sustained"] > 0 and pk["hbm_gbs"] > 0
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        assert pk["source"] == "measured"


@pytest.mark.parametrize("workload,metric,unit", [("c1", "train ray-samples/sec (fwd+bwd+Adam)", "ray-samples/s"),
                                                  ("render", "render rays/sec (fused forward)", "rays/s")])
def test_reference_arm_prints_one_contract_line(workload, metric, unit):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "1"]
    if workload == "c1":
        cmd += ["--rays", "256"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                       # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == unit
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert isinstance(d["config"], dict) and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    staged = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "src"))
    assert cb["kind"] == ("reference" if staged else "port")


def test_reference_arm_other_ranks_exit_without_work():
    """under torchrun rank 0 alone runs the CPU arm; the other ranks exit 0 and print nothing"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])
