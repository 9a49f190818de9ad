"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm (`--impl reference`)
runs here -- it is the unmodified reference (or the oracle port) on host cores -- and must print ONE JSON line with the keys
and the `config` / `steps` / `warmup` our own arm prints; the FLOP accounting is the one DESIGN.md quotes; and our arm
must fail loudly when there is no GPU (no CPU path)."""
import argparse
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _run(*flags, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *flags], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"expected one JSON line, got {len(lines)}"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mcts_simulations_per_second" and d["unit"] == "sims/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    assert "reference_sample" in d and "reference_sample" not in d["config"]          # config holds the workload only
    # the same config our arm prints for the same flags (the driver compares the two lines)
    import bench
    ns = argparse.Namespace(games=1000, threads=2, batch=16, sims=800, channels=128, blocks=9)
    assert d["config"] == bench.workload_config(ns, 1)
    assert set(d["config"]) == {"workload", "games_per_gpu", "sims_per_move", "leaf_batch", "l2", "parallelism"}


def test_reference_arm_other_ranks_exit_silently():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_flop_accounting_is_the_documented_one():
    import bench
    f_all, f_conv = bench.flops_per_eval()                       # 128x10b, SURVEY 8d: dense 2 x MAC, padding taps counted
    assert f_all == 342_327_808                                   # SURVEY.md 8d / DESIGN.md section 3.1: per network evaluation
    assert f_conv == 2 * 576 * 128 * (17 + 18 * 128)              # the 19 convolutions of the tower
    assert 0 < f_all - f_conv < 0.01 * f_conv                     # the heads are < 1 % of an evaluation
    f_all5, f_conv5 = bench.flops_per_eval(c=256, blocks=19, hv=256)
    assert f_all5 == 2_874_065_664                                # the config5 sub-record's flop_per_eval
    assert 0 < bench.ISSUED_MMA_FRACTION < 1


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU behaviour")
def test_our_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-extra-configs", timeout=300)
    assert r.returncode != 0
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines()), "no bench line may be printed without a GPU"
    assert "GPU" in r.stderr or "CUDA" in r.stderr or "cuda" in r.stderr
