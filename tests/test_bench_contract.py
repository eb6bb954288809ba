"""CPU: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm
(`--impl reference`: the reference's own modules from oracle/_ref on the host cores, or the oracle port
when the staged copy is missing) prints ONE JSON line with the keys the driver reads; ranks other than 0
stay silent; the engine arm refuses to run without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.pop("RANK", None)
    e.pop("WORLD_SIZE", None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, env=e, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_contract_line():
    # batch 1 per step keeps this at a few seconds; the driver's run uses the default batch 16 (BASELINE configs[2])
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--ref-batch", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("mel-frames/sec full CycleGAN train step") and d["unit"] == "mel-frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["value"] - 1 * 64 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "batch 64 per GPU" in d["config"]["workload"]      # the engine arm's config
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == (os.cpu_count() or 1)
    assert cb["value"] == d["value"] and cb["unit"] == d["unit"] and "batch 1" in cb["sample"]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mask_cyclegan_vc", "model.py")):
        assert cb["kind"] == "reference"        # the staged, unmodified reference modules are the ones timed
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_silently():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--ref-batch", "1"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])


def test_engine_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present: the engine arm would run")
    r = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert r.returncode != 0
    assert "needs a CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]      # no JSON line from a run that measured nothing
