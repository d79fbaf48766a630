"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, non-zero
ranks of a torchrun launch stay silent, and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

import parity_common as pc

BENCH = os.path.join(pc.REPO, "bench.py")


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, cwd=pc.REPO, capture_output=True, text=True, env=e, timeout=600)


@pytest.mark.skipif(not (os.path.exists(pc.REF_LIB) or os.path.exists(pc.PORT_LIB)), reason="no CPU engine built")
def test_reference_arm_prints_one_json_line():
    out = run(["--impl", "reference", "--workload", "battle1", "--steps", "2", "--warmup", "3", "--cpu-steps", "20"])
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "dtype",
                "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in j, key
    assert j["impl"] == "reference" and j["gpu_launches"] == 0 and j["value"] > 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"]


def test_reference_arm_other_ranks_stay_silent():
    out = run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = run(["--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e", "--workload", "battle1"])
    assert out.returncode != 0
    assert out.stdout.strip() == "" or "value" not in out.stdout
