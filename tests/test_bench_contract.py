"""CPU tier: the parts of bench.py's contract that need no GPU -- the reference arm (`--impl reference`: the compiled,
unmodified reference timed on the host cores) prints one JSON line with the contract's keys; ranks other than 0 print
nothing; the B200 arm fails loudly without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("workload", ["config2_small", "curvature3_small"])
def test_reference_arm_line(workload):
    """(smoke-size workloads: the default one is 512^3 x 5 variables, minutes of host time)"""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3", "--workload", workload],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gcells/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and "workload" in d["config"]
    # the reference arm runs the WHOLE workload (same grid, every variable): its config object is the B200 arm's
    import bench
    spec = bench.workload_spec(workload)
    assert d["config"] == bench.config_of(spec, spec["build"](fill=False))
    assert d["steps"] == 2 and d["metric"] == bench.metric_name(spec["kind"])
    # under torchrun only rank 0 runs the reference
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    q = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert q.returncode == 0 and q.stdout.strip() == ""


def test_b200_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--no-extras", "--no-cpu-baseline"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert p.returncode != 0 and not any(l.startswith("{") for l in p.stdout.splitlines())
