"""The bench line contract, checked on the one arm that runs without a GPU: `bench.py --impl reference` (the reference's
step on the host cores, a bounded sample: the unmodified reference code when oracle/_ref/refpy_cpu is built, else the
oracle port).  Keys and invariants are the ones the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import reference_step


@pytest.mark.parametrize("kind", ["reference", "port"])
def test_reference_arm_prints_one_contract_line(kind):
    if kind == "reference" and not reference_step.available():
        pytest.skip("oracle/_ref/refpy_cpu not built (python -m oracle.build_ref)")
    env = dict(os.environ, GG_CPU_KIND=kind)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                           "--warmup", "0", "--cpu-batch", "1"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and lines[0].startswith("{"), "stdout must carry exactly ONE JSON line:\n" + proc.stdout[-2000:]
    line = json.loads(lines[0])
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["impl"] == "reference"
    assert line["metric"] == "gangealing_train_images_per_sec_256" and "images/sec at 256" in baseline["metric"]
    assert line["unit"] == "images/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["n_gpus"] == 1 and line["steps"] >= 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["vs_baseline"] is None and line["dtype"] == "f32" and line["data"] == "synthetic"
    assert isinstance(line["config"], dict) and "workload" in line["config"] and "model" not in line["config"]
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == kind and cpu["cores"] >= 1 and cpu["value"] == line["value"] and cpu["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_non_zero_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert proc.returncode == 0
    assert not [l for l in proc.stdout.splitlines() if l.startswith("{")]
