"""bench.py's contract, as far as a box without a GPU can check it: the reference arm runs the reference's own CPU
renderer and prints one JSON line with the agreed keys; the product arm refuses to run without a CUDA device (there
is no CPU fallback to time by accident)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _run(*args):
    env = dict(os.environ, RTM_QUIET="1")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          timeout=600, cwd=ROOT, env=env)


def test_reference_arm_line(pyoracle):
    r = _run("--impl", "reference", "--workload", "C1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "C1" and d["config"]["width"] == 512 and d["config"]["spp"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["kind"] == ("reference" if pyoracle.have_ref() else "port")
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--workload", "C1", "--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and not [l for l in r.stdout.splitlines() if l.startswith("{")]
