"""bench.py's output contract on CPU: the reference arm (the only arm that runs without a GPU) prints exactly ONE JSON
line on stdout with the keys the driver reads, and the product arm refuses to run without a CUDA device instead of
falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
  e = dict(os.environ)
  e.pop("WORLD_SIZE", None); e.pop("RANK", None)
  e.update(env or {})
  return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                        timeout=600)


def test_reference_arm_prints_one_json_line():
  r = _run("--impl", "reference", "--steps", "3", "--warmup", "1")
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.strip()]
  assert len(lines) == 1, r.stdout
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["unit"] == "maps/s" and d["higher_is_better"] is True
  assert d["value"] > 0 and d["steps"] >= 3 and d["dtype"] == "f32" and d["data"] == "synthetic"
  # the unmodified reference is the line whenever it is importable (/root/reference here, baseline/_ref on the GPU
  # box), with the C / OpenMP port next to it; otherwise the port is the line
  assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
  if d["cpu_baseline"]["kind"] == "reference":
    assert d["port"]["kind"] == "port" and d["port"]["value"] > 0
  else:
    assert "reference_note" in d
  assert "units_per_step" not in d["config"]       # same config keys as the product arm's line
  assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
  assert "config 2" in d["config"]["workload"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
  r = _run("--impl", "reference", "--gpus", "2", "--steps", "3", env={"RANK": "1", "WORLD_SIZE": "2"})
  assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_has_no_cpu_path():
  r = _run("--steps", "3")
  assert r.returncode != 0 and "needs a CUDA device" in (r.stderr + r.stdout)
