"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference`: the reference's own CPU code from oracle/_ref, or the plain-C port) prints exactly ONE JSON
line on stdout with the keys the driver reads, also under a multi-rank launch (rank 0 prints, the others exit 0)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2s",
                           "--steps", "2", "--warmup", "1", "--ref-sample", "300"] + extra,
                          capture_output=True, text=True, env=e, timeout=900)


@pytest.mark.parametrize("mode", ["locate", "count"])
def test_reference_arm_prints_one_json_line(mode):
    out = _run(["--mode", mode])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == ("occ/s" if mode == "locate" else "patterns/s")
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if cb["kind"] == "reference" and "cached .ri" in cb["index"]:
        assert d["repo_libraries_loaded"] == []      # the reference arm runs without any of this repo's native code


def test_reference_arm_other_ranks_exit_quietly():
    out = _run([], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_default_workloads_and_shared_config():
    import bench
    assert bench.job_size("c5", 8) == bench.job_size("c5", 1) == 100_000          # strong: one fixed job
    assert bench.job_size("c3", 8) == 8 * bench.job_size("c3", 1) == 1_000_000    # the config's 1M patterns on 8 GPUs
    a = bench.shared_config("c5", 8, 4_000_000_001, 676_376)
    assert a["workload"].startswith("C5") and a["job_patterns"] == 100_000 and a["gpus"] == 8 and "model" not in a
