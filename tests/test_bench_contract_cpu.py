"""CPU: the JSON contract of ``bench.py --impl reference`` (the arm the driver runs next to the GPU arm): one line on rank 0 with
the GPU arm's metric / unit / direction, ``impl``, a ``cpu_baseline`` describing the run and an ``e2e`` object with zero copy
bytes; the other ranks of a torchrun launch print nothing and exit 0.  Runs the small configs[0] workload (4 layers, batch 8)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line_on_rank_0_only():
    r = _run({"RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("audio-sec/sec train") and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["steps"] == 1 and d["n_gpus"] == 1 and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "utterances" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    other = _run({"RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""
