"""bench.py contract pieces that can be checked without a GPU: the reference (CPU) arm prints one JSON
line with the required keys, and non-zero ranks of that arm exit 0 without work."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "gpt2-tiny",
                           "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)  # fmt: skip


def test_reference_arm_json_line():
    res = run()
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):  # fmt: skip
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "samples_per_sec" and line["value"] > 0
    # "reference" when baseline/_ref (pip install --target of /root/reference) is present, else the oracle port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["vs_baseline"] is None  # BASELINE.md publishes no number for this metric


def test_reference_arm_other_ranks_exit_quietly():
    res = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""
