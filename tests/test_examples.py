"""The user-facing example scripts run end to end on CPU (two processes over gloo)."""

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd, nproc=1, port=29681):
    env = dict(os.environ, ACCELERATE_USE_CPU="yes", OMP_NUM_THREADS="1")
    if nproc > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", str(port)] + cmd  # fmt: skip
    else:
        cmd = [sys.executable] + cmd
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    return out.stdout


def test_nlp_example_two_processes_uneven_eval_tail():
    # 37 evaluation samples over 2 ranks: gather_for_metrics must drop the duplicated tail (the script asserts the count)
    out = run(["examples/nlp_example.py", "--cpu", "--tiny", "--num_epochs", "1", "--train_samples", "48", "--eval_samples", "37"],
              nproc=2)  # fmt: skip
    assert out.count("epoch 0:") == 1, "only the main process prints"


def test_gpt2_example_two_processes_with_accumulation():
    out = run(["examples/gpt2_data_parallel.py", "--model", "tiny", "--cpu", "--steps", "4", "--batch", "2", "--seq", "32",
               "--accumulate", "2"], nproc=2, port=29682)  # fmt: skip
    assert "samples/s" in out and "step 3:" in out
