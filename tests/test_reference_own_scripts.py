"""The reference's OWN distributed test scripts (src/accelerate/test_utils/scripts/*.py, read from the reference tree
where it lies) executed against this package: tests/refshim/run_reference_script.py aliases ``import accelerate`` to
``accelerate_b200``.  CPU/gloo; build container only."""

import os
import subprocess
import sys

import pytest

from conftest import run_distributed

REF = "/root/reference/src/accelerate/test_utils/scripts"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = {"ACCELERATE_USE_CPU": "yes", "PYTHONDONTWRITEBYTECODE": "1"}


def launch(script, world):
    runner = os.path.join("..", "refshim", "run_reference_script.py")  # run_distributed resolves under tests/scripts
    out = run_distributed(runner, world, script, env=ENV, timeout=900)
    assert out.count(f"REFERENCE_SCRIPT_OK {script}") == world, out[-3000:]


@pytest.mark.parametrize("world", [2])
def test_reference_test_sync(world):
    """test_sync.py: mean-of-gradients, no_sync, accumulate, dataloader-driven sync, schedulers."""
    launch("test_sync.py", world)


@pytest.mark.parametrize("world", [4])
def test_reference_test_ops(world):
    """test_ops.py: gather / gather_object / broadcast / pad_across_processes / reduce, debug-mode shape checks."""
    launch("test_ops.py", world)


@pytest.mark.parametrize("world", [2])
def test_reference_test_script(world):
    """test_script.py: process control, RNG sync, dataloader preparation, training_check (bf16 / fp16-on-CPU), splits."""
    launch("test_script.py", world)


def test_reference_test_distributed_data_loop():
    launch("test_distributed_data_loop.py", 2)


def test_reference_scripts_single_process():
    for script in ("test_script.py",):
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "refshim", "run_reference_script.py"), script],
                             env=dict(os.environ, **ENV), capture_output=True, text=True, timeout=600, cwd=ROOT)  # fmt: skip
        assert res.returncode == 0 and f"REFERENCE_SCRIPT_OK {script}" in res.stdout, res.stderr[-3000:]
