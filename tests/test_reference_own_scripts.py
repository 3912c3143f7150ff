"""The reference's OWN distributed test scripts (src/accelerate/test_utils/scripts/*.py, read from the reference tree
where it lies) executed against this package: tests/refshim/run_reference_script.py aliases ``import accelerate`` to
``accelerate_b200``.  CPU/gloo; build container only."""

import os
import subprocess
import sys

import pytest

from conftest import run_concurrently, run_distributed

REF = "/root/reference/src/accelerate/test_utils/scripts"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = {"ACCELERATE_USE_CPU": "yes", "PYTHONDONTWRITEBYTECODE": "1"}


def launch(script, world):
    runner = os.path.join("..", "refshim", "run_reference_script.py")  # run_distributed resolves under tests/scripts
    out = run_distributed(runner, world, script, env=ENV, timeout=900)
    assert out.count(f"REFERENCE_SCRIPT_OK {script}") == world, out[-3000:]


def _single_process(script):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "refshim", "run_reference_script.py"), script],
                         env=dict(os.environ, **ENV), capture_output=True, text=True, timeout=600, cwd=ROOT)  # fmt: skip
    assert res.returncode == 0 and f"REFERENCE_SCRIPT_OK {script}" in res.stdout, res.stderr[-3000:]


def test_reference_distributed_scripts():
    """test_sync.py (mean-of-gradients, no_sync, accumulate, schedulers; 2 ranks), test_ops.py (gather / gather_object /
    broadcast / pad / reduce, debug-mode shape checks; 4 ranks), test_script.py (process control, RNG sync, dataloader
    preparation, training_check incl. bf16 and fp16-on-CPU, splits; 2 ranks and 1 process), test_distributed_data_loop.py
    (2 ranks) — run concurrently."""
    jobs = {
        "test_sync.py x2": lambda: launch("test_sync.py", 2),
        "test_ops.py x4": lambda: launch("test_ops.py", 4),
        "test_script.py x2": lambda: launch("test_script.py", 2),
        "test_distributed_data_loop.py x2": lambda: launch("test_distributed_data_loop.py", 2),
        "test_script.py x1": lambda: _single_process("test_script.py"),
    }
    run_concurrently(jobs, max_workers=3)


# ---- the reference's unit-test FILES (reference/tests/*.py) under pytest, same import alias --------------------------
# Deselected, with the reason: environment artefacts of this container, or features outside the data-parallel scope.
UNIT_FILES = {
    "tests/test_data_loader.py": [],  # all 31, the torchdata StatefulDataLoader cases included
    "tests/test_optimizer.py": [],
    "tests/test_scheduler.py": [],
    "tests/test_state_checkpointing.py": [],
    "tests/test_grad_sync.py": [],
    "tests/test_cpu.py": [],  # test_ops + test_script under debug_launcher
    "tests/test_kwargs_handlers.py": [],
    "tests/test_logging.py": [],
    "tests/test_memory_utils.py": [],
    "tests/test_utils.py": [
        "::UtilsTester::test_patch_environment_key_exists",  # this image exports CC=/opt/gcc/bin/gcc; the test uses "CC"
        "::UtilsTester::test_convert_to_fp32",  # torch.compile / inductor cannot build in this container
    ],
    "tests/test_accelerator.py": [
        "::AcceleratorTester::test_prepare_model_8bit_cpu_offload_raises_valueerror_not_typeerror",  # bitsandbytes path
        # compares psutil's SYSTEM-wide available RAM before/after: meaningless while other tests run concurrently
        "::AcceleratorTester::test_free_memory_dereferences_prepared_components",
    ],
}


def _run_unit_file(path):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "refshim", "run_reference_script.py"), "--pytest", path]
    for node in UNIT_FILES[path]:
        cmd += ["--deselect", node]
    res = subprocess.run(cmd, env=dict(os.environ, **ENV), capture_output=True, text=True, timeout=900, cwd="/tmp")
    summary = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
    ok = res.returncode == 0 and " passed" in summary and " failed" not in summary
    return path, ok, summary, res.stdout[-3000:] + res.stderr[-2000:]


def test_reference_unit_test_files():
    """Eleven files of the reference's unit-test suite, run concurrently (each in its own interpreter)."""
    def checked(path):
        _, ok, summary, tail = _run_unit_file(path)
        assert ok, tail
        return summary

    results = run_concurrently({path: (lambda p=path: checked(p)) for path in sorted(UNIT_FILES)}, max_workers=4)
    print("\n".join(f"{path}: {summary}" for path, summary in results.items()))
