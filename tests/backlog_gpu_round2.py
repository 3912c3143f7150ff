"""Round-2 GPU validation backlog — NOT collected by default (the file name does not match ``test_*.py``).

Multi-process CUDA legs of features that landed after round 1's GPU budget was spent; their host logic is covered on
gloo by tests/test_dist_cpu.py, the CUDA legs below still have to be seen green once on a GPU box:

    python -m pytest tests/backlog_gpu_round2.py -m gpu -x -q

Every launch bounds the device-side flag waits to 30 s (B200DP_TIMEOUT_MS) so a divergence fails instead of hanging.
When they pass, move them into tests/test_gpu_dist.py.
"""

import pytest

from conftest import run_distributed

pytestmark = pytest.mark.gpu
ENV = {"B200DP_TIMEOUT_MS": "30000"}


@pytest.mark.parametrize("find_unused", ["0", "1"])
def test_join_uneven_inputs_cuda(find_unused):
    out = run_distributed("dist_join.py", 2, "--device", "cuda", env=dict(ENV, JOIN_FIND_UNUSED=find_unused), timeout=300)
    assert out.count("JOIN_OK") == 2, out[-3000:]


def test_unused_parameters_cuda():
    out = run_distributed("dist_unused.py", 2, "--device", "cuda", env=ENV, timeout=300)
    assert out.count("UNUSED_OK") == 2, out[-3000:]


def test_user_and_powersgd_hooks_cuda():
    out = run_distributed("dist_comm_hook.py", 2, "--device", "cuda", env=dict(ENV, B200DP_TEST_CUDA_HOOKS="1"), timeout=600)
    assert out.count("COMM_HOOK_OK") == 2, out[-3000:]
