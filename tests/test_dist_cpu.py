"""World-size-2 and -4 runs on CPU (gloo): the N>1 host path — bucket table, autograd hooks, forward-time
sync latch, accumulate/no_sync, trigger_sync_in_backward, parameter broadcast, collectives — checked
against the golden outputs of the reference (tests/scripts/dist_check.py)."""

import pytest

from conftest import run_distributed


@pytest.mark.parametrize("world", [2, 4])
def test_dist_check_cpu(world):
    out = run_distributed("dist_check.py", world, "--device", "cpu", timeout=600)
    assert out.count("DIST_CHECK_OK") == world, out[-3000:]


def test_comm_hooks_cpu():
    out = run_distributed("dist_comm_hook.py", 2, "--device", "cpu", timeout=600)
    assert out.count("COMM_HOOK_OK") == 2, out[-3000:]


@pytest.mark.parametrize("world", [2, 4])
def test_training_matches_single_process(world):
    out = run_distributed("dist_training.py", world, "--device", "cpu", timeout=600)
    assert out.count("TRAINING_OK") == world, out[-3000:]


def test_local_sgd_cpu():
    out = run_distributed("dist_local_sgd.py", 2, "--device", "cpu", timeout=600)
    assert out.count("LOCAL_SGD_OK") == 2, out[-3000:]


def test_bert_accumulation_clip_buffers_cpu():
    """BASELINE.json config 4 shape (scaled down): bf16, gradient_accumulation_steps=4, clip on the sync step, buffers."""
    out = run_distributed("dist_bert.py", 2, "--device", "cpu", timeout=600)
    assert out.count("BERT_CHECK_OK") == 2, out[-3000:]


def test_unused_parameters_match_torch_ddp():
    """find_unused_parameters=True: grad=None for globally unused parameters, accumulated gradients reduced as they
    are — bit-equal to torch DistributedDataParallel on the same process group."""
    out = run_distributed("dist_unused.py", 2, timeout=300)
    assert out.count("UNUSED_OK") == 2, out[-3000:]


@pytest.mark.parametrize("seed,unused,hook", [(7, "0", ""), (3, "0", "bf16"), (11, "1", ""), (5, "0", "fp16")])
def test_differential_against_torch_ddp(seed, unused, hook):
    """A random, rank-consistent sequence of sync / no_sync / zero_grad / eval / repeated-forward steps on a model with
    tied weights, frozen and bf16 parameters, a module applied twice and BatchNorm buffers — with ``unused`` the ranks
    also take different branches under find_unused_parameters=True, with ``hook`` both sides compress (torch's
    fp16/bf16_compress_hook vs the wire-format path): gradients (including which are None) AND buffers
    bit-equal to torch DistributedDataParallel (gloo, N=2)."""
    out = run_distributed("dist_ddp_diff.py", 2, timeout=300, env={"DDP_DIFF_SEED": str(seed), "DDP_DIFF_UNUSED": unused, "DDP_DIFF_HOOK": hook,
                               "DDP_DIFF_NO_BCAST": "1" if seed == 3 else "0"})  # fmt: skip
    assert out.count("DDP_DIFF_OK") == 2, out[-3000:]


def test_differential_against_torch_ddp_reentrant_checkpointing():
    """with ``checkpoint(use_reentrant=True)`` the first gradient hook of a backward fires inside a
    nested graph task; the end-of-backward finalizer must wait for the OUTER task.  Same random action sequence as
    above, two checkpointed segments, bit-equal to torch DDP."""
    out = run_distributed("dist_ddp_diff.py", 2, timeout=300, env={"DDP_DIFF_SEED": "13", "DDP_DIFF_UNUSED": "0", "DDP_DIFF_HOOK": "",
                               "DDP_DIFF_NO_BCAST": "0", "DDP_DIFF_CKPT": "1"})  # fmt: skip
    assert out.count("DDP_DIFF_OK") == 2, out[-3000:]


@pytest.mark.parametrize("world,find_unused", [(2, "0"), (2, "1"), (4, "0")])
def test_join_uneven_inputs_matches_torch_ddp(world, find_unused):
    """torch's Join context (behind the reference's join_uneven_inputs): ranks train 3+2r iterations, early finishers
    shadow the collectives; parameters after every step and the final synced model equal torch DDP's."""
    out = run_distributed("dist_join.py", world, timeout=300, env={"JOIN_FIND_UNUSED": find_unused})
    assert out.count("JOIN_OK") == world, out[-3000:]


def test_frozen_parameters_and_buffers_are_synchronised():
    """every parameter — frozen ones too — and every buffer equals rank 0's after prepare (torch's
    `_sync_module_states`), and LocalSGD averages the frozen ones like the reference does."""
    out = run_distributed("dist_frozen.py", 2, timeout=300)
    assert out.count("FROZEN_OK") == 2, out[-3000:]
