"""-m gpu: multi-rank parity of the P2P kernels through the public API, against the reference's golden
outputs and the oracle.  With fewer GPUs than ranks the ranks share a GPU (PartialState picks gloo for
the rendezvous group, the data path is unchanged)."""

import pytest
import torch

from conftest import run_distributed

pytestmark = pytest.mark.gpu

ENV = {"B200DP_TIMEOUT_MS": "60000"}


@pytest.mark.parametrize("kind", ["fd", "ipc"])
def test_two_ranks_against_golden(kind):
    out = run_distributed("dist_check.py", 2, "--device", "cuda", env={**ENV, "B200DP_SEG_KIND": kind}, timeout=900)
    assert out.count("DIST_CHECK_OK") == 2, out[-3000:]


def test_four_ranks_against_golden():
    out = run_distributed("dist_check.py", 4, "--device", "cuda", env=ENV, timeout=1200)
    assert out.count("DIST_CHECK_OK") == 4, out[-3000:]


def test_gpt2_small_grads_two_ranks():
    """BASELINE.json config 2 shape: GPT-2-small-like block stack, fp32, 2 ranks, grads vs oracle + torch DDP semantics."""
    out = run_distributed("dist_gpt2.py", 2, env=ENV, timeout=1200)
    assert out.count("GPT2_CHECK_OK") == 2, out[-3000:]


def test_comm_hooks_two_ranks():
    """fp16/bf16 wire format (DDPCommunicationHookType) vs the fp32 wire, reference tolerance 1e-2."""
    out = run_distributed("dist_comm_hook.py", 2, "--device", "cuda", env=ENV, timeout=900)
    assert out.count("COMM_HOOK_OK") == 2, out[-3000:]


def test_training_matches_single_process_two_ranks():
    """reference test_script.py:449-620 `training_check`: distributed SGD == single-process SGD (1e-6)."""
    out = run_distributed("dist_training.py", 2, "--device", "cuda", env=ENV, timeout=900)
    assert out.count("TRAINING_OK") == 2, out[-3000:]


def test_local_sgd_two_ranks():
    """LocalSGD parameter averaging through the fusion buffer (pack, one allreduce, unpack, divide)."""
    out = run_distributed("dist_local_sgd.py", 2, "--device", "cuda", env=ENV, timeout=900)
    assert out.count("LOCAL_SGD_OK") == 2, out[-3000:]


def test_bert_large_shaped_accumulation_and_clip_two_ranks():
    """BASELINE.json config 4: BERT-large-shaped seq-cls (1024/16/4096), bf16, accumulate(4) + fused clip + buffer broadcast."""
    out = run_distributed("dist_bert.py", 2, "--device", "cuda", env=ENV, timeout=1200)
    assert out.count("BERT_CHECK_OK") == 2, out[-3000:]


# ---- CUDA legs of features whose host logic is pinned on gloo by tests/test_dist_cpu.py (round-1 backlog) ----
SHORT = {"B200DP_TIMEOUT_MS": "30000"}  # a divergence fails in 30 s instead of hanging


@pytest.mark.parametrize("find_unused", ["0", "1"])
def test_join_uneven_inputs_cuda(find_unused):
    """reference accelerator.py:1299-1380 `join_uneven_inputs` (pinned by test_distributed_data_loop.py:162-183): ranks
    train 3+2r iterations, early finishers shadow every bucket reduction with zeros through the same kernels."""
    out = run_distributed("dist_join.py", 2, "--device", "cuda", env=dict(SHORT, JOIN_FIND_UNUSED=find_unused), timeout=300)
    assert out.count("JOIN_OK") == 2, out[-3000:]


def test_unused_parameters_cuda():
    """find_unused_parameters=True (dataclasses.py:181-195): the used-map allreduce runs over b200dp_push."""
    out = run_distributed("dist_unused.py", 2, "--device", "cuda", env=SHORT, timeout=300)
    assert out.count("UNUSED_OK") == 2, out[-3000:]


def test_user_and_powersgd_hooks_cuda():
    """PowerSGD / wrapped / user hooks through torch's plug point on CUDA ranks (dataclasses.py:200-237; tolerance of
    test_ddp_comm_hook.py:57)."""
    out = run_distributed("dist_comm_hook.py", 2, "--device", "cuda", env=dict(SHORT, B200DP_TEST_CUDA_HOOKS="1"), timeout=600)
    assert out.count("COMM_HOOK_OK") == 2, out[-3000:]


def test_three_ranks_two_shot_bit_exact():
    """A non-power-of-two world never takes the NVLS kernel (its scale-after-sum differs from the oracle's pre-divide
    in the last ulp): bit-exact vs the oracle at N=3."""
    out = run_distributed("dist_allreduce_check.py", 3, env=ENV, timeout=600)
    assert out.count("ALLREDUCE_CHECK_OK") == 3, out[-3000:]


@pytest.mark.parametrize("world,pipe_chunk_mb,fused", [(2, "1", "1"), (4, "8", "1"), (2, "8", "0")])
def test_large_gather_and_reduce_fused_kernels(world, pipe_chunk_mb, fused):
    """accelerator.gather / reduce above the latency path: one launch per staging window — b200dp_gather (stage + pull),
    b200dp_reduce (fused below 8 MiB, the pipelined comm-CTA / copy-CTA kernel above; 1 MiB chunks force a dozen
    hand-offs per call) — ragged sizes, odd source offsets, all float dtypes, mean, messages larger than the window
    (32 MiB staging forces several launches).  fused=0: the round-1 three-launch sequences the sweep compares against."""
    env = dict(ENV, B200DP_STAGING_MB="32", B200DP_PIPE_CHUNK_MB=pipe_chunk_mb, B200DP_FUSED_GATHER=fused, B200DP_FUSED_REDUCE=fused,
               B200DP_PIPE_REDUCE="2" if pipe_chunk_mb == "1" else "1",  # "2": the pipelined kernel for two-shot worlds too
               B200DP_FUSED_GATHER_MIN_MB="0",  # the fused gather kernel at every size above the latency path
               B200DP_PUSH2_ROUND_KB="16" if pipe_chunk_mb == "1" else "128")  # two ranks: several push rounds per CTA
    out = run_distributed("dist_collectives_large.py", world, env=env, timeout=900)
    assert out.count("COLLECTIVES_LARGE_OK") == world, out[-3000:]


@pytest.mark.parametrize("world", [1, 2])
def test_save_state_load_state_resume_cuda(world):
    """SURVEY.md §8f row 4 on the GPU path: mid-run save_state, resume into fresh objects, replay == uninterrupted run
    bit for bit (weights, fused-Adam moments and step counters, scheduler), reference file names on disk."""
    out = run_distributed("dist_resume_gpu.py", world, env=ENV, timeout=600)
    assert out.count("RESUME_GPU_OK") == world, out[-3000:]
