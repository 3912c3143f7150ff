"""-m gpu, single process: the communication-hook plug point on a CUDA module (torch's
``hook(state, bucket) -> Future`` protocol, reference call site dataclasses.py:234-237).  Runs last among the GPU
files on purpose: this path was added after round 1's final GPU run."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_user_comm_hook_on_cuda_single_process():
    from accelerate_b200 import Accelerator, B200DataParallel

    acc = Accelerator()
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(64, 257), torch.nn.GELU(), torch.nn.Linear(257, 33)).cuda()
    ref = torch.nn.Sequential(torch.nn.Linear(64, 257), torch.nn.GELU(), torch.nn.Linear(257, 33)).cuda()
    ref.load_state_dict(model.state_dict())
    ddp = B200DataParallel(model, device_ids=[0], bucket_cap_mb=0.02)
    assert len(ddp.reducer.buckets) > 1
    calls = []

    def hook(state, bucket):
        calls.append((bucket.index(), bucket.is_last(), bucket.buffer().is_cuda))
        fut = torch.futures.Future()
        fut.set_result(bucket.buffer() * state["factor"])  # a new tensor: the reducer copies it back into the bucket
        return fut

    ddp.register_comm_hook({"factor": 3.0}, hook)
    x = torch.randn(8, 64, device="cuda")
    for step in range(2):  # first step adopts fresh gradients into the fusion buffer, second finds them there
        calls.clear()
        ddp.zero_grad(set_to_none=(step == 0))
        ref.zero_grad()
        ddp(x).square().mean().backward()
        ref(x).square().mean().backward()
        torch.cuda.synchronize()
        assert [c[0] for c in calls] == list(range(len(ddp.reducer.buckets))) and calls[-1][1] and all(c[2] for c in calls)
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad * 3.0, rtol=1e-6, atol=1e-9), step
    # the fused statistics are not available behind a user hook: clipping takes the multi-tensor kernels
    assert ddp.reducer.stats_valid is False
    total = acc.clip_grad_norm_(ddp.parameters(), 0.1)
    want = torch.sqrt(sum((q.grad.double() * 3.0).pow(2).sum() for q in ref.parameters()))
    assert acc.last_clip_mode == "multi_tensor" and abs(total.item() - want.item()) <= 1e-5 * want.item()
    from accelerate_b200.state import PartialState

    PartialState().runtime.ctx.check_status()
