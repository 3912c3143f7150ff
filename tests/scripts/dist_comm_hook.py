"""Compress hooks (reference: src/accelerate/test_utils/scripts/test_ddp_comm_hook.py:39-57): gradients
produced with DistributedDataParallelKwargs(comm_hook=FP16|BF16) (or comm_wrapper=...) must match the
no-hook gradients within rtol=atol=1e-2, and be identical on all ranks."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from scenarios import make_batch, make_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    args = ap.parse_args()
    cpu = args.device == "cpu"
    if cpu:
        os.environ["ACCELERATE_USE_CPU"] = "yes"
        torch.set_num_threads(1)
    from accelerate_b200 import Accelerator, DDPCommunicationHookType, DistributedDataParallelKwargs
    from accelerate_b200.state import AcceleratorState, GradientState, PartialState

    results = {}
    for name, kw in [("no", {}), ("fp16", {"comm_hook": DDPCommunicationHookType.FP16}),
                     ("bf16", {"comm_hook": DDPCommunicationHookType.BF16}),
                     ("bf16_wrapper", {"comm_wrapper": DDPCommunicationHookType.BF16})]:  # fmt: skip
        AcceleratorState._reset_state()
        GradientState._reset_state()
        acc = Accelerator(cpu=cpu, kwargs_handlers=[DistributedDataParallelKwargs(**kw)])
        rank, world, dev = acc.process_index, acc.num_processes, acc.device
        model = acc.prepare(make_model())
        assert model.reducer.wire_dtype == {"no": None, "fp16": torch.float16}.get(name, torch.bfloat16)
        for step in range(2):  # second step: grads are already views of the fusion buffer
            x, y = (t.to(dev) for t in make_batch(rank, step))
            model.zero_grad(set_to_none=(step == 0))
            acc.backward(torch.nn.functional.mse_loss(model(x), y))
            grads = [p.grad.detach().float().cpu().numpy().copy() for p in model.parameters()]
            everyone = acc.gather(torch.cat([p.grad.reshape(-1) for p in model.parameters()]).reshape(1, -1)).cpu().numpy()
            assert all(everyone[0].tobytes() == everyone[r].tobytes() for r in range(world)), (name, "ranks differ")
            results[(name, step)] = grads
        if not cpu:  # the fused clip still works on the widened fp32 buffer
            ref = [p.grad.detach().clone() for p in model.parameters()]
            tn = acc.clip_grad_norm_(model.parameters(), 0.01)
            want = torch.sqrt(sum((g.double() ** 2).sum() for g in ref)).item()
            assert acc.last_clip_mode == "fused" and abs(tn.item() - want) <= 1e-5 * want, (name, tn.item(), want)
    for (name, step), grads in results.items():
        for g, base in zip(grads, results[("no", step)]):
            assert np.allclose(g, base, rtol=1e-2, atol=1e-2), (name, step, np.abs(g - base).max())
            if name != "no":
                assert not np.array_equal(g, base) or np.all(base == 0), (name, "hook had no effect")
    if not cpu:
        PartialState().runtime.ctx.check_status()
    acc.wait_for_everyone()
    print(f"COMM_HOOK_OK rank {acc.process_index}", flush=True)
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
