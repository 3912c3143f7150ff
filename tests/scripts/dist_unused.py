"""find_unused_parameters=True against torch DistributedDataParallel on the same process group (gloo, CPU): a
globally unused parameter keeps grad=None, a partially used one gets the mean with zeros, a gradient accumulated under
no_sync is reduced as it is, and an existing zero gradient stays zero (torch Reducer semantics)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


class TwoHeads(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(4, 4)
        self.b = torch.nn.Linear(4, 4)

    def forward(self, x, use_b):
        return self.b(x) if use_b else self.a(x)


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    cuda = ap.parse_args().device == "cuda"
    if not cuda:
        os.environ["ACCELERATE_USE_CPU"] = "yes"
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel

    from accelerate_b200 import B200DataParallel
    from accelerate_b200.state import PartialState

    state = PartialState(cpu=not cuda)
    rank = state.process_index
    torch.manual_seed(1 + rank)
    x = torch.randn(3, 4)
    if cuda:  # ours on the GPU (used-map allreduce through b200dp_push); torch DDP on the CPU over its own gloo group
        dev = state.device
        ours = B200DataParallel(TwoHeads().to(dev), device_ids=[dev.index], find_unused_parameters=True)
        ref = DistributedDataParallel(TwoHeads(), find_unused_parameters=True, process_group=dist.new_group(backend="gloo"))
    else:
        ours = B200DataParallel(TwoHeads(), find_unused_parameters=True)
        ref = DistributedDataParallel(TwoHeads(), find_unused_parameters=True)

    def same(tag):
        for (n, p), q in zip(ours.module.named_parameters(), ref.module.parameters()):
            assert (p.grad is None) == (q.grad is None), (tag, n, p.grad is None, q.grad is None)
            if p.grad is not None:
                got = p.grad.cpu()
                ok = torch.allclose(got, q.grad, rtol=1e-5, atol=1e-6) if cuda else torch.equal(got, q.grad)
                assert ok, (tag, n, (got - q.grad).abs().max())

    for m in (ours, ref):  # A: nobody uses b
        m(x, False).sum().backward()
    assert ours.module.b.weight.grad is None
    same("A")
    for m in (ours, ref):  # B: rank 0 uses a, rank 1 uses b -> both defined, mean with zeros
        m.zero_grad(set_to_none=True)
        m(x, rank == 1).sum().backward()
    assert ours.module.a.weight.grad is not None and ours.module.b.weight.grad is not None
    same("B")
    for m in (ours, ref):  # C: micro-step 1 (no_sync) uses b, micro-step 2 (sync) uses a
        m.zero_grad(set_to_none=True)
        with m.no_sync():
            m(x, True).sum().backward()
        m(x, False).sum().backward()
    assert float(ours.module.b.weight.grad.abs().sum()) > 0
    same("C")
    for m in (ours, ref):  # D: nobody uses b, but a (zeroed) gradient exists
        m.zero_grad(set_to_none=False)
        m(x, False).sum().backward()
    assert float(ours.module.b.weight.grad.abs().sum()) == 0.0
    same("D")
    if cuda:
        state.runtime.ctx.check_status()
    print(f"UNUSED_OK rank {rank}", flush=True)
    state.destroy_process_group()


if __name__ == "__main__":
    main()
