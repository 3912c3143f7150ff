"""LocalSGD (reference: src/accelerate/local_sgd.py, examples/by_feature/local_sgd.py): no gradient sync inside the
context, parameters averaged every K steps and on exit == mean of the ranks' parameters (oracle.reduce "mean")."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from scenarios import make_batch, make_model  # noqa: E402

from oracle import ddp_oracle as oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    args = ap.parse_args()
    cpu = args.device == "cpu"
    if cpu:
        os.environ["ACCELERATE_USE_CPU"] = "yes"
        torch.set_num_threads(1)
    from accelerate_b200 import Accelerator, LocalSGD
    from accelerate_b200.state import PartialState

    acc = Accelerator(cpu=cpu)
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    model = acc.prepare(make_model())
    opt = acc.prepare(torch.optim.SGD(model.parameters(), lr=0.05))

    exact = world == 2  # one commutative add; with more ranks gloo's ring / the NVSwitch add in their own order

    def same(a, b):
        return a.tobytes() == b.tobytes() if exact else np.allclose(a, b, rtol=1e-6, atol=1e-8)

    def snapshot():
        return [acc.gather(p.detach().reshape(1, -1)).cpu().numpy() for p in model.parameters()]

    with LocalSGD(accelerator=acc, model=model, local_sgd_steps=2, enabled=True) as local_sgd:
        for step in range(3):
            x, y = (t.to(dev) for t in make_batch(rank, 20 + step))
            acc.backward(torch.nn.functional.mse_loss(model(x), y))
            assert model.reducer.num_reductions == 0, "LocalSGD must not reduce gradients"
            opt.step()
            opt.zero_grad()
            before = snapshot()
            local_sgd.step()
            after = snapshot()
            for b, a in zip(before, after):
                if step == 1:  # K = 2: averaged now
                    want = oracle.reduce([b[r] for r in range(world)], "mean")
                    for r in range(world):
                        assert same(a[r], want), "parameters are not the rank mean"
                    assert all(a[0].tobytes() == a[r].tobytes() for r in range(world)), "ranks must agree bit for bit"
                else:  # untouched, and (after local steps on different data) different per rank
                    assert a.tobytes() == b.tobytes()
                    assert not all(b[0].tobytes() == b[r].tobytes() for r in range(1, world)) or step == 2
        before = snapshot()
    after = snapshot()  # averaged on exit
    for b, a in zip(before, after):
        want = oracle.reduce([b[r] for r in range(world)], "mean")
        assert all(same(a[r], want) for r in range(world))
    if not cpu:
        PartialState().runtime.ctx.check_status()
    acc.wait_for_everyone()
    print(f"LOCAL_SGD_OK rank {rank}", flush=True)
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
