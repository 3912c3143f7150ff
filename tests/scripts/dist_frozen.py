"""torch DDP's `_sync_module_states` makes EVERY parameter and buffer equal to rank 0's at wrap time,
frozen parameters included; the Join exit (`_sync_final_model`) and LocalSGD's averaging cover them too (the reference's
LocalSGD averages all `model.parameters()`, local_sgd.py:99-107).  Ranks build the model with DIFFERENT seeds."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    os.environ["ACCELERATE_USE_CPU"] = "yes"
    torch.set_num_threads(1)
    from accelerate_b200 import Accelerator
    from accelerate_b200.local_sgd import LocalSGD

    acc = Accelerator(cpu=True)
    rank, world = acc.process_index, acc.num_processes
    torch.manual_seed(100 + rank)  # per-rank initialisation on purpose
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.BatchNorm1d(8), torch.nn.Linear(8, 4))
    for p in model[0].parameters():
        p.requires_grad_(False)  # frozen feature extractor
    model[1].running_mean.fill_(float(rank))
    model = acc.prepare(model)

    def all_equal(t):
        everyone = acc.gather(t.detach().reshape(1, -1))
        return bool((everyone == everyone[0]).all())

    for name, p in model.module.named_parameters():
        assert all_equal(p), f"{name} differs across ranks after prepare"
    for name, b in model.module.named_buffers():
        assert all_equal(b.float()), f"buffer {name} differs across ranks after prepare"
    # LocalSGD: the frozen parameters take part in the average as well (here: ranks nudge them apart by hand)
    opt = acc.prepare(torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.1))
    with LocalSGD(accelerator=acc, model=model, local_sgd_steps=2, enabled=True) as local_sgd:
        for step in range(2):
            x = torch.randn(6, 8, generator=torch.Generator().manual_seed(7 * rank + step))
            with acc.accumulate(model):
                acc.backward(model(x).square().mean())
                opt.step()
                opt.zero_grad()
            with torch.no_grad():
                model.module[0].weight.add_(0.01 * (rank + 1))
            local_sgd.step()
    for name, p in model.module.named_parameters():
        assert all_equal(p), f"{name} differs across ranks after LocalSGD averaging"
    print(f"FROZEN_OK rank {rank}/{world}", flush=True)
    acc.end_training()


if __name__ == "__main__":
    main()
