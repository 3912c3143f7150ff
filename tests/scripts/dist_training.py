"""End-to-end training equals single-process training (reference pin: test_utils/scripts/test_script.py:449-620
`training_check`, ATOL=RTOL=1e-6): 3 epochs of SGD on y = a*x + b, (1) per-rank batches dealt by the prepared
dataloader, (2) split_batches, (3) with gradient accumulation + scheduler (test_sync.py:248-303, LR equality),
(4) bf16 autocast.  Every rank must end with the baseline's parameters and identical bits across ranks."""
import argparse
import os
import sys

import torch
from torch.utils.data import DataLoader

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


class RegressionDataset:
    def __init__(self, a=2.0, b=3.0, length=96, seed=42):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.randn(length, generator=g)
        self.y = a * self.x + b + 0.1 * torch.randn(length, generator=g)

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return {"x": self.x[i], "y": self.y[i]}


class RegressionModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(0.0))
        self.b = torch.nn.Parameter(torch.tensor(0.0))

    def forward(self, x):
        return x * self.a + self.b


def baseline(batch_size, epochs=3, accumulate=1, with_sched=False):
    torch.manual_seed(0)
    model = RegressionModel()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 0.9**s) if with_sched else None
    dl = DataLoader(RegressionDataset(), batch_size=batch_size)
    lrs = []
    for _ in range(epochs):
        for i, batch in enumerate(dl):
            loss = torch.nn.functional.mse_loss(model(batch["x"]), batch["y"]) / accumulate
            loss.backward()
            if (i + 1) % accumulate == 0 or i == len(dl) - 1:
                opt.step()
                if sched is not None:
                    sched.step()
                opt.zero_grad()
            lrs.append(opt.param_groups[0]["lr"])
    return model.a.item(), model.b.item(), lrs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    args = ap.parse_args()
    cpu = args.device == "cpu"
    if cpu:
        os.environ["ACCELERATE_USE_CPU"] = "yes"
        torch.set_num_threads(1)
    from accelerate_b200 import Accelerator, DataLoaderConfiguration
    from accelerate_b200.state import AcceleratorState, GradientState, PartialState

    def run(split, precision="no", accumulate=1, with_sched=False):
        AcceleratorState._reset_state()
        GradientState._reset_state()
        acc = Accelerator(cpu=cpu, mixed_precision=precision, gradient_accumulation_steps=accumulate,
                          dataloader_config=DataLoaderConfiguration(split_batches=split))  # fmt: skip
        world = acc.num_processes
        torch.manual_seed(0)
        model = RegressionModel()
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 0.9**s) if with_sched else None
        dl = DataLoader(RegressionDataset(), batch_size=16 if split else 16 // world)
        if with_sched:
            model, opt, dl, sched = acc.prepare(model, opt, dl, sched)
        else:
            model, opt, dl = acc.prepare(model, opt, dl)
        lrs = []
        for _ in range(3):
            for batch in dl:
                with acc.accumulate(model):
                    loss = torch.nn.functional.mse_loss(model(batch["x"]), batch["y"])
                    acc.backward(loss)
                    opt.step()
                    if sched is not None:
                        sched.step()
                    opt.zero_grad()
                lrs.append(opt.param_groups[0]["lr"])
        inner = acc.unwrap_model(model)
        ab = torch.stack([inner.a.detach(), inner.b.detach()]).reshape(1, 2)
        everyone = acc.gather(ab).cpu()
        assert all(torch.equal(everyone[0], everyone[r]) for r in range(world)), "ranks diverged"
        return inner.a.item(), inner.b.item(), lrs, acc

    a0, b0, _ = baseline(16)
    for split in (False, True):
        a, b, _, acc = run(split)
        assert abs(a - a0) <= 1e-6 + 1e-6 * abs(a0) and abs(b - b0) <= 1e-6 + 1e-6 * abs(b0), (split, a, a0, b, b0)
    a, b, _, acc = run(False, precision="bf16")
    assert abs(a - a0) < 2e-2 and abs(b - b0) < 2e-2, ("bf16", a, a0, b, b0)
    # accumulation (2 micro-steps) + scheduler: the scheduler steps `num_processes` times per optimizer step
    # unless split_batches, so compare with a baseline whose lambda is evaluated at the same step counts
    world = acc.num_processes
    a2, b2, lrs2 = baseline(16, accumulate=2, with_sched=False)
    a, b, lrs, acc = run(True, accumulate=2, with_sched=False)
    assert abs(a - a2) <= 1e-5 and abs(b - b2) <= 1e-5, ("accum", a, a2, b, b2)
    a3, b3, lrs3 = baseline(16, accumulate=2, with_sched=True)
    a, b, lrs, acc = run(True, accumulate=2, with_sched=True)
    assert abs(a - a3) <= 1e-5 and abs(b - b3) <= 1e-5, ("accum+sched", a, a3, b, b3)
    assert all(abs(x - y) <= 1e-9 for x, y in zip(lrs, lrs3)), "learning-rate trajectories differ"
    if not cpu:
        PartialState().runtime.ctx.check_status()
    acc.wait_for_everyone()
    print(f"TRAINING_OK rank {acc.process_index}", flush=True)
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
