"""SURVEY.md §8f row 4 on CUDA ranks (reference accelerator.py:3579-3895 save_state / load_state, checkpointing.py:63-318):
train through the fused path (B200DataParallel + fused clip + fused AdamW whose moments live in flat buffers and whose
`step` counters are views of one vector), save mid-run, keep training, then resume into FRESH objects and replay: the
resumed run must reproduce the uninterrupted one bit for bit (weights, optimizer moments, step counters, learning rate),
on every rank, with the reference's file names on disk."""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def build():
    torch.manual_seed(11)
    model = torch.nn.Sequential(torch.nn.Linear(32, 129), torch.nn.GELU(), torch.nn.Linear(129, 64), torch.nn.GELU(), torch.nn.Linear(64, 8))
    opt = torch.optim.AdamW(model.parameters(), lr=3e-3, weight_decay=0.05)
    sched = torch.optim.lr_scheduler.StepLR(opt, 1, gamma=0.9)
    return model, opt, sched


def main():
    from accelerate_b200 import Accelerator, DistributedDataParallelKwargs
    from accelerate_b200.state import AcceleratorState, GradientState, PartialState

    def fresh():
        AcceleratorState._reset_state()
        GradientState._reset_state()
        acc = Accelerator(mixed_precision="bf16", kwargs_handlers=[DistributedDataParallelKwargs(bucket_cap_mb=0.01)])
        model, opt, sched = acc.prepare(*build())
        return acc, model, opt, sched

    def steps(acc, model, opt, sched, first, count):
        for i in range(first, first + count):
            g = torch.Generator().manual_seed(1000 * acc.process_index + i)
            x = torch.randn(16, 32, generator=g).to(acc.device)
            acc.backward(model(x).float().square().mean())
            acc.clip_grad_norm_(model.parameters(), 0.5)
            opt.step()
            sched.step()
            opt.zero_grad()

    acc, model, opt, sched = fresh()
    rank, world = acc.process_index, acc.num_processes
    root = tempfile.mkdtemp() if rank == 0 else None
    box = [root]
    if world > 1:
        torch.distributed.broadcast_object_list(box, src=0)
    ckpt = os.path.join(box[0], "ckpt")
    steps(acc, model, opt, sched, 0, 3)
    acc.save_state(ckpt)
    acc.wait_for_everyone()
    names = sorted(os.listdir(ckpt))
    assert "model.safetensors" in names and "optimizer.bin" in names and "scheduler.bin" in names, names
    assert all(f"random_states_{r}.pkl" in names for r in range(world)), names
    steps(acc, model, opt, sched, 3, 3)
    assert getattr(opt, "_fused_plan", None) is not None, "the fused AdamW step must be the one under test"
    want = {k: v.detach().clone() for k, v in acc.unwrap_model(model).state_dict().items()}
    want_opt = opt.state_dict()
    want_lr = sched.get_last_lr()

    if world > 1:
        model.reducer.close()
    acc2, model2, opt2, sched2 = fresh()
    acc2.load_state(ckpt)
    steps(acc2, model2, opt2, sched2, 3, 3)
    got = acc2.unwrap_model(model2).state_dict()
    for k, v in want.items():
        assert torch.equal(got[k], v), (k, (got[k] - v).abs().max().item())
    got_opt = opt2.state_dict()
    assert sched2.get_last_lr() == want_lr
    for idx, st in want_opt["state"].items():
        for key in ("step", "exp_avg", "exp_avg_sq"):
            assert torch.equal(got_opt["state"][idx][key].cpu(), st[key].cpu()), (idx, key)
    PartialState().runtime.ctx.check_status()
    acc2.wait_for_everyone()
    print(f"RESUME_GPU_OK rank {rank}/{world}", flush=True)
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
