#!/usr/bin/env python
"""Where does a training step spend its time?  CUDA-event phase times of the bench.py step (GPT-2-medium, bf16,
per-GPU batch 8 x 512) on the compute stream, max over ranks:

    forward | backward (incl. the join with the reduction stream = exposed communication) | clip | optimizer | zero_grad

run three ways: our step, our step under ``no_sync`` (no reduction at all: the difference in the backward phase is
what the gradient reduction really costs), and stock torch DDP + NCCL.  One JSON line on rank 0.

    python benchmarks/step_breakdown.py                                   # 1 GPU
    torchrun --nproc-per-node 8 benchmarks/step_breakdown.py --steps 30   # 8 GPUs
"""
import argparse
import contextlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PHASES = ["forward", "backward", "clip", "optimizer", "zero_grad"]


def build(model_name):
    from transformers import GPT2Config, GPT2LMHeadModel

    shape = {"gpt2": (768, 12, 12), "gpt2-medium": (1024, 24, 16), "tiny": (128, 2, 4)}[model_name]
    torch.manual_seed(42)
    return GPT2LMHeadModel(GPT2Config(n_embd=shape[0], n_layer=shape[1], n_head=shape[2]))


def run_phases(step_fns, batches, steps, warmup, reduce_max):
    """step_fns: list of callables, one per phase, each taking the batch index; returns mean ms per phase."""
    n = len(step_fns)
    for i in range(warmup):
        for fn in step_fns:
            fn(i)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(steps)]
    for i in range(steps):
        ev[i][0].record()
        for k, fn in enumerate(step_fns):
            fn(i)
            ev[i][k + 1].record()
    torch.cuda.synchronize()
    ms = [sum(ev[i][k].elapsed_time(ev[i][k + 1]) for i in range(steps)) / steps for k in range(n)]
    total = sum(ev[i][0].elapsed_time(ev[i][n]) for i in range(steps)) / steps
    return reduce_max(ms + [total])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt2-medium")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-ddp", action="store_true")
    args = ap.parse_args()
    from accelerate_b200 import Accelerator, DistributedDataParallelKwargs

    acc = Accelerator(mixed_precision="bf16", kwargs_handlers=[DistributedDataParallelKwargs(bucket_cap_mb=25)])
    dev, world, rank = acc.device, acc.num_processes, acc.process_index
    g = torch.Generator().manual_seed(1234 + rank)
    batches = [torch.randint(0, 50257, (args.batch, args.seq), generator=g).to(dev) for _ in range(8)]

    def reduce_max(values):
        t = torch.tensor(values, device=dev, dtype=torch.float32).reshape(1, -1)
        return (acc.gather(t) if world > 1 else t).max(dim=0).values.tolist()

    result = {"n_gpus": world, "model": args.model, "per_gpu_batch": args.batch, "seq_len": args.seq, "steps": args.steps,
              "unit": "ms per step, max over ranks", "phases": PHASES + ["total"]}  # fmt: skip

    # ---- ours ----
    model = build(args.model).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    model, opt = acc.prepare(model, opt)
    box = {}

    def make_steps(sync_ctx):
        def forward(i):
            box["ctx"] = sync_ctx()
            box["ctx"].__enter__()
            box["loss"] = model(input_ids=batches[i % 8], labels=batches[i % 8]).loss

        def backward(i):
            acc.backward(box["loss"])
            box["ctx"].__exit__(None, None, None)

        def clip(i):
            acc.clip_grad_norm_(model.parameters(), 1.0)

        def optimizer(i):
            opt.step()

        def zero(i):
            opt.zero_grad()

        return [forward, backward, clip, optimizer, zero]

    result["ours"] = run_phases(make_steps(contextlib.nullcontext), batches, args.steps, args.warmup, reduce_max)
    if world > 1:
        result["ours_no_sync"] = run_phases(make_steps(lambda: acc.no_sync(model)), batches, args.steps, args.warmup, reduce_max)
        result["reduction_cost_ms"] = result["ours"][1] - result["ours_no_sync"][1]
    result["optimizer_step"] = "fused" if getattr(opt, "_fused_plan", None) is not None else "torch"
    result["clip_mode"] = acc.last_clip_mode
    del model, opt
    torch.cuda.empty_cache()

    # ---- stock torch DDP + NCCL (single GPU: plain module) ----
    import torch.distributed as dist

    if not args.no_ddp and (world == 1 or dist.get_backend() == "nccl"):
        ref = build(args.model).to(dev).train()
        if world > 1:
            ref = torch.nn.parallel.DistributedDataParallel(ref, device_ids=[dev.index], bucket_cap_mb=25)
        ropt = torch.optim.AdamW(ref.parameters(), lr=1e-4)

        def r_forward(i):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                box["loss"] = ref(input_ids=batches[i % 8], labels=batches[i % 8]).loss

        steps = [r_forward, lambda i: box["loss"].backward(), lambda i: torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0),
                 lambda i: ropt.step(), lambda i: ropt.zero_grad()]  # fmt: skip
        result["torch_ddp_nccl"] = run_phases(steps, batches, args.steps, args.warmup, reduce_max)
    acc.wait_for_everyone()
    if rank == 0:
        print(json.dumps(result), flush=True)
    acc.end_training()


if __name__ == "__main__":
    main()
