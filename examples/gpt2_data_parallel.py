"""BASELINE.json config 3 as a user script: GPT-2-medium, bf16 autocast, synthetic 512-token batches, AdamW,
``clip_grad_norm_(1.0)`` every step, one process per GPU.

    torchrun --nproc-per-node 8 examples/gpt2_data_parallel.py --steps 50
    python examples/gpt2_data_parallel.py --model tiny --cpu --steps 3          # plumbing check without a GPU

What runs underneath on B200s: gradients are views of a symmetric fusion buffer, one fused NVLink allreduce kernel
per 25 MiB bucket overlaps with backward, clip_grad_norm_ reads the statistics that kernel left behind, and the AdamW
step is one multi-tensor kernel over the same buffer.  ``bench.py`` is the measured version of this.
"""

import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from accelerate_b200 import Accelerator, DistributedDataParallelKwargs  # noqa: E402
from accelerate_b200.utils import set_seed  # noqa: E402

SHAPES = {"gpt2": (768, 12, 12), "gpt2-medium": (1024, 24, 16), "tiny": (64, 2, 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt2-medium", choices=sorted(SHAPES))
    ap.add_argument("--batch", type=int, default=8, help="per-GPU micro-batch (weak scaling)")
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--accumulate", type=int, default=1)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    from transformers import GPT2Config, GPT2LMHeadModel

    accelerator = Accelerator(
        cpu=args.cpu, mixed_precision=None if args.cpu else "bf16", gradient_accumulation_steps=args.accumulate,
        kwargs_handlers=[DistributedDataParallelKwargs(bucket_cap_mb=25)],
    )  # fmt: skip
    set_seed(42)  # identical initial weights everywhere (prepare() broadcasts rank 0's anyway)
    n_embd, n_layer, n_head = SHAPES[args.model]
    vocab = 50257 if args.model != "tiny" else 1000
    model = GPT2LMHeadModel(GPT2Config(n_embd=n_embd, n_layer=n_layer, n_head=n_head, vocab_size=vocab, n_positions=max(args.seq, 64)))
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-4)
    model, optimizer = accelerator.prepare(model, optimizer)
    g = torch.Generator().manual_seed(1234 + accelerator.process_index)  # every rank its own data
    model.train()
    grad_norm = float("nan")
    t0 = time.perf_counter()
    for step in range(args.steps):
        ids = torch.randint(0, vocab, (args.batch, args.seq), generator=g).to(accelerator.device)
        with accelerator.accumulate(model):
            loss = model(input_ids=ids, labels=ids).loss
            accelerator.backward(loss)
            if accelerator.sync_gradients:
                grad_norm = accelerator.clip_grad_norm_(model.parameters(), 1.0)
            optimizer.step()
            optimizer.zero_grad()
        if step % 5 == 0 or step == args.steps - 1:
            mean_loss = accelerator.reduce(loss.detach(), "mean")  # one small P2P allreduce
            accelerator.print(f"step {step}: loss {mean_loss.item():.4f} grad_norm {float(grad_norm):.3f}")
    accelerator.wait_for_everyone()
    dt = time.perf_counter() - t0
    accelerator.print(f"{args.steps * args.batch * accelerator.num_processes / dt:.1f} samples/s (wall clock, incl. warm-up)")
    accelerator.end_training()


if __name__ == "__main__":
    main()
