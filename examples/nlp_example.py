"""The training loop of the reference's ``examples/nlp_example.py`` (lines 149-189) on synthetic MRPC-shaped
batches (there is no network here for GLUE / the BERT checkpoint): BERT-base-cased-shaped sequence classifier,
AdamW lr 2e-5, linear warm-up of 100 steps, seed 42, accuracy over ``gather_for_metrics``.

    python examples/nlp_example.py --cpu --tiny                       # BASELINE.json config 1 (plumbing, 1 process)
    torchrun --nproc-per-node 8 examples/nlp_example.py --mixed_precision bf16

Only the import differs from a script written for the reference.
"""

import argparse
import os
import sys

import torch
from torch.utils.data import DataLoader, Dataset

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from accelerate_b200 import Accelerator  # noqa: E402  (the reference: `from accelerate import Accelerator`)
from accelerate_b200.utils import set_seed  # noqa: E402

MAX_GPU_BATCH_SIZE = 16
EVAL_BATCH_SIZE = 32


class SyntheticMRPC(Dataset):
    """Sentence pairs of 40..128 tokens with a binary label that is a (learnable) function of the tokens."""

    def __init__(self, n, vocab, seed):
        g = torch.Generator().manual_seed(seed)
        self.lengths = torch.randint(40, 129, (n,), generator=g).tolist()
        self.ids = [torch.randint(5, vocab, (length,), generator=g) for length in self.lengths]
        self.labels = [int(ids[:8].sum() % 2) for ids in self.ids]

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, i):
        ids = self.ids[i]
        first = len(ids) // 2
        return {"input_ids": ids, "token_type_ids": torch.cat([torch.zeros(first), torch.ones(len(ids) - first)]).long(),
                "labels": self.labels[i]}  # fmt: skip


def collate(samples):  # pad to the longest sequence in the batch, like tokenizer.pad(padding="longest")
    longest = max(len(s["input_ids"]) for s in samples)

    def pad(t):
        return torch.cat([t, t.new_zeros(longest - len(t))])

    return {
        "input_ids": torch.stack([pad(s["input_ids"]) for s in samples]),
        "token_type_ids": torch.stack([pad(s["token_type_ids"]) for s in samples]),
        "attention_mask": torch.stack([pad(torch.ones(len(s["input_ids"]), dtype=torch.long)) for s in samples]),
        "labels": torch.tensor([s["labels"] for s in samples]),
    }


def training_function(config, args):
    from transformers import BertConfig, BertForSequenceClassification, get_linear_schedule_with_warmup

    accelerator = Accelerator(cpu=args.cpu, mixed_precision=args.mixed_precision)
    lr, num_epochs, seed, batch_size = config["lr"], int(config["num_epochs"]), int(config["seed"]), int(config["batch_size"])

    gradient_accumulation_steps = 1
    if batch_size > MAX_GPU_BATCH_SIZE:
        gradient_accumulation_steps = batch_size // MAX_GPU_BATCH_SIZE
        batch_size = MAX_GPU_BATCH_SIZE

    set_seed(seed)
    if args.tiny:
        cfg = BertConfig(vocab_size=2000, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128)
    else:
        cfg = BertConfig(vocab_size=28996)  # bert-base-cased shape, random init
    train_dl = DataLoader(SyntheticMRPC(args.train_samples, cfg.vocab_size, 1), shuffle=True, collate_fn=collate,
                          batch_size=batch_size)  # fmt: skip
    eval_dl = DataLoader(SyntheticMRPC(args.eval_samples, cfg.vocab_size, 2), shuffle=False, collate_fn=collate,
                         batch_size=EVAL_BATCH_SIZE)  # fmt: skip
    model = BertForSequenceClassification(cfg)
    model = model.to(accelerator.device)
    optimizer = torch.optim.AdamW(params=model.parameters(), lr=lr)
    lr_scheduler = get_linear_schedule_with_warmup(
        optimizer=optimizer, num_warmup_steps=100,
        num_training_steps=(len(train_dl) * num_epochs) // gradient_accumulation_steps,
    )  # fmt: skip
    model, optimizer, train_dl, eval_dl, lr_scheduler = accelerator.prepare(model, optimizer, train_dl, eval_dl, lr_scheduler)

    for epoch in range(num_epochs):
        model.train()
        for step, batch in enumerate(train_dl):
            # (the prepared dataloader already placed the batch on accelerator.device)
            outputs = model(**batch)
            loss = outputs.loss / gradient_accumulation_steps
            accelerator.backward(loss)
            if step % gradient_accumulation_steps == 0:
                optimizer.step()
                lr_scheduler.step()
                optimizer.zero_grad()

        model.eval()
        correct = total = 0
        for batch in eval_dl:
            with torch.no_grad():
                outputs = model(**batch)
            predictions = outputs.logits.argmax(dim=-1)
            predictions, references = accelerator.gather_for_metrics((predictions, batch["labels"]))
            correct += int((predictions == references).sum())
            total += references.numel()
        assert total == args.eval_samples, (total, args.eval_samples)  # gather_for_metrics drops the duplicated tail
        accelerator.print(f"epoch {epoch}:", {"accuracy": correct / total, "loss": float(loss.detach())})
    accelerator.end_training()
    return correct / total


def main():
    parser = argparse.ArgumentParser(description="Simple example of training script.")
    parser.add_argument("--mixed_precision", type=str, default=None, choices=["no", "fp16", "bf16"])
    parser.add_argument("--cpu", action="store_true", help="If passed, will train on the CPU.")
    parser.add_argument("--tiny", action="store_true", help="2-layer model: seconds on a CPU")
    parser.add_argument("--num_epochs", type=int, default=3)
    parser.add_argument("--train_samples", type=int, default=3668)  # MRPC train / validation sizes
    parser.add_argument("--eval_samples", type=int, default=408)
    args = parser.parse_args()
    config = {"lr": 2e-5, "num_epochs": args.num_epochs, "seed": 42, "batch_size": 16}
    training_function(config, args)


if __name__ == "__main__":
    main()
