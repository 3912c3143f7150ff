"""ONE user training script, run twice: with the reference (``import accelerate`` from /root/reference/src) and with
this package (``accelerate_b200``), 2 ranks on CPU/gloo.  Everything a loop observes is written to a JSON file per
rank — sync_gradients flags, losses, learning rates, gradient norms returned by clip_grad_norm_, step_was_skipped,
gathered evaluation outputs, a digest of the final parameters — and tests/test_reference_loop_parity.py compares the
two files field by field.  The loop mixes accumulate(), an uneven last batch, clip, scheduler, seeded shuffling,
gather_for_metrics and a second epoch; the variant is chosen by --variant.
"""
import argparse
import hashlib
import json
import os
import sys

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ours", "reference"], required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--variant", default="all")
    args = ap.parse_args()
    os.environ["ACCELERATE_USE_CPU"] = "yes"
    torch.set_num_threads(1)
    if args.impl == "reference":
        sys.path.insert(0, "/root/reference/src")
        import accelerate as lib
        from accelerate.state import AcceleratorState, GradientState
        from accelerate.utils import DataLoaderConfiguration, GradientAccumulationPlugin, set_seed
    else:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        import accelerate_b200 as lib
        from accelerate_b200.state import AcceleratorState, GradientState
        from accelerate_b200.utils import DataLoaderConfiguration, GradientAccumulationPlugin, set_seed

    variants = [
        dict(accum=2, n=37, bs=4, split=False, dispatch=False, sync_each_batch=False, seedable=True, drop_last=False),
        dict(accum=3, n=40, bs=4, split=True, dispatch=False, sync_each_batch=False, seedable=False, drop_last=True),
        dict(accum=2, n=29, bs=3, split=False, dispatch=True, sync_each_batch=True, seedable=False, drop_last=False),
        dict(accum=1, n=26, bs=5, split=False, dispatch=False, sync_each_batch=False, seedable=True, drop_last=False),
        dict(accum=2, n=33, bs=4, split=False, dispatch=False, sync_each_batch=False, seedable=True, drop_last=False, mp="bf16"),
        dict(accum=3, n=31, bs=2, split=False, dispatch=False, sync_each_batch=False, seedable=False, drop_last=False,
             sync_with_dataloader=False, adjust_scheduler=False),
        dict(accum=2, n=24, bs=4, split=True, dispatch=True, sync_each_batch=False, seedable=False, drop_last=False,
             step_with_opt=False),
    ]
    chosen = range(len(variants)) if args.variant == "all" else [int(args.variant)]
    logs = {}
    for vi in chosen:
        if args.impl == "reference":
            AcceleratorState._reset_state(True)
        else:
            AcceleratorState._reset_state()
        GradientState._reset_state()
        logs[str(vi)] = run_variant(lib, variants[vi], DataLoaderConfiguration, GradientAccumulationPlugin, set_seed)
    import torch.distributed as dist

    rank = dist.get_rank()
    dist.barrier()
    with open(f"{args.out}.rank{rank}.json", "w") as f:
        json.dump(logs, f)
    print(f"LOOP_DONE {args.impl} rank {rank}", flush=True)
    dist.destroy_process_group()


def run_variant(lib, v, DataLoaderConfiguration, GradientAccumulationPlugin, set_seed):
    accelerator = lib.Accelerator(
        cpu=True, mixed_precision=v.get("mp"), step_scheduler_with_optimizer=v.get("step_with_opt", True),
        gradient_accumulation_plugin=GradientAccumulationPlugin(
            num_steps=v["accum"], sync_each_batch=v["sync_each_batch"], sync_with_dataloader=v.get("sync_with_dataloader", True),
            adjust_scheduler=v.get("adjust_scheduler", True)),
        dataloader_config=DataLoaderConfiguration(split_batches=v["split"], dispatch_batches=v["dispatch"],
                                                  use_seedable_sampler=v["seedable"]),
    )  # fmt: skip
    set_seed(42)
    rank = accelerator.process_index
    g = torch.Generator().manual_seed(5)
    xs, ws = torch.randn(v["n"], 12, generator=g), torch.randn(12, 3, generator=g)
    ys = (xs @ ws).argmax(dim=1)
    train = torch.utils.data.TensorDataset(xs, ys, torch.arange(v["n"]))
    evald = torch.utils.data.TensorDataset(xs[: v["n"] - 2], ys[: v["n"] - 2])
    model = torch.nn.Sequential(torch.nn.Linear(12, 20), torch.nn.Tanh(), torch.nn.BatchNorm1d(20), torch.nn.Linear(20, 3))
    optimizer = torch.optim.AdamW(model.parameters(), lr=3e-2, weight_decay=0.01)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=2, gamma=0.7)
    train_dl = torch.utils.data.DataLoader(train, batch_size=v["bs"], shuffle=True, drop_last=v["drop_last"])
    eval_dl = torch.utils.data.DataLoader(evald, batch_size=v["bs"])
    model, optimizer, train_dl, eval_dl, scheduler = accelerator.prepare(model, optimizer, train_dl, eval_dl, scheduler)

    log = {"variant": v, "world": accelerator.num_processes, "steps": [], "eval": []}
    loss_fn = torch.nn.CrossEntropyLoss()
    for epoch in range(2):
        model.train()
        if hasattr(train_dl, "set_epoch"):
            train_dl.set_epoch(epoch)
        for x, y, idx in train_dl:
            with accelerator.accumulate(model):
                loss = loss_fn(model(x), y)
                accelerator.backward(loss)
                norm = None
                if accelerator.sync_gradients:
                    norm = float(accelerator.clip_grad_norm_(model.parameters(), 0.5))
                optimizer.step()
                scheduler.step()
                optimizer.zero_grad()
            log["steps"].append({
                "epoch": epoch, "idx": idx.tolist(), "sync": bool(accelerator.sync_gradients), "loss": float(loss.detach()).hex(),
                "norm": None if norm is None else norm.hex(), "lr": scheduler.get_last_lr()[0].hex(),
                "skipped": bool(accelerator.optimizer_step_was_skipped),
                "end": bool(accelerator.gradient_state.end_of_dataloader),
            })  # fmt: skip
        model.eval()
        preds, refs = [], []
        for x, y in eval_dl:
            with torch.no_grad():
                p = model(x).argmax(dim=1)
            p, r = accelerator.gather_for_metrics((p, y))
            preds += p.tolist()
            refs += r.tolist()
        log["eval"].append({"preds": preds, "refs": refs})
    unwrapped = accelerator.unwrap_model(model)
    digest = hashlib.sha256()
    for name, t in sorted(unwrapped.state_dict().items()):
        digest.update(name.encode())
        digest.update(t.detach().cpu().contiguous().numpy().tobytes())
    log["final_state_sha256"] = digest.hexdigest()
    log["reduced_probe"] = [float(a).hex() for a in accelerator.reduce(torch.tensor([1.5 + rank, -2.0 * rank]), "mean").tolist()]
    accelerator.wait_for_everyone()
    return log


if __name__ == "__main__":
    main()
