"""Generate the golden fixtures by running the REFERENCE (huggingface/accelerate) itself.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONPATH=/root/reference/src PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/sampler_shards.json   BatchSamplerShard / IterableDatasetShard index tables
  tests/golden/ops_w{2,4}.npz        gather / reduce / broadcast / pad_across_processes results (gloo)
  tests/golden/ddp_w{2,4}_{fp32,bf16}.npz   per-rank local grads, DDP-reduced grads, grads after
                                     accelerator.clip_grad_norm_, total norm, accumulation case

Everything goes through the reference's public API (Accelerator(cpu=True), accelerator.prepare,
accelerator.backward, accelerator.clip_grad_norm_, accelerator.gather/reduce/pad_across_processes,
utils.broadcast) with torch's stock DistributedDataParallel underneath.
"""

import json
import os
import socket
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from scenarios import SAMPLER_GRID, make_batch, make_model, ops_inputs  # noqa: E402


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# ------------------------------------------------------------------ samplers (single process)
def gen_samplers():
    from torch.utils.data import BatchSampler, SequentialSampler

    from accelerate.data_loader import BatchSamplerShard, IterableDatasetShard

    table = {"batch_sampler": [], "iterable": []}
    for n, length, bs, drop_last, split, even in SAMPLER_GRID:
        if split and bs % n:
            continue
        inner = BatchSampler(SequentialSampler(range(length)), batch_size=bs, drop_last=drop_last)
        shards = [
            [list(b) for b in BatchSamplerShard(inner, n, r, split_batches=split, even_batches=even)] for r in range(n)
        ]
        lens = [len(BatchSamplerShard(inner, n, r, split_batches=split, even_batches=even)) for r in range(n)]
        table["batch_sampler"].append({"n": n, "length": length, "bs": bs, "drop_last": drop_last, "split": split,
                                       "even": even, "shards": shards, "lens": lens})  # fmt: skip
        if even:  # IterableDatasetShard has no even_batches knob
            class Stream(torch.utils.data.IterableDataset):
                def __iter__(self_inner):
                    return iter(range(length))

            shards = [
                list(IterableDatasetShard(Stream(), batch_size=bs, drop_last=drop_last, num_processes=n,
                                          process_index=r, split_batches=split))
                for r in range(n)
            ]  # fmt: skip
            table["iterable"].append({"n": n, "length": length, "bs": bs, "drop_last": drop_last, "split": split,
                                      "shards": shards})  # fmt: skip
    with open(os.path.join(HERE, "sampler_shards.json"), "w") as f:
        json.dump(table, f, separators=(",", ":"))
    print("sampler cases:", len(table["batch_sampler"]), len(table["iterable"]))


# ------------------------------------------------------------------ distributed workers
def _worker(rank, world, port, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ACCELERATE_USE_CPU="yes")  # fmt: skip
    torch.set_num_threads(1)
    from accelerate import Accelerator
    from accelerate.state import AcceleratorState, GradientState, PartialState
    from accelerate.utils import broadcast

    out = {}
    # ---- ops ----
    acc = Accelerator(cpu=True)
    for name, t in ops_inputs(rank, world).items():
        out[f"gather/{name}"] = acc.gather(t.clone()).numpy()
        if t.dtype != torch.bool:
            out[f"reduce_sum/{name}"] = acc.reduce(t.clone(), "sum").numpy()
            if t.is_floating_point():
                out[f"reduce_mean/{name}"] = acc.reduce(t.clone(), "mean").numpy()
        b = t.clone()
        broadcast(b, from_process=world - 1)
        out[f"broadcast/{name}"] = b.numpy()
    ragged = torch.arange((rank + 2) * 3, dtype=torch.float32).reshape(rank + 2, 3) + 100 * rank
    out["pad/dim0"] = acc.pad_across_processes(ragged, dim=0).numpy()
    out["pad/dim0_first"] = acc.pad_across_processes(ragged, dim=0, pad_index=-1, pad_first=True).numpy()
    ragged1 = torch.arange(2 * (rank + 1), dtype=torch.int64).reshape(2, rank + 1) + 10 * rank
    out["pad/dim1"] = acc.pad_across_processes(ragged1, dim=1, pad_index=7).numpy()
    out["pad/dimneg"] = acc.pad_across_processes(ragged1, dim=-1, pad_index=7, pad_first=True).numpy()
    np.savez(os.path.join(outdir, f"ops_r{rank}.npz"), **out)

    # ---- DDP gradient path ----
    for precision in ("no", "bf16"):
        AcceleratorState._reset_state()
        GradientState._reset_state()
        res = {}
        acc = Accelerator(cpu=True, mixed_precision=precision)
        model = make_model()
        local = make_model()  # unwrapped twin: this rank's own gradient
        model = acc.prepare(model)
        x, y = make_batch(rank, 0)
        with acc.autocast():
            loss_local = torch.nn.functional.mse_loss(local(x), y)
        loss_local.backward()
        for i, p in enumerate(local.parameters()):
            res[f"local/{i}"] = p.grad.numpy().copy()
        loss = torch.nn.functional.mse_loss(model(x), y)
        acc.backward(loss)
        for i, p in enumerate(model.parameters()):
            res[f"ddp/{i}"] = p.grad.numpy().copy()
        tn = acc.clip_grad_norm_(model.parameters(), 0.05)
        res["clip/total_norm"] = np.asarray(tn.item(), dtype=np.float32)
        for i, p in enumerate(model.parameters()):
            res[f"clip/{i}"] = p.grad.numpy().copy()
        model.zero_grad()

        # accumulation: 2 micro-steps under accelerator.accumulate, sync on the second
        AcceleratorState._reset_state()
        GradientState._reset_state()
        acc = Accelerator(cpu=True, mixed_precision=precision, gradient_accumulation_steps=2)
        model2 = acc.prepare(make_model())
        local2 = make_model()
        for micro in range(2):
            x, y = make_batch(rank, 1 + micro)
            with acc.autocast():
                (torch.nn.functional.mse_loss(local2(x), y) / 2).backward()
            with acc.accumulate(model2):
                acc.backward(torch.nn.functional.mse_loss(model2(x), y))
            res[f"accum/sync_flag/{micro}"] = np.asarray(int(acc.sync_gradients))
            if micro == 0:
                for i, p in enumerate(model2.parameters()):
                    res[f"accum/micro0/{i}"] = p.grad.numpy().copy()
        for i, p in enumerate(local2.parameters()):
            res[f"accum/local/{i}"] = p.grad.numpy().copy()
        for i, p in enumerate(model2.parameters()):
            res[f"accum/ddp/{i}"] = p.grad.numpy().copy()
        tag = "fp32" if precision == "no" else "bf16"
        np.savez(os.path.join(outdir, f"ddp_{tag}_r{rank}.npz"), **res)
    # ---- prepared dataloaders: per-rank batches, end-of-epoch flags, gather_for_metrics truncation ----
    from scenarios import LOADER_GRID, loader_dataset

    loaders = []
    for length, bs, drop_last, split, dispatch, iterable in LOADER_GRID:
        if split and bs % world:
            continue
        AcceleratorState._reset_state()
        GradientState._reset_state()
        from accelerate.utils import DataLoaderConfiguration

        acc = Accelerator(cpu=True, dataloader_config=DataLoaderConfiguration(split_batches=split, dispatch_batches=dispatch))
        dl = torch.utils.data.DataLoader(loader_dataset(length, iterable), batch_size=bs, drop_last=drop_last)
        dl = acc.prepare(dl)
        rec = {"length": length, "bs": bs, "drop_last": drop_last, "split": split, "dispatch": dispatch,
               "iterable": iterable, "batches": [], "end_flags": [], "metrics": []}
        for batch in dl:
            rec["batches"].append(batch.tolist())
            rec["end_flags"].append(bool(acc.gradient_state.end_of_dataloader))
            rec["metrics"].append(acc.gather_for_metrics(batch).tolist())
        rec["remainder"] = int(dl.remainder)
        try:
            rec["len"] = len(dl)
        except TypeError:
            rec["len"] = None
        loaders.append(rec)
    with open(os.path.join(outdir, f"loaders_r{rank}.json"), "w") as f:
        json.dump(loaders, f)
    PartialState().destroy_process_group()


def gen_distributed(world):
    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as tmp:
        port = free_port()
        mp.start_processes(_worker, args=(world, port, tmp), nprocs=world, start_method="fork")
        ops = {}
        for r in range(world):
            with np.load(os.path.join(tmp, f"ops_r{r}.npz")) as z:
                for k in z.files:
                    ops[f"r{r}/{k}"] = z[k]
        np.savez_compressed(os.path.join(HERE, f"ops_w{world}.npz"), **ops)
        for tag in ("fp32", "bf16"):
            merged = {}
            for r in range(world):
                with np.load(os.path.join(tmp, f"ddp_{tag}_r{r}.npz")) as z:
                    for k in z.files:
                        merged[f"r{r}/{k}"] = z[k]
            np.savez_compressed(os.path.join(HERE, f"ddp_w{world}_{tag}.npz"), **merged)
        per_rank = [json.load(open(os.path.join(tmp, f"loaders_r{r}.json"))) for r in range(world)]
        with open(os.path.join(HERE, f"loaders_w{world}.json"), "w") as f:
            json.dump(per_rank, f, separators=(",", ":"))
    print("distributed fixtures for world", world)


if __name__ == "__main__":
    import accelerate

    assert "/root/reference" in accelerate.__file__, accelerate.__file__
    gen_samplers()
    for w in (2, 4):
        gen_distributed(w)
