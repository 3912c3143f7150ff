"""Prepared dataloaders under the reference and under this package (CPU/gloo), many configurations in one process
group: every batch each rank sees over two epochs, the end_of_dataloader / remainder bookkeeping, what
gather_for_metrics returns per batch, len(), total_dataset_length, and skip_first_batches — dumped to JSON per rank
and compared by tests/test_reference_loop_parity.py."""
import argparse
import itertools
import json
import os
import sys

import torch


class Stream(torch.utils.data.IterableDataset):
    def __init__(self, n):
        self.n = n

    def __iter__(self):
        return iter(torch.arange(self.n).tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ours", "reference"], required=True)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    os.environ["ACCELERATE_USE_CPU"] = "yes"
    torch.set_num_threads(1)
    if args.impl == "reference":
        sys.path.insert(0, "/root/reference/src")
        import accelerate as lib
        from accelerate.state import AcceleratorState, GradientState
        from accelerate.utils import DataLoaderConfiguration, set_seed
    else:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        import accelerate_b200 as lib
        from accelerate_b200.state import AcceleratorState, GradientState
        from accelerate_b200.utils import DataLoaderConfiguration, set_seed

    grid = []
    for length, bs, drop_last, split, mode, shuffle, even in itertools.product(
        (9, 16, 23), (2, 4), (False, True), (False, True), ("shard", "dispatch", "iterable", "iterable_dispatch"),
        (False, True), (True, False),
    ):
        if mode != "shard" and (shuffle or not even):
            continue  # iterable datasets cannot shuffle; the dispatcher has no even_batches=False
        if not even and (split or drop_last):
            continue
        grid.append((length, bs, drop_last, split, mode, shuffle, even))
    log = {"configs": len(grid), "runs": []}
    world = int(os.environ.get("WORLD_SIZE", "1"))
    for length, bs, drop_last, split, mode, shuffle, even in grid:
        if "dispatch" in mode and drop_last and length < (bs if split else bs * world):
            continue  # not one full global batch: the reference's dispatcher crashes broadcasting `None`
        AcceleratorState._reset_state(True) if args.impl == "reference" else AcceleratorState._reset_state()
        GradientState._reset_state()
        cfg = DataLoaderConfiguration(split_batches=split, dispatch_batches=(True if "dispatch" in mode else False),
                                      even_batches=even, use_seedable_sampler=shuffle)  # fmt: skip
        acc = lib.Accelerator(cpu=True, dataloader_config=cfg)
        world = acc.num_processes
        if split and bs % world:
            continue
        if os.environ.get("DL_TRACE") and acc.process_index == 0:
            print("CFG", length, bs, drop_last, split, mode, shuffle, even, file=sys.stderr, flush=True)
        set_seed(1234)
        data = Stream(length) if mode.startswith("iterable") else list(range(length))
        dl = torch.utils.data.DataLoader(data, batch_size=bs, drop_last=drop_last, shuffle=shuffle)
        dl = acc.prepare(dl)
        run = {"cfg": [length, bs, drop_last, split, mode, shuffle, even], "epochs": []}
        try:
            run["len"] = len(dl)
        except TypeError:
            run["len"] = None
        for key, attr in (("total", "total_dataset_length"), ("total_bs", "total_batch_size")):
            try:
                run[key] = getattr(dl, attr)
            except (TypeError, AttributeError) as err:  # e.g. the reference's property calls len() on an IterableDataset
                run[key] = type(err).__name__
        for epoch in range(2):
            if hasattr(dl, "set_epoch"):
                dl.set_epoch(epoch)
            seen = []
            for batch in dl:
                gs = acc.gradient_state
                entry = {"batch": batch.tolist(), "end": bool(gs.end_of_dataloader), "remainder": gs.remainder}
                if even:  # gather needs equal shapes on every rank
                    entry["metrics"] = acc.gather_for_metrics(batch).tolist()
                seen.append(entry)
            run["epochs"].append(seen)
            run.setdefault("after", []).append([bool(acc.gradient_state.end_of_dataloader), acc.gradient_state.remainder,
                                                bool(acc.gradient_state.in_dataloader)])  # fmt: skip
        if even and "dispatch" not in mode:  # (the reference's dispatcher crashes when skipping exhausts a short dataset)
            run["skipped"] = [b.tolist() for b in acc.skip_first_batches(dl, 1)]
        log["runs"].append(run)
    rank = acc.process_index
    acc.wait_for_everyone()
    with open(f"{args.out}.rank{rank}.json", "w") as f:
        json.dump(log, f)
    print(f"LOADER_DONE {args.impl} rank {rank} ({len(log['runs'])} configurations)", flush=True)
    acc.end_training()


if __name__ == "__main__":
    main()
