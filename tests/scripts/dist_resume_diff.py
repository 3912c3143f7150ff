"""Mid-epoch checkpoint / resume across the two libraries, in ONE process group (CPU/gloo, 2 ranks): for saver and
loader in {reference, this package}: train to the middle of epoch 0, save_state; build everything afresh with the
loader library, load_state, skip_first_batches, finish — the final weights must equal an uninterrupted run's, bit for
bit, in all four combinations (on-disk format compatibility)."""
import argparse
import hashlib
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STOP_AT = 3  # optimizer steps taken in epoch 0 before the checkpoint


def libs():
    sys.path.insert(0, "/root/reference/src")
    sys.path.insert(0, ROOT)
    out = {}
    for name, pkg in (("reference", "accelerate"), ("ours", "accelerate_b200")):
        lib = importlib.import_module(pkg)
        state = importlib.import_module(pkg + ".state")
        utils = importlib.import_module(pkg + ".utils")
        out[name] = (lib, state, utils)
    return out


def fresh(lib, state, utils, name):
    state.AcceleratorState._reset_state(True) if name == "reference" else state.AcceleratorState._reset_state()
    state.GradientState._reset_state()
    acc = lib.Accelerator(cpu=True, dataloader_config=utils.DataLoaderConfiguration(use_seedable_sampler=True))
    utils.set_seed(42)
    g = torch.Generator().manual_seed(5)
    xs = torch.randn(40, 6, generator=g)
    ys = xs.sum(dim=1, keepdim=True)
    model = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    opt = torch.optim.AdamW(model.parameters(), lr=2e-2)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=3, gamma=0.5)
    dl = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xs, ys), batch_size=4, shuffle=True)
    model, opt, dl, sched = acc.prepare(model, opt, dl, sched)
    return acc, model, opt, dl, sched


def train(acc, model, opt, sched, batches):
    for x, y in batches:
        acc.backward(torch.nn.functional.mse_loss(model(x), y))
        acc.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        sched.step()
        opt.zero_grad()


def digest(acc, model, sched):
    h = hashlib.sha256()
    for k, t in sorted(acc.unwrap_model(model).state_dict().items()):
        h.update(k.encode())
        h.update(t.detach().numpy().tobytes())
    h.update(repr(sched.get_last_lr()).encode())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", required=True)
    args = ap.parse_args()
    os.environ["ACCELERATE_USE_CPU"] = "yes"
    torch.set_num_threads(1)
    L = libs()
    full = {}
    for name, (lib, state, utils) in L.items():
        acc, model, opt, dl, sched = fresh(lib, state, utils, name)
        for epoch in range(2):
            dl.set_epoch(epoch)
            train(acc, model, opt, sched, dl)
        full[name] = digest(acc, model, sched)
    assert full["reference"] == full["ours"], full
    for saver, (lib, state, utils) in L.items():
        acc, model, opt, dl, sched = fresh(lib, state, utils, saver)
        dl.set_epoch(0)
        train(acc, model, opt, sched, (b for i, b in enumerate(dl) if i < STOP_AT))
        ckpt = os.path.join(args.dir, f"by_{saver}")
        acc.save_state(ckpt)
        acc.wait_for_everyone()
    rank = acc.process_index
    for saver in L:
        for loader, (lib, state, utils) in L.items():
            acc, model, opt, dl, sched = fresh(lib, state, utils, loader)
            acc.load_state(os.path.join(args.dir, f"by_{saver}"))
            dl.set_epoch(0)
            train(acc, model, opt, sched, acc.skip_first_batches(dl, STOP_AT))
            dl.set_epoch(1)
            train(acc, model, opt, sched, dl)
            got = digest(acc, model, sched)
            assert got == full["ours"], (saver, loader, "resumed run diverged from the uninterrupted one")
            acc.wait_for_everyone()
    files = sorted(os.listdir(os.path.join(args.dir, "by_ours")))
    assert files == sorted(os.listdir(os.path.join(args.dir, "by_reference"))), files
    print(f"RESUME_OK rank {rank} files {files}", flush=True)
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
