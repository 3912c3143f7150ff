"""Uneven inputs under torch's ``Join`` context (the machinery behind the reference's ``join_uneven_inputs``,
accelerator.py:1299-1372), differentially against torch DistributedDataParallel on the same process group (gloo, CPU):
rank r trains 3 + 2r iterations — some under no_sync — on a model with BatchNorm buffers; the ranks that finish early
shadow the others' collectives.  After every step the still-training ranks, and at the end all ranks, must hold exactly
the parameters / buffers torch DDP produces."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def net():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(6, 10), torch.nn.BatchNorm1d(10), torch.nn.Tanh(), torch.nn.Linear(10, 2))


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    args = ap.parse_args()
    cuda = args.device == "cuda"
    if not cuda:
        os.environ["ACCELERATE_USE_CPU"] = "yes"
    torch.set_num_threads(1)
    import torch.distributed as dist
    from torch.distributed.algorithms.join import Join
    from torch.nn.parallel import DistributedDataParallel

    from accelerate_b200 import B200DataParallel
    from accelerate_b200.state import PartialState

    state = PartialState(cpu=not cuda)
    rank, world = state.process_index, state.num_processes
    find_unused = os.environ.get("JOIN_FIND_UNUSED", "0") == "1"
    dev = state.device
    if cuda:  # ours on the GPU through the kernels; torch DDP stays on the CPU over its own gloo group
        ours = B200DataParallel(net().to(dev), device_ids=[dev.index], bucket_cap_mb=0.0005, find_unused_parameters=find_unused)
        ref = DistributedDataParallel(net(), bucket_cap_mb=0.0005, find_unused_parameters=find_unused,
                                      process_group=dist.new_group(backend="gloo"))  # fmt: skip
    else:
        ours = B200DataParallel(net(), bucket_cap_mb=0.0005, find_unused_parameters=find_unused)
        ref = DistributedDataParallel(net(), bucket_cap_mb=0.0005, find_unused_parameters=find_unused)
    trace = {}
    for name, model in (("ref", ref), ("ours", ours)):
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.5)
        g = torch.Generator().manual_seed(50 + rank)
        snaps = []
        with (model.join() if find_unused else Join([model])):  # both spellings of the context
            for it in range(3 + 2 * rank):
                x, y = torch.randn(5, 6, generator=g), torch.randn(5, 2, generator=g)
                if cuda and name == "ours":
                    x, y = x.to(dev), y.to(dev)
                if it % 3 == 1:  # an accumulation micro-step: the joined ranks must learn that nobody reduces
                    with model.no_sync():
                        torch.nn.functional.mse_loss(model(x), y).backward()
                    continue
                torch.nn.functional.mse_loss(model(x), y).backward()
                opt.step()
                opt.zero_grad()
                snaps.append([p.detach().cpu().clone() for p in model.parameters()])
        final = [t.detach().cpu().clone() for t in list(model.parameters()) + list(model.buffers())]
        trace[name] = (snaps, final)
    def same(p, q):  # two ranks: the mean is exact; more: gloo's ring associates differently from our rank order
        if cuda:  # GPU vs CPU arithmetic
            return torch.allclose(p, q, rtol=1e-3, atol=1e-5)
        return torch.equal(p, q) if world == 2 else torch.allclose(p, q, rtol=1e-4, atol=1e-6)

    for k, (a, b) in enumerate(zip(trace["ref"][0], trace["ours"][0])):
        for p, q in zip(a, b):
            assert same(p, q), (rank, "step", k, (p - q).abs().max().item())
    for p, q in zip(trace["ref"][1], trace["ours"][1]):
        assert same(p, q), (rank, "final", (p - q).abs().max().item())
    # everyone ends with the model of the rank that trained longest
    flat = torch.cat([t.reshape(-1).float() for t in trace["ours"][1]])
    gloo = dist.new_group(backend="gloo") if cuda else None
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat, group=gloo)
    assert all(torch.equal(gathered[0], t) for t in gathered)
    if cuda:
        state.runtime.ctx.check_status()
    print(f"JOIN_OK rank {rank} ({len(trace['ours'][0])} optimizer steps)", flush=True)
    state.destroy_process_group()


if __name__ == "__main__":
    main()
