"""Differential test against torch DistributedDataParallel itself (gloo, CPU, 2 ranks: every float operation of the
mean is exact or commutative there, so the two wrappers must agree bit for bit): a random but rank-consistent sequence
of sync steps, no_sync micro-steps, zero_grad flavours, eval forwards and repeated forwards, on a model with tied
weights, an embedding, LayerNorm and BatchNorm buffers."""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.emb = torch.nn.Embedding(50, 16)
        self.norm = torch.nn.LayerNorm(16)
        self.fc = torch.nn.Linear(16, 16)
        self.bn = torch.nn.BatchNorm1d(16)
        self.head = torch.nn.Linear(16, 50, bias=False)
        self.head.weight = self.emb.weight  # tied, like GPT-2's wte / lm_head
        self.frozen = torch.nn.Linear(16, 16)  # never trained: not the reducer's business
        for p in self.frozen.parameters():
            p.requires_grad_(False)
        self.gain = torch.nn.Parameter(torch.ones(16, dtype=torch.bfloat16))  # a second gradient dtype (own buckets)
        self.side = torch.nn.Linear(16, 16)  # only used when `branch` says so (find_unused_parameters runs)
        self.branch = 0
        # reentrant activation checkpointing (HF gradient_checkpointing_enable() in many releases): the checkpointed
        # segments re-run autograd from inside a Function's backward, so gradient hooks fire in NESTED graph tasks —
        # the first one of a backward included (the tail segment is the first thing backward reaches)
        self.ckpt = os.environ.get("DDP_DIFF_CKPT", "0") == "1"

    def _mid(self, h):
        return self.fc(torch.tanh(self.fc(h)))  # one module applied twice: its gradients accumulate inside one backward

    def _tail(self, h):
        return self.bn(h + 0.1 * self.frozen(h)) * self.gain.float()

    def forward(self, ids):
        from torch.utils.checkpoint import checkpoint

        h = self.norm(self.emb(ids)).mean(dim=1)
        h = checkpoint(self._mid, h, use_reentrant=True) if self.ckpt else self._mid(h)
        if self.branch == 1:
            h = h + self.side(h)
        elif self.branch == 2:  # skips `fc` entirely on top of it
            h = self.side(self.norm(self.emb(ids)).mean(dim=1))
        h = checkpoint(self._tail, h, use_reentrant=True) if self.ckpt else self._tail(h)
        return self.head(h)


def main():
    os.environ["ACCELERATE_USE_CPU"] = "yes"
    torch.set_num_threads(1)
    from torch.nn.parallel import DistributedDataParallel

    from accelerate_b200 import B200DataParallel
    from accelerate_b200.state import PartialState

    state = PartialState(cpu=True)
    rank, world = state.process_index, state.num_processes
    assert world == 2
    unused = os.environ.get("DDP_DIFF_UNUSED", "0") == "1"  # ranks take different branches: locally unused parameters
    bcast = os.environ.get("DDP_DIFF_NO_BCAST", "0") != "1"  # broadcast_buffers=False: BatchNorm statistics stay per rank
    ours = B200DataParallel(Net(), bucket_cap_mb=0.0005, find_unused_parameters=unused, broadcast_buffers=bcast)  # several buckets
    ref = DistributedDataParallel(Net(), bucket_cap_mb=0.0005, find_unused_parameters=unused, broadcast_buffers=bcast)

    hook = os.environ.get("DDP_DIFF_HOOK", "")  # "fp16" / "bf16": torch's compress hooks vs the wire-format path
    if hook:
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks

        from accelerate_b200 import DDPCommunicationHookType

        ours.register_comm_hook(None, DDPCommunicationHookType(hook))
        ref.register_comm_hook(None, getattr(default_hooks, f"{hook}_compress_hook"))

    def pick_branch():
        choice = plan.choice([1, 1, 2, (1, 2), (2, 1), (1, 1)]) if unused else 1
        branch = choice[rank] if isinstance(choice, tuple) else choice
        ours.module.branch = ref.module.branch = branch
    assert len(ours.reducer.buckets) > 1
    data = torch.Generator().manual_seed(100 + rank)  # per-rank data
    plan = random.Random(int(os.environ.get("DDP_DIFF_SEED", "7")))  # the same sequence of actions on both ranks
    loss_fn = torch.nn.CrossEntropyLoss()

    def batch():
        return torch.randint(0, 50, (6, 5), generator=data), torch.randint(0, 50, (6,), generator=data)

    def compare(tag):
        for (n, p), q in zip(ours.module.named_parameters(), ref.module.parameters()):
            assert (p.grad is None) == (q.grad is None), (tag, n)
            if p.grad is not None:
                assert torch.equal(p.grad, q.grad), (tag, n, (p.grad - q.grad).abs().max().item())
        for (n, b), c in zip(ours.module.named_buffers(), ref.module.buffers()):
            assert torch.equal(b, c), (tag, "buffer", n)

    syncs = 0
    for it in range(40):
        action = plan.choice(["sync", "sync", "accumulate", "eval", "double_forward", "zero_keep", "zero_none", "sgd"])
        pick_branch()
        if action == "sync":
            x, y = batch()
            for m in (ours, ref):
                loss_fn(m(x), y).backward()
            syncs += 1
            compare((it, action))
        elif action == "accumulate":
            k = plan.randint(1, 3)
            micro = [batch() for _ in range(k + 1)]
            for m in (ours, ref):
                with m.no_sync():
                    for x, y in micro[:-1]:
                        loss_fn(m(x), y).backward()
                x, y = micro[-1]
                loss_fn(m(x), y).backward()
            syncs += 1
            compare((it, action, k))
        elif action == "eval":
            x, _ = batch()
            for m in (ours, ref):
                m.eval()
                with torch.no_grad():
                    m(x)
                m.train()
            compare((it, action))
        elif action == "double_forward":  # a no_grad forward, then the training forward and its backward
            # (the other order is an error in torch DDP itself: its buffer broadcast bumps BatchNorm's saved tensors)
            (x1, _), (x2, y2) = batch(), batch()
            for m in (ours, ref):
                with torch.no_grad():
                    m(x1)
                loss_fn(m(x2), y2).backward()
            syncs += 1
            compare((it, action))
        elif action == "sgd":  # parameters move identically -> later comparisons stay meaningful
            for m in (ours, ref):
                with torch.no_grad():
                    for p in m.parameters():
                        if p.grad is not None:
                            p.add_(p.grad, alpha=-0.05)
        else:
            for m in (ours, ref):
                m.zero_grad(set_to_none=(action == "zero_none"))
    assert syncs >= 5 and ours.reducer.num_reductions == syncs, (syncs, ours.reducer.num_reductions)
    print(f"DDP_DIFF_OK rank {rank} ({syncs} reductions compared)", flush=True)
    state.destroy_process_group()


if __name__ == "__main__":
    main()
