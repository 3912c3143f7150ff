"""BASELINE.json config 2: GPT-2-small fp32 through accelerator.prepare on 2 ranks, synthetic 512-token
batches; gradients after one accelerator.backward must equal (bit for bit) the oracle's
sum_r fl(g_r / N) of the ranks' local gradients, be identical on all ranks, and — when every rank has
its own GPU so NCCL can run — match stock torch DistributedDataParallel + NCCL (what the reference
wraps with, accelerator.py:1892) within rtol=1e-3/atol=1e-5 (bit-exact at N=2)."""

import os
import sys

os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ddp_oracle as oracle  # noqa: E402


def build(seed=42, layers=12):
    from transformers import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(seed)
    cfg = GPT2Config(n_layer=layers, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, attn_implementation="eager")  # eager attention: deterministic backward
    return GPT2LMHeadModel(cfg)


def main():
    from accelerate_b200 import Accelerator, B200DataParallel
    from accelerate_b200.state import PartialState

    layers = int(os.environ.get("GPT2_LAYERS", "12"))
    torch.use_deterministic_algorithms(True, warn_only=True)
    acc = Accelerator()
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    model = acc.prepare(build(layers=layers))
    assert isinstance(model, B200DataParallel)
    twin = build(layers=layers).to(dev)
    g = torch.Generator().manual_seed(1234 + rank)
    ids = torch.randint(0, 50257, (2, 512), generator=g).to(dev)

    twin(input_ids=ids, labels=ids).loss.backward()
    loss = model(input_ids=ids, labels=ids).loss
    acc.backward(loss)
    nb = len(model.reducer.buckets)
    assert nb >= 10, nb  # 497.8 MB of fp32 grads in ~25 MiB buckets
    checked = 0
    for (name, p), tp in zip(model.module.named_parameters(), twin.parameters()):
        local_all = acc.gather(tp.grad.reshape(1, -1)).cpu().numpy()
        want = oracle.ddp_mean_reduce([local_all[r] for r in range(world)]).reshape(p.shape)
        got = p.grad.cpu().numpy()
        if model.reducer.effective_algo() == "nvls" and world > 2:
            assert np.allclose(got, want, rtol=1e-6, atol=1e-9), f"{name}: NVLS result beyond fp32 rounding of the oracle"
        else:
            assert got.tobytes() == want.tobytes(), f"{name}: reduced grad differs from oracle (max abs diff {np.abs(got - want).max()})"
        checked += got.size
    everyone = acc.gather(torch.stack([p.grad.double().sum() for p in model.parameters()]).reshape(1, -1)).cpu().numpy()
    assert all(everyone[0].tobytes() == everyone[r].tobytes() for r in range(world))
    assert checked == 124_439_808 or layers != 12

    # fused clip vs torch on the same gradients
    ref_grads = [p.grad.detach().clone() for p in model.parameters()]
    tn = acc.clip_grad_norm_(model.parameters(), 1.0)
    assert acc.last_clip_mode == "fused"
    ref_params = [torch.nn.Parameter(torch.empty_like(g_)) for g_ in ref_grads]
    for rp, rg in zip(ref_params, ref_grads):
        rp.grad = rg.clone()  # clip_grad_norm_ works in place: keep ref_grads untouched for the DDP comparison
    ref_tn = torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    assert abs(tn.item() - ref_tn.item()) <= 1e-5 * ref_tn.item(), (tn.item(), ref_tn.item())
    for p, rp in zip(model.parameters(), ref_params):
        assert torch.allclose(p.grad, rp.grad, rtol=1e-5, atol=1e-9)

    # stock torch DDP + NCCL, only when NCCL can run (one GPU per rank)
    if torch.cuda.device_count() >= world and PartialState().backend == "nccl":
        ddp = torch.nn.parallel.DistributedDataParallel(build(layers=layers).to(dev), device_ids=[dev.index])
        ddp(input_ids=ids, labels=ids).loss.backward()
        for (name, p), q in zip(model.module.named_parameters(), ddp.module.parameters()):
            unclipped = ref_grads.pop(0)
            if world == 2:
                assert torch.equal(unclipped, q.grad), f"{name}: not bit-identical to torch DDP+NCCL at N=2"
            assert torch.allclose(unclipped, q.grad, rtol=1e-3, atol=1e-5), name
        print(f"compared against torch DDP+NCCL on rank {rank}", flush=True)
    PartialState().runtime.ctx.check_status()
    torch.cuda.synchronize()
    acc.wait_for_everyone()
    print(f"GPT2_CHECK_OK rank {rank}/{world} buckets {nb}", flush=True)
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
