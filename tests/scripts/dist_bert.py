"""BASELINE.json config 4 shape: BERT-large-shaped sequence classifier (hidden 1024 / 16 heads / 4096 FFN, fewer layers
to keep the test short), bf16 autocast, gradient_accumulation_steps=4 through accelerator.accumulate, and
accelerator.clip_grad_norm_ on the sync step only.  Checks: no communication on the 3 non-sync micro-steps, the
reduced gradient equals the oracle's mean of the ranks' ACCUMULATED local gradients (bit-exact for two-shot), the
fused clip matches torch.nn.utils.clip_grad_norm_, module buffers (BERT's position_ids / token_type_ids) are
re-broadcast from rank 0, all ranks end with identical parameters."""
import argparse
import os
import sys

os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ddp_oracle as oracle  # noqa: E402


def build(layers, hidden, heads, ffn, seed=42):
    from transformers import BertConfig, BertForSequenceClassification

    torch.manual_seed(seed)
    cfg = BertConfig(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads, intermediate_size=ffn,
                     vocab_size=30522, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                     attn_implementation="eager")  # fmt: skip
    return BertForSequenceClassification(cfg)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    args = ap.parse_args()
    cpu = args.device == "cpu"
    if cpu:
        os.environ["ACCELERATE_USE_CPU"] = "yes"
        torch.set_num_threads(2)
    from accelerate_b200 import Accelerator, B200DataParallel
    from accelerate_b200.state import PartialState

    shape = (2, 64, 4, 128) if cpu else (3, 1024, 16, 4096)
    seq = 32 if cpu else 128
    torch.use_deterministic_algorithms(True, warn_only=True)
    acc = Accelerator(cpu=cpu, mixed_precision="bf16", gradient_accumulation_steps=4)
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    model = build(*shape)
    nbuf = len(list(model.buffers()))
    assert nbuf >= 1, "BERT registers position_ids / token_type_ids buffers"
    if rank != 0:  # perturb a buffer on non-zero ranks: the wrapper must restore rank 0's at wrap time
        next(model.buffers()).add_(1)
    model = acc.prepare(model)
    assert isinstance(model, B200DataParallel)
    everyone = acc.gather(next(model.module.buffers()).reshape(1, -1).clone()).cpu()
    assert all(torch.equal(everyone[0], everyone[r]) for r in range(world)), "buffers not broadcast from rank 0"
    twin = build(*shape).to(dev)
    opt = acc.prepare(torch.optim.SGD(model.parameters(), lr=0.1))

    g = torch.Generator().manual_seed(77 + rank)
    for micro in range(4):
        ids = torch.randint(0, 30522, (2, seq), generator=g).to(dev)
        labels = torch.randint(0, 2, (2,), generator=g).to(dev)
        with acc.autocast():
            (twin(input_ids=ids, labels=labels).loss / 4).backward()
        with acc.accumulate(model):
            acc.backward(model(input_ids=ids, labels=labels).loss)
            if acc.sync_gradients:
                before_clip = [p.grad.detach().clone() for p in model.parameters()]
                total = acc.clip_grad_norm_(model.parameters(), 0.25)
            opt.step()  # gated: only steps on the sync micro-step
        assert acc.sync_gradients == (micro == 3)
        assert model.reducer.num_reductions == (1 if micro == 3 else 0), "non-sync micro-steps must not communicate"
    # gradients before clipping == oracle mean of the ranks' accumulated local gradients
    exact = model.reducer.effective_algo() in ("twoshot", "gloo") and (not cpu or world == 2)
    for (name, p), tp, got in zip(model.module.named_parameters(), twin.parameters(), before_clip):
        assert p.grad is None or p.grad.dtype == torch.float32
        locals_all = acc.gather(tp.grad.reshape(1, -1)).cpu().numpy()
        want = oracle.ddp_mean_reduce([locals_all[r] for r in range(world)]).reshape(got.shape)
        have = got.cpu().numpy()
        if exact:
            assert have.tobytes() == want.tobytes(), f"{name}: differs from the oracle (max {np.abs(have - want).max()})"
        else:
            assert np.allclose(have, want, rtol=1e-6, atol=1e-8), name
    # fused clip == torch's clip on the same gradients
    ref = [torch.nn.Parameter(torch.empty_like(x)) for x in before_clip]
    for rp, x in zip(ref, before_clip):
        rp.grad = x.clone()
    ref_total = torch.nn.utils.clip_grad_norm_(ref, 0.25)
    assert abs(total.item() - ref_total.item()) <= 1e-5 * ref_total.item(), (total.item(), ref_total.item())
    assert acc.last_clip_mode == ("torch" if cpu else "fused")
    # identical parameters everywhere after the (single) optimizer step
    digest = torch.stack([p.detach().double().sum() for p in model.parameters()]).reshape(1, -1)
    allp = acc.gather(digest).cpu()
    assert all(torch.equal(allp[0], allp[r]) for r in range(world)), "ranks diverged after the step"
    if not cpu:
        PartialState().runtime.ctx.check_status()
        torch.cuda.synchronize()
    acc.wait_for_everyone()
    print(f"BERT_CHECK_OK rank {rank} buckets {len(model.reducer.buckets)} buffers {nbuf}", flush=True)
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
