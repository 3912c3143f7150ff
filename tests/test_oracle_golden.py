"""Pin the oracle: every function against the golden outputs the REFERENCE produced
(tests/golden/make_golden.py ran huggingface/accelerate on gloo at 2 and 4 ranks), and against
torch's own routines for the pieces no reference test pins (clip / unscale)."""

import math
import os

import numpy as np
import pytest
import torch

from oracle import ddp_oracle as oracle
from scenarios import ops_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_ddp_mean_reduce_matches_reference_ddp(world, tag):
    z = np.load(os.path.join(GOLD, f"ddp_w{world}_{tag}.npz"))
    nparams = len([k for k in z.files if k.startswith("r0/ddp/")])
    assert nparams == 4
    for i in range(nparams):
        locals_ = [z[f"r{r}/local/{i}"] for r in range(world)]
        want = z[f"r0/ddp/{i}"]
        got = oracle.ddp_mean_reduce(locals_)
        assert got.dtype == np.float32 and want.dtype == np.float32  # grads are fp32 even under bf16 autocast
        if world == 2:
            assert bits(got, want), f"N=2 must be bit-identical to reference DDP (param {i})"
        assert np.allclose(got, want, rtol=1e-3, atol=1e-5)
        for r in range(1, world):  # reference: identical on every rank
            assert bits(z[f"r{r}/ddp/{i}"], want)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_clip_matches_reference(world, tag):
    z = np.load(os.path.join(GOLD, f"ddp_w{world}_{tag}.npz"))
    grads = [z[f"r0/ddp/{i}"] for i in range(4)]
    clipped, tn = oracle.clip_grad_norm(grads, 0.05)
    assert abs(tn - z["r0/clip/total_norm"]) <= 1e-6 * tn
    for i in range(4):
        assert np.allclose(clipped[i], z[f"r0/clip/{i}"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("world", [2, 4])
def test_accumulate_then_reduce_matches_reference(world):
    z = np.load(os.path.join(GOLD, f"ddp_w{world}_fp32.npz"))
    for i in range(4):
        # the twin model accumulated (loss/2) over both micro-steps -> its grad is the local accumulated grad
        locals_ = [z[f"r{r}/accum/local/{i}"] for r in range(world)]
        got = oracle.ddp_mean_reduce(locals_)
        assert np.allclose(got, z[f"r0/accum/ddp/{i}"], rtol=1e-3, atol=1e-5)
        # micro-step 0 ran under no_sync: ranks hold DIFFERENT (local) gradients
        assert not bits(z[f"r0/accum/micro0/{i}"], z[f"r1/accum/micro0/{i}"])
    assert int(z["r0/accum/sync_flag/0"]) == 0 and int(z["r0/accum/sync_flag/1"]) == 1


@pytest.mark.parametrize("world", [2, 4])
def test_collectives_match_reference(world):
    z = np.load(os.path.join(GOLD, f"ops_w{world}.npz"))
    per_rank = [ops_inputs(r, world) for r in range(world)]
    for name in per_rank[0]:
        ins = [per_rank[r][name].numpy() for r in range(world)]
        for r in range(world):
            assert bits(oracle.gather(ins), z[f"r{r}/gather/{name}"]), name
            assert bits(oracle.broadcast(ins, src=world - 1)[r], z[f"r{r}/broadcast/{name}"]), name
            if ins[0].dtype != np.bool_:
                got = oracle.reduce(ins, "sum")
                ref = z[f"r{r}/reduce_sum/{name}"]
                assert bits(got, ref) if (world == 2 or ins[0].dtype.kind == "i") else np.allclose(got, ref, 1e-6, 1e-6), name
                if ins[0].dtype.kind == "f":
                    got = oracle.reduce(ins, "mean")
                    ref = z[f"r{r}/reduce_mean/{name}"]
                    assert bits(got, ref) if world == 2 else np.allclose(got, ref, 1e-6, 1e-6), name
    ragged = [np.arange((r + 2) * 3, dtype=np.float32).reshape(r + 2, 3) + 100 * r for r in range(world)]
    ragged1 = [np.arange(2 * (r + 1), dtype=np.int64).reshape(2, r + 1) + 10 * r for r in range(world)]
    for r in range(world):
        assert bits(oracle.pad_across_processes(ragged, 0)[r], z[f"r{r}/pad/dim0"])
        assert bits(oracle.pad_across_processes(ragged, 0, -1, True)[r], z[f"r{r}/pad/dim0_first"])
        assert bits(oracle.pad_across_processes(ragged1, 1, 7)[r], z[f"r{r}/pad/dim1"])
        assert bits(oracle.pad_across_processes(ragged1, -1, 7, True)[r], z[f"r{r}/pad/dimneg"])


def test_clip_and_unscale_match_torch_routines():
    g = torch.Generator().manual_seed(0)
    grads = [torch.randn(s, generator=g) * 10 for s in (1, 17, 1000, 33 * 7)]
    for max_norm, norm_type in [(1.0, 2.0), (1e6, 2.0), (0.3, math.inf)]:
        params = [torch.nn.Parameter(t.clone()) for t in grads]
        for p, t in zip(params, grads):
            p.grad = t.clone()
        tn = torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type=norm_type)
        got, otn = oracle.clip_grad_norm([t.numpy() for t in grads], max_norm, norm_type)
        assert abs(otn - tn.item()) <= 1e-6 * tn.item()
        for p, o in zip(params, got):
            assert np.allclose(p.grad.numpy(), o, rtol=1e-6, atol=0)
    # unscale: torch's fused routine on CPU
    scale = torch.tensor(65536.0)
    ts = [t.clone() * 65536 for t in grads]
    ts[2][5] = float("inf")
    found = torch.zeros(())
    inv = scale.double().reciprocal().float()
    ref = [t.clone() for t in ts]
    torch._amp_foreach_non_finite_check_and_unscale_(ref, found, inv)
    got, ofound = oracle.unscale([t.numpy() for t in ts], 65536.0)
    assert ofound == found.item() == 1.0
    for r, o in zip(ref, got):
        assert bits(r.numpy(), o)
    assert oracle.inv_scale_of(3.0) == np.float32(1.0 / 3.0)


def test_bucket_assignment_matches_torch():
    import torch.distributed as dist

    g = torch.Generator().manual_seed(1)
    counts = torch.randint(1, 3_000_000, (200,), generator=g).tolist()
    tensors = [torch.empty(n, device="meta") for n in counts]
    sizes = [t.numel() * 4 for t in tensors]
    for first, cap in [(1 << 20, 25 << 20), (15 << 20, 15 << 20), (1, 1)]:
        want, _ = dist._compute_bucket_assignment_by_size(tensors, [first, cap])
        assert oracle.bucket_assignment(sizes, first, cap) == want


@pytest.mark.parametrize("cls,decoupled", [(torch.optim.AdamW, True), (torch.optim.Adam, False)])
def test_adam_step_matches_torch_foreach(cls, decoupled):
    """The oracle's optimizer restatement against torch's own foreach Adam/AdamW on CPU, several steps, changing lr."""
    g = torch.Generator().manual_seed(2)
    for wd, betas in [(0.01, (0.9, 0.999)), (0.0, (0.9, 0.999)), (0.1, (0.4, 0.95))]:
        p = torch.nn.Parameter(torch.randn(257, generator=g))
        opt = cls([p], lr=1e-2, weight_decay=wd, betas=betas, foreach=True)
        op, om, ov = p.detach().numpy().copy(), np.zeros(257, np.float32), np.zeros(257, np.float32)
        for step in range(1, 6):
            lr = 1e-2 * 0.7**step
            opt.param_groups[0]["lr"] = lr
            grad = torch.randn(257, generator=g)
            p.grad = grad.clone()
            opt.step()
            op, om, ov = oracle.adam_step(op, grad.numpy(), om, ov, step, lr, betas[0], betas[1], 1e-8, wd, decoupled)
            assert np.allclose(op, p.detach().numpy(), rtol=1e-5, atol=1e-6), (cls, wd, step)
            st = opt.state[p]
            assert np.allclose(om, st["exp_avg"].numpy(), rtol=1e-5, atol=1e-6)
            assert np.allclose(ov, st["exp_avg_sq"].numpy(), rtol=1e-5, atol=1e-8)


def test_c_adam_oracle_agrees_with_the_numpy_restatement():
    """oracle.c (true single-rounding fmaf) vs ddp_oracle.adam_step (fma emulated through float64): the same values,
    bit for bit except where the emulation's double rounding shows (rare), for Adam and AdamW."""
    from oracle import c_oracle

    rng = np.random.default_rng(5)
    n = 50_000
    p, g, m = (rng.standard_normal(n).astype(np.float32) * s_ for s_ in (0.5, 1e-2, 1e-2))
    v = rng.random(n).astype(np.float32) * 1e-4 + 1e-8
    for decoupled in (True, False):
        c = c_oracle.adam_step(p, g, m, v, 7, 3e-4, 0.9, 0.999, 1e-8, 0.1, decoupled=decoupled, mult=0.625)
        py = oracle.adam_step(p, g, m, v, 7, 3e-4, 0.9, 0.999, 1e-8, 0.1, decoupled=decoupled, mult=0.625)
        for a, b in zip(c, py):
            np.testing.assert_allclose(a, b, rtol=3e-7, atol=0)
            assert (a != b).mean() < 1e-3
