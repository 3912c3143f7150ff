"""-m gpu: the public API on one GPU — clip paths, gradients-as-views life cycle, fp16 loss scaling with
the fused unscale/clip (reference pins: tests/test_optimizer.py:40-82 overflow => skipped step), unused
parameters.  World size 1 exercises the full reducer (pack -> fused allreduce+stats -> clip) when the
wrapper is constructed by hand."""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def mlp(seed=0, dtype=torch.float32):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(64, 257), torch.nn.GELU(), torch.nn.Linear(257, 33)).cuda().to(dtype)


def assert_tracks_torch(ours, theirs, lr, what, rtol=1e-5, atol=1e-7):
    """Parameters after a fused Adam step vs torch's foreach step.  Bit-equal is the normal outcome; otherwise the
    two may differ by float32 rounding — except on the rare elements whose effective gradient cancels to ~eps, where
    Adam's update g/(|g|+eps) amplifies a last-bit difference (bounded by one full step of size ~lr)."""
    if torch.equal(ours, theirs):
        return True
    diff = (ours - theirs).abs()
    bad = diff > atol + rtol * theirs.abs()
    assert bad.float().mean().item() <= 1e-4 and diff.max().item() <= 2.5 * lr, (what, int(bad.sum()), diff.max().item())
    return False


def torch_clip_reference(grads, max_norm, norm_type=2.0):
    params = [torch.nn.Parameter(torch.empty_like(g)) for g in grads]
    for p, g in zip(params, grads):
        p.grad = g.clone()
    tn = torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type=norm_type)
    return tn, [p.grad for p in params]


@pytest.mark.parametrize("norm_type", [2.0, math.inf])
def test_multi_tensor_clip_matches_torch(norm_type):
    from accelerate_b200 import Accelerator

    acc = Accelerator()
    model = acc.prepare(mlp())
    x = torch.randn(16, 64, device="cuda")
    acc.backward(model(x).square().mean() * 100)
    grads = [p.grad.clone() for p in model.parameters()]
    ref_tn, ref = torch_clip_reference(grads, 0.5, norm_type)
    tn = acc.clip_grad_norm_(model.parameters(), 0.5, norm_type)
    assert acc.last_clip_mode == "multi_tensor"
    assert torch.allclose(tn, ref_tn, rtol=1e-5)
    for p, r in zip(model.parameters(), ref):
        assert torch.allclose(p.grad, r, rtol=1e-5, atol=1e-9)
    # subset of parameters + no clipping needed (coef clamps to 1: grads untouched, bit for bit)
    sub = list(model.parameters())[:2]
    before = [p.grad.clone() for p in sub]
    tn2 = acc.clip_grad_norm_(sub, 1e9)
    assert all(torch.equal(p.grad, b) for p, b in zip(sub, before))
    assert torch.allclose(tn2, torch_clip_reference(before, 1e9)[0], rtol=1e-5)


def test_gradients_become_views_and_are_repacked():
    from accelerate_b200 import Accelerator, B200DataParallel
    from accelerate_b200.state import PartialState

    acc = Accelerator()
    model = mlp()
    ddp = B200DataParallel(model, device_ids=[0], bucket_cap_mb=0.02)  # several buckets
    assert len(ddp.reducer.buckets) >= 2 and ddp.bucket_bytes_cap == int(0.02 * 2**20)
    opt = acc.prepare(torch.optim.SGD(model.parameters(), lr=0.1))
    rt = PartialState().runtime
    x = torch.randn(8, 64, device="cuda")
    flat = next(iter(ddp.reducer.groups.values())).flat

    def step(set_to_none):
        ddp(x).square().mean().backward()
        for b in ddp.reducer.buckets:
            for p, ptr in zip(b.params, b.ptrs):
                assert p.grad.data_ptr() == ptr, "grad must live in the fusion buffer after a sync backward"
        ref_tn, ref = torch_clip_reference([p.grad for p in model.parameters()], 0.01)
        tn = acc.clip_grad_norm_(ddp.parameters(), 0.01)
        assert acc.last_clip_mode == "fused"
        assert torch.allclose(tn, ref_tn, rtol=1e-5)
        for p, r in zip(model.parameters(), ref):
            assert torch.allclose(p.grad, r, rtol=1e-5, atol=1e-10)
        opt.step()
        opt.zero_grad(set_to_none=set_to_none)

    step(True)
    assert all(p.grad is None for p in model.parameters())
    launches = rt.ctx.launches
    step(False)  # fresh grads -> one pack per bucket + one allreduce per bucket + finalize + scale
    nb = len(ddp.reducer.buckets)
    assert rt.ctx.launches - launches >= 2 * nb + 3
    assert all(p.grad is not None and p.grad.abs().sum() == 0 for p in model.parameters())
    launches = rt.ctx.launches
    step(False)  # grads are already views: autograd accumulates in place, NO pack launches
    assert rt.ctx.launches - launches == nb + 3, "nb allreduces + fence + finalize + scale; NO pack when grads already live in the fusion buffer"
    assert ddp.reducer.num_reductions == 3
    # second backward without a forward in between must not reduce again (expect_autograd_hooks semantics)
    out = ddp(x).square().mean()
    out.backward(retain_graph=True)
    n = ddp.reducer.num_reductions
    out.backward()
    assert ddp.reducer.num_reductions == n
    # clip_grad_value_ on the fusion buffer
    acc.clip_grad_value_(ddp.parameters(), 1e-4)
    assert all(p.grad.abs().max() <= 1e-4 for p in model.parameters())
    assert flat.abs().max() <= 1e-4
    rt.ctx.check_status()


def test_unused_parameters():
    from accelerate_b200 import B200DataParallel
    from accelerate_b200.state import PartialState

    PartialState()

    class TwoHeads(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(8, 8)
            self.b = torch.nn.Linear(8, 8)

        def forward(self, x, use_b):
            return self.b(x) if use_b else self.a(x)

    x = torch.randn(4, 8, device="cuda")
    strict = B200DataParallel(TwoHeads().cuda(), device_ids=[0])
    with pytest.raises(RuntimeError, match="did not receive a gradient"):
        strict(x, False).sum().backward()
    strict.reducer.remove_hooks()
    lenient = B200DataParallel(TwoHeads().cuda(), device_ids=[0], find_unused_parameters=True)
    lenient(x, False).sum().backward()
    assert lenient.module.b.weight.grad is None, "globally unused: grad stays undefined, as with torch DDP"
    assert lenient.module.a.weight.grad.abs().sum() > 0
    with lenient.no_sync():
        lenient(x, True).sum().backward()
    kept = lenient.module.b.weight.grad.clone()
    lenient(x, False).sum().backward()  # b unused in the sync backward: its accumulated gradient is reduced as it is
    assert torch.equal(lenient.module.b.weight.grad, kept) and kept.abs().sum() > 0


def test_fp16_scaler_fused_unscale_clip_and_overflow_skip():
    from accelerate_b200 import Accelerator, B200DataParallel, GradScalerKwargs

    acc = Accelerator(mixed_precision="fp16", kwargs_handlers=[GradScalerKwargs(init_scale=1024.0, growth_interval=1)])
    model = mlp()
    ddp = B200DataParallel(model, device_ids=[0])
    opt = acc.prepare(torch.optim.SGD(model.parameters(), lr=0.01))
    x = torch.randn(8, 64, device="cuda")
    with acc.autocast():
        loss = ddp(x).float().square().mean()
    acc.backward(loss)
    scaled = [p.grad.clone() for p in model.parameters()]
    ref_tn, ref = torch_clip_reference([g / 1024.0 for g in scaled], 0.05)
    tn = acc.clip_grad_norm_(ddp.parameters(), 0.05)  # ONE pass: unscale and clip
    assert acc.last_clip_mode == "fused"
    assert torch.allclose(tn, ref_tn, rtol=1e-5)
    for p, r in zip(model.parameters(), ref):
        assert torch.allclose(p.grad, r, rtol=1e-5, atol=1e-10)
    before = [p.detach().clone() for p in model.parameters()]
    opt.step()
    assert not opt.step_was_skipped and acc.scaler.get_scale() == 2048.0  # growth_interval=1
    assert any(not torch.equal(b, p) for b, p in zip(before, model.parameters()))
    opt.zero_grad()

    # overflow: a non-finite gradient on the wire -> found_inf from the allreduce statistics -> step skipped
    with acc.autocast():
        loss = ddp(x).float().square().mean() * float("inf")
    acc.backward(loss)
    before = [p.detach().clone() for p in model.parameters()]
    opt.step()  # scaler.step -> unscale_ (fused override) -> found_inf -> no optimizer.step
    assert opt.step_was_skipped and acc.optimizer_step_was_skipped
    assert all(torch.equal(b, p) for b, p in zip(before, model.parameters()))
    assert acc.scaler.get_scale() == 1024.0  # backoff 0.5
    opt.zero_grad()

    # explicit unscale_gradients() (no clip): one multiply pass by 1/scale, found_inf from the statistics.
    # (Calling clip_grad_norm_ afterwards would raise in the reference too: GradScaler.unscale_ may run
    # only once per step, accelerator.py:2944/3007.)
    with acc.autocast():
        loss = ddp(x).float().square().mean()
    acc.backward(loss)
    scaled = [p.grad.clone() for p in model.parameters()]
    acc.unscale_gradients()
    for p, g in zip(model.parameters(), scaled):
        assert torch.equal(p.grad, g * (1.0 / 1024.0))
    with pytest.raises(RuntimeError, match="already been called"):
        acc.clip_grad_norm_(ddp.parameters(), 0.05)
    opt.step()
    assert not opt.step_was_skipped


@pytest.mark.parametrize("cls", [torch.optim.AdamW, torch.optim.Adam])
@pytest.mark.parametrize("defer", ["0", "1"])
def test_fused_adam_consumption_matches_torch(cls, defer, monkeypatch):
    """SURVEY.md §8f row 1 (B200DP_FUSED_ADAM, on by default): one multi-tensor kernel instead of torch's foreach Adam/AdamW,
    optionally consuming the deferred clip multiplier; parameters and optimizer state must track torch's own
    optimizer (fp32 rounding) over several steps with a scheduler and two parameter groups."""
    from accelerate_b200 import Accelerator, B200DataParallel
    from accelerate_b200.state import PartialState

    monkeypatch.setenv("B200DP_DEFER_CLIP", defer)  # fused Adam itself is on by default
    acc = Accelerator()
    model, ref = mlp(3), mlp(3)
    ddp = B200DataParallel(model, device_ids=[0], bucket_cap_mb=0.02)

    def groups(m):
        decay = [p for p in m.parameters() if p.ndim > 1]
        rest = [p for p in m.parameters() if p.ndim <= 1]
        return [{"params": decay, "weight_decay": 0.1}, {"params": rest, "weight_decay": 0.0}]

    opt = cls(groups(model), lr=3e-3, betas=(0.9, 0.95))
    ref_opt = cls(groups(ref), lr=3e-3, betas=(0.9, 0.95), foreach=True)
    sched = torch.optim.lr_scheduler.StepLR(opt, 1, gamma=0.8)
    ref_sched = torch.optim.lr_scheduler.StepLR(ref_opt, 1, gamma=0.8)
    opt, sched = acc.prepare(opt, sched)
    rt = PartialState().runtime
    g = torch.Generator(device="cuda").manual_seed(0)
    bits = []
    for step in range(5):
        x = torch.randn(8, 64, device="cuda", generator=g)
        ddp(x).square().mean().backward()
        ref(x).square().mean().backward()
        tn = acc.clip_grad_norm_(ddp.parameters(), 0.05)
        ref_tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        assert torch.allclose(tn, ref_tn, rtol=1e-5)
        launches = rt.ctx.launches
        opt.step()
        assert opt._fused_plan is not None and rt.ctx.launches - launches == 2, "one fused launch per parameter group"
        ref_opt.step()
        sched.step()
        ref_sched.step()
        opt.zero_grad(set_to_none=False)  # grads stay views of the fusion buffer
        ref_opt.zero_grad()
        for p, q in zip(model.parameters(), ref.parameters()):
            bits.append(assert_tracks_torch(p, q, 3e-3, step))
    for p, q in zip(model.parameters(), ref.parameters()):
        a, b = opt.optimizer.state[p], ref_opt.state[q]
        assert float(a["step"]) == float(b["step"]) == 5.0
        assert torch.allclose(a["exp_avg"], b["exp_avg"], rtol=1e-5, atol=1e-8)
        assert torch.allclose(a["exp_avg_sq"], b["exp_avg_sq"], rtol=1e-5, atol=1e-10)
    # the optimizer is still the user's: state_dict round-trips (load_state_dict replaces the state tensors; the plan
    # re-adopts them), and an ineligible step falls back to torch's own
    sd = opt.state_dict()
    assert len(sd["state"]) == 4 and sd["param_groups"][0]["lr"] == ref_opt.param_groups[0]["lr"]
    import copy

    opt.load_state_dict(copy.deepcopy(sd))
    ref_opt.load_state_dict(copy.deepcopy(ref_opt.state_dict()))
    ddp(x).square().mean().backward()
    ref(x).square().mean().backward()
    opt.step()
    ref_opt.step()
    opt.zero_grad(set_to_none=False)
    ref_opt.zero_grad()
    for p, q in zip(model.parameters(), ref.parameters()):
        assert_tracks_torch(p, q, 3e-3, "after load_state_dict")
        assert float(opt.optimizer.state[p]["step"]) == 6.0
    opt.zero_grad(set_to_none=True)
    ddp(x).square().mean().backward()
    for p in model.parameters():
        p.grad = p.grad.clone()  # no longer views -> plan declines
    before = rt.ctx.launches
    opt.step()
    assert rt.ctx.launches == before, "fallback must not launch the fused kernel"
    rt.ctx.check_status()


@pytest.mark.parametrize("cls", [torch.optim.AdamW, torch.optim.Adam])
def test_fused_adam_loose_gradients_single_process(cls):
    """The same fused step without a reducer (one process: the reference does not wrap the model either): gradient
    pointers are collected per step from ``param.grad``; a step with a missing gradient falls back to torch."""
    from accelerate_b200 import Accelerator
    from accelerate_b200.state import PartialState

    acc = Accelerator()
    model, ref = mlp(5), mlp(5)
    opt = cls(model.parameters(), lr=2e-3, weight_decay=0.05)
    ref_opt = cls(ref.parameters(), lr=2e-3, weight_decay=0.05, foreach=True)
    model, opt = acc.prepare(model, opt)
    assert not hasattr(model, "reducer"), "world size 1 leaves the model unwrapped"
    rt = PartialState().runtime
    g = torch.Generator(device="cuda").manual_seed(1)
    bits = []
    for step in range(4):
        x = torch.randn(8, 64, device="cuda", generator=g)
        acc.backward(model(x).square().mean())
        ref(x).square().mean().backward()
        tn = acc.clip_grad_norm_(model.parameters(), 0.05)
        ref_tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        assert torch.allclose(tn, ref_tn, rtol=1e-5)
        before = rt.ctx.launches
        opt.step()
        assert opt._fused_plan is not None and opt._fused_plan.reducer is None
        assert rt.ctx.launches - before == 1, "one fused launch for the single parameter group"
        ref_opt.step()
        opt.zero_grad()  # set_to_none=True: next step's gradients are fresh tensors
        ref_opt.zero_grad()
        for p, q in zip(model.parameters(), ref.parameters()):
            bits.append(assert_tracks_torch(p, q, 2e-3, step))
    for p, q in zip(model.parameters(), ref.parameters()):
        a, b = opt.optimizer.state[p], ref_opt.state[q]
        assert float(a["step"]) == float(b["step"]) == 4.0
        assert torch.allclose(a["exp_avg"], b["exp_avg"], rtol=1e-5, atol=1e-8)
        assert torch.allclose(a["exp_avg_sq"], b["exp_avg_sq"], rtol=1e-5, atol=1e-10)
    # a parameter without a gradient: torch skips it, so must we -> the plan declines and torch's step runs
    acc.backward(model(x).square().mean())
    ref(x).square().mean().backward()
    first, ref_first = next(iter(model.parameters())), next(iter(ref.parameters()))
    first.grad = None
    ref_first.grad = None
    before = rt.ctx.launches
    opt.step()
    ref_opt.step()
    assert rt.ctx.launches == before
    for p, q in zip(model.parameters(), ref.parameters()):
        assert_tracks_torch(p, q, 2e-3, "fallback step")
    rt.ctx.check_status()


def test_deferred_clip_is_invisible_and_consumed_by_the_fused_step(monkeypatch):
    """VERDICT r1 #6 / SURVEY.md §8f row 1: by default ``clip_grad_norm_`` leaves the coefficient pending
    (accelerate_b200/lazy_grad.py).  (a) The fused AdamW step consumes it: no scale pass is launched and the update
    equals torch's clip + AdamW.  (b) Anything that looks at ``p.grad`` first sees the reference's clipped values
    (the pass is triggered by the access) and the following step is still right.  (c) A non-fused optimizer gets
    clipped gradients too."""
    from accelerate_b200 import Accelerator, B200DataParallel, lazy_grad
    from accelerate_b200.state import PartialState

    monkeypatch.delenv("B200DP_DEFER_CLIP", raising=False)
    acc = Accelerator()
    model, ref = mlp(7), mlp(7)
    ddp = B200DataParallel(model, device_ids=[0], bucket_cap_mb=0.02)
    opt = acc.prepare(torch.optim.AdamW(model.parameters(), lr=2e-3))
    ref_opt = torch.optim.AdamW(ref.parameters(), lr=2e-3, foreach=True)
    rt = PartialState().runtime
    g = torch.Generator(device="cuda").manual_seed(3)
    for step in range(4):
        x = torch.randn(8, 64, device="cuda", generator=g)
        ddp(x).square().mean().backward()
        ref(x).square().mean().backward()
        unclipped = [p.grad.clone() for p in model.parameters()]  # nothing is pending before the clip: plain reads
        tn = acc.clip_grad_norm_(ddp.parameters(), 0.05)
        ref_tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        assert acc.last_clip_mode == "fused" and torch.allclose(tn, ref_tn, rtol=1e-4)
        assert ddp.reducer._pending_mult is not None, "the multiply pass must not have run yet"
        launches = rt.ctx.launches
        if step % 2 == 1:  # (b) observe the gradients before the step: torch's clip of the very same values
            _, want = torch_clip_reference(unclipped, 0.05)
            for p, w in zip(model.parameters(), want):
                assert torch.allclose(p.grad, w, rtol=1e-5, atol=1e-10)
            assert ddp.reducer._pending_mult.done and rt.ctx.launches - launches == 1, "exactly one multiply pass"
            launches = rt.ctx.launches
        opt.step()
        assert ddp.reducer._pending_mult is None
        assert rt.ctx.launches - launches == 1, "only the fused Adam launch: the clip multiplier rides along"
        ref_opt.step()
        opt.zero_grad(set_to_none=step < 2)
        ref_opt.zero_grad()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert_tracks_torch(p, q, 2e-3, step)
    # gradients kept across steps (no zero_grad): the pending coefficient must be in them before the next backward adds
    x = torch.randn(8, 64, device="cuda", generator=g)
    for m in (ddp, ref):
        m(x).square().mean().backward()
    acc.clip_grad_norm_(ddp.parameters(), 0.05)
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
    for m in (ddp, ref):  # second backward accumulates on top of the clipped gradients
        m(x).square().mean().backward()
    for p, q in zip(model.parameters(), ref.parameters()):
        # (the two models differ by fp32 rounding of the steps above; a coefficient that was NOT applied before the
        # accumulation would show up as a factor-of-many difference, not as 1e-3)
        assert torch.allclose(p.grad, q.grad, rtol=1e-3, atol=1e-7)
    # (c) a plain SGD step reads p.grad from Python -> clipped values
    sgd = acc.prepare(torch.optim.SGD(model.parameters(), lr=0.1))
    ref_sgd = torch.optim.SGD(ref.parameters(), lr=0.1)
    opt.zero_grad()
    ref_opt.zero_grad()
    for m in (ddp, ref):
        m(x).square().mean().backward()
    acc.clip_grad_norm_(ddp.parameters(), 0.05)
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
    sgd.step()
    ref_sgd.step()
    for p, q in zip(model.parameters(), ref.parameters()):
        assert_tracks_torch(p, q, 2e-3, "sgd after deferred clip")
    rt.ctx.check_status()


def test_deferred_clip_loose_gradients_single_process(monkeypatch):
    """The same laziness without a reducer (world size 1): multi_stats + finalize, the multi_scale pass stays pending and
    the loose-layout fused AdamW consumes it."""
    from accelerate_b200 import Accelerator, lazy_grad
    from accelerate_b200.state import PartialState

    monkeypatch.delenv("B200DP_DEFER_CLIP", raising=False)
    acc = Accelerator()
    model, ref = mlp(9), mlp(9)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3)
    ref_opt = torch.optim.AdamW(ref.parameters(), lr=2e-3, foreach=True)
    model, opt = acc.prepare(model, opt)
    rt = PartialState().runtime
    g = torch.Generator(device="cuda").manual_seed(4)
    for step in range(3):
        x = torch.randn(8, 64, device="cuda", generator=g)
        acc.backward(model(x).square().mean())
        ref(x).square().mean().backward()
        tn = acc.clip_grad_norm_(model.parameters(), 0.05)
        ref_tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        assert acc.last_clip_mode == "multi_tensor" and torch.allclose(tn, ref_tn, rtol=1e-4)
        assert lazy_grad.has_pending()
        launches = rt.ctx.launches
        if step == 1:
            for p, q in zip(model.parameters(), ref.parameters()):
                assert torch.allclose(p.grad, q.grad, rtol=1e-3, atol=1e-8)  # the two models differ by fp32 rounding of step 0
            assert not lazy_grad.has_pending()
            launches = rt.ctx.launches
        opt.step()
        assert not lazy_grad.has_pending() and rt.ctx.launches - launches == 1
        ref_opt.step()
        opt.zero_grad()
        ref_opt.zero_grad()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert_tracks_torch(p, q, 2e-3, step)


def test_stale_statistics_are_not_trusted():
    """the fused clip may use the allreduce statistics only while they describe the buffer.  A
    gradient edited through torch (``p.grad.mul_``), or a second backward accumulating into the views, bumps the
    fusion buffer's version counter: the clip must then recompute the norm from the live gradients."""
    from accelerate_b200 import Accelerator, B200DataParallel

    acc = Accelerator()
    model = mlp(11)
    ddp = B200DataParallel(model, device_ids=[0], bucket_cap_mb=0.02)
    x = torch.randn(8, 64, device="cuda")
    ddp(x).square().mean().backward()
    with torch.no_grad():
        for p in model.parameters():
            p.grad.mul_(3.0)
    ref_tn, ref = torch_clip_reference([p.grad for p in model.parameters()], 0.05)
    tn = acc.clip_grad_norm_(ddp.parameters(), 0.05)
    assert acc.last_clip_mode == "multi_tensor", "stale statistics must not be used"
    assert torch.allclose(tn, ref_tn, rtol=1e-5)
    for p, r in zip(model.parameters(), ref):
        assert torch.allclose(p.grad, r, rtol=1e-5, atol=1e-10)
    # a second backward without a forward accumulates into the views
    for p in model.parameters():
        p.grad = None
    out = ddp(x).square().mean()
    out.backward(retain_graph=True)
    out.backward()
    ref_tn, ref = torch_clip_reference([p.grad for p in model.parameters()], 0.05)
    tn = acc.clip_grad_norm_(ddp.parameters(), 0.05)
    assert acc.last_clip_mode == "multi_tensor" and torch.allclose(tn, ref_tn, rtol=1e-5)


def test_free_memory_releases_segments_and_slots():
    """``Accelerator.free_memory()`` + re-prepare (the find_executable_batch_size pattern) must
    not leak the gradient-sized symmetric buffer or the statistics slots."""
    from accelerate_b200 import Accelerator, B200DataParallel
    from accelerate_b200.state import PartialState

    acc = Accelerator()
    rt = PartialState().runtime
    x = torch.randn(8, 64, device="cuda")
    slots0 = rt._slot_next
    seg_ids = []
    for attempt in range(4):
        model = mlp(attempt)
        ddp = B200DataParallel(model, device_ids=[0], bucket_cap_mb=0.02)
        acc._wrappers.append(ddp)
        ddp(x).square().mean().backward()
        kept = [p.grad.clone() for p in model.parameters()]
        seg_ids.append(next(iter(ddp.reducer.groups.values())).segment.seg)
        assert ddp.reducer.slot_base == slots0, "slots of the closed reducer are reused"
        acc.free_memory()
        assert rt._slot_next == slots0
        for p, k in zip(model.parameters(), kept):  # gradients survive as ordinary tensors
            assert torch.equal(p.grad, k)
    assert len(set(seg_ids)) == 1, f"segment table slot must be reused, got {seg_ids}"
    rt.ctx.check_status()


def test_fp16_fused_adam_step_skips_on_device(monkeypatch):
    """VERDICT r1 #8 (reference optimizer.py:163-177): under a GradScaler the fused Adam kernel reads found-inf on the
    device — no host sync in ``step`` — and the skipped step is settled lazily.  Parameters, `step` counters, the
    loss scale and ``step_was_skipped`` follow torch's GradScaler + AdamW over a sequence with an overflow."""
    from accelerate_b200 import Accelerator, B200DataParallel, GradScalerKwargs, lazy_grad
    from accelerate_b200.state import PartialState

    monkeypatch.delenv("B200DP_DEFER_CLIP", raising=False)
    acc = Accelerator(mixed_precision="fp16", kwargs_handlers=[GradScalerKwargs(init_scale=256.0, growth_interval=2)])
    model, ref = mlp(13), mlp(13)
    ddp = B200DataParallel(model, device_ids=[0], bucket_cap_mb=0.02)
    opt = acc.prepare(torch.optim.AdamW(model.parameters(), lr=2e-3))
    ref_opt = torch.optim.AdamW(ref.parameters(), lr=2e-3, foreach=True)
    ref_scaler = torch.amp.GradScaler("cuda", init_scale=256.0, growth_interval=2)
    rt = PartialState().runtime
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(6):
        x = torch.randn(8, 64, device="cuda", generator=g)
        boom = float("inf") if step == 2 else 1.0
        with acc.autocast():
            loss = ddp(x).float().square().mean() * boom
        acc.backward(loss)
        with torch.autocast("cuda", dtype=torch.float16):
            ref_loss = ref(x).float().square().mean() * boom
        ref_scaler.scale(ref_loss).backward()
        if step % 2 == 0:  # clip on some steps (unscale+clip fused), plain step on the others
            tn = acc.clip_grad_norm_(ddp.parameters(), 0.05)
            ref_scaler.unscale_(ref_opt)
            ref_tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
            if step != 2:
                assert torch.allclose(tn, ref_tn, rtol=1e-4)
        launches = rt.ctx.launches
        opt.step()
        assert opt._fused_plan is not None, "AdamW under a GradScaler must take the fused step"
        assert ddp.reducer._pending_mult is None and not lazy_grad.has_pending(), "1/scale (and the clip) ride along with the step"
        assert rt.ctx.launches - launches == (2 if step % 2 else 1), "finalize (unscale) + ONE fused Adam launch; no multiply pass"
        ref_scaler.step(ref_opt)
        ref_scaler.update()
        assert opt.step_was_skipped == (step == 2)
        assert acc.scaler.get_scale() == ref_scaler.get_scale()
        opt.zero_grad()
        ref_opt.zero_grad()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert_tracks_torch(p, q, 2e-3, step, rtol=1e-4)
    for p, q in zip(model.parameters(), ref.parameters()):
        assert float(opt.optimizer.state[p]["step"]) == float(ref_opt.state[q]["step"]) == 5.0
    rt.ctx.check_status()
