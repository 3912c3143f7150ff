"""Host logic of the reducer without any communication (one process, CPU transport): bucket table vs torch's own
assignment, fusion-buffer layout, gradients-as-views life cycle, forward-time sync latch, unused parameters,
wrapper attributes the reference's tests read (tests/test_kwargs_handlers.py:173-202)."""

import pytest
import torch
import torch.distributed as dist

from accelerate_b200.reducer import B200DataParallel, compute_bucket_assignment


def test_bucket_assignment_matches_torch_mixed_dtypes():
    g = torch.Generator().manual_seed(3)
    counts = torch.randint(1, 2_000_000, (120,), generator=g).tolist()
    dtypes = [torch.float32 if i % 5 else torch.bfloat16 for i in range(len(counts))]
    tensors = [torch.empty(n, dtype=dt, device="meta") for n, dt in zip(counts, dtypes)]
    sizes = [t.numel() * t.element_size() for t in tensors]
    for first, cap in [(1 << 20, 25 << 20), (15 << 20, 15 << 20), (1 << 10, 1 << 12)]:
        want, _ = dist._compute_bucket_assignment_by_size(tensors, [first, cap])
        assert compute_bucket_assignment(sizes, dtypes, first, cap) == want


def small_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))


def test_wrapper_attributes_and_layout():
    ddp = B200DataParallel(small_model(), bucket_cap_mb=15, find_unused_parameters=True, broadcast_buffers=False,
                           gradient_as_bucket_view=True)  # fmt: skip
    assert ddp.bucket_bytes_cap // 2**20 == 15 and ddp.find_unused_parameters and ddp.dim == 0
    assert ddp.broadcast_buffers is False and ddp.gradient_as_bucket_view is True
    assert ddp.require_backward_grad_sync and ddp.require_forward_param_sync
    assert list(ddp.state_dict()) == ["module.0.weight", "module.0.bias", "module.2.weight", "module.2.bias"]
    default = B200DataParallel(small_model())
    assert default.bucket_bytes_cap == 25 * 2**20 and default.bucket_bytes_cap_default
    red = default.reducer
    assert len(red.buckets) == 1 and [p.shape for p in red.buckets[0].params] == [(3,), (3, 7), (7,), (7, 5)]
    for b in red.buckets:  # every slice starts on a 32-byte boundary, padding is zero
        assert all(off % 8 == 0 for off in b.offsets) and b.count % 8 == 0
    assert float(next(iter(red.groups.values())).flat.abs().sum()) == 0.0
    with pytest.raises(RuntimeError, match="doesn't have any parameter"):
        B200DataParallel(torch.nn.ReLU())


def test_views_no_sync_and_latch():
    ddp = B200DataParallel(small_model())
    red = ddp.reducer
    x = torch.randn(4, 5)
    ddp(x).sum().backward()
    assert red.num_reductions == 1
    flat = next(iter(red.groups.values())).flat
    for b in red.buckets:
        for p, ptr in zip(b.params, b.ptrs):
            assert p.grad.data_ptr() == ptr  # grads live in the fusion buffer now
    want = [p.grad.clone() for p in ddp.parameters()]
    # accumulate in place into the views (no zero_grad): values double, still the same memory
    ddp(x).sum().backward()
    assert all(torch.allclose(p.grad, 2 * w) for p, w in zip(ddp.parameters(), want)) and red.num_reductions == 2
    assert flat.abs().sum() > 0
    # forward under no_sync, backward outside: NOT reduced
    with ddp.no_sync():
        out = ddp(x)
    out.sum().backward()
    assert red.num_reductions == 2
    # reducer.prepare_for_backward re-arms (what Accelerator.trigger_sync_in_backward reaches into)
    with ddp.no_sync():
        out = ddp(x)
    red.prepare_for_backward([])
    out.sum().backward()
    assert red.num_reductions == 3
    # eval / no_grad forwards never arm
    with torch.no_grad():
        ddp(x)
    assert red._armed is False
    red.remove_hooks()


def test_unused_parameters_policy():
    class Two(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = torch.nn.Linear(2, 2), torch.nn.Linear(2, 2)

        def forward(self, x, use_b=False):
            return self.b(x) if use_b else self.a(x)

    strict = B200DataParallel(Two())
    with pytest.raises(RuntimeError, match="did not receive a gradient"):
        strict(torch.randn(3, 2)).sum().backward()
    strict.reducer.remove_hooks()
    lenient = B200DataParallel(Two(), find_unused_parameters=True)
    lenient(torch.randn(3, 2)).sum().backward()
    # torch DDP: a parameter no rank used keeps grad=None (Reducer "globally unused")
    assert lenient.module.b.weight.grad is None and lenient.module.a.weight.grad is not None
    # a gradient accumulated under no_sync is reduced as it is by a sync backward that does not touch the parameter
    with lenient.no_sync():
        lenient(torch.randn(3, 2), use_b=True).sum().backward()
    kept = lenient.module.b.weight.grad.clone()
    lenient(torch.randn(3, 2)).sum().backward()
    assert torch.equal(lenient.module.b.weight.grad, kept) and float(kept.abs().sum()) > 0


def test_comm_hook_registration_surface():
    from accelerate_b200 import DDPCommunicationHookType, DistributedDataParallelKwargs

    ddp = B200DataParallel(small_model())
    DistributedDataParallelKwargs(comm_hook=DDPCommunicationHookType.BF16).register_comm_hook(ddp)
    assert ddp.reducer.wire_dtype == torch.bfloat16
    ddp(torch.randn(4, 5)).sum().backward()
    full = B200DataParallel(small_model())
    full(torch.randn(4, 5)).sum().backward()  # different input: only check representability
    for p in ddp.parameters():
        assert torch.equal(p.grad, p.grad.to(torch.bfloat16).float()), "wire format must round gradients to bf16"
    # one hook per wrapper, as in torch
    with pytest.raises(RuntimeError, match="only be called once"):
        DistributedDataParallelKwargs(comm_hook=DDPCommunicationHookType.POWER_SGD).register_comm_hook(ddp)
    with pytest.raises(RuntimeError, match="only be called once"):
        ddp.register_comm_hook(None, lambda state, bucket: None)
    # as in the reference, a wrapper without a hook registers nothing
    plain = B200DataParallel(small_model())
    DistributedDataParallelKwargs(comm_wrapper=DDPCommunicationHookType.BF16).register_comm_hook(plain)
    assert plain.reducer.wire_dtype is None and plain.reducer.user_hook is None


def test_user_comm_hook_protocol_single_process():
    """torch's plug point: hook(state, bucket) -> Future[Tensor], once per bucket per sync backward, in bucket order,
    with the UN-divided flat bucket; the future's value becomes the gradients (in place or as a new tensor)."""
    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks

    calls = []

    def doubling_hook(state, bucket):
        calls.append((bucket.index(), bucket.is_last(), bucket.buffer().numel(), len(bucket.gradients()), len(bucket.parameters())))
        for g, p in zip(bucket.gradients(), bucket.parameters()):
            assert g.shape == p.shape and g.data_ptr() >= bucket.buffer().data_ptr()
        fut = torch.futures.Future()
        fut.set_result(bucket.buffer() * state["factor"])  # a NEW tensor: must be copied back into the bucket
        return fut

    torch.manual_seed(0)
    ddp = B200DataParallel(small_model(), bucket_cap_mb=0.0001)
    ref = B200DataParallel(small_model(), bucket_cap_mb=0.0001)
    ref.load_state_dict(ddp.state_dict())
    ddp.register_comm_hook({"factor": 2.0}, doubling_hook)
    x = torch.randn(4, 5)
    for _ in range(2):  # second pass: gradients are already views of the fusion buffer
        ddp.zero_grad(set_to_none=False)
        ref.zero_grad(set_to_none=False)
        calls.clear()
        ddp(x).sum().backward()
        ref(x).sum().backward()
        assert [c[0] for c in calls] == list(range(len(ddp.reducer.buckets))) and len(calls) > 1
        assert [c[1] for c in calls] == [False] * (len(calls) - 1) + [True]
        for p, q in zip(ddp.parameters(), ref.parameters()):
            assert torch.equal(p.grad, q.grad * 2.0)
    assert ddp.reducer.stats_valid is False or not ddp.reducer.is_cuda
    with ddp.no_sync():
        calls.clear()
        ddp(x).sum().backward()
        assert calls == [], "no_sync: the hook must not run"
    with pytest.raises(TypeError, match="torch.futures.Future"):
        bad = B200DataParallel(small_model())
        bad.register_comm_hook(None, lambda state, bucket: bucket.buffer())
        bad(x).sum().backward()
    # torch's canonical hooks map onto the native paths
    fast = B200DataParallel(small_model())
    fast.register_comm_hook(None, default_hooks.allreduce_hook)
    assert fast.reducer.user_hook is None and fast.reducer.wire_dtype is None
    half = B200DataParallel(small_model())
    half.register_comm_hook(None, default_hooks.bf16_compress_hook)
    assert half.reducer.wire_dtype == torch.bfloat16 and half.reducer.user_hook is None


class _Switchable(torch.nn.Module):
    """Two branches applied in an order chosen per call: the gradient-ready order inside a bucket flips with it."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(1)
        self.a = torch.nn.Linear(6, 6)
        self.b = torch.nn.Linear(6, 6)
        self.c = torch.nn.Linear(6, 64)  # large enough to fill a bucket of its own: a and b then share the next one
        self.order = "ab"

    def forward(self, x):
        first, second = (self.a, self.b) if self.order == "ab" else (self.b, self.a)
        return self.c(torch.tanh(second(torch.tanh(first(x)))))


def test_fast_hooks_learn_the_firing_order_and_survive_a_change():
    """Round-2 fast hooks: per bucket only the parameter that fired last in the previous iteration runs Python, every other
    hook is `list.append`.  When the order changes, the bucket must still be reduced exactly once per backward with all of
    its gradients final (late, at a later trigger or at the end of backward), and the new order is learned."""
    import copy

    ddp = B200DataParallel(_Switchable(), bucket_cap_mb=0.0003)  # ~300-byte buckets: [c.bias, c.weight], [b.*, a.*]
    ref = copy.deepcopy(ddp.module)
    red = ddp.reducer
    assert red._fast and [len(b.params) for b in red.buckets] == [2, 4]
    g = torch.Generator().manual_seed(0)
    relearned = 0
    for it, order in enumerate(["ab", "ab", "ba", "ba", "ab", "ba", "ba"]):
        ddp.module.order = ref.order = order
        x = torch.randn(5, 6, generator=g)
        triggers_before = [b.trigger for b in red.buckets]
        for m in (ddp, ref):
            m.zero_grad(set_to_none=(it % 2 == 0))
            m(x).square().mean().backward()
        assert red.num_reductions == it + 1
        for p, q in zip(ddp.module.parameters(), ref.parameters()):
            assert torch.equal(p.grad, q.grad), (it, order)
        for b in red.buckets:  # gradients live in the fusion buffer afterwards
            for p, ptr in zip(b.params, b.ptrs):
                assert p.grad.data_ptr() == ptr
        relearned += sum(t != b.trigger for t, b in zip(triggers_before, red.buckets))
    assert relearned >= 1, "flipping the order must have moved at least one bucket's trigger"
    # a backward under no_sync leaves the reducer idle and the bookkeeping clean
    with ddp.no_sync():
        ddp(x).square().mean().backward()
    assert red.num_reductions == 7
    ddp(x).square().mean().backward()
    assert red.num_reductions == 8
