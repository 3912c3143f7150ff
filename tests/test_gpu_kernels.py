"""-m gpu: every kernel behind the C ABI against the oracle (numpy) on seeded inputs, one process.

World size 1 still runs the real kernels: the cross-rank barrier degenerates to a flag write/read on
the local pad, the allreduce is "load own copy, pre-divide, store".  Multi-rank parity lives in
test_gpu_dist.py.  Integer/byte work is bit-exact; float work is bit-exact where the arithmetic is
defined (scale, pre-divide) and within 1e-6 relative for reductions whose association order differs.
"""

import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ddp_oracle as oracle  # noqa: E402


@pytest.fixture(scope="module")
def rt():
    from accelerate_b200 import native
    from accelerate_b200.state import PartialState

    assert torch.cuda.is_available()
    state = PartialState()
    assert state.runtime is not None, "native runtime must be up on a CUDA box (no fallback)"
    return state.runtime


def stream():
    return torch.cuda.current_stream().cuda_stream


def read_out(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("numel", [4 * 8, 1000 * 8, 1 << 20, (1 << 22) + 8 * 37])
def test_allreduce_bucket_world1_and_stats(rt, dtype, numel):
    from accelerate_b200 import native

    seg = rt.alloc_segment(numel * 4 + 64, multicast=False)
    try:
        g = torch.Generator(device="cpu").manual_seed(numel)
        x = (torch.randn(numel, generator=g) * 3).to(dtype)
        view = seg.tensor(dtype, numel, byte_offset=16)
        view.copy_(x)
        slot = rt.reserve_slots(1)
        es = x.element_size()
        rt.ctx.allreduce_bucket(seg.seg, 16 // es, numel, DT[dtype], 0.25, slot, native.F_STATS, native.ALGO_TWOSHOT,
                                native.CH_GRAD, 0, native.NO_STREAM, stream())
        out = torch.empty(3, device=rt.device)
        rt.ctx.finalize_norm(slot, 1, True, 2.0, -1.0, None, out.data_ptr(), stream())
        got = view.float().cpu()
        want = (x.float() * 0.25).to(dtype).float()  # x * 2^-2 is exact in every format
        assert torch.equal(got, want)
        res = read_out(out)
        ref = math.sqrt(float((want.double() ** 2).sum()))
        assert abs(res[0] - ref) <= 1e-6 * ref + 1e-12, (res, ref)
        assert res[1] == 1.0 and res[2] == 0.0
        # inf-norm + found_inf
        view[numel // 2] = float("inf")
        rt.ctx.allreduce_bucket(seg.seg, 16 // es, numel, DT[dtype], 1.0, slot, native.F_STATS, native.ALGO_TWOSHOT,
                                native.CH_GRAD, 0, native.NO_STREAM, stream())
        rt.ctx.finalize_norm(slot, 1, True, math.inf, -1.0, None, out.data_ptr(), stream())
        res = read_out(out)
        assert res[2] == 1.0 and math.isinf(res[0])
        rt.ctx.check_status()
    finally:
        rt.free_segment(seg)


def test_finalize_and_scale_match_torch_clip(rt):
    from accelerate_b200 import native

    n = 3_000_000
    seg = rt.alloc_segment(n * 4, multicast=False)
    try:
        x = torch.randn(n, generator=torch.Generator().manual_seed(3)) * 0.01
        v = seg.tensor(torch.float32, n)
        v.copy_(x)
        slot = rt.reserve_slots(1)
        rt.ctx.allreduce_bucket(seg.seg, 0, n, 0, 1.0, slot, native.F_STATS, native.ALGO_TWOSHOT, native.CH_GRAD, 0, native.NO_STREAM, stream())
        out = torch.empty(3, device=rt.device)
        rt.ctx.finalize_norm(slot, 1, True, 2.0, 1.0, None, out.data_ptr(), stream())
        rt.ctx.scale_flat(v.data_ptr(), n, 0, out.data_ptr() + 4, 0, stream())
        res = read_out(out)
        p = torch.nn.Parameter(x.clone())
        p.grad = x.clone()
        tn = torch.nn.utils.clip_grad_norm_([p], 1.0)
        want_clipped, want_tn = oracle.clip_grad_norm([x.numpy()], 1.0)
        assert abs(res[0] - want_tn) <= 1e-6 * want_tn  # oracle accumulates in float64
        assert abs(res[0] - tn.item()) <= 2e-4 * tn.item()  # torch CPU sums 3M squares in float32
        # given the device's multiplier the scaling itself is bit-exact
        assert np.array_equal(v.cpu().numpy(), x.numpy() * np.float32(res[1]))
        assert np.allclose(v.cpu().numpy(), p.grad.numpy(), rtol=2e-4, atol=0)
        # multiplier 1 -> untouched bits, no traffic
        one = torch.ones(1, device=rt.device)
        before = v.clone()
        rt.ctx.scale_flat(v.data_ptr(), n, 0, one.data_ptr(), 0, stream())
        assert torch.equal(before, v)
        # unscale via scale_ptr: inv_scale = float(1/double(scale))
        scale = torch.tensor([1024.0], device=rt.device)
        rt.ctx.finalize_norm(slot, 1, True, 2.0, -1.0, scale.data_ptr(), out.data_ptr(), stream())
        res2 = read_out(out)
        assert res2[1] == np.float32(1.0 / 1024.0) and abs(res2[0] - res[0] / 1024.0) <= 1e-6 * res2[0]
        # clamp
        rt.ctx.clamp_flat(v.data_ptr(), n, 0, 1e-4, 0, stream())
        assert np.array_equal(v.cpu().numpy(), oracle.clip_grad_value([before.cpu().numpy()], 1e-4)[0])
    finally:
        rt.free_segment(seg)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_multi_tensor_stats_and_scale(rt, dtype):
    sizes = [1, 3, 0, 8191, 8192, 8193, 100_003, 5, 1 << 20]
    g = torch.Generator().manual_seed(11)
    host = [(torch.randn(s, generator=g) * 2).to(dtype) for s in sizes]
    base = [torch.empty(s + 3, dtype=dtype, device=rt.device) for s in sizes]
    dev = []
    for i, (h, b) in enumerate(zip(host, base)):
        v = b[(i % 3) : (i % 3) + h.numel()]  # deliberately misaligned views
        v.copy_(h)
        dev.append(v)
    slot = rt.reserve_slots(2)
    used = rt.ctx.multi_stats([d.data_ptr() for d in dev], [d.numel() for d in dev], DT[dtype], slot, stream())
    assert used == 1
    out = torch.empty(3, device=rt.device)
    rt.ctx.finalize_norm(slot, used, False, 2.0, 0.5, None, out.data_ptr(), stream())
    rt.ctx.multi_scale([d.data_ptr() for d in dev], [d.numel() for d in dev], DT[dtype], out.data_ptr() + 4, stream())
    res = read_out(out)
    want_tn = oracle.total_norm([h.float().numpy() for h in host])
    assert abs(res[0] - want_tn) <= 2e-6 * want_tn, (res[0], want_tn)
    coef = oracle.clip_coef(res[0], 0.5)
    assert res[1] == coef
    for h, d in zip(host, dev):
        want = (h.float() * float(res[1])).to(dtype)
        assert torch.equal(d.cpu(), want), "scale must be a single correctly-rounded multiply"
    # inf norm
    rt.ctx.multi_stats([d.data_ptr() for d in dev], [d.numel() for d in dev], DT[dtype], slot, stream())
    rt.ctx.finalize_norm(slot, used, False, math.inf, -1.0, None, out.data_ptr(), stream())
    res = read_out(out)
    assert res[0] == oracle.total_norm([d.float().cpu().numpy() for d in dev], math.inf)


def test_multi_tensor_more_than_table(rt):
    from accelerate_b200 import native

    n = native.MT_MAX + 50
    ts = [torch.full((7,), float(i % 5), device=rt.device) for i in range(n)]
    slot = rt.reserve_slots(2)
    used = rt.ctx.multi_stats([t.data_ptr() for t in ts], [7] * n, 0, slot, stream())
    assert used == 2
    out = torch.empty(3, device=rt.device)
    rt.ctx.finalize_norm(slot, used, False, 2.0, -1.0, None, out.data_ptr(), stream())
    want = math.sqrt(sum(7 * float(i % 5) ** 2 for i in range(n)))
    assert abs(read_out(out)[0] - want) < 1e-5 * want


@pytest.mark.parametrize("bulk", [0, 1])
def test_pack_bit_exact(rt, bulk):
    g = torch.Generator().manual_seed(5)
    sizes = [0, 1, 15, 16, 17, 4096, 32768, 32769, 65536 + 48, 1_000_003, 8 << 20]
    srcs, dsts, keep = [], [], []
    for i, s in enumerate(sizes):
        a = torch.randint(0, 256, (s + 32,), dtype=torch.uint8, generator=g).to(rt.device)
        b = torch.zeros(s + 64, dtype=torch.uint8, device=rt.device)
        so, do = (0, 16) if i % 2 == 0 else (i % 7, (i * 3) % 5)  # aligned and misaligned spans
        srcs.append(a[so : so + s])
        dsts.append(b[do : do + s])
        keep.append((a, b, so, do, s))
    rt.ctx.pack([t.data_ptr() for t in srcs], [t.data_ptr() for t in dsts], sizes, bulk, stream())
    torch.cuda.synchronize()
    rt.ctx.check_status()
    for a, b, so, do, s in keep:
        assert torch.equal(b[do : do + s], a[so : so + s])
        assert int(b[:do].sum()) == 0 and int(b[do + s :].sum()) == 0, "wrote outside the span"


def test_pack_many_tensors(rt):
    from accelerate_b200 import native

    n = native.MT_MAX * 2 + 3
    a = torch.arange(n * 20, dtype=torch.float32, device=rt.device)
    b = torch.zeros_like(a)
    rt.ctx.pack([a.data_ptr() + 80 * i for i in range(n)], [b.data_ptr() + 80 * i for i in range(n)], [80] * n, 1, stream())
    assert torch.equal(a, b)


@pytest.mark.parametrize("dtype,pad", [(torch.float32, 0), (torch.int64, -100), (torch.bool, 1), (torch.bfloat16, 2.5),
                                       (torch.uint8, 255), (torch.float64, -1.5)])  # fmt: skip
@pytest.mark.parametrize("shape,dim,new", [((3, 5), 0, 7), ((3, 5), 1, 6), ((2, 3, 4), 1, 5), ((4,), 0, 9), ((2, 0, 3), 1, 2)])
@pytest.mark.parametrize("pad_first", [False, True])
def test_pad_copy_matches_oracle(rt, dtype, pad, shape, dim, new, pad_first):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(shape, generator=g) * 10).to(dtype)
    if dtype == torch.bool:
        x = torch.randn(shape, generator=g) > 0
    dx = x.to(rt.device)
    new_shape = list(shape)
    new_shape[dim] = new
    out = torch.empty(new_shape, dtype=dtype, device=rt.device)
    es = x.element_size()
    bits = int.from_bytes(torch.tensor([pad], dtype=dtype).view(torch.uint8).numpy().tobytes(), "little")
    outer = int(np.prod(shape[:dim])) if dim else 1
    inner = int(np.prod(shape[dim + 1 :])) if dim + 1 < len(shape) else 1
    rt.ctx.pad_copy(dx.data_ptr() if dx.numel() else None, out.data_ptr(), outer, shape[dim], new, inner * es, es, bits, pad_first, stream())
    # oracle: a second, longer "rank" forces the padding
    other_shape = list(shape)
    other_shape[dim] = new
    ref_in = x.float().numpy() if dtype == torch.bfloat16 else x.numpy()
    want = oracle.pad_across_processes([ref_in, np.zeros(other_shape, dtype=ref_in.dtype)], dim, pad, pad_first)[0]
    got = out.float().cpu().numpy() if dtype == torch.bfloat16 else out.cpu().numpy()
    assert np.array_equal(got, want)


def test_oneshot_pull_exchange_world1(rt):
    from accelerate_b200 import native

    if rt.world != 1:
        pytest.skip("single-process test")
    st = rt.ensure_staging()
    for dtype, dt in [(torch.float32, 0), (torch.float64, 3), (torch.int32, 4), (torch.int64, 5), (torch.bfloat16, 1)]:
        x = (torch.randn(1003, generator=torch.Generator().manual_seed(9)) * 7).to(dtype).to(rt.device)
        rt.ctx.pack([x.data_ptr()], [st.ptr], [x.numel() * x.element_size()], 1, stream())
        out = torch.empty_like(x)
        rt.ctx.allreduce_oneshot(st.seg, 0, x.numel(), dt, 1.0, out.data_ptr(), native.CH_USER, 0, stream())
        assert torch.equal(out, x)
        out2 = torch.zeros(x.numel() * x.element_size() + 3, dtype=torch.uint8, device=rt.device)[3:]  # misaligned dst
        rt.ctx.pull(st.seg, 0, out2.numel(), out2.data_ptr(), 0, 0, 1, native.CH_USER, 0, stream())
        assert torch.equal(out2, x.view(torch.uint8))
    assert rt.ctx.exchange_small(b"\x01\x02\x03\x04\x05", stream()) == [b"\x01\x02\x03\x04\x05"]
    assert rt.ctx.exchange_small(bytes(range(128)), stream()) == [bytes(range(128))]
    rt.ctx.check_status()


@pytest.mark.parametrize("narrow", [torch.bfloat16, torch.float16])
def test_cast_kernel_matches_torch_compress_hook(rt, narrow):
    """f32 -> narrow with divide (torch: buffer.to(dtype).div_(N)) and back; ragged and misaligned spans."""
    g = torch.Generator().manual_seed(8)
    for n, off in [(1, 0), (7, 0), (8, 0), (1000, 0), (1003, 1), (1 << 20, 0), ((1 << 20) + 5, 3)]:
        base = (torch.randn(n + 8, generator=g) * 50).to(rt.device)
        x = base[off : off + n]
        for world in (1, 2, 3, 8):
            out = torch.empty(n + 8, dtype=narrow, device=rt.device)[off : off + n]
            rt.ctx.cast(x.data_ptr(), DT[torch.float32], out.data_ptr(), DT[narrow], n, float(world), 0, stream())
            want = x.to(narrow).div_(world) if world != 1 else x.to(narrow)
            assert torch.equal(out, want), (n, off, world)
            back = torch.empty(n + 8, device=rt.device)[off : off + n]
            rt.ctx.cast(out.data_ptr(), DT[narrow], back.data_ptr(), DT[torch.float32], n, 1.0, 0, stream())
            assert torch.equal(back, want.float())


def test_push_latency_path_world1(rt):
    """b200dp_push (one launch, one barrier): gather = copy, reduce = identity / divisor at world size 1;
    ragged sizes, misaligned source, window boundaries, double-buffer parity over repeated calls."""
    from accelerate_b200 import native

    if rt.world != 1:
        pytest.skip("single-process test")
    st = rt.ensure_staging()
    limit = rt.ctx.push_max_bytes()
    assert limit == 32 * 16384 and st.nbytes >= 2 * rt.ctx.push_area_bytes()
    g = torch.Generator().manual_seed(4)
    for rep in range(3):
        for nbytes in (1, 15, 16, 4096, 16384, 16385, 100_000, limit):
            raw = torch.randint(0, 256, (nbytes + 8,), dtype=torch.uint8, generator=g).to(rt.device)
            for off in (0, 3):
                src = raw[off : off + nbytes]
                out = torch.zeros(nbytes, dtype=torch.uint8, device=rt.device)
                rt.ctx.push(st.seg, 0, src.data_ptr(), nbytes, native.U8, 1.0, out.data_ptr(), native.CH_PUSH, stream())
                assert torch.equal(out, src), (rep, nbytes, off)
    for dtype, dt in [(torch.float32, native.F32), (torch.float64, native.F64), (torch.int64, native.I64),
                      (torch.bfloat16, native.BF16), (torch.float16, native.F16), (torch.int32, native.I32)]:  # fmt: skip
        x = (torch.randn(10_001, generator=g) * 9).to(dtype).to(rt.device)
        out = torch.empty_like(x)
        rt.ctx.push(st.seg, 1, x.data_ptr(), x.numel() * x.element_size(), dt, 1.0, out.data_ptr(), native.CH_PUSH, stream())
        assert torch.equal(out, x)
        if x.is_floating_point():
            rt.ctx.push(st.seg, 1, x.data_ptr(), x.numel() * x.element_size(), dt, 4.0, out.data_ptr(), native.CH_PUSH, stream())
            assert torch.equal(out, x / 4)
    with pytest.raises(native.B200dpError, match="exceed"):
        big = torch.zeros(limit + 16, dtype=torch.uint8, device=rt.device)
        rt.ctx.push(st.seg, 0, big.data_ptr(), limit + 16, native.U8, 1.0, big.data_ptr(), native.CH_PUSH, stream())
    rt.ctx.check_status()


def test_public_ops_single_process(rt):
    """DistributedType.NO semantics of the reference: gather/pad return the input, reduce clones."""
    from accelerate_b200 import Accelerator

    acc = Accelerator()
    t = torch.arange(6.0, device=acc.device)
    assert acc.gather(t) is t
    r = acc.reduce(t, "sum")
    assert r is not t and torch.equal(r, t)
    assert acc.pad_across_processes(t) is t


def test_full_size_properties(rt):
    """GPT-2-medium-sized fusion buffer (354.8M fp32 = 1.42 GB): linearity and checksum properties."""
    from accelerate_b200 import native

    n = 354_823_168
    seg = rt.alloc_segment(n * 4, multicast=False)
    try:
        v = seg.tensor(torch.float32, n)
        v.normal_(generator=torch.Generator(device=rt.device).manual_seed(0))
        ref_sumsq = float((v.double() ** 2).sum())
        ref_sum = float(v.double().sum())
        slot = rt.reserve_slots(1)
        rt.ctx.allreduce_bucket(seg.seg, 0, n, 0, 0.5, slot, native.F_STATS, native.ALGO_TWOSHOT, native.CH_GRAD, 0, native.NO_STREAM, stream())
        out = torch.empty(3, device=rt.device)
        rt.ctx.finalize_norm(slot, 1, True, 2.0, 1.0, None, out.data_ptr(), stream())
        res = read_out(out)
        want = math.sqrt(ref_sumsq) * 0.5
        assert abs(res[0] - want) <= 1e-6 * want
        assert abs(float(v.double().sum()) - 0.5 * ref_sum) <= 1e-9 * abs(ref_sum) + 1e-3  # exact halving
        rt.ctx.scale_flat(v.data_ptr(), n, 0, out.data_ptr() + 4, 0, stream())
        after = math.sqrt(float((v.double() ** 2).sum()))
        assert abs(after - 1.0) < 1e-4, after  # clipped to max_norm=1
        rt.ctx.check_status()
    finally:
        rt.free_segment(seg)


@pytest.mark.parametrize("decoupled", [True, False])
def test_multi_adam_matches_oracle(rt, decoupled):
    """K9 through the C ABI vs oracle.adam_step (itself pinned on torch.optim.Adam/AdamW): ragged and misaligned
    tensors, more tensors than one table holds, the device-side multiplier and the found-inf skip.  float32 with
    identical operation order; the oracle emulates fma through float64 (rare last-bit differences), hence 2e-6."""
    from accelerate_b200 import native

    sizes = [1, 3, 4, 257, 4096 + 5, 70000] + [9] * (native.MT_MAX + 20)
    g = torch.Generator(device="cpu").manual_seed(11)
    host = {k: [] for k in "pgmv"}
    for n in sizes:
        host["p"].append(torch.randn(n, generator=g) * 0.5)
        host["g"].append(torch.randn(n, generator=g) * 1e-2)
        host["m"].append(torch.randn(n, generator=g) * 1e-2)
        host["v"].append(torch.rand(n, generator=g) * 1e-4 + 1e-8)
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += n  # moments deliberately NOT padded: odd offsets take the scalar path
    m_flat = torch.cat(host["m"]).to(rt.device)
    v_flat = torch.cat(host["v"]).to(rt.device)
    # parameters / gradients: separate allocations, the second one misaligned by one element
    p_dev = [t.to(rt.device) for t in host["p"]]
    g_dev = [t.to(rt.device) for t in host["g"]]
    odd = torch.empty(sizes[1] + 1, device=rt.device)
    odd[1:].copy_(host["g"][1])
    g_dev[1] = odd[1:]
    lr, beta1, beta2, eps, wd, step = 3e-4, 0.9, 0.999, 1e-8, 0.1, 3
    mult = torch.tensor([0.0, 0.625, 0.0], device=rt.device)  # finalize_norm's layout: [norm, multiplier, found_inf]
    table = rt.ctx.make_adam_table([t.data_ptr() for t in p_dev], sizes, offs, [t.data_ptr() for t in g_dev])
    decay_mul = (1 - lr * wd) if decoupled else 1.0
    l2 = 0.0 if decoupled else wd
    bc1, bc2_sqrt = 1 - beta1**step, (1 - beta2**step) ** 0.5

    def launch(found_ptr):
        rt.ctx.multi_adam(table, m_flat.data_ptr(), v_flat.data_ptr(), decay_mul, l2, beta1, beta2, -(lr / bc1), bc2_sqrt,
                          eps, mult.data_ptr() + 4, found_ptr, stream())  # fmt: skip

    # found_inf set: nothing may change
    mult[2] = 1.0
    launch(mult.data_ptr() + 8)
    torch.cuda.synchronize()
    assert torch.equal(m_flat.cpu(), torch.cat(host["m"])) and torch.equal(v_flat.cpu(), torch.cat(host["v"]))
    assert all(torch.equal(d.cpu(), h) for d, h in zip(p_dev, host["p"]))
    mult[2] = 0.0
    launch(mult.data_ptr() + 8)
    torch.cuda.synchronize()
    m_got, v_got = m_flat.cpu().numpy(), v_flat.cpu().numpy()
    for i, n in enumerate(sizes):
        want_p, want_m, want_v = oracle.adam_step(
            host["p"][i].numpy(), host["g"][i].numpy(), host["m"][i].numpy(), host["v"][i].numpy(), step, lr, beta1, beta2,
            eps, wd, decoupled=decoupled, mult=0.625,
        )  # fmt: skip
        np.testing.assert_allclose(p_dev[i].cpu().numpy(), want_p, rtol=2e-6, atol=1e-9, err_msg=f"param {i}")
        np.testing.assert_allclose(m_got[offs[i] : offs[i] + n], want_m, rtol=2e-6, atol=1e-10, err_msg=f"exp_avg {i}")
        np.testing.assert_allclose(v_got[offs[i] : offs[i] + n], want_v, rtol=2e-6, atol=1e-14, err_msg=f"exp_avg_sq {i}")
    # ... and BIT-exact against the C oracle, whose fmaf / sqrtf / division round exactly like the kernel's intrinsics
    from oracle import c_oracle

    for i, n in enumerate(sizes[:8]):
        want_p, want_m, want_v = c_oracle.adam_step(
            host["p"][i].numpy(), host["g"][i].numpy(), host["m"][i].numpy(), host["v"][i].numpy(), step, lr, beta1, beta2,
            eps, wd, decoupled=decoupled, mult=0.625,
        )  # fmt: skip
        assert np.array_equal(p_dev[i].cpu().numpy(), want_p), f"param {i} differs from the C oracle in some bits"
        assert np.array_equal(m_got[offs[i] : offs[i] + n], want_m) and np.array_equal(v_got[offs[i] : offs[i] + n], want_v)
    rt.ctx.check_status()
