#!/usr/bin/env python
"""Single-GPU probe that launches each hot kernel a few times on realistic sizes so `ncu --set full` can capture
it (multi-rank commands must not be wrapped in ncu).  World size 1: the allreduce kernel's NVLink side is absent,
its HBM side (read S, write S, statistics) is what the capture shows.

    ncu --set full --clock-control none --import-source on \
        -k regex:"allreduce_inplace|pack_bulk|pack_kernel|scale_flat|push_kernel|multi_adam" \
        -c 14 -o gpurun_out/prof_kernels python benchmarks/kernel_probe.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def adam_full_size():
    """The fused Adam/AdamW step at the HEADLINE size: GPT-2-medium's 292 parameter tensors (354,823,168 elements,
    loose gradients), one launch = 28 B x 354.8 M = 9.93 GB of algorithmic traffic.

        ncu --set full --clock-control none --import-source on -k regex:multi_adam -s 1 -c 1 \
            -o gpurun_out/r02_multi_adam python benchmarks/kernel_probe.py --adam
    """
    from transformers import GPT2Config, GPT2LMHeadModel

    from accelerate_b200.state import PartialState

    rt = PartialState().runtime
    st = torch.cuda.current_stream().cuda_stream
    with torch.device("meta"):
        shapes = [tuple(p.shape) for p in GPT2LMHeadModel(GPT2Config(n_embd=1024, n_layer=24, n_head=16)).parameters()]
    params = [torch.randn(s, device=rt.device) * 0.02 for s in shapes]
    grads = [torch.randn(s, device=rt.device) * 1e-3 for s in shapes]
    numels = [p.numel() for p in params]
    assert sum(numels) == 354_823_168, sum(numels)
    moffs, total = [], 0
    for k in numels:
        moffs.append(total)
        total += (k + 3) & ~3
    m_flat = torch.zeros(total, device=rt.device)
    v_flat = torch.zeros(total, device=rt.device)
    table = rt.ctx.make_adam_table([p.data_ptr() for p in params], numels, moffs, [g.data_ptr() for g in grads])
    mult = torch.tensor([0.7], device=rt.device)
    for rep in range(3):
        rt.ctx.multi_adam(table, m_flat.data_ptr(), v_flat.data_ptr(), 1 - 1e-4 * 0.01, 0.0, 0.9, 0.999, -1e-4 / (1 - 0.9 ** (rep + 1)),
                          (1 - 0.999 ** (rep + 1)) ** 0.5, 1e-8, mult.data_ptr(), None, st)  # fmt: skip
    torch.cuda.synchronize()
    rt.ctx.check_status()
    print("adam probe ok", sum(numels), "elements")


def main():
    if "--adam" in sys.argv:
        return adam_full_size()
    from accelerate_b200 import native
    from accelerate_b200.state import PartialState

    rt = PartialState().runtime
    st = torch.cuda.current_stream().cuda_stream
    n = 25 * 1024 * 1024 // 4  # one 25 MiB fp32 bucket
    seg = rt.alloc_segment(8 * n * 4, multicast=False)
    flat = seg.tensor(torch.float32, 8 * n)
    flat.normal_()
    slot = rt.reserve_slots(8)
    out = torch.empty(3, device=rt.device)
    grads = [torch.randn(1024, 1024, device=rt.device) for _ in range(24)] + [torch.randn(1024, device=rt.device) for _ in range(8)]
    sizes = [g.numel() * 4 for g in grads]
    dsts, off = [], 0
    for s in sizes:
        dsts.append(flat.data_ptr() + off)
        off += (s + 31) // 32 * 32
    # fused Adam/AdamW step on a GPT-2-medium-like tensor list (24 blocks' worth of 1024x1024 matrices + vectors):
    # gradients = the tensors above, moments in two flat buffers, 28 algorithmic bytes per element
    params = [torch.randn_like(g) for g in grads]
    numels = [g.numel() for g in grads]
    moffs, total = [], 0
    for k in numels:
        moffs.append(total)
        total += (k + 3) & ~3
    m_flat = torch.zeros(total, device=rt.device)
    v_flat = torch.zeros(total, device=rt.device)
    adam_table = rt.ctx.make_adam_table([p.data_ptr() for p in params], numels, moffs, [g.data_ptr() for g in grads])
    small = torch.randn(4096, device=rt.device)
    staging = rt.ensure_staging()
    small_out = torch.empty_like(small)
    for rep in range(3):
        for j in range(4):  # the steady-state sequence of a backward: stats on, no closing barrier
            rt.ctx.allreduce_bucket(seg.seg, j * n, n, native.F32, 0.125, slot + j, native.F_STATS | native.F_NO_END_BARRIER,
                                    native.ALGO_TWOSHOT, native.CH_GRAD, 16, native.NO_STREAM, st)
        rt.ctx.fence(native.CH_GRAD, st)
        rt.ctx.finalize_norm(slot, 4, True, 2.0, 1.0, None, out.data_ptr(), st)
        rt.ctx.scale_flat(flat.data_ptr(), 4 * n, native.F32, out.data_ptr() + 4, 0, st)
        rt.ctx.pack([g.data_ptr() for g in grads], dsts, sizes, 1, st)
        rt.ctx.pack([g.data_ptr() for g in grads], dsts, sizes, 0, st)
        rt.ctx.push(staging.seg, 1, small.data_ptr(), small.numel() * 4, native.F32, 1.0, small_out.data_ptr(), native.CH_PUSH, st)
        rt.ctx.multi_adam(adam_table, m_flat.data_ptr(), v_flat.data_ptr(), 1 - 1e-4 * 0.01, 0.0, 0.9, 0.999, -1e-4 / (1 - 0.9 ** (rep + 1)),
                          (1 - 0.999 ** (rep + 1)) ** 0.5, 1e-8, out.data_ptr() + 4, None, st)  # fmt: skip
    torch.cuda.synchronize()
    rt.ctx.check_status()
    print("probe ok")


if __name__ == "__main__":
    main()
