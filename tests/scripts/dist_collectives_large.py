"""Large-message path of accelerator.gather / accelerator.reduce (the fused one-launch kernels b200dp_gather /
b200dp_reduce; reference semantics: operations.py:316-352 gather, :839-878 reduce): sizes above the 512 KB latency path,
ragged byte counts, every float dtype, "mean", a source tensor that starts at an odd offset, and tensors larger than the
staging window (several launches).  gather must be bit-exact; reduce must equal the rank-order fp32 sum rounded once to
the tensor's dtype (two-shot) bit for bit — NVLS sums in the switch's order: float32-rounding close — and `mean`
divides in the tensor's dtype like the reference's `/= N`; the input must stay untouched."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from accelerate_b200 import Accelerator
    from accelerate_b200.state import PartialState

    acc = Accelerator()
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    rt = PartialState().runtime
    staging = rt.ensure_staging()
    nvls = bool(staging.multicast and world >= 4)
    # elements: ragged sizes (fused / pipelined kernels; the third exceeds a 16 MiB window in fp32) and 16-byte-exact ones
    # (two ranks: the push kernel b200dp_reduce uses for aligned tensors; the last one needs two windows)
    sizes = [(1 << 20) + 3, 3 * (1 << 20) + 17, 5 * (1 << 20) + 1, 3 * (1 << 20), 6 * (1 << 20)]
    for dtype in (torch.float32, torch.bfloat16, torch.float16):
        for numel in sizes:
            g = torch.Generator().manual_seed(7919 * rank + numel % 1009)
            base = (torch.randn(numel + 1, generator=g) * 2).to(dtype).to(dev)
            for x in (base[:numel], base[1:]):  # second view: data pointer is only element-aligned
                keep = x.clone()
                gathered = acc.gather(x)  # 1-D: concatenation of the ranks' tensors
                assert gathered.shape == (world * numel,) and gathered.dtype == dtype
                parts = gathered.view(world, numel)
                assert torch.equal(parts[rank], x), ("gather: own block", dtype, numel)
                # every rank must see the same concatenation: compare checksums through the latency path
                chk = parts.float().sum(dim=1, dtype=torch.float64)
                allchk = acc.gather(chk.reshape(1, -1))
                assert torch.equal(allchk, allchk[0].expand_as(allchk)), ("gather: ranks disagree", dtype, numel)
                total = acc.reduce(x, "sum")
                assert torch.equal(x, keep), "reduce must not mutate its input"
                want = parts[0].float()
                for r in range(1, world):
                    want = want + parts[r].float()  # rank order, fp32 accumulate
                want = want.to(dtype)
                if nvls:
                    # the switch adds the N values in its own fixed order: any summation order stays within the
                    # forward-error bound (N-1) * 2^-24 * sum_r |x_r| of the exact sum (so 2x that between two orders),
                    # plus one rounding to the tensor's dtype
                    bound = 2 * (world - 1) * 2.0**-24 * parts.float().abs().sum(dim=0)
                    ulp = {torch.float32: 2.0**-23, torch.bfloat16: 2.0**-7, torch.float16: 2.0**-10}[dtype]
                    err = (total.float() - want.float()).abs()
                    limit = bound + ulp * want.float().abs() + 1e-30
                    assert bool((err <= limit).all()), ("reduce nvls", dtype, numel, (err / limit).max().item())
                else:
                    assert torch.equal(total, want), ("reduce", dtype, numel, (total.float() - want.float()).abs().max().item())
                mean = acc.reduce(x, "mean")
                want_mean = (total.float() / world).to(dtype)
                assert torch.equal(mean, want_mean), ("mean", dtype, numel, (mean.float() - want_mean.float()).abs().max().item())
                del gathered, parts, total, mean
            del base
    # non-float dtypes keep the staged one-shot path at any size
    n_int = (1 << 19) + 5
    xi = torch.arange(rank, rank + n_int, device=dev, dtype=torch.int64)
    si = acc.reduce(xi, "sum")
    assert torch.equal(si, sum(torch.arange(r, r + n_int, device=dev, dtype=torch.int64) for r in range(world)))
    gi = acc.gather(xi)
    assert torch.equal(gi.view(world, -1)[world - 1], torch.arange(world - 1, world - 1 + n_int, device=dev, dtype=torch.int64))
    rt.ctx.check_status()
    acc.wait_for_everyone()
    print(f"COLLECTIVES_LARGE_OK rank {rank}/{world} nvls={nvls}", flush=True)
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
