#!/usr/bin/env python
"""CTA-count sweep of the fused large-message kernels (b200dp_gather / b200dp_reduce) next to NCCL, to fix their default
grid sizes.  NVLink-bound kernels stream fastest with FEW CTAs that each keep many 16-byte requests in flight (r01 found
the same for the bucket allreduce: 16 CTAs beat 100).

    torchrun --nproc-per-node N benchmarks/collective_sweep.py [--out gpurun_out/r02/collective_sweep_nN.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--blocks", default="8,16,24,32,48,64,96,148")
    ap.add_argument("--sizes-mib", default="16,64,256")
    args = ap.parse_args()
    os.environ.setdefault("B200DP_STAGING_MB", "288")
    from accelerate_b200 import Accelerator, native
    from accelerate_b200.state import PartialState

    acc = Accelerator()
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    rt = PartialState().runtime
    nccl = PartialState().backend == "nccl"
    staging = rt.ensure_staging()
    st = torch.cuda.current_stream().cuda_stream
    algo = native.ALGO_NVLS if (staging.multicast and world >= 4) else native.ALGO_TWOSHOT

    def timeit(fn, nbytes):
        iters = 30 if nbytes <= (64 << 20) else 8
        for _ in range(4):
            fn()
        acc.wait_for_everyone()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return acc.gather(torch.tensor([s.elapsed_time(e) / iters], device=dev)).max().item()

    rows = []
    for mib in [int(x) for x in args.sizes_mib.split(",")]:
        nbytes = mib << 20
        n = nbytes // 4
        t = torch.randn(n, device=dev)
        out_r = torch.empty_like(t)
        row = {"MiB": mib}
        ar, ag = 2 * (world - 1) / world, world - 1
        gather_ok = nbytes * world <= (6 << 30)
        out_g = torch.empty(n * world, device=dev) if gather_ok else None
        for blocks in [int(b) for b in args.blocks.split(",")]:
            ms = timeit(lambda: rt.ctx.reduce(staging.seg, 0, t.data_ptr(), n, native.F32, 1.0, out_r.data_ptr(), algo, native.CH_USER, blocks, st), nbytes)
            row[f"reduce_b{blocks}_busbw"] = nbytes * ar / (ms * 1e-3) / 1e9
            if gather_ok:
                ms = timeit(lambda: rt.ctx.gather(staging.seg, 0, t.data_ptr(), nbytes, out_g.data_ptr(), nbytes, native.CH_USER, blocks, st), nbytes)
                row[f"gather_b{blocks}_busbw"] = nbytes * ag / (ms * 1e-3) / 1e9
        # the round-1 sequences the fused kernels replace: pack -> in-place allreduce on staging -> pack out ; pack -> pull
        def unfused_reduce():
            rt.ctx.pack([t.data_ptr()], [staging.ptr], [nbytes], rt.bulk_for(nbytes), st)
            rt.ctx.allreduce_bucket(staging.seg, 0, n, native.F32, 1.0, -1, 0, algo, native.CH_USER, 0, native.NO_STREAM, st)
            rt.ctx.pack([staging.ptr], [out_r.data_ptr()], [nbytes], rt.bulk_for(nbytes), st)

        def unfused_gather():
            rt.ctx.pack([t.data_ptr()], [staging.ptr], [nbytes], rt.bulk_for(nbytes), st)
            rt.ctx.pull(staging.seg, 0, nbytes, out_g.data_ptr(), nbytes, 0, world, native.CH_USER, 0, st)

        ms = timeit(unfused_reduce, nbytes)
        row["unfused_reduce_busbw"] = nbytes * ar / (ms * 1e-3) / 1e9
        if gather_ok:
            ms = timeit(unfused_gather, nbytes)
            row["unfused_gather_busbw"] = nbytes * ag / (ms * 1e-3) / 1e9
        if nccl:
            ms = timeit(lambda: dist.all_reduce(out_r), nbytes)
            row["nccl_allreduce_busbw"] = nbytes * ar / (ms * 1e-3) / 1e9
            if gather_ok:
                ms = timeit(lambda: dist.all_gather_into_tensor(out_g, t), nbytes)
                row["nccl_allgather_busbw"] = nbytes * ag / (ms * 1e-3) / 1e9
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del t, out_r, out_g
    rt.ctx.check_status()
    if rank == 0 and args.out:
        json.dump({"world": world, "algo": "nvls" if algo == native.ALGO_NVLS else "twoshot", "rows": rows}, open(args.out, "w"), indent=1)
    acc.wait_for_everyone()
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
