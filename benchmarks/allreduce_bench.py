#!/usr/bin/env python
"""BASELINE.json config 5: allreduce / allgather bus GB/s vs message size, ours vs torch.distributed NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        benchmarks/allreduce_bench.py [--max-log2 30] [--blocks 0,32,148] [--out gpurun_out/allreduce_nN.json]

For each size S = 2^k bytes (fp32, or bf16 with --dtype bf16): (a) the in-place fused bucket allreduce on a symmetric segment — the
gradient path (two-shot and, where bound, NVLS), optionally for several CTA counts; (b) accelerator.reduce(t)
end to end (staging copy + kernel + copy out); (c) accelerator.gather(t); (d) NCCL all_reduce /
all_gather_into_tensor through torch.distributed (the reference's path, operations.py:871,342).
busbw = algbw * 2(N-1)/N (allreduce), algbw * (N-1)/N with S = total output (allgather): nccl-tests convention.
CUDA-event timed on the launching stream, max over ranks.
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
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=30)
    ap.add_argument("--blocks", default="0")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"], help="element type of the message (config 5: both)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from accelerate_b200 import Accelerator, native
    from accelerate_b200.state import PartialState

    acc = Accelerator()
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    state = PartialState()
    rt = state.runtime
    nccl = state.backend == "nccl"
    max_bytes = 1 << args.max_log2
    tdtype, ndtype, esize = {"fp32": (torch.float32, native.F32, 4), "bf16": (torch.bfloat16, native.BF16, 2)}[args.dtype]
    seg = rt.alloc_segment(max_bytes, multicast=True)
    flat = seg.tensor(tdtype, max_bytes // esize)
    flat.normal_()
    st = torch.cuda.current_stream().cuda_stream
    blocks_list = [int(b) for b in args.blocks.split(",")]
    algos = [("twoshot", native.ALGO_TWOSHOT)] + ([("nvls", native.ALGO_NVLS)] if seg.multicast else [])
    rows = []

    def timeit(fn, nbytes):
        iters = 100 if nbytes <= (64 << 20) else (20 if nbytes <= (256 << 20) else 8)
        warm = 20 if nbytes <= (64 << 20) else 3
        for _ in range(warm):
            fn()
        acc.wait_for_everyone()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / iters], device=dev)
        return acc.gather(ms).max().item()

    for k in range(args.min_log2, args.max_log2 + 1):
        nbytes = 1 << k
        n = nbytes // esize
        row = {"bytes": nbytes, "dtype": args.dtype}
        for name, algo in algos:
            for blocks in blocks_list:
                ms = timeit(lambda: rt.ctx.allreduce_bucket(seg.seg, 0, n, ndtype, 1.0 / world, 0, native.F_STATS, algo,
                                                            native.CH_USER, blocks, native.NO_STREAM, st), nbytes)  # fmt: skip
                row[f"{name}_b{blocks}_us"] = ms * 1e3
                row[f"{name}_b{blocks}_busbw"] = nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
            flat[:n].normal_()  # keep values bounded: the in-place mean keeps shrinking them otherwise
        t = torch.randn(n, device=dev).to(tdtype)
        ms = timeit(lambda: acc.reduce(t, "sum"), nbytes)
        row["api_reduce_us"] = ms * 1e3
        row["api_reduce_busbw"] = nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
        if nbytes * world <= (4 << 30):
            ms = timeit(lambda: acc.gather(t), nbytes)
            row["api_gather_us"] = ms * 1e3
            row["api_gather_busbw"] = nbytes * (world - 1) / (ms * 1e-3) / 1e9
        if nccl:
            ms = timeit(lambda: dist.all_reduce(t), nbytes)
            row["nccl_allreduce_us"] = ms * 1e3
            row["nccl_allreduce_busbw"] = nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
            if nbytes * world <= (4 << 30):
                out = torch.empty(n * world, device=dev, dtype=tdtype)
                ms = timeit(lambda: dist.all_gather_into_tensor(out, t), nbytes)
                row["nccl_allgather_us"] = ms * 1e3
                row["nccl_allgather_busbw"] = nbytes * (world - 1) / (ms * 1e-3) / 1e9
                del out
        del t
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    rt.ctx.check_status()
    if rank == 0 and args.out:
        json.dump({"world": world, "dtype": args.dtype, "multicast": bool(seg.multicast), "kind": rt.kind, "rows": rows}, open(args.out, "w"), indent=1)
    acc.wait_for_everyone()
    state.destroy_process_group()


if __name__ == "__main__":
    main()
