#!/usr/bin/env python
"""Why did the bucket pack run at 3.2 TB/s inside bench.py (destination = the symmetric fusion buffer) but at 5.2 TB/s
in the stand-alone sweep (destination = an ordinary torch allocation)?  Same kernel, same sources: time it for every
combination of source / destination memory kind.

    python benchmarks/pack_dst_probe.py                  # 1 GPU: torch memory vs cuMemCreate segment
    torchrun --nproc-per-node 2 benchmarks/pack_dst_probe.py   # adds peer-mapped and multicast-bound segments
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from accelerate_b200 import Accelerator
    from accelerate_b200.state import PartialState

    acc = Accelerator()
    rt = PartialState().runtime
    dev = acc.device
    st = torch.cuda.current_stream().cuda_stream
    shapes = [(4096, 1024), (4096,), (1024, 4096), (1024,), (1024, 1024), (1024,), (3072, 1024), (3072,), (1024,), (1024,)]
    nbytes = [4 * int(torch.tensor(s).prod()) for s in shapes]
    total = sum(nbytes)
    span = (total + 64 * len(shapes) + (2 << 20)) // (2 << 20) * (2 << 20)

    def spans(base):
        out, off = [], base
        for n in nbytes:
            out.append(off)
            off += (n + 31) // 32 * 32
        return out

    kinds = {}
    t_src = torch.randn(span // 4, device=dev)
    t_dst = torch.empty(span // 4, device=dev)
    kinds["torch"] = (t_src.data_ptr(), t_dst.data_ptr())
    seg_a = rt.alloc_segment(span, multicast=False)
    seg_b = rt.alloc_segment(span, multicast=False)
    kinds["segment"] = (seg_a.ptr, seg_b.ptr)
    if rt.multicast_possible:
        mc_a = rt.alloc_segment(span, multicast=True)
        mc_b = rt.alloc_segment(span, multicast=True)
        kinds["segment_multicast_bound"] = (mc_a.ptr, mc_b.ptr)
    rows = {}
    for sk, (sp, _) in kinds.items():
        for dk, (_, dp) in kinds.items():
            srcs, dsts = spans(sp), spans(dp)
            for label, bulk in (("ldg_stg", 0), ("bulk_tma", 1)):
                t = bench.event_time(lambda: rt.ctx.pack(srcs, dsts, nbytes, bulk, st), reps=20)
                rows[f"{sk}->{dk}:{label}"] = round(2 * total / (t * 1e-3) / 1e9)
    # torch's own copy between the same kinds of memory
    for sk, (sp, _) in kinds.items():
        for dk, (_, dp) in kinds.items():
            a = torch.as_tensor(bench_span(sp, total, rt), device=dev)
            b = torch.as_tensor(bench_span(dp, total, rt), device=dev)
            t = bench.event_time(lambda: b.copy_(a), reps=20)
            rows[f"{sk}->{dk}:torch_copy"] = round(2 * total / (t * 1e-3) / 1e9)
    if acc.process_index == 0:
        print(json.dumps({"bytes": total, "world": acc.num_processes, "GBps": rows}), flush=True)
    acc.wait_for_everyone()
    PartialState().destroy_process_group()


def bench_span(ptr, nbytes, keep):
    from accelerate_b200.runtime import _DeviceSpan

    return _DeviceSpan(ptr, nbytes, keep)


if __name__ == "__main__":
    main()
