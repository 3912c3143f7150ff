#!/usr/bin/env python
"""Single-GPU sweep of the bucket pack kernels on a real GPT-2-medium bucket's tensor list (one transformer block:
~12.6 M elements in 12 tensors, padded to ~32 MiB with the next block's first tensors), next to torch's own copy of
the same bytes.  Each TMA configuration runs in its own process (the configuration is read once per process).

    python benchmarks/pack_sweep.py            # prints one JSON line per configuration
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch

    import bench
    from accelerate_b200.state import PartialState

    rt = PartialState().runtime
    st = torch.cuda.current_stream().cuda_stream
    shapes = [(1024,), (1024,), (1024, 3072), (3072,), (1024, 1024), (1024,), (1024,), (1024,), (1024, 4096), (4096,), (4096, 1024), (1024,)]
    out = {}
    for label, reps in (("one_block_50MB", 1), ("small_bucket_4MB", 0)):
        these = shapes * reps if reps else [(1024, 1024), (1024,)]
        fresh = [torch.randn(s, device=rt.device) for s in these]
        peaks, _ = bench.measured_peaks()
        out[label] = bench.pack_probe(rt, fresh, st, float(peaks["hbm_gbs"]))
    # the whole gradient set in one launch (upper bound for this kernel at any size)
    big = [torch.randn(1024, 1024, device=rt.device) for _ in range(256)]
    peaks, _ = bench.measured_peaks()
    out["256_x_4MB_1GB"] = bench.pack_probe(rt, big, st, float(peaks["hbm_gbs"]))
    print("PACK_SWEEP " + json.dumps({"cfg": {k: os.environ.get(k) for k in ("B200DP_BULK_WARPS", "B200DP_BULK_STAGES", "B200DP_BULK_STAGE_KB", "B200DP_BULK_CTAS_PER_SM")},
                                       "result": out}), flush=True)


def main():
    if os.environ.get("PACK_SWEEP_CHILD") == "1":
        return child()
    cfgs = [(4, 3, 16, 1), (4, 4, 8, 1), (8, 3, 8, 1), (4, 2, 8, 2), (2, 4, 16, 1), (4, 3, 8, 2), (8, 2, 4, 2), (1, 4, 32, 1)]
    for warps, stages, kb, per_sm in cfgs:
        env = dict(os.environ, PACK_SWEEP_CHILD="1", B200DP_BULK_WARPS=str(warps), B200DP_BULK_STAGES=str(stages),
                   B200DP_BULK_STAGE_KB=str(kb), B200DP_BULK_CTAS_PER_SM=str(per_sm))
        res = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=300)
        lines = [l for l in res.stdout.splitlines() if l.startswith("PACK_SWEEP ")]
        print(lines[-1][len("PACK_SWEEP "):] if lines else json.dumps({"cfg": [warps, stages, kb, per_sm], "error": res.stderr[-800:]}), flush=True)


if __name__ == "__main__":
    main()
