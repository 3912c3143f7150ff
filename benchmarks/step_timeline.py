#!/usr/bin/env python
"""Timeline of ONE training step of the bench.py workload (GPT-2-medium, bf16, per-GPU batch 8 x 512), from a CUPTI
activity trace taken through torch.profiler (nsys is not in the image; Kineto records every kernel of the process,
including the ones libb200dp.so launches on its own stream).  VERDICT r1 asked for exactly this evidence: the backward
span, every bucket kernel's start/end on the side stream, the pack spans, and the tail after the last backward kernel.

    python benchmarks/step_timeline.py --out profiles/r02_timeline_n1.json
    torchrun --nproc-per-node 8 benchmarks/step_timeline.py --out profiles/r02_timeline_n8.json

Output (rank 0's trace; every rank runs the same step): a JSON summary + a text table next to it.  Three variants are
traced back to back: our step, our step under ``no_sync`` (no reduction: the backward kernels alone) and stock torch
DDP + NCCL, so that "what the reduction costs" splits into (a) slower backward kernels while bucket kernels are
resident (interference), (b) the exposed tail after the last backward kernel, (c) extra kernels on the compute stream.
Times under a profiler are NOT bench numbers; they are used for attribution only.
"""
import argparse
import contextlib
import gzip
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OWN = ("allreduce_inplace_kernel", "pack_kernel", "pack_bulk_kernel", "fence_kernel", "finalize_norm_kernel", "scale_flat_kernel",
       "multi_adam_kernel", "multi_stats_kernel", "multi_scale_kernel", "cast_kernel", "pull_kernel", "push_kernel",
       "exchange_small_kernel", "allreduce_oneshot_kernel", "pad_copy_kernel", "clamp_flat_kernel")  # fmt: skip
PHASES = ("fwd", "bwd", "clip", "opt", "zero")


def short(name):
    for o in OWN:
        if o in name:
            tail = name[name.index(o) + len(o):]
            return o + (tail.split("(")[0] if tail.startswith("<") else "")
    if "nccl" in name.lower():
        return "nccl:" + name.split("(")[0][:48]
    return None


def trace_steps(step_phases, steps_to_trace, warmup):
    """step_phases: list of (phase name, callable).  Returns the parsed chrome-trace events of the traced steps."""
    from torch.profiler import ProfilerActivity, profile, record_function

    for _ in range(warmup):
        for _, fn in step_phases:
            fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for s in range(steps_to_trace):
            with record_function(f"STEP{s}"):
                for name, fn in step_phases:
                    with record_function(f"PH_{name}"):
                        fn()
        torch.cuda.synchronize()
    fd, path = tempfile.mkstemp(suffix=".json")
    os.close(fd)
    prof.export_chrome_trace(path)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        events = json.load(f)["traceEvents"]
    os.unlink(path)
    return events


def analyse(events, which_step):
    """Split one traced step into phases by the CPU-side launch time of each kernel (correlation id -> runtime call)."""
    cpu_ranges = {}  # name -> (ts, end)
    for e in events:
        if e.get("cat") == "user_annotation" and e.get("ph") == "X":
            cpu_ranges.setdefault(e["name"], []).append((e["ts"], e["ts"] + e["dur"]))
    step_rng = cpu_ranges[f"STEP{which_step}"][0]
    phase_rng = {}
    for ph in PHASES:
        for ts, end in cpu_ranges.get(f"PH_{ph}", []):
            if step_rng[0] <= ts <= step_rng[1]:
                phase_rng[ph] = (ts, end)
    launch_ts = {}
    for e in events:
        if e.get("cat") in ("cuda_runtime", "cuda_driver") and "args" in e and "correlation" in e["args"]:
            launch_ts[e["args"]["correlation"]] = e["ts"]
    kernels = []
    for e in events:
        if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and e.get("ph") == "X":
            lt = launch_ts.get(e.get("args", {}).get("correlation"))
            if lt is None or not (step_rng[0] <= lt <= step_rng[1]):
                continue
            phase = next((ph for ph, (a, b) in phase_rng.items() if a <= lt <= b), "?")
            kernels.append({"name": e["name"], "ts": e["ts"], "dur": e["dur"], "stream": e["args"].get("stream"), "phase": phase,
                            "grid": e["args"].get("grid"), "block": e["args"].get("block")})  # fmt: skip
    kernels.sort(key=lambda k: k["ts"])
    if not kernels:
        return None
    by_stream = {}
    for k in kernels:
        by_stream.setdefault(k["stream"], []).append(k)
    compute = max(by_stream, key=lambda s: len(by_stream[s]))
    t0 = kernels[0]["ts"]
    out = {"step_span_us": kernels[-1]["ts"] + kernels[-1]["dur"] - t0, "compute_stream": compute,
           "streams": {str(s): len(v) for s, v in by_stream.items()}, "phases": {}}  # fmt: skip
    for ph in PHASES:
        ks = [k for k in by_stream[compute] if k["phase"] == ph]
        if ks:
            out["phases"][ph] = {"start_us": ks[0]["ts"] - t0, "end_us": ks[-1]["ts"] + ks[-1]["dur"] - t0,
                                 "span_us": ks[-1]["ts"] + ks[-1]["dur"] - ks[0]["ts"], "kernels": len(ks),
                                 "busy_us": sum(k["dur"] for k in ks)}  # fmt: skip
    own = [dict(k, short=short(k["name"])) for k in kernels if short(k["name"])]
    bwd = out["phases"].get("bwd")
    side = [k for k in own if k["stream"] != compute]
    out["own_kernel_totals_us"] = {}
    for k in own:
        agg = out["own_kernel_totals_us"].setdefault(k["short"], {"n": 0, "us": 0.0, "stream": k["stream"]})
        agg["n"] += 1
        agg["us"] += k["dur"]
    if bwd:
        bwd_end = t0 + bwd["end_us"]
        bwd_compute = [k for k in by_stream[compute] if k["phase"] == "bwd" and not short(k["name"])]
        last_compute_end = max((k["ts"] + k["dur"] for k in bwd_compute), default=bwd_end)
        side_bwd = [k for k in side if k["phase"] in ("bwd",)]
        out["backward"] = {
            "autograd_kernels": len(bwd_compute), "autograd_busy_us": sum(k["dur"] for k in bwd_compute),
            "last_autograd_kernel_end_us": last_compute_end - t0,
            "side_stream_kernels": len(side_bwd), "side_stream_busy_us": sum(k["dur"] for k in side_bwd),
            "side_stream_last_end_us": max((k["ts"] + k["dur"] for k in side_bwd), default=last_compute_end) - t0,
        }
        out["backward"]["exposed_tail_us"] = max(0.0, out["backward"]["side_stream_last_end_us"] - out["backward"]["last_autograd_kernel_end_us"])
        nxt = [k for k in by_stream[compute] if k["phase"] in ("clip", "opt") ]
        if nxt:
            out["backward"]["gap_to_next_phase_kernel_us"] = nxt[0]["ts"] - last_compute_end
        # overlap: how much of the side-stream busy time lies inside the autograd span
        first_compute = min((k["ts"] for k in bwd_compute), default=bwd_end)
        hidden = sum(max(0.0, min(k["ts"] + k["dur"], last_compute_end) - max(k["ts"], first_compute)) for k in side_bwd)
        out["backward"]["side_stream_hidden_us"] = hidden
        out["side_stream_timeline"] = [
            {"k": k["short"], "start_us": round(k["ts"] - t0, 1), "dur_us": round(k["dur"], 1), "after_bwd_end": k["ts"] >= last_compute_end}
            for k in side_bwd]  # fmt: skip
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt2-medium")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-ddp", action="store_true")
    args = ap.parse_args()
    import bench
    from accelerate_b200 import Accelerator, DistributedDataParallelKwargs
    from accelerate_b200.state import PartialState

    acc = Accelerator(mixed_precision="bf16", kwargs_handlers=[DistributedDataParallelKwargs(bucket_cap_mb=25)])
    dev, world, rank = acc.device, acc.num_processes, acc.process_index
    g = torch.Generator().manual_seed(1234 + rank)
    batches = [torch.randint(0, 50257, (args.batch, 512), generator=g).to(dev) for _ in range(4)]
    result = {"n_gpus": world, "model": args.model, "per_gpu_batch": args.batch, "unit": "us (under the profiler: attribution only)"}
    box = {"i": 0}

    model = bench.build_model(args.model).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    model, opt = acc.prepare(model, opt)

    def phases(sync_ctx):
        def fwd():
            box["ctx"] = sync_ctx()
            box["ctx"].__enter__()
            ids = batches[box["i"] % 4]
            box["i"] += 1
            box["loss"] = model(input_ids=ids, labels=ids).loss

        def bwd():
            acc.backward(box["loss"])
            box["ctx"].__exit__(None, None, None)

        return [("fwd", fwd), ("bwd", bwd), ("clip", lambda: acc.clip_grad_norm_(model.parameters(), 1.0)),
                ("opt", opt.step), ("zero", opt.zero_grad)]  # fmt: skip

    def cpu_phase_times(step_phases, steps=12):
        """Host time spent INSIDE each phase call (no synchronisation: pure enqueue cost) next to the GPU time per step,
        without any profiler attached.  If the sum of the host times is close to the step time, the step is
        launch-bound and only host-side savings show up in the throughput."""
        import time

        for _ in range(3):
            for _, fn in step_phases:
                fn()
        torch.cuda.synchronize()
        acc.wait_for_everyone()
        host = {name: 0.0 for name, _ in step_phases}
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        t_all = time.perf_counter()
        for _ in range(steps):
            for name, fn in step_phases:
                t0 = time.perf_counter()
                fn()
                host[name] += time.perf_counter() - t0
        t_all = time.perf_counter() - t_all
        e.record()
        torch.cuda.synchronize()
        out = {name: round(v / steps * 1e3, 3) for name, v in host.items()}
        out["host_total_ms"] = round(t_all / steps * 1e3, 3)
        out["gpu_step_ms"] = round(s.elapsed_time(e) / steps, 3)
        return out

    def run(tag, step_phases):
        acc.wait_for_everyone()
        host = cpu_phase_times(step_phases)
        events = trace_steps(step_phases, 3, args.warmup)
        result[tag] = analyse(events, 1)
        if result[tag] is not None:
            result[tag]["host_enqueue_ms_per_phase (no profiler)"] = host
        acc.wait_for_everyone()

    run("ours", phases(contextlib.nullcontext))
    if world > 1:
        run("ours_no_sync", phases(lambda: acc.no_sync(model)))
        red = model.reducer
        result["reducer"] = {"algo": red.effective_algo(), "buckets": len(red.buckets), "ctas": red.blocks_for(red._pick_algo(next(iter(red.groups.values())).segment)),
                             "bucket_MiB": [round(b.count * b.group.elem_size / 2**20, 1) for b in red.buckets]}  # fmt: skip
    del model, opt
    torch.cuda.empty_cache()

    if not args.no_ddp and (world == 1 or PartialState().backend == "nccl"):
        ref = bench.build_model(args.model).to(dev).train()
        if world > 1:
            ref = torch.nn.parallel.DistributedDataParallel(ref, device_ids=[dev.index], bucket_cap_mb=25)
        ropt = torch.optim.AdamW(ref.parameters(), lr=1e-4, fused=True)

        def r_fwd():
            ids = batches[box["i"] % 4]
            box["i"] += 1
            with torch.autocast("cuda", dtype=torch.bfloat16):
                box["loss"] = ref(input_ids=ids, labels=ids).loss

        run("torch_ddp_nccl_fused_adam", [("fwd", r_fwd), ("bwd", lambda: box["loss"].backward()),
                                          ("clip", lambda: torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)),
                                          ("opt", ropt.step), ("zero", ropt.zero_grad)])  # fmt: skip
    if rank == 0:
        ours = result.get("ours") or {}
        ns = result.get("ours_no_sync") or {}
        if ours.get("backward") and ns.get("backward"):
            a, b = ours["backward"], ns["backward"]
            result["attribution_us"] = {
                "step_span_delta_vs_no_sync": ours["step_span_us"] - ns["step_span_us"],
                "autograd_kernels_slower_by (interference while bucket kernels are resident)": a["autograd_busy_us"] - b["autograd_busy_us"],
                "exposed_tail_after_last_autograd_kernel": a["exposed_tail_us"],
                "side_stream_busy": a["side_stream_busy_us"], "side_stream_hidden_under_autograd": a["side_stream_hidden_us"],
            }  # fmt: skip
        text = json.dumps(result)
        print(text, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(result, f, indent=1)
            with open(os.path.splitext(args.out)[0] + ".txt", "w") as f:
                for tag in ("ours", "ours_no_sync", "torch_ddp_nccl_fused_adam"):
                    r = result.get(tag)
                    if not r:
                        continue
                    f.write(f"== {tag}: step span {r['step_span_us']:.0f} us, streams {r['streams']}\n")
                    for ph, v in r["phases"].items():
                        f.write(f"   {ph:5s} start {v['start_us']:9.0f}  span {v['span_us']:9.0f}  busy {v['busy_us']:9.0f}  kernels {v['kernels']}\n")
                    if "backward" in r:
                        f.write(f"   backward: {json.dumps(r['backward'])}\n")
                    f.write(f"   host enqueue ms/phase (no profiler): {json.dumps(r.get('host_enqueue_ms_per_phase (no profiler)'))}\n")
                    for name, agg in sorted(r["own_kernel_totals_us"].items(), key=lambda kv: -kv[1]["us"]):
                        f.write(f"   own {name:50s} n={agg['n']:3d}  total {agg['us']:9.1f} us  stream {agg['stream']}\n")
                    for k in r.get("side_stream_timeline", []):
                        f.write(f"      side {k['k']:46s} @{k['start_us']:9.1f}  {k['dur_us']:8.1f} us {'TAIL' if k['after_bwd_end'] else ''}\n")
                if "attribution_us" in result:
                    f.write("== attribution: " + json.dumps(result["attribution_us"]) + "\n")
    acc.wait_for_everyone()
    acc.end_training()


if __name__ == "__main__":
    main()
