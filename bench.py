#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric: samples/sec of a GPT-2-medium bf16 data-parallel training
step (forward, accelerator.backward with the fused bucket allreduce, clip_grad_norm_(1.0), AdamW step,
zero_grad) on synthetic 512-token batches, at N = 1/2/4/8 B200s of one node (weak scaling: fixed
per-GPU micro-batch).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference      # the same step on the host CPU cores (oracle port of the path)

One JSON line on rank 0.  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = the same
through the public API with pinned host batches (H2D inside the timed region) and a D2H loss read per
step; `roofline` = the dominant kernel of OUR path timed with CUDA events; `cpu_baseline` = the path on
the host cores (bounded sample); `torch_ddp_nccl` = stock torch DDP+NCCL+clip_grad_norm_ (what the
reference delegates to, accelerator.py:1892/3007) on the same box in the same run.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ = 512
VOCAB = 50257
GRAD_BYTES = 354_823_168 * 4  # GPT-2-medium fp32 gradients


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("BENCH_BATCH", "8")), help="per-GPU micro-batch")
    ap.add_argument("--model", default=os.environ.get("BENCH_MODEL", "gpt2-medium"))
    ap.add_argument("--no-ddp-leg", action="store_true", help="skip the stock torch DDP+NCCL comparison leg")
    ap.add_argument("--no-cpu-leg", action="store_true")
    ap.add_argument("--defer-clip", action="store_true", help="A/B only: let the fused AdamW kernel apply the clip multiplier")
    ap.add_argument("--torch-adam", action="store_true", help="A/B only: B200DP_FUSED_ADAM=0 (torch's foreach AdamW)")
    return ap.parse_args()


def model_config(name):
    from transformers import GPT2Config

    if name == "gpt2-medium":
        return GPT2Config(n_embd=1024, n_layer=24, n_head=16)
    if name == "gpt2-small":
        return GPT2Config()
    if name == "gpt2-tiny":  # plumbing checks only; never a reported number
        return GPT2Config(n_embd=128, n_layer=2, n_head=4)
    raise ValueError(name)


def build_model(name, seed=42):
    import torch
    from transformers import GPT2LMHeadModel

    torch.manual_seed(seed)
    return GPT2LMHeadModel(model_config(name))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}  # fmt: skip


# ------------------------------------------------------------------------------------------------
# CPU arm: the path on the host cores (oracle port), also used as cpu_baseline
# ------------------------------------------------------------------------------------------------
def cpu_leg(model_name, steps, warmup, budget_s=25.0, micro_batch=1):
    """Training step of the same model on the host CPU: torch CPU bf16-autocast forward/backward, then
    the ORACLE's restatement of the path (world size 1: clip by global norm), then AdamW.  Each step is a
    bounded sample of the workload: `micro_batch` sequences of 512 tokens."""
    import numpy as np
    import psutil
    import torch

    from oracle import c_oracle

    cores = psutil.cpu_count(logical=False) or os.cpu_count()
    os.environ.setdefault("ORACLE_THREADS", str(cores))
    torch.set_num_threads(cores)
    model = build_model(model_name)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1234)
    times = []
    t_start = time.perf_counter()
    for step in range(warmup + steps):
        ids = torch.randint(0, VOCAB, (micro_batch, SEQ), generator=g)
        t0 = time.perf_counter()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss = model(input_ids=ids, labels=ids).loss
        loss.backward()
        params = [p for p in model.parameters() if p.grad is not None]
        c_oracle.clip_grad_norm_inplace([p.grad.numpy() for p in params], 1.0)  # in place on the grads' memory
        opt.step()
        opt.zero_grad()
        dt = time.perf_counter() - t0
        if step >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    mean = sum(times) / len(times)
    return {"value": micro_batch / mean, "unit": "samples/s", "cores": cores, "kind": "port",
            "n_params": sum(p.numel() for p in model.parameters()),
            "sample": f"{len(times)} steps x {micro_batch} sequence(s) of {SEQ} tokens, {model_name}, torch CPU bf16 autocast "
                      f"fwd/bwd + oracle (C, pthreads) clip_grad_norm + AdamW; {mean * 1e3:.0f} ms/step"}  # fmt: skip


def workload_config(model_name, n_params, per_gpu_batch, world):
    """The `config` both arms report: BASELINE.json's headline workload at this GPU count."""
    return {
        "workload": f"{model_name} ({n_params} params) bf16-autocast train step: fwd + bwd + gradient bucket allreduce "
                    "+ clip_grad_norm_(1.0) + AdamW + zero_grad",
        "global_batch": per_gpu_batch * world, "per_gpu_batch": per_gpu_batch, "seq_len": SEQ, "parallelism": f"dp{world}",
        "bucket_cap_mb": 25,
    }  # fmt: skip


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    leg = cpu_leg(args.model, max(1, args.steps), max(0, min(args.warmup, 2)), budget_s=150.0)
    n_params = leg.pop("n_params")
    line = {
        "impl": "reference", "metric": "samples_per_sec", "value": leg["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / leg["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        # the GPU arm's workload and configuration; every CPU step is a bounded sample of it (`cpu_baseline.sample`)
        "config": dict(workload_config(args.model, n_params, args.batch, args.gpus),
                       cpu_path="torch CPU bf16 autocast fwd/bwd + oracle (C, pthreads) clip_grad_norm + AdamW on the host cores",
                       per_step_batch=1),
        "cpu_baseline": leg,
        "e2e": {"value": leg["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, transformers notices) write to fd 1; the contract is ONE JSON line
    on stdout, so everything else is routed to stderr and the JSON goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch

    from accelerate_b200 import Accelerator, DistributedDataParallelKwargs
    from accelerate_b200.state import PartialState

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    if args.defer_clip:
        os.environ["B200DP_DEFER_CLIP"] = "1"
    if args.torch_adam:
        os.environ["B200DP_FUSED_ADAM"] = "0"
    acc = Accelerator(mixed_precision="bf16", kwargs_handlers=[DistributedDataParallelKwargs(bucket_cap_mb=25)])
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE is {world}: launch with torchrun --nproc-per-node {args.gpus}"
    rt = PartialState().runtime
    B = args.batch
    steps, warmup = args.steps, max(args.warmup, 3)

    model = build_model(args.model).to(dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    model, opt = acc.prepare(model, opt)
    n_params = sum(p.numel() for p in model.parameters())

    g = torch.Generator().manual_seed(1234 + rank)
    n_batches = 8
    host_batches = [torch.randint(0, VOCAB, (B, SEQ), generator=g).pin_memory() for _ in range(n_batches)]
    dev_batches = [b.to(dev) for b in host_batches]

    def step_resident(i):
        ids = dev_batches[i % n_batches]
        loss = model(input_ids=ids, labels=ids).loss
        acc.backward(loss)
        acc.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        return loss

    def step_e2e(i):
        ids = host_batches[i % n_batches].to(dev, non_blocking=True)  # H2D from pinned memory, every step
        loss = model(input_ids=ids, labels=ids).loss
        acc.backward(loss)
        acc.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        return loss.item()  # D2H read of the step's result

    def timed(fn, count):
        acc.wait_for_everyone()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = rt.ctx.launches
        start.record()
        last = None
        for i in range(count):
            last = fn(i)
        end.record()
        torch.cuda.synchronize()
        acc.wait_for_everyone()
        ms = torch.tensor([start.elapsed_time(end)], device=dev)
        if world > 1:
            ms = acc.gather(ms).max().reshape(1)  # max over ranks, through our own allgather kernel
        return ms.item(), rt.ctx.launches - launches0, last

    for i in range(warmup):
        step_resident(i)
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    total_ms, launches, last_loss = timed(step_resident, steps)
    clocks = sampler.stop() if rank == 0 else None
    for i in range(2):
        step_e2e(i)
    e2e_ms, _, _ = timed(step_e2e, steps)
    rt.ctx.check_status()
    clip_mode = acc.last_clip_mode
    plan = getattr(opt, "_fused_plan", None)
    if plan is None:
        optimizer_step = "torch foreach AdamW"
    else:
        optimizer_step = "b200dp_multi_adam, one launch per group, " + ("fusion-buffer gradients" if plan.reducer is not None else "loose gradients")
        if args.defer_clip and plan.reducer is not None:
            optimizer_step += " + deferred clip"

    # ---- roofline of the dominant kernel of OUR path, CUDA events on the launching stream ----
    peaks, peak_src = measured_peaks()
    roofline, extra = kernel_roofline(acc, model, rt, world, peaks, peak_src)

    # ---- stock torch DDP + NCCL + torch clip_grad_norm_ on the same box (what the reference delegates to) ----
    # ---- the fused optimizer step (K9), timed alone; rank-local, never allowed to take the result down ----
    try:
        adam_probe = adam_kernel_probe(opt, rt, peaks, peak_src)
    except Exception as err:
        adam_probe = {"kernel": "multi_adam_kernel", "error": repr(err)}

    ddp_leg = None
    if not args.no_ddp_leg and (world == 1 or PartialState().backend == "nccl"):
        del opt
        torch.cuda.empty_cache()
        ddp_leg = torch_ddp_leg(args, dev, world, rank, dev_batches, steps, warmup, acc)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_leg:
        try:
            cpu = cpu_leg(args.model, steps=4, warmup=1, budget_s=25.0)
            cpu.pop("n_params", None)
        except Exception as err:  # the CPU leg must never take the GPU result down
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {err}"}

    if rank == 0:
        gb = B * world
        line = {
            "metric": "samples_per_sec", "value": gb * steps / (total_ms / 1e3), "unit": "samples/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(
                workload_config(args.model, n_params, B, world),
                l2="per-step working set (1.4 GB grads + weights + activations) exceeds the 126 MB L2; no flush needed",
                clip_mode=clip_mode, final_loss=float(last_loss.detach()), optimizer_step=optimizer_step,
            ),
            "e2e": {"value": gb * steps / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": B * SEQ * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
        }  # fmt: skip
        line.update(extra)
        if adam_probe is not None:
            line["roofline_adam_kernel"] = adam_probe
        if ddp_leg is not None:
            line["torch_ddp_nccl"] = ddp_leg
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    acc.wait_for_everyone()
    PartialState().destroy_process_group()


def adam_kernel_probe(opt, rt, peaks, peak_src):
    """CUDA-event time of one fused Adam/AdamW step (all parameter groups) on the optimizer's own tables and moment
    buffers, after the measured steps.  Launched with lr = 0 (neg_step_size 0, no decay), so parameters keep their
    values; algorithmic bytes = 28 per element (read p, g, m, v; write p, m, v)."""
    import torch

    plan = getattr(opt, "_fused_plan", None)
    if plan is None:
        return None
    st = torch.cuda.current_stream().cuda_stream
    groups = [(params, table) for params, table in plan.tables if params]
    elems = sum(p.numel() for params, _ in groups for p in params)
    if plan.reducer is None:  # loose layout: gradients are separate tensors, freed by zero_grad -> scratch ones
        scratch = [[torch.zeros_like(p) for p in params] for params, _ in groups]
        gptrs = [[g_.data_ptr() for g_ in gs] for gs in scratch]
    else:
        gptrs = [None] * len(groups)

    def launch():
        for (_, table), g_ in zip(groups, gptrs):
            rt.ctx.multi_adam(table, plan.m_flat.data_ptr(), plan.v_flat.data_ptr(), 1.0, 0.0, 0.9, 0.999, 0.0, 1.0, 1e-8,
                              None, None, st, grads=g_)  # fmt: skip

    reps = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
    for _ in range(3):
        launch()
    for i in range(reps):
        ev[2 * i].record()
        launch()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps)) / reps
    nbytes = 28 * elems
    peak = float(peaks["hbm_gbs"])
    achieved = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "multi_adam_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})", "traffic": None,
            "bytes_per_step": nbytes, "ms_per_step": ms, "launches_per_step": len(groups),
            "gradients": "fusion buffer" if plan.reducer is not None else "loose tensors"}  # fmt: skip


def kernel_roofline(acc, model, rt, world, peaks, peak_src):
    """Time the dominant kernel of our path in place, on the stream it is launched on.

    N=1: the path is clip_grad_norm_ = multi_stats (reads S) + finalize + multi_scale (reads+writes S);
         the scale kernel dominates: algorithmic bytes 2*S per launch, HBM-bound.
    N>1: the fused bucket allreduce dominates: per 25 MiB bucket each GPU must move 2(N-1)/N * bytes over
         NVLink (nccl-tests bus-bandwidth convention), NVLink-bound (770 GB/s measured peer copy, 900 nominal).
    """
    import torch

    from accelerate_b200 import native
    from accelerate_b200.runtime import native_dtype

    dev = acc.device
    st = torch.cuda.current_stream().cuda_stream
    extra = {}
    reps = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
    if world == 1:
        grads = [torch.randn_like(p) for p in model.parameters()]
        ptrs, numels = [g_.data_ptr() for g_ in grads], [g_.numel() for g_ in grads]
        nbytes = sum(numels) * 4
        mult = torch.tensor([0.999], device=dev)
        slot = rt.reserve_slots(2)
        for _ in range(3):
            rt.ctx.multi_scale(ptrs, numels, native.F32, mult.data_ptr(), st)
        for i in range(reps):
            ev[2 * i].record()
            rt.ctx.multi_scale(ptrs, numels, native.F32, mult.data_ptr(), st)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        t_scale = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps)) / reps
        for i in range(reps):
            ev[2 * i].record()
            rt.ctx.multi_stats(ptrs, numels, native.F32, slot, st)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        t_stats = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps)) / reps
        achieved = 2 * nbytes / (t_scale * 1e-3) / 1e9
        peak = float(peaks["hbm_gbs"])
        roof = {"kernel": "multi_scale_kernel<float>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one `ncu --set full`
                # capture (profiles/r01_n1_clip_kernels_ncu.txt): 1.419 GB + 1.360 GB -> no re-reads
                "traffic": 2_779_191_000 if nbytes == GRAD_BYTES else None,
                "traffic_source": "profiles/r01_n1_clip_kernels_ncu.txt",
                "bytes_per_launch": 2 * nbytes, "ms_per_launch": t_scale}  # fmt: skip
        extra["roofline_stats_kernel"] = {"kernel": "multi_stats_kernel<float>", "bound": "hbm",
                                          "achieved": nbytes / (t_stats * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                          "frac": nbytes / (t_stats * 1e-3) / 1e9 / peak, "bytes_per_launch": nbytes,
                                          "ms_per_launch": t_stats}  # fmt: skip
        return roof, extra
    # N > 1: one 25 MiB bucket of the real fusion buffer, in place, all ranks in lock-step
    red = model.reducer
    group = next(iter(red.groups.values()))
    count = min(25 * 1024 * 1024 // 4, group.total_elems // 4 * 4)
    algo = red._pick_algo(group.segment)
    sizes = {}
    blocks = red.blocks_for(algo)
    flags_iso = native.F_STATS
    flags_seq = native.F_STATS | native.F_NO_END_BARRIER
    K = 8  # buckets back to back, the way a backward pass issues them (closing barrier replaced by ONE fence)

    def isolated(cnt):
        rt.ctx.allreduce_bucket(group.segment.seg, 0, cnt, native.F32, 1.0 / world, red.slot_base, flags_iso, algo,
                                native.CH_GRAD, blocks, native.NO_STREAM, st)  # fmt: skip

    def sequence(cnt):
        for j in range(K):
            rt.ctx.allreduce_bucket(group.segment.seg, j * cnt, cnt, native.F32, 1.0 / world, red.slot_base + j % len(red.buckets),
                                    flags_seq, algo, native.CH_GRAD, blocks, native.NO_STREAM, st)  # fmt: skip
        rt.ctx.fence(native.CH_GRAD, st)

    def clock(fn, arg, per=1):
        acc.wait_for_everyone()
        for _ in range(3):
            fn(arg)
        for i in range(reps):
            ev[2 * i].record()
            fn(arg)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        t = torch.tensor([sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps)) / reps / per], device=dev)
        return acc.gather(t).max().item()

    def entry(nbytes, ms):
        return {"bytes": nbytes, "ms": ms, "busbw_GBps": nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9,
                "algbw_GBps": nbytes / (ms * 1e-3) / 1e9}  # fmt: skip

    assert K * count <= group.total_elems
    sizes["bucket_25MiB_steady"] = entry(count * 4, clock(sequence, count, per=K))
    sizes["bucket_25MiB_isolated"] = entry(count * 4, clock(isolated, count))
    whole = group.total_elems // 4 * 4
    sizes["all_grads"] = entry(whole * 4, clock(isolated, whole))
    sizes["ctas"] = blocks
    # bucket pack (what replaces DDP's per-parameter copy-in): TMA bulk vs LDG/STG, on a mid-size real bucket
    mid = sorted(red.buckets, key=lambda bk: bk.count)[len(red.buckets) // 2]
    fresh = [torch.randn_like(p) for p in mid.params]
    pack = {"bucket_bytes": sum(mid.nbytes), "tensors": len(mid.params)}
    for label, bulk in (("bulk_tma", 1), ("ldg_stg", 0)):
        srcs, dsts = [f.data_ptr() for f in fresh], list(mid.ptrs)
        for _ in range(3):
            rt.ctx.pack(srcs, dsts, mid.nbytes, bulk, st)
        for i in range(reps):
            ev[2 * i].record()
            rt.ctx.pack(srcs, dsts, mid.nbytes, bulk, st)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        t = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps)) / reps
        pack[label] = {"ms": t, "GBps": 2 * sum(mid.nbytes) / (t * 1e-3) / 1e9}
    extra["pack_kernel"] = pack
    group.flat.zero_()
    b = sizes["bucket_25MiB_steady"]
    roof = {"kernel": f"allreduce_inplace_kernel<float,{'NVLS' if algo == native.ALGO_NVLS else 'twoshot'}>", "bound": "nvlink",
            "achieved": b["busbw_GBps"], "peak": 770.0, "unit": "GB/s", "frac": b["busbw_GBps"] / 770.0,
            "peak_source": "B200_PROFILING.md measured peer copy 770 GB/s/dir (900 nominal)", "traffic": None,
            "bytes_per_launch": b["bytes"], "ms_per_launch": b["ms"], "frac_of_900_nominal": b["busbw_GBps"] / 900.0,
            "how": f"{K} x 25 MiB buckets back to back on one stream exactly as a backward pass issues them (statistics on, "
                   f"closing barrier replaced by one fence), {blocks} CTAs; time per bucket, max over ranks; bus bandwidth = "
                   "S*2(N-1)/N/t (nccl-tests convention; NVLS moves ~S per GPU so it can exceed the link rate)"}  # fmt: skip
    extra["allreduce_sizes"] = sizes
    try:  # HBM-side kernel of the fused clip at N > 1: one multiply pass over the fusion buffer (2*S bytes)
        mult = torch.tensor([0.999], device=dev)
        n_el = group.total_elems
        for _ in range(3):
            rt.ctx.scale_flat(group.flat.data_ptr(), n_el, native.F32, mult.data_ptr(), 0, st)
        for i in range(reps):
            ev[2 * i].record()
            rt.ctx.scale_flat(group.flat.data_ptr(), n_el, native.F32, mult.data_ptr(), 0, st)
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        t_scale = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps)) / reps
        hbm = 2 * n_el * 4 / (t_scale * 1e-3) / 1e9
        extra["roofline_hbm"] = {"kernel": "scale_flat_kernel<float>", "bound": "hbm", "achieved": hbm, "peak": float(peaks["hbm_gbs"]),
                                 "unit": "GB/s", "frac": hbm / float(peaks["hbm_gbs"]), "bytes_per_launch": 2 * n_el * 4,
                                 "ms_per_launch": t_scale, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})"}  # fmt: skip
    except Exception as err:  # never let an auxiliary measurement take the headline down
        extra["roofline_hbm"] = {"error": str(err)}
    return roof, extra


def torch_ddp_leg(args, dev, world, rank, dev_batches, steps, warmup, acc):
    """The stack the reference delegates to, stock: torch DDP (bucket_cap_mb=25) + NCCL, autocast bf16,
    torch.nn.utils.clip_grad_norm_, AdamW.  Same model, seeds, batches and step count."""
    import torch

    model = build_model(args.model).to(dev)
    model.train()
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], bucket_cap_mb=25)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)

    def step(i):
        ids = dev_batches[i % len(dev_batches)]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(input_ids=ids, labels=ids).loss
        loss.float().backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()

    for i in range(max(warmup, 10 if world > 1 else warmup)):  # covers DDP's bucket rebuild
        step(i)
    acc.wait_for_everyone()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        step(i)
    end.record()
    torch.cuda.synchronize()
    acc.wait_for_everyone()
    ms = torch.tensor([start.elapsed_time(end)], device=dev)
    if world > 1:
        ms = acc.gather(ms).max().reshape(1)
    ms = ms.item()
    del model, opt
    torch.cuda.empty_cache()
    return {"value": args.batch * world * steps / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms / steps,
            "what": "torch DistributedDataParallel + NCCL + torch clip_grad_norm_ + AdamW (same model/batches/steps)"}  # fmt: skip


if __name__ == "__main__":
    main()
