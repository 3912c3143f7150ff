#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric: samples/sec of a GPT-2-medium bf16 data-parallel training
step (forward, accelerator.backward with the fused bucket allreduce, clip_grad_norm_(1.0), AdamW step,
zero_grad) on synthetic 512-token batches, at N = 1/2/4/8 B200s of one node (weak scaling: fixed
per-GPU micro-batch).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference            # the reference's own CPU path on the host cores
    ... bench.py --model bert-large --accum 4   # BASELINE.json config 4 (accumulate + clip on sync steps)
    ... bench.py --microbench                   # BASELINE.json config 5 (reduce / gather 1 KiB - 1 GiB vs NCCL)

One JSON line on rank 0.  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = the same
through the public API with pinned host batches (H2D inside the timed region) and a D2H loss read per
step; `parity` (N > 1) = gradients of one sync backward through our reducer vs stock torch DDP+NCCL on the
same inputs (rtol 1e-3 / atol 1e-5), vs the rank-order sum of the local gradients, and bit-identity across
ranks, computed before the timed region and asserted; `roofline` = the dominant kernel of OUR path timed with
CUDA events (NVLink byte counters from NVML as `traffic` at N > 1); `cpu_baseline` = the reference on the host
cores (bounded sample); `torch_ddp_nccl*` / `reference_accelerate_ddp` = what the reference delegates to
(accelerator.py:1892/3007) on the same box in the same run.
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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")  # `pip install --target baseline/_ref /root/reference`

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
    ap.add_argument("--accum", type=int, default=1, help="gradient_accumulation_steps (config 4: --model bert-large --accum 4)")
    ap.add_argument("--microbench", action="store_true", help="config 5: accelerator.reduce / gather 1 KiB - 1 GiB vs NCCL")
    ap.add_argument("--max-log2", type=int, default=30, help="--microbench: largest message, log2 bytes")
    ap.add_argument("--no-ddp-leg", action="store_true", help="skip the stock torch DDP+NCCL / reference-Accelerate legs")
    ap.add_argument("--no-cpu-leg", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the N>1 gradient parity block")
    ap.add_argument("--no-defer-clip", action="store_true", help="A/B only: B200DP_DEFER_CLIP=0 (multiply pass inside clip_grad_norm_)")
    ap.add_argument("--torch-adam", action="store_true", help="A/B only: B200DP_FUSED_ADAM=0 (torch's foreach AdamW)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# models and batches
# ------------------------------------------------------------------------------------------------
def model_config(name, dropout=None):
    from transformers import BertConfig, GPT2Config

    drop = {} if dropout is None else {"resid_pdrop": dropout, "embd_pdrop": dropout, "attn_pdrop": dropout}
    if name == "gpt2-medium":
        return GPT2Config(n_embd=1024, n_layer=24, n_head=16, **drop)
    if name == "gpt2-small":
        return GPT2Config(**drop)
    if name == "gpt2-tiny":  # plumbing checks only; never a reported number
        return GPT2Config(n_embd=128, n_layer=2, n_head=4, **drop)
    bdrop = {} if dropout is None else {"hidden_dropout_prob": dropout, "attention_probs_dropout_prob": dropout}
    if name == "bert-large":  # BASELINE.json config 4: BERT-large seq-cls, 335,143,938 parameters
        return BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, num_labels=2, **bdrop)
    if name == "bert-tiny":
        return BertConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256, num_labels=2, **bdrop)
    raise ValueError(name)


def build_model(name, seed=42, dropout=None):
    import torch
    from transformers import BertForSequenceClassification, GPT2LMHeadModel

    torch.manual_seed(seed)
    cfg = model_config(name, dropout)
    return (BertForSequenceClassification if name.startswith("bert") else GPT2LMHeadModel)(cfg)


def vocab_of(name):
    return 30522 if name.startswith("bert") else VOCAB


def forward_loss(model, ids, name):
    if name.startswith("bert"):
        return model(input_ids=ids, labels=ids[:, 0] % 2).loss
    return model(input_ids=ids, labels=ids).loss


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


class NvlinkCounters:
    """NVLink payload byte counters of this process's GPU, read from NVML (the hardware counters `nvidia-smi nvlink
    -gt d` prints): NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX summed over all links (scope UINT_MAX), KiB.  ncu cannot
    profile a flag-synchronised multi-rank kernel (it replays each launch in isolation), so this is the NVLink-side
    counter evidence next to the CUDA-event bus bandwidth: bytes that actually crossed this GPU's links."""

    TX, RX = 138, 139  # NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX

    def __init__(self, device):
        self.ok, self.why = False, ""
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.ok = self.read() is not None
            if not self.ok:
                self.why = "NVML returned no NVLink throughput field values"
        except Exception as err:  # noqa: BLE001 - an absent counter must never take the bench down
            self.why = repr(err)

    def read(self):
        try:
            vals = self.nv.nvmlDeviceGetFieldValues(self.h, [(self.TX, 0xFFFFFFFF), (self.RX, 0xFFFFFFFF)])
            out = []
            for v in vals:
                if v.nvmlReturn != 0:
                    return None
                out.append(int(v.value.ullVal) * 1024)
            return out  # [tx_bytes, rx_bytes]
        except Exception:
            return None


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path on the host cores (baseline/_ref), else the oracle port; also cpu_baseline
# ------------------------------------------------------------------------------------------------
def cpu_leg(model_name, steps, warmup, budget_s=25.0, micro_batch=1):
    """One training step of the same model on the host CPU, a bounded sample of the workload (`micro_batch`
    sequences of 512 tokens per step).

    kind "reference": the UNMODIFIED reference (huggingface/accelerate installed into baseline/_ref) through its own
    public API — ``Accelerator(cpu=True, mixed_precision="bf16")``, ``prepare``, ``backward``, ``clip_grad_norm_``,
    AdamW — i.e. exactly the calls the GPU arm makes on this package.  kind "port" (baseline/_ref not importable):
    torch CPU bf16-autocast forward/backward, the ORACLE's restatement of the clip, AdamW."""
    import psutil
    import torch

    cores = psutil.cpu_count(logical=False) or os.cpu_count()
    os.environ.setdefault("ORACLE_THREADS", str(cores))
    torch.set_num_threads(cores)
    model = build_model(model_name)
    model.train()
    n_params = sum(p.numel() for p in model.parameters())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    kind, ref_acc, why_port = "port", None, ""
    # this leg is ONE process on the host cores: a torchrun environment (RANK/WORLD_SIZE of the GPU job) must not make
    # the reference wait for peers
    scrubbed = {k: os.environ.pop(k) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK",
                                                "ROLE_RANK", "ROLE_WORLD_SIZE") if k in os.environ}  # fmt: skip
    if os.path.isdir(os.path.join(REF_DIR, "accelerate")) and os.environ.get("BENCH_CPU_PORT", "0") != "1":
        try:
            sys.path.insert(0, REF_DIR)
            import accelerate as ref_accelerate

            # the GPU comparison leg may have initialised the reference's singletons for CUDA in this process
            ref_accelerate.state.AcceleratorState._reset_state(reset_partial_state=True)
            ref_acc = ref_accelerate.Accelerator(cpu=True, mixed_precision="bf16")
            model, opt = ref_acc.prepare(model, opt)
            kind = "reference"
        except Exception as err:  # noqa: BLE001
            why_port = f" (reference import failed: {err!r})"
            ref_acc = None
        finally:
            if REF_DIR in sys.path:
                sys.path.remove(REF_DIR)
    os.environ.update(scrubbed)
    if ref_acc is None:
        from oracle import c_oracle
    g = torch.Generator().manual_seed(1234)
    times = []
    t_start = time.perf_counter()
    for step in range(warmup + steps):
        ids = torch.randint(0, vocab_of(model_name), (micro_batch, SEQ), generator=g)
        t0 = time.perf_counter()
        if ref_acc is not None:
            loss = forward_loss(model, ids, model_name)
            ref_acc.backward(loss)
            ref_acc.clip_grad_norm_(model.parameters(), 1.0)
        else:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                loss = forward_loss(model, ids, model_name)
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
    path = ("huggingface/accelerate 1.14.0.dev0 (baseline/_ref), Accelerator(cpu=True, mixed_precision='bf16'): prepare + "
            "backward + clip_grad_norm_ + torch AdamW" if kind == "reference" else
            "torch CPU bf16 autocast fwd/bwd + oracle (C, pthreads) clip_grad_norm + AdamW" + why_port)  # fmt: skip
    return {"value": micro_batch / mean, "unit": "samples/s", "cores": cores, "kind": kind, "n_params": n_params, "path": path,
            "sample": f"{len(times)} steps x {micro_batch} sequence(s) of {SEQ} tokens, {model_name}; {mean * 1e3:.0f} ms/step"}  # fmt: skip


def workload_config(model_name, n_params, per_gpu_batch, world, accum=1):
    """The `config` both arms report: BASELINE.json's workload at this GPU count."""
    cfg = {
        "workload": f"{model_name} ({n_params} params) bf16-autocast train step: fwd + bwd + gradient bucket allreduce "
                    "+ clip_grad_norm_(1.0) + AdamW + zero_grad",
        "global_batch": per_gpu_batch * world * accum, "per_gpu_batch": per_gpu_batch, "seq_len": SEQ, "parallelism": f"dp{world}",
        "bucket_cap_mb": 25,
    }  # fmt: skip
    if accum > 1:
        cfg["gradient_accumulation_steps"] = accum
    return cfg


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
        "config": dict(workload_config(args.model, n_params, args.batch, args.gpus, args.accum), cpu_path=leg["path"], per_step_batch=1),
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
    if args.no_defer_clip:
        os.environ["B200DP_DEFER_CLIP"] = "0"
    if args.torch_adam:
        os.environ["B200DP_FUSED_ADAM"] = "0"
    acc = Accelerator(mixed_precision="bf16", gradient_accumulation_steps=args.accum,
                      kwargs_handlers=[DistributedDataParallelKwargs(bucket_cap_mb=25)])  # fmt: skip
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE is {world}: launch with torchrun --nproc-per-node {args.gpus}"
    rt = PartialState().runtime
    if args.microbench:
        microbench(args, acc, rt, rank, world, dev)
        acc.wait_for_everyone()
        PartialState().destroy_process_group()
        return
    B, A = args.batch, args.accum
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- gradient parity on THIS configuration, before anything is timed ----
    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_block(args, acc, rt, rank, world, dev)

    model = build_model(args.model).to(dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    model, opt = acc.prepare(model, opt)
    n_params = sum(p.numel() for p in model.parameters())

    g = torch.Generator().manual_seed(1234 + rank)
    n_batches = 8
    host_batches = [torch.randint(0, vocab_of(args.model), (B, SEQ), generator=g).pin_memory() for _ in range(n_batches)]
    dev_batches = [b.to(dev) for b in host_batches]

    def micro(ids):
        with acc.accumulate(model):
            loss = forward_loss(model, ids, args.model)
            acc.backward(loss)
            if acc.sync_gradients:
                acc.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            opt.zero_grad()
        return loss

    def step_resident(i):
        for a in range(A):
            loss = micro(dev_batches[(i * A + a) % n_batches])
        return loss

    def step_e2e(i):
        for a in range(A):
            ids = host_batches[(i * A + a) % n_batches].to(dev, non_blocking=True)  # H2D from pinned memory, every micro-step
            loss = micro(ids)
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
    from accelerate_b200 import lazy_grad

    if plan is None:
        optimizer_step = "torch foreach AdamW"
    else:
        optimizer_step = "b200dp_multi_adam, one launch per group, " + ("fusion-buffer gradients" if plan.reducer is not None else "loose gradients")
        if lazy_grad.enabled():
            optimizer_step += " + clip coefficient applied on the fly (deferred clip, lazy_grad.py)"

    # ---- rooflines of OUR kernels, CUDA events on the launching stream ----
    peaks, peak_src = measured_peaks()
    roofline, extra = kernel_roofline(acc, model, rt, world, peaks, peak_src)
    try:  # the fused optimizer step (K9), timed alone; rank-local, never allowed to take the result down
        adam_probe = adam_kernel_probe(opt, rt, peaks, peak_src, n_params)
    except Exception as err:  # noqa: BLE001
        adam_probe = {"kernel": "multi_adam_kernel", "error": repr(err)}
    if world == 1 and adam_probe and "frac" in adam_probe:
        # the longest own kernel of the single-GPU step is the fused optimizer step: it is THE roofline line there
        extra["roofline_clip_scale_kernel"] = roofline
        roofline = adam_probe
    elif adam_probe is not None:
        extra["roofline_adam_kernel"] = adam_probe

    # ---- what the reference delegates to, stock, on the same box in the same run ----
    legs = {}
    if not args.no_ddp_leg and (world == 1 or PartialState().backend == "nccl"):
        del opt
        model = None
        torch.cuda.empty_cache()
        legs["torch_ddp_nccl"] = torch_ddp_leg(args, dev, world, dev_batches, steps, warmup, acc, fused_adam=False)
        legs["torch_ddp_nccl_fused_adam"] = torch_ddp_leg(args, dev, world, dev_batches, steps, warmup, acc, fused_adam=True)
        legs["reference_accelerate_ddp"] = reference_accelerate_leg(args, dev, world, dev_batches, steps, warmup, acc)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_leg:
        try:
            cpu = cpu_leg(args.model, steps=4, warmup=1, budget_s=25.0)
            cpu.pop("n_params", None)
        except Exception as err:  # the CPU leg must never take the GPU result down
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {err}"}

    if rank == 0:
        gb = B * A * world
        line = {
            "metric": "samples_per_sec", "value": gb * steps / (total_ms / 1e3), "unit": "samples/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(
                workload_config(args.model, n_params, B, world, A),
                l2="per-step working set (1.4 GB grads + weights + activations) exceeds the 126 MB L2; no flush needed",
                clip_mode=clip_mode, final_loss=float(last_loss.detach()), optimizer_step=optimizer_step,
            ),
            "e2e": {"value": gb * steps / (e2e_ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": B * A * SEQ * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
        }  # fmt: skip
        line.update(extra)
        line.update(legs)
        if parity is not None:
            line["parity"] = parity
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    acc.wait_for_everyone()
    PartialState().destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit(f"gradient parity FAILED: {json.dumps(parity)}")


# ------------------------------------------------------------------------------------------------
# parity block (N > 1)
# ------------------------------------------------------------------------------------------------
def parity_block(args, acc, rt, rank, world, dev):
    """Gradients of ONE sync backward on identical per-rank batches (dropout 0, same initial weights) through

    * ``B200DataParallel`` (``accelerator.prepare``, the kernel that carries the timed number — NVLS from 4 ranks up), and
    * stock ``torch.nn.parallel.DistributedDataParallel(bucket_cap_mb=25)`` + NCCL — the call the reference makes at
      ``src/accelerate/accelerator.py:1892-1894``,

    compared at the north star's tolerance (rtol 1e-3, atol 1e-5).  Then the exactness checks no second library can
    give: a ``no_sync`` backward leaves the LOCAL gradients L_r in ``.grad``, a sync backward of ``0 * loss`` adds exact
    zeros and reduces them, so the result must equal ``sum_r fl(L_r * 1/N)`` in rank order — bit for bit for the
    two-shot kernel, within float32 rounding of it for NVLS (the switch adds in its own fixed order) — with the L_r
    gathered by NCCL (independent transport) and the sum formed by separate torch mul / add kernels (no FMA).  Finally
    a checksum of the whole fusion buffer must be identical on every rank (gathered through ``accelerator.gather``)."""
    import torch
    import torch.distributed as dist

    from accelerate_b200.state import PartialState

    t0 = time.perf_counter()
    nccl = PartialState().backend == "nccl"
    name = args.model
    ids = torch.randint(0, vocab_of(name), (args.batch, SEQ), generator=torch.Generator().manual_seed(4321 + rank)).to(dev)
    ours = acc.prepare(build_model(name, dropout=0.0).to(dev))
    ours.train()
    red = ours.reducer
    algo = red.effective_algo()
    params = list(ours.module.parameters())
    result = {"config": f"{name}, bf16 autocast, {world} ranks, per-rank batch {args.batch} x {SEQ}, dropout 0, bucket_cap_mb 25",
              "algo": algo, "buckets": len(red.buckets), "rtol": 1e-3, "atol": 1e-5}  # fmt: skip

    # ---- (1) one ordinary sync backward through the public API ----
    acc.backward(forward_loss(ours, ids, name) * float(args.accum))  # backward() divides by the accumulation steps
    torch.cuda.synchronize()
    r1 = [p.grad.detach().clone() for p in params]
    checksum = torch.zeros(2, dtype=torch.int64, device=dev)
    weights = (torch.arange(1 << 20, device=dev, dtype=torch.int64) % 251) + 1
    for group in red.groups.values():
        words = group.flat.view(torch.int32 if group.elem_size == 4 else torch.int16)
        for c, chunk in enumerate(words.split(1 << 20)):
            w = chunk.to(torch.int64)
            checksum[0] += w.sum()
            checksum[1] += (w * weights[: w.numel()]).sum() * (c + 1)
    everyone = acc.gather(checksum.reshape(1, 2))
    result["all_ranks_bit_identical"] = bool((everyone == everyone[0]).all().item())
    for p in params:
        p.grad = None

    # ---- (2) exactness: local gradients, then a sync backward that adds exact zeros ----
    with ours.no_sync():
        acc.backward(forward_loss(ours, ids, name) * float(args.accum))
    local = [p.grad.detach().clone() for p in params]
    acc.backward(forward_loss(ours, ids, name) * 0.0)
    torch.cuda.synchronize()
    inv = 1.0 / world
    bad_exact = 0
    max_abs = max_rel = 0.0
    viol = 0
    for p, mine in zip(params, local):
        flat = mine.reshape(-1)
        allr = torch.empty(world * flat.numel(), dtype=flat.dtype, device=dev)
        if nccl:
            dist.all_gather_into_tensor(allr, flat)
        else:
            allr = acc.gather(flat.reshape(1, -1)).reshape(-1)
        allr = allr.view(world, -1)
        want = allr[0] * inv
        for r in range(1, world):
            want = want + allr[r] * inv  # separate mul and add kernels: rounds like the oracle (no FMA contraction)
        got = p.grad.reshape(-1)
        if algo == "twoshot":
            bad_exact += int((got != want).sum().item())
        diff = (got - want).abs()
        max_abs = max(max_abs, diff.max().item())
        max_rel = max(max_rel, (diff / want.abs().clamp_min(1e-12)).max().item() if diff.max().item() > 0 else 0.0)
        viol += int((diff > 1e-8 + 1e-5 * want.abs()).sum().item())
        del allr, want
    result["vs_rank_order_sum"] = {"max_abs": max_abs, "max_rel": max_rel, "elements_outside_rtol1e-5_atol1e-8": viol,
                                   "bit_exact": (bad_exact == 0) if algo == "twoshot" else None,
                                   "what": "reduce(L_r + 0) vs sum_r fl(L_r/N) in rank order; bit-exact required for two-shot, "
                                           "float32-rounding close for NVLS (switch-order sum)"}  # fmt: skip
    exact_ok = (bad_exact == 0) if algo == "twoshot" else (viol == 0)
    r2 = [p.grad.detach().clone() for p in params]
    # noise floor of the comparison itself: two backward passes of THIS model on the same batch (bf16 autocast, atomics
    # in the embedding / attention backward) do not produce the same local gradients bit for bit
    nf_abs, nf_bad = 0.0, 0
    for a, b_ in zip(r1, r2):
        diff = (a - b_).abs()
        nf_abs = max(nf_abs, diff.max().item())
        nf_bad += int((diff > 1e-5 + 1e-3 * b_.abs()).sum().item())
    result["repeat_backward_noise_floor"] = {"max_abs": nf_abs, "elements_outside_tolerance": nf_bad,
                                             "what": "ours vs ours, two backward passes of the same batch: what any comparison of "
                                                     "independently computed gradients is subject to"}  # fmt: skip
    ours.reducer.close()
    acc._wrappers = [w for w in acc._wrappers if w is not ours]
    acc._models = [m for m in acc._models if m is not ours.module]
    del ours
    torch.cuda.empty_cache()

    # ---- (3) stock torch DDP + NCCL ----
    vs_ddp = {"unavailable": "needs one GPU per rank (NCCL)"}
    ddp_ok = True
    if nccl:
        ref = torch.nn.parallel.DistributedDataParallel(build_model(name, dropout=0.0).to(dev), device_ids=[dev.index], bucket_cap_mb=25)
        ref.train()
        rparams = list(ref.module.parameters())

        def compare(ours_list, rtol, atol):
            m_abs = m_rel = 0.0
            bad = total = 0
            for a, q in zip(ours_list, rparams):
                b_ = q.grad
                diff = (a - b_).abs()
                m_abs = max(m_abs, diff.max().item())
                m_rel = max(m_rel, (diff / b_.abs().clamp_min(1e-5)).max().item())
                bad += int((diff > atol + rtol * b_.abs()).sum().item())
                total += b_.numel()
            return {"max_abs": m_abs, "max_rel_over_floor_1e-5": m_rel, "elements_outside_tolerance": bad, "elements": total}

        # (3a) the reduction alone: BOTH wrappers reduce the very same local gradients L_r.  DDP gets them as pre-existing
        # .grad (what a no_sync micro-step leaves behind) and reduces L_r + 0 in a sync backward of 0 * loss — its own
        # bucket copy-in, ncclAllReduce and copy-back (accelerator.py:1892-1894 -> torch reducer.cpp).
        for q, mine in zip(rparams, local):
            q.grad = mine.clone()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = forward_loss(ref, ids, name) * 0.0
        loss.float().backward()
        torch.cuda.synchronize()
        same = compare(r2, 1e-3, 1e-5)
        same["what"] = ("same local gradients through both wrappers: B200DataParallel vs torch DistributedDataParallel(bucket_cap_mb=25) "
                        "+ NCCL, rtol 1e-3 / atol 1e-5 (north star); isolates the reduction from backward non-determinism")  # fmt: skip
        # (3b) end to end: an independent forward/backward through DDP on the same weights and batch (subject to the
        # noise floor above, reported for completeness; not part of the verdict)
        for q in rparams:
            q.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = forward_loss(ref, ids, name)
        loss.float().backward()
        torch.cuda.synchronize()
        indep = compare(r1, 1e-3, 1e-5)
        indep["what"] = "independent forward/backward through each wrapper (same weights, batch, dropout 0): includes the noise floor"
        vs_ddp = {"same_local_gradients": same, "independent_backward": indep}
        ddp_ok = same["elements_outside_tolerance"] == 0
        del ref, rparams
        torch.cuda.empty_cache()
    del local, r1, r2
    result["vs_torch_ddp_nccl"] = vs_ddp
    flags = torch.tensor([int(result["all_ranks_bit_identical"]), int(exact_ok), int(ddp_ok)], device=dev)
    flags = acc.gather(flags.reshape(1, 3)).min(dim=0).values  # every rank must agree
    result["ok"] = bool(flags.min().item() == 1)
    result["seconds"] = round(time.perf_counter() - t0, 1)
    return result


# ------------------------------------------------------------------------------------------------
# kernel rooflines
# ------------------------------------------------------------------------------------------------
def adam_kernel_probe(opt, rt, peaks, peak_src, n_params):
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
    traffic, traffic_src = None, None
    summary = os.path.join(ROOT, "profiles", "r02_multi_adam_ncu.json")  # dram__bytes_read+write.sum of one `ncu --set full` capture
    if elems == 354_823_168 and os.path.exists(summary):
        try:
            traffic = json.load(open(summary))["dram_bytes_per_launch"]
            traffic_src = "profiles/r02_multi_adam_ncu.json"
        except Exception:
            pass
    return {"kernel": "multi_adam_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})", "traffic": traffic,
            "traffic_source": traffic_src, "bytes_per_launch": nbytes, "ms_per_launch": ms, "launches_per_step": len(groups),
            "gradients": "fusion buffer" if plan.reducer is not None else "loose tensors",
            "why_this_kernel": "longest own kernel of the step (28 B/element: read p, g, m, v; write p, m, v)"}  # fmt: skip


def event_time(fn, reps=10, warm=3):
    import torch

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
    for _ in range(warm):
        fn()
    for i in range(reps):
        ev[2 * i].record()
        fn()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    return sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps)) / reps


def kernel_roofline(acc, model, rt, world, peaks, peak_src):
    """Time the kernels of our path in place, on the stream they are launched on.

    N=1: clip_grad_norm_ = multi_stats (reads S) + finalize; the multi_scale pass (2*S) is what the deferred clip
         removes from the step — still measured, it is the pass any non-fused consumer triggers.
    N>1: the fused bucket allreduce: per 25 MiB bucket each GPU moves 2(N-1)/N * bytes over NVLink in the nccl-tests
         bus-bandwidth convention; by wire bytes NVLS moves S(1+1/N) per direction.  Peak = a peer copy measured here.
    """
    import torch

    from accelerate_b200 import native

    dev = acc.device
    st = torch.cuda.current_stream().cuda_stream
    extra = {}
    hbm_peak = float(peaks["hbm_gbs"])
    if world == 1:
        grads = [torch.randn_like(p) for p in model.parameters()]
        ptrs, numels = [g_.data_ptr() for g_ in grads], [g_.numel() for g_ in grads]
        nbytes = sum(numels) * 4
        mult = torch.tensor([0.999], device=dev)
        slot = rt.reserve_slots(2)
        t_scale = event_time(lambda: rt.ctx.multi_scale(ptrs, numels, native.F32, mult.data_ptr(), st))
        t_stats = event_time(lambda: rt.ctx.multi_stats(ptrs, numels, native.F32, slot, st))
        achieved = 2 * nbytes / (t_scale * 1e-3) / 1e9
        roof = {"kernel": "multi_scale_kernel<float>", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one `ncu --set full`
                # capture (profiles/r01_n1_clip_kernels_ncu.txt): 1.419 GB + 1.360 GB -> no re-reads
                "traffic": 2_779_191_000 if nbytes == GRAD_BYTES else None,
                "traffic_source": "profiles/r01_n1_clip_kernels_ncu.txt",
                "bytes_per_launch": 2 * nbytes, "ms_per_launch": t_scale,
                "note": "runs only when something other than the fused Adam step reads a gradient after clip_grad_norm_"}  # fmt: skip
        extra["roofline_stats_kernel"] = {"kernel": "multi_stats_kernel<float>", "bound": "hbm",
                                          "achieved": nbytes / (t_stats * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                          "frac": nbytes / (t_stats * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": nbytes,
                                          "ms_per_launch": t_stats}  # fmt: skip
        extra["pack_kernel"] = pack_probe(rt, [g_ for g_ in grads if g_.numel() * 4 <= (8 << 20)][:16] or grads[:4], st, hbm_peak)
        return roof, extra

    # N > 1: one 25 MiB bucket of the real fusion buffer, in place, all ranks in lock-step
    red = model.reducer
    group = next(iter(red.groups.values()))
    count = min(25 * 1024 * 1024 // 4, group.total_elems // 4 * 4)
    algo = red._pick_algo(group.segment)
    nvls = algo == native.ALGO_NVLS
    sizes = {}
    blocks = red.blocks_for(algo)
    flags_iso = native.F_STATS
    flags_seq = native.F_STATS | native.F_NO_END_BARRIER
    K = 8  # buckets back to back, the way a backward pass issues them (closing barrier replaced by ONE fence)
    reps = 10

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
        t = torch.tensor([event_time(lambda: fn(arg), reps=reps) / per], device=dev)
        return acc.gather(t).max().item()

    def entry(nbytes, ms):
        return {"bytes": nbytes, "ms": ms, "busbw_GBps": nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9,
                "algbw_GBps": nbytes / (ms * 1e-3) / 1e9}  # fmt: skip

    # ---- NVLink peak measured here: every rank pulls 256 MiB from its right neighbour at the same time ----
    span = min(256 << 20, group.total_elems * 4 // 16 * 16)
    sink = torch.empty(span, dtype=torch.uint8, device=dev)
    nbr = (acc.process_index + 1) % world
    by_ctas = {}
    for ctas in (16, 32, 64, 148):
        t_copy = clock(lambda _: rt.ctx.pull(group.segment.seg, 0, span, sink.data_ptr(), 0, nbr, 1, native.CH_USER, ctas, st), None)
        by_ctas[ctas] = span / (t_copy * 1e-3) / 1e9
    t_ce = clock(lambda _: rt.ctx.peer_copy(group.segment.seg, nbr, 0, sink.data_ptr(), span, st), None)
    by_ctas["copy_engine"] = span / (t_ce * 1e-3) / 1e9
    nvlink_peak = max(by_ctas.values())
    extra["nvlink_peer_copy"] = {"GBps_per_direction": nvlink_peak, "bytes": span, "GBps_by_method": by_ctas,
                                 "how": "every rank reads 256 MiB from rank+1's symmetric segment into local HBM, all ranks at once "
                                        "(each GPU's links carry one stream in and one out): pull_kernel with 16/32/64/148 CTAs and a "
                                        "copy-engine cudaMemcpyAsync over the peer mapping; best method, max over ranks"}  # fmt: skip
    del sink

    counters = NvlinkCounters(dev)
    assert K * count <= group.total_elems
    before = counters.read() if counters.ok else None
    sizes["bucket_25MiB_steady"] = entry(count * 4, clock(sequence, count, per=K))
    torch.cuda.synchronize()
    after = counters.read() if counters.ok else None
    launches_counted = K * (reps + 3)
    traffic = None
    if before is not None and after is not None:
        tx, rx = (after[0] - before[0]) / launches_counted, (after[1] - before[1]) / launches_counted
        traffic = {"nvlink_tx_bytes_per_launch": tx, "nvlink_rx_bytes_per_launch": rx, "launches": launches_counted,
                   "source": "NVML NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX (all links of this GPU), delta around the bucket sequence"}  # fmt: skip
    sizes["bucket_25MiB_isolated"] = entry(count * 4, clock(isolated, count))
    whole = group.total_elems // 4 * 4
    sizes["all_grads"] = entry(whole * 4, clock(isolated, whole))
    sizes["ctas"] = blocks
    mid = sorted(red.buckets, key=lambda bk: bk.count)[len(red.buckets) // 2]
    extra["pack_kernel"] = pack_probe(rt, [torch.randn_like(p) for p in mid.params], st, hbm_peak, dsts=list(mid.ptrs))
    group.flat.zero_()
    b = sizes["bucket_25MiB_steady"]
    wire = b["bytes"] * ((1 + 1 / world) if nvls else 2 * (world - 1) / world)  # bytes that cross this GPU's links, per direction
    wire_rate = wire / (b["ms"] * 1e-3) / 1e9
    roof = {"kernel": f"allreduce_inplace_kernel<float,{'NVLS' if nvls else 'twoshot'}>", "bound": "nvlink",
            "achieved": b["busbw_GBps"], "peak": nvlink_peak, "unit": "GB/s", "frac": b["busbw_GBps"] / nvlink_peak,
            "peak_source": "peer copy measured in this run (nvlink_peer_copy); NVLink 5 nominal is 900 GB/s/dir",
            "traffic": traffic["nvlink_tx_bytes_per_launch"] + traffic["nvlink_rx_bytes_per_launch"] if traffic else None,
            "traffic_detail": traffic if traffic else {"unavailable": counters.why or "NVML NVLink counters not readable"},
            "bytes_per_launch": b["bytes"], "ms_per_launch": b["ms"], "frac_of_900_nominal": b["busbw_GBps"] / 900.0,
            "by_wire_bytes": {"bytes_per_direction": wire, "GBps": wire_rate, "frac_of_measured_peak": wire_rate / nvlink_peak,
                              "frac_of_900_nominal": wire_rate / 900.0,
                              "what": "S(1+1/N) per direction for NVLS (the switch reduces and replicates), 2(N-1)/N*S for two-shot"},
            "how": f"{K} x 25 MiB buckets back to back on one stream exactly as a backward pass issues them (statistics on, "
                   f"closing barrier replaced by one fence), {blocks} CTAs; time per bucket, max over ranks; achieved = bus bandwidth "
                   "S*2(N-1)/N/t (nccl-tests convention; NVLS moves ~S per GPU so it can exceed the link rate)"}  # fmt: skip
    extra["allreduce_sizes"] = sizes
    try:  # HBM-side kernel of a materialised clip at N > 1: one multiply pass over the fusion buffer (2*S bytes)
        mult = torch.tensor([0.999], device=dev)
        n_el = group.total_elems
        t_scale = event_time(lambda: rt.ctx.scale_flat(group.flat.data_ptr(), n_el, native.F32, mult.data_ptr(), 0, st))
        hbm = 2 * n_el * 4 / (t_scale * 1e-3) / 1e9
        extra["roofline_hbm"] = {"kernel": "scale_flat_kernel<float>", "bound": "hbm", "achieved": hbm, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": hbm / hbm_peak, "bytes_per_launch": 2 * n_el * 4,
                                 "ms_per_launch": t_scale, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                                 "note": "not in the default step any more: the fused Adam kernel applies the clip coefficient"}  # fmt: skip
    except Exception as err:  # never let an auxiliary measurement take the headline down
        extra["roofline_hbm"] = {"error": str(err)}
    return roof, extra


def pack_probe(rt, fresh, st, hbm_peak, dsts=None):
    """Bucket pack (what replaces DDP's per-parameter copy-in): both variants on a real bucket's tensor list, next to
    torch's own copy of the same bytes (what an HBM copy of this size can reach at all)."""
    import torch

    nbytes = [f.numel() * f.element_size() for f in fresh]
    total = sum(nbytes)
    if dsts is None:
        sink = torch.empty(total + 64 * len(fresh), dtype=torch.uint8, device=fresh[0].device)
        dsts, off = [], sink.data_ptr()
        for n in nbytes:
            dsts.append(off)
            off += (n + 31) // 32 * 32
    srcs = [f.data_ptr() for f in fresh]
    out = {"bucket_bytes": total, "tensors": len(fresh), "default": "bulk_tma" if rt.bulk_for(total) else "ldg_stg"}
    for label, bulk in (("bulk_tma", 1), ("ldg_stg", 0)):
        t = event_time(lambda: rt.ctx.pack(srcs, dsts, nbytes, bulk, st))
        out[label] = {"ms": t, "GBps": 2 * total / (t * 1e-3) / 1e9, "frac_of_hbm_peak": 2 * total / (t * 1e-3) / 1e9 / hbm_peak}
    a = torch.empty(total, dtype=torch.uint8, device=fresh[0].device)
    b = torch.empty_like(a)
    t = event_time(lambda: b.copy_(a))
    out["torch_copy_same_bytes"] = {"ms": t, "GBps": 2 * total / (t * 1e-3) / 1e9, "frac_of_hbm_peak": 2 * total / (t * 1e-3) / 1e9 / hbm_peak}
    return out


# ------------------------------------------------------------------------------------------------
# comparators: what the reference delegates to, on the same box
# ------------------------------------------------------------------------------------------------
def _time_leg(step, steps, warmup, world, acc, dev, args):
    import torch

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
    return {"value": args.batch * args.accum * world * steps / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms / steps}


def torch_ddp_leg(args, dev, world, dev_batches, steps, warmup, acc, fused_adam):
    """The stack the reference delegates to, stock: torch DDP (bucket_cap_mb=25) + NCCL, autocast bf16,
    torch.nn.utils.clip_grad_norm_, AdamW (foreach — torch's default — or fused=True).  Same model, seeds, batches, step count."""
    import contextlib

    import torch

    model = build_model(args.model).to(dev)
    model.train()
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], bucket_cap_mb=25)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True) if fused_adam else torch.optim.AdamW(model.parameters(), lr=1e-4)
    A = args.accum

    def step(i):
        for a in range(A):
            ids = dev_batches[(i * A + a) % len(dev_batches)]
            sync = a == A - 1
            ctx = contextlib.nullcontext() if (sync or world == 1) else model.no_sync()
            with ctx:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = forward_loss(model, ids, args.model)
                (loss.float() / A).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()

    leg = _time_leg(step, steps, warmup, world, acc, dev, args)
    del model, opt
    torch.cuda.empty_cache()
    leg["what"] = ("torch DistributedDataParallel + NCCL + torch clip_grad_norm_ + AdamW("
                   + ("fused=True" if fused_adam else "foreach") + ") (same model/batches/steps)")  # fmt: skip
    return leg


def reference_accelerate_leg(args, dev, world, dev_batches, steps, warmup, acc):
    """the UNMODIFIED reference (huggingface/accelerate, installed into baseline/_ref) through its
    own public API on the GPUs: ``Accelerator(mixed_precision="bf16").prepare`` (-> torch DDP + NCCL at N > 1),
    ``accumulate``, ``backward``, ``clip_grad_norm_``, AdamW.  Same process group, model, seeds, batches, step count."""
    import torch

    if not os.path.isdir(os.path.join(REF_DIR, "accelerate")):
        return {"unavailable": "baseline/_ref/accelerate not present on this box"}
    try:
        sys.path.insert(0, REF_DIR)
        import accelerate as ref_accelerate

        racc = ref_accelerate.Accelerator(
            mixed_precision="bf16", gradient_accumulation_steps=args.accum,
            kwargs_handlers=[ref_accelerate.DistributedDataParallelKwargs(bucket_cap_mb=25)])  # fmt: skip
        model = build_model(args.model).to(dev)
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        model, opt = racc.prepare(model, opt)
        A = args.accum

        def step(i):
            for a in range(A):
                ids = dev_batches[(i * A + a) % len(dev_batches)]
                with racc.accumulate(model):
                    loss = forward_loss(model, ids, args.model)
                    racc.backward(loss)
                    if racc.sync_gradients:
                        racc.clip_grad_norm_(model.parameters(), 1.0)
                    opt.step()
                    opt.zero_grad()

        leg = _time_leg(step, steps, warmup, world, acc, dev, args)
        leg["what"] = (f"huggingface/accelerate {ref_accelerate.__version__} from baseline/_ref, unmodified: Accelerator.prepare "
                       f"(wrapper: {type(model).__name__}) + accumulate + backward + clip_grad_norm_ + torch AdamW")  # fmt: skip
        del model, opt
        torch.cuda.empty_cache()
        return leg
    except Exception as err:  # noqa: BLE001 - the comparison leg must never take our own result down
        return {"unavailable": f"reference Accelerate leg failed: {err!r}"}
    finally:
        if REF_DIR in sys.path:
            sys.path.remove(REF_DIR)


# ------------------------------------------------------------------------------------------------
# config 5: collective microbench
# ------------------------------------------------------------------------------------------------
def microbench(args, acc, rt, rank, world, dev):
    """BASELINE.json config 5: ``accelerator.reduce(t, "sum")`` and ``accelerator.gather(t)`` on fp32 tensors of 2^k bytes,
    next to torch.distributed NCCL (the reference's path, operations.py:871,342) and the in-place bucket allreduce the
    gradient path uses.  busbw = algbw * 2(N-1)/N (allreduce), algbw * (N-1)/N with S = total output (allgather)."""
    import torch
    import torch.distributed as dist

    from accelerate_b200 import native
    from accelerate_b200.state import PartialState

    nccl = PartialState().backend == "nccl" and world > 1
    st = torch.cuda.current_stream().cuda_stream
    max_bytes = 1 << args.max_log2
    seg = rt.alloc_segment(max_bytes, multicast=True)
    flat = seg.tensor(torch.float32, max_bytes // 4)
    red_algo = native.ALGO_NVLS if (seg.multicast and world >= 4 and world & (world - 1) == 0) else native.ALGO_TWOSHOT

    def timeit(fn, nbytes):
        iters = 50 if nbytes <= (64 << 20) else (10 if nbytes <= (256 << 20) else 5)
        for _ in range(10 if nbytes <= (64 << 20) else 2):
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
        return acc.gather(ms).max().item() if world > 1 else ms.item()

    rows = []
    for k in range(10, args.max_log2 + 1, 1 if args.max_log2 <= 24 else 2):
        nbytes = 1 << k
        n = nbytes // 4
        ar, ag = 2 * (world - 1) / world, (world - 1)
        row = {"bytes": nbytes}
        flat[:n].normal_()
        ms = timeit(lambda: rt.ctx.allreduce_bucket(seg.seg, 0, n, native.F32, 1.0 / world, 0, native.F_STATS, red_algo,
                                                    native.CH_USER, 0, native.NO_STREAM, st), nbytes)  # fmt: skip
        row["bucket_allreduce_us"], row["bucket_allreduce_busbw"] = ms * 1e3, nbytes * ar / (ms * 1e-3) / 1e9
        t = torch.randn(n, device=dev)
        ms = timeit(lambda: acc.reduce(t, "sum"), nbytes)
        row["reduce_us"], row["reduce_busbw"] = ms * 1e3, nbytes * ar / (ms * 1e-3) / 1e9
        if nbytes * world <= (4 << 30):
            ms = timeit(lambda: acc.gather(t), nbytes)
            row["gather_us"], row["gather_busbw"] = ms * 1e3, nbytes * ag / (ms * 1e-3) / 1e9
        if nccl:
            ms = timeit(lambda: dist.all_reduce(t), nbytes)
            row["nccl_allreduce_us"], row["nccl_allreduce_busbw"] = ms * 1e3, nbytes * ar / (ms * 1e-3) / 1e9
            if nbytes * world <= (4 << 30):
                out = torch.empty(n * world, device=dev)
                ms = timeit(lambda: dist.all_gather_into_tensor(out, t), nbytes)
                row["nccl_allgather_us"], row["nccl_allgather_busbw"] = ms * 1e3, nbytes * ag / (ms * 1e-3) / 1e9
                del out
        del t
        rows.append(row)
    rt.ctx.check_status()
    if rank == 0:
        big = [r for r in rows if r["bytes"] >= (16 << 20) and "nccl_allreduce_busbw" in r]
        emit({
            "metric": "allreduce_bus_GBps_vs_size", "unit": "GB/s", "n_gpus": world, "higher_is_better": True, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "accelerator.reduce / accelerator.gather on fp32 tensors of 2^k bytes (BASELINE.json config 5)",
                                            "bucket_allreduce_algo": "nvls" if red_algo == native.ALGO_NVLS else "twoshot"},
            "value": max((r["reduce_busbw"] for r in rows), default=None),
            "wins_vs_nccl_at_16MiB_and_up": {"reduce": sum(r["reduce_busbw"] >= r["nccl_allreduce_busbw"] for r in big),
                                             "gather": sum(r.get("gather_busbw", 0) >= r.get("nccl_allgather_busbw", 1e30) for r in big),
                                             "sizes": len(big)},
            "rows": rows,
        })  # fmt: skip


if __name__ == "__main__":
    main()
