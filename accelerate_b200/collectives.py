"""Tensor collectives on the P2P kernels: allgather, allreduce, broadcast, pad-to-max.

Each function takes ONE contiguous CUDA tensor and follows the reference's per-leaf semantics
(``src/accelerate/utils/operations.py``): ``gather`` allocates and concatenates on dim 0
(:316-352), ``reduce`` never mutates its input and divides only for "mean" (:839-878),
``broadcast`` writes in place (:520-525), ``pad_across_processes`` returns the *same object* when
no padding is needed (:743-796).  Data moves through the symmetric staging segment:

    stage (b200dp_pack, one launch)  ->  peers pull / reduce over NVLink (one launch)

so a small gather costs two launches and zero NCCL calls.  Tensors larger than the staging
segment are processed in chunks.
"""

from __future__ import annotations

import torch

from . import native
from .runtime import SymmRuntime, native_dtype

import os

# below this many bytes per rank the out-of-place one-shot allreduce wins (latency bound)
ONESHOT_MAX_BYTES = 512 * 1024
# large messages: one fused launch per staging window (b200dp_gather / b200dp_reduce).  "0" selects the round-1 sequence
# (pack -> pull / in-place allreduce -> pack) that the fused kernels are measured against (benchmarks/collective_sweep.py).
FUSED_GATHER = os.environ.get("B200DP_FUSED_GATHER", "1") != "0"
FUSED_REDUCE = os.environ.get("B200DP_FUSED_REDUCE", "1") != "0"
# gather: below this many bytes per rank the two-launch sequence (pack with the whole chip, then pull) is still a few
# per cent faster than the fused kernel, above it the fused kernel wins (r02 sweeps: 2 GPUs 430 vs 409 GB/s at 64 MiB,
# 8 GPUs 614 vs 603 at 256 MiB but 568 vs 580 at 64 MiB).  B200DP_FUSED_GATHER_MIN_MB overrides (0 = always fused).
_FUSED_GATHER_MIN_MB = os.environ.get("B200DP_FUSED_GATHER_MIN_MB")


def _fused_gather_min(world: int) -> int:
    if _FUSED_GATHER_MIN_MB is not None:
        return int(_FUSED_GATHER_MIN_MB) << 20
    return (32 if world <= 2 else 128) << 20


def _bulk_capacity(rt, staging) -> int:
    """Bytes of the staging segment usable by the stage-then-pull/reduce path (the tail belongs to b200dp_push)."""
    return (staging.nbytes - rt.ctx.push_area_bytes()) // 16 * 16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_cuda(t: torch.Tensor, rt: SymmRuntime, op: str):
    if t.device != rt.device:
        raise RuntimeError(
            f"{op}: tensor is on {t.device} but this process drives {rt.device}; move it there first "
            "(there is no CPU fallback for the CUDA collectives)"
        )


def all_gather(t: torch.Tensor) -> torch.Tensor:
    """out[r*b:(r+1)*b] = rank r's tensor; 0-dim inputs become shape [N]. Bit-exact, any dtype."""
    rt = SymmRuntime.get()
    _check_cuda(t, rt, "gather")
    if t.ndim == 0:
        t = t.clone()[None]
    if not t.is_contiguous():
        t = t.contiguous()
    world = rt.world
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return out
    staging = rt.ensure_staging()
    st = _stream()
    if nbytes <= rt.push_max:  # latency path: one launch, one cross-GPU barrier
        rt.ctx.push(staging.seg, 0, t.data_ptr(), nbytes, native.U8, 1.0, out.data_ptr(), native.CH_PUSH, st)
        return out
    # large path: ONE fused launch per staging-window-sized piece — CTA b stages its portion and pulls the same portion
    # from every peer as soon as their CTA b has staged theirs (b200dp_gather); no separate pack, no round trip
    cap = _bulk_capacity(rt, staging)
    src, dst = t.data_ptr(), out.data_ptr()
    done = 0
    fused = FUSED_GATHER and nbytes >= _fused_gather_min(world)
    while done < nbytes:
        n = min(cap, nbytes - done)
        if fused:
            rt.ctx.gather(staging.seg, 0, src + done, n, dst + done, nbytes, native.CH_USER, 0, st)
        else:
            rt.ctx.pack([src + done], [staging.ptr], [n], rt.bulk_for(n), st)
            rt.ctx.pull(staging.seg, 0, n, dst + done, nbytes, 0, world, native.CH_USER, 0, st)
        done += n
    return out


def all_reduce(t: torch.Tensor, mean: bool = False) -> torch.Tensor:
    """Return sum over ranks of ``t`` (a fresh tensor; ``t`` is untouched), divided by N if ``mean``."""
    rt = SymmRuntime.get()
    _check_cuda(t, rt, "reduce")
    world = rt.world
    src = t if t.is_contiguous() else t.contiguous()
    dt = native_dtype(src.dtype)
    if dt is None:
        raise TypeError(f"reduce: dtype {src.dtype} is not supported by the P2P allreduce kernels")
    if src.dtype in (torch.bool, torch.int8):
        raise TypeError(f"reduce: dtype {src.dtype} has no well-defined sum")
    out = torch.empty_like(src)
    numel, es = src.numel(), src.element_size()
    if numel == 0:
        return out
    if mean and not src.is_floating_point():
        # reference: `cloned_tensor /= N` raises for integer tensors
        raise RuntimeError("result type Float can't be cast to the desired output type " + str(src.dtype))
    staging = rt.ensure_staging()
    st = _stream()
    nbytes = numel * es
    divisor = float(world) if mean else 1.0
    fused_types = (torch.float32, torch.bfloat16, torch.float16)
    if nbytes <= rt.push_max:  # latency path: one launch, one cross-GPU barrier
        rt.ctx.push(staging.seg, 1, src.data_ptr(), nbytes, dt, divisor, out.data_ptr(), native.CH_PUSH, st)
        return out.view(t.shape)
    cap_elems = _bulk_capacity(rt, staging) // 16 * 16 // es
    large = nbytes > ONESHOT_MAX_BYTES and src.dtype in fused_types
    algo = native.ALGO_NVLS if (staging.multicast and world >= 4) else native.ALGO_TWOSHOT
    done = 0
    while done < numel:
        n = min(cap_elems, numel - done)
        if large and FUSED_REDUCE:
            # ONE fused launch (b200dp_reduce): CTA b stages its portion, the owners reduce their share of it over all
            # ranks' staged copies (two-shot by peer loads, or NVLS in the switch) and write the result into every
            # rank's staging, CTA b copies its portion out — dividing on the way for "mean", like the reference's `/= N`
            rt.ctx.reduce(staging.seg, 0, src.data_ptr() + done * es, n, dt, divisor, out.data_ptr() + done * es, algo,
                          native.CH_USER, 0, st)  # fmt: skip
        elif large:
            nb = n * es
            rt.ctx.pack([src.data_ptr() + done * es], [staging.ptr], [nb], rt.bulk_for(nb), st)
            padded = (nb + 15) // 16 * 16 // es
            rt.ctx.allreduce_bucket(staging.seg, 0, padded, dt, 1.0, -1, 0, algo, native.CH_USER, 0, native.NO_STREAM, st)
            rt.ctx.pack([staging.ptr], [out.data_ptr() + done * es], [nb], rt.bulk_for(nb), st)
        else:
            rt.ctx.pack([src.data_ptr() + done * es], [staging.ptr], [n * es], rt.bulk_for(n * es), st)
            rt.ctx.allreduce_oneshot(staging.seg, 0, n, dt, divisor, out.data_ptr() + done * es, native.CH_USER, 0, st)
        done += n
    if mean and large and not FUSED_REDUCE:
        out /= world  # same op the reference applies (`/= N`)
    return out.view(t.shape)


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    """In place: every rank's ``t`` becomes rank ``src``'s."""
    rt = SymmRuntime.get()
    _check_cuda(t, rt, "broadcast")
    work = t if t.is_contiguous() else t.contiguous()
    nbytes = work.numel() * work.element_size()
    if nbytes:
        staging = rt.ensure_staging()
        st = _stream()
        cap = _bulk_capacity(rt, staging)
        done = 0
        while done < nbytes:
            n = min(cap, nbytes - done)
            if rt.rank == src:
                rt.ctx.pack([work.data_ptr() + done], [staging.ptr], [n], rt.bulk_for(n), st)
            rt.ctx.pull(staging.seg, 0, n, work.data_ptr() + done, 0, src, 1, native.CH_USER, 0, st)
            done += n
    if work is not t:
        t.copy_(work)
    return t


def exchange_shape(shape) -> list[list[int]]:
    """All ranks' shapes as host ints with one tiny kernel (no NCCL, no D2H copy, no stream sync)."""
    import struct

    rt = SymmRuntime.get()
    payload = struct.pack(f"<{len(shape)}q", *shape) if len(shape) else b""
    if len(payload) > native.SMALL_MAX:
        raise ValueError(f"tensor rank {len(shape)} too large for the shape exchange")
    if not payload:
        return [[] for _ in range(rt.world)]
    parts = rt.ctx.exchange_small(payload, _stream())
    return [list(struct.unpack(f"<{len(shape)}q", p)) for p in parts]


def pad_to_max(t: torch.Tensor, dim: int, pad_index, pad_first: bool) -> torch.Tensor:
    rt = SymmRuntime.get()
    _check_cuda(t, rt, "pad_across_processes")
    sizes = exchange_shape(tuple(t.shape))
    max_size = max(s[dim] for s in sizes)
    if max_size == t.shape[dim]:
        return t
    if isinstance(pad_index, torch.Tensor) and pad_index.numel() == 1:
        pad_index = pad_index.item()  # e.g. split_between_processes(apply_padding=True) passes `inputs[-1]`
    # `tensor.new_zeros(..) + pad_index` may promote (int tensor, float pad): follow torch's rule
    res_dtype = torch.result_type(t, pad_index)
    src = t if t.dtype == res_dtype else t.to(res_dtype)
    if not src.is_contiguous():
        src = src.contiguous()
    new_shape = list(src.shape)
    new_shape[dim] = max_size
    out = torch.empty(new_shape, dtype=src.dtype, device=src.device)
    es = src.element_size()
    if es > 8 or isinstance(pad_index, torch.Tensor):
        # complex128 has no native pattern and a tensor-valued pad (a whole row) broadcasts: rare enough for the 3-op route
        out = src.new_zeros(new_shape) + pad_index
        idx = [slice(None)] * len(new_shape)
        idx[dim] = slice(max_size - src.shape[dim], max_size) if pad_first else slice(0, src.shape[dim])
        out[tuple(idx)] = src
        return out
    pad_bits = int.from_bytes(torch.tensor([pad_index], dtype=src.dtype).view(torch.uint8).numpy().tobytes(), "little")
    outer = 1
    for s in src.shape[:dim]:
        outer *= s
    inner = 1
    for s in src.shape[dim + 1 :]:
        inner *= s
    rt.ctx.pad_copy(src.data_ptr() if src.numel() else None, out.data_ptr(), outer, src.shape[dim], max_size, inner * es,
                    es, pad_bits, pad_first, _stream())  # fmt: skip
    return out


# ---------------------------------------------------------------------------------------------
# object collectives: pickled bytes over the same kernels (reference:
# operations.py:492-517 all_gather_object, :669-690 broadcast_object_list)
# ---------------------------------------------------------------------------------------------
def _exchange_lengths(n: int) -> list[int]:
    import struct

    rt = SymmRuntime.get()
    return [struct.unpack("<q", p)[0] for p in rt.ctx.exchange_small(struct.pack("<q", n), _stream())]


def gather_object(obj) -> list:
    """[rank 0's obj, rank 1's obj, ...]: lengths by the 1-CTA shape exchange, payload by the allgather kernels."""
    import pickle

    rt = SymmRuntime.get()
    payload = pickle.dumps(obj)
    lengths = _exchange_lengths(len(payload))
    width = (max(lengths) + 15) // 16 * 16
    buf = torch.zeros(width, dtype=torch.uint8)
    buf[: len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    gathered = all_gather(buf.to(rt.device).reshape(1, width)).cpu()
    return [pickle.loads(gathered[r, : lengths[r]].numpy().tobytes()) for r in range(rt.world)]


def broadcast_object(obj, src: int = 0):
    """Every rank returns rank ``src``'s object."""
    import pickle

    rt = SymmRuntime.get()
    payload = pickle.dumps(obj) if rt.rank == src else b""
    n = _exchange_lengths(len(payload))[src]
    width = (n + 15) // 16 * 16
    buf = torch.zeros(width, dtype=torch.uint8)
    if rt.rank == src:
        buf[:n] = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    dev = buf.to(rt.device)
    broadcast_(dev, src)
    return obj if rt.rank == src else pickle.loads(dev.cpu()[:n].numpy().tobytes())
