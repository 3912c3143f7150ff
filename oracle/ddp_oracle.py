"""ORACLE — test infrastructure only.

A CPU restatement (numpy, float32 arithmetic) of the data-parallel gradient path that the reference
delegates to PyTorch.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
(``accelerate_b200/``) never does.

Where the arithmetic lives: NOT in ``/root/reference`` (pure Python, zero native code) but in its
dependency **torch** (``setup.py:85`` pins ``torch>=2.0.0``; installed here: 2.11.0+cu128, NCCL
2.28.9).  Each function cites the reference call site that reaches the torch routine it restates,
and the torch routine itself.

Pinning: ``tests/golden/*.npz|json`` were produced by running the *reference itself*
(``PYTHONPATH=/root/reference/src``, gloo, 2 and 4 ranks) with ``tests/golden/make_golden.py``;
``tests/test_oracle_golden.py`` checks every function below against them, and against torch's own
``clip_grad_norm_`` / ``_amp_foreach_non_finite_check_and_unscale_`` on CPU.
"""

from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------
# a2: DDP gradient reduction  (reference: accelerator.py:1892-1894 -> torch DistributedDataParallel;
# torch reducer.hpp:276 mark_variable_ready_dense `bucket_view = grad * (1/N)`, :524 allreduce SUM)
# --------------------------------------------------------------------------------------------
def ddp_mean_reduce(per_rank_grads, dtype=np.float32):
    """grad <- sum_r fl(g_r * fl(1/N)), accumulated in rank order in float32, result cast to ``dtype``.

    Pre-divide THEN sum (torch default hook: "divide first to avoid overflow").  At N=2 this is
    bit-identical to DDP on any backend (one commutative add); for N>2 the association order of
    NCCL/gloo rings differs, which the parity tests allow for with rtol=1e-3/atol=1e-5
    (BASELINE.json), while all *our* ranks must match this definition bit for bit.
    """
    n = len(per_rank_grads)
    inv = F32(1.0) / F32(n)
    acc = np.asarray(per_rank_grads[0], dtype=F32) * inv
    for g in per_rank_grads[1:]:
        acc = acc + np.asarray(g, dtype=F32) * inv
    return acc.astype(dtype)


def ddp_accumulate_then_reduce(per_rank_microbatch_grads):
    """no_sync semantics (torch distributed.py:1474-1500): micro-step grads accumulate locally
    (plain float32 adds, autograd AccumulateGrad), only the accumulated tensor is reduced."""
    locally = []
    for micro in per_rank_microbatch_grads:
        acc = np.asarray(micro[0], dtype=F32).copy()
        for g in micro[1:]:
            acc = acc + np.asarray(g, dtype=F32)
        locally.append(acc)
    return ddp_mean_reduce(locally)


# --------------------------------------------------------------------------------------------
# a6: GradScaler.unscale_   (reference: accelerator.py:2944 -> torch/amp/grad_scaler.py
# `_unscale_grads_` -> `_amp_foreach_non_finite_check_and_unscale_`)
# --------------------------------------------------------------------------------------------
def inv_scale_of(scale) -> np.float32:
    """torch: ``self._scale.double().reciprocal().float()``."""
    return F32(1.0 / np.float64(F32(scale)))


def unscale(grads, scale):
    """-> (unscaled grads, found_inf).  g * inv_scale in float32; found_inf = any non-finite input."""
    inv = inv_scale_of(scale)
    found = 0.0
    out = []
    for g in grads:
        g = np.asarray(g, dtype=F32)
        if not np.all(np.isfinite(g)):
            found = 1.0
        with np.errstate(invalid="ignore", over="ignore"):
            out.append(g * inv)
    return out, found


# --------------------------------------------------------------------------------------------
# a7: clip_grad_norm_   (reference: accelerator.py:3007 -> torch/nn/utils/clip_grad.py)
# --------------------------------------------------------------------------------------------
def total_norm(grads, norm_type=2.0) -> np.float32:
    """torch: norms = _foreach_norm(grads, p); total = vector_norm(stack(norms), p)."""
    if len(grads) == 0:
        return F32(0.0)
    if np.isinf(norm_type):
        return F32(max(np.max(np.abs(np.asarray(g, dtype=F32))) if np.size(g) else 0.0 for g in grads))
    per = [np.sqrt(np.sum(np.asarray(g, dtype=np.float64) ** 2)) for g in grads]  # float64 accumulation ...
    return F32(np.sqrt(np.sum(np.asarray(per, dtype=np.float64) ** 2)))  # ... rounded once to float32


def clip_coef(total, max_norm) -> np.float32:
    """torch: clamp(max_norm / (total_norm + 1e-6), max=1.0), all in float32."""
    coef = F32(max_norm) / (F32(total) + F32(1e-6))
    return F32(min(coef, F32(1.0))) if not np.isnan(coef) else coef


def clip_grad_norm(grads, max_norm, norm_type=2.0):
    """-> (clipped grads, total_norm).  g * coef in float32 (torch multiplies even when coef == 1)."""
    tn = total_norm(grads, norm_type)
    coef = clip_coef(tn, max_norm)
    return [np.asarray(g, dtype=F32) * coef for g in grads], tn


def clip_grad_value(grads, clip_value):
    """reference: accelerator.py:3034 -> torch clip_grad_value_ = clamp_(-v, v)."""
    v = F32(clip_value)
    return [np.clip(np.asarray(g, dtype=F32), -v, v) for g in grads]


# --------------------------------------------------------------------------------------------
# a9-a12: tensor collectives   (reference: utils/operations.py)
# --------------------------------------------------------------------------------------------
def gather(per_rank):
    """operations.py:316-352: 0-dim -> [1]; out = concat over ranks on dim 0.  Bit-exact, any dtype."""
    parts = [np.asarray(t)[None] if np.ndim(t) == 0 else np.asarray(t) for t in per_rank]
    return np.concatenate(parts, axis=0)


def reduce(per_rank, reduction="mean"):
    """operations.py:839-878: clone, all_reduce(SUM) in the tensor's dtype, then ``/= N`` iff "mean"."""
    acc = np.asarray(per_rank[0]).copy()
    for t in per_rank[1:]:
        acc = acc + np.asarray(t)
    if reduction == "mean":
        acc = (acc / acc.dtype.type(len(per_rank))).astype(acc.dtype)
    return acc


def broadcast(per_rank, src=0):
    """operations.py:520-525: every rank ends with rank ``src``'s values."""
    return [np.asarray(per_rank[src]).copy() for _ in per_rank]


def pad_across_processes(per_rank, dim=0, pad_index=0, pad_first=False):
    """operations.py:743-796.  Returns one array per rank; a rank already at the max size gets its
    input back unchanged (the reference returns the same object)."""
    outs = []
    nd = np.ndim(per_rank[0])
    if dim >= nd or dim < -nd:
        return [np.asarray(t) for t in per_rank]
    if dim < 0:
        dim += nd
    max_size = max(np.shape(t)[dim] for t in per_rank)
    for t in per_rank:
        t = np.asarray(t)
        if t.shape[dim] == max_size:
            outs.append(t)
            continue
        new_shape = list(t.shape)
        new_shape[dim] = max_size
        out = np.zeros(new_shape, dtype=t.dtype) + np.asarray(pad_index).astype(t.dtype)
        window = [slice(None)] * nd
        window[dim] = slice(max_size - t.shape[dim], max_size) if pad_first else slice(0, t.shape[dim])
        out[tuple(window)] = t
        outs.append(out)
    return outs


# --------------------------------------------------------------------------------------------
# bucket layout   (torch reducer.hpp:30-31 kDefaultFirstBucketBytes / kDefaultBucketBytesCap,
# torch.distributed._compute_bucket_assignment_by_size; configured by dataclasses.py:183)
# --------------------------------------------------------------------------------------------
def bucket_assignment(sizes_bytes, first_cap=1 << 20, cap=25 << 20):
    """Single-dtype restatement: walk tensors in order, close a bucket when its bytes reach the limit
    (first_cap for the first bucket, cap afterwards); tensors are never split."""
    buckets, cur, total, limit = [], [], 0, first_cap
    for i, nbytes in enumerate(sizes_bytes):
        cur.append(i)
        total += nbytes
        if total >= limit:
            buckets.append(cur)
            cur, total, limit = [], 0, cap
    if cur:
        buckets.append(cur)
    return buckets


# --------------------------------------------------------------------------------------------
# a8 / §8f row 1: the optimizer step the reduced gradients feed   (reference: optimizer.py:145-181 ->
# torch/optim/adam.py `_multi_tensor_adam`, non-capturable branch; AdamW = decoupled weight decay)
# --------------------------------------------------------------------------------------------
def _fma(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64; the sum is rounded to float64 and
    then to float32 (double rounding can differ from a true fma in the last bit, rarely: the tests use a tolerance)."""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64) + np.asarray(c, dtype=np.float64)).astype(F32)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decoupled=True, mult=1.0):
    """One Adam/AdamW update in float32, operation by operation as torch's foreach implementation does it on CUDA
    (``a + alpha*b`` inside one functor is a single fma there).  ``step`` is the 1-based step count AFTER the
    increment.  Returns new (p, m, v)."""
    p, g, m, v = (np.asarray(a, dtype=F32).copy() for a in (p, g, m, v))
    g = g * F32(mult)
    if weight_decay != 0:
        if decoupled:
            p = p * F32(1 - lr * weight_decay)
        else:
            g = _fma(F32(weight_decay), p, g)
    w = F32(1 - beta1)
    diff = g - m
    m = _fma(w, diff, m) if abs(float(w)) < 0.5 else g - diff * (F32(1.0) - w)
    v = v * F32(beta2)
    v = _fma(F32(1 - beta2), g * g, v)
    bc1 = 1 - beta1**step
    bc2_sqrt = (1 - beta2**step) ** 0.5
    denom = np.sqrt(v) / F32(bc2_sqrt) + F32(eps)
    p = _fma(F32(-(lr / bc1)), m / denom, p)
    return p, m, v


# --------------------------------------------------------------------------------------------
# one whole step of the path, for smoke tests and the CPU baseline
# --------------------------------------------------------------------------------------------
def reduce_and_clip_step(per_rank_flat_grads, max_norm=1.0):
    """mean-reduce a flat gradient vector over ranks, then clip by global 2-norm."""
    reduced = ddp_mean_reduce(per_rank_flat_grads)
    clipped, tn = clip_grad_norm([reduced], max_norm)
    return clipped[0], tn
