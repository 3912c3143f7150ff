"""ctypes view of oracle/liboracle.so (ORACLE — test infrastructure only; see oracle.c)."""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        lib = ctypes.CDLL(_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.oracle_ddp_mean_reduce_f32.argtypes = [ctypes.POINTER(fp), ctypes.c_int, ctypes.c_size_t, fp]
        lib.oracle_sumsq_f32.argtypes = [fp, ctypes.c_size_t]
        lib.oracle_sumsq_f32.restype = ctypes.c_double
        lib.oracle_maxabs_f32.argtypes = [fp, ctypes.c_size_t]
        lib.oracle_maxabs_f32.restype = ctypes.c_float
        lib.oracle_has_nonfinite_f32.argtypes = [fp, ctypes.c_size_t]
        lib.oracle_clip_coef.argtypes = [ctypes.c_float, ctypes.c_float]
        lib.oracle_clip_coef.restype = ctypes.c_float
        lib.oracle_inv_scale.argtypes = [ctypes.c_float]
        lib.oracle_inv_scale.restype = ctypes.c_float
        lib.oracle_scale_f32.argtypes = [fp, ctypes.c_size_t, ctypes.c_float]
        lib.oracle_clamp_f32.argtypes = [fp, ctypes.c_size_t, ctypes.c_float]
        d = ctypes.c_double
        lib.oracle_adam_step_f32.argtypes = [fp, fp, fp, fp, ctypes.c_size_t, d, d, d, d, d, d, d, ctypes.c_float]
        lib.oracle_adam_step_f32.restype = None
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def ddp_mean_reduce(per_rank):
    lib = load()
    arrs = [np.ascontiguousarray(g, dtype=np.float32).reshape(-1) for g in per_rank]
    out = np.empty_like(arrs[0])
    ptrs = (ctypes.POINTER(ctypes.c_float) * len(arrs))(*[_fp(a) for a in arrs])
    lib.oracle_ddp_mean_reduce_f32(ptrs, len(arrs), arrs[0].size, _fp(out))
    return out.reshape(np.shape(per_rank[0]))


def clip_grad_norm_inplace(grads, max_norm):
    """grads: list of writable contiguous float32 arrays (e.g. views of torch CPU grads). -> total_norm"""
    lib = load()
    total = 0.0
    for g in grads:
        total += lib.oracle_sumsq_f32(_fp(g), g.size)
    tn = np.float32(np.sqrt(total))
    coef = lib.oracle_clip_coef(tn, np.float32(max_norm))
    for g in grads:
        lib.oracle_scale_f32(_fp(g), g.size, coef)
    return tn


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decoupled=True, mult=1.0):
    """Same contract as ``ddp_oracle.adam_step`` (returns new p, m, v) with TRUE single-rounding fmas (C ``fmaf``),
    i.e. the exact roundings of torch's CUDA foreach functors and of ``b200dp_multi_adam``."""
    lib = load()
    p, m, v = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1).copy() for a in (p, m, v))
    g = np.ascontiguousarray(g, dtype=np.float32).reshape(-1)
    decay_mul = (1 - lr * weight_decay) if (decoupled and weight_decay != 0) else 1.0
    l2 = weight_decay if (not decoupled and weight_decay != 0) else 0.0
    lib.oracle_adam_step_f32(_fp(p), _fp(g), _fp(m), _fp(v), p.size, decay_mul, l2, beta1, beta2,
                             -(lr / (1 - beta1**step)), (1 - beta2**step) ** 0.5, eps, np.float32(mult))  # fmt: skip
    return p, m, v
