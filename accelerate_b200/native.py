"""ctypes binding of ``libb200dp.so`` (C ABI declared in ``include/b200dp.h``).

This is the only place the Python host layer touches native code.  There is deliberately no
fallback: if the shared library is missing or a call fails, a :class:`B200dpError` is raised.  The
library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``), which
``__graft_entry__.build()`` calls.
"""

from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libb200dp.so")
CSRC_DIR = os.path.join(_HERE, "csrc")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

# --- constants mirrored from include/b200dp.h (tests/test_cabi.py checks they agree) ---
MAX_WORLD = 16
MAX_BLOCKS = 304
MAX_SLOTS = 1024
SMALL_MAX = 128
MT_MAX = 640
F32, BF16, F16, F64, I32, I64, U8 = range(7)
HANDLE_CUDAIPC, HANDLE_POSIXFD = 0, 1
CH_GRAD, CH_USER, CH_SMALL, CH_PUSH = 0, 1, 2, 3
ALGO_AUTO, ALGO_TWOSHOT, ALGO_NVLS, ALGO_ONESHOT = 0, 1, 2, 3
F_STATS = 1
F_NO_END_BARRIER = 2
Q_LAUNCHES, Q_MULTICAST_OK, Q_VMM_OK, Q_STATUS, Q_SM_COUNT, Q_COMM_STREAM = range(6)
ETIMEOUT = -4
NO_STREAM = ctypes.c_void_p(-1).value  # B200DP_NO_STREAM: stream 0 is a real stream, so NULL cannot mean 'none'

EXPORTS = [
    "b200dp_version", "b200dp_last_error", "b200dp_init", "b200dp_finalize", "b200dp_query",
    "b200dp_set_timeout_ms", "b200dp_reset_status", "b200dp_seg_create", "b200dp_seg_export",
    "b200dp_seg_import", "b200dp_seg_peer_ptr", "b200dp_seg_destroy", "b200dp_mc_create",
    "b200dp_mc_attach", "b200dp_mc_bind", "b200dp_seg_mc_ptr", "b200dp_control_bytes",
    "b200dp_set_control", "b200dp_fdx_open", "b200dp_fdx_send", "b200dp_fdx_recv", "b200dp_fdx_close",
    "b200dp_stream_wait", "b200dp_fence", "b200dp_allreduce_bucket", "b200dp_allreduce_oneshot", "b200dp_finalize_norm",
    "b200dp_scale_flat", "b200dp_clamp_flat", "b200dp_cast", "b200dp_multi_stats", "b200dp_multi_scale", "b200dp_multi_adam", "b200dp_pack",
    "b200dp_pull", "b200dp_gather", "b200dp_reduce", "b200dp_peer_copy", "b200dp_push_max_bytes", "b200dp_push_area_bytes", "b200dp_push", "b200dp_exchange_small",
    "b200dp_pad_copy",
]  # fmt: skip


class B200dpError(RuntimeError):
    """A native call failed (no CPU fallback exists for the CUDA data path)."""


def nvcc_path() -> str | None:
    cand = shutil.which("nvcc")
    if cand:
        return cand
    for p in ("/usr/local/cuda/bin/nvcc",):
        if os.path.exists(p):
            return p
    return None


def _sources():
    return [os.path.join(CSRC_DIR, f) for f in ("b200dp.cu", "kernels.cuh", "device_utils.cuh")] + [
        os.path.join(INCLUDE_DIR, "b200dp.h")
    ]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/b200dp.cu for sm_100a into accelerate_b200/lib/libb200dp.so (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = nvcc_path()
    if nvcc is None:
        raise B200dpError("nvcc not found: cannot build libb200dp.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
        "-o", tmp, os.path.join(CSRC_DIR, "b200dp.cu"), "-ldl", "-lpthread", "-lrt",
    ]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout, res.stderr)
    if res.returncode != 0:
        raise B200dpError(f"nvcc failed ({res.returncode}):\n{res.stderr[-4000:]}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None
_lib_lock = threading.Lock()


def load():
    """Load the shared library and declare argument types.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200dpError(
                f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc). The CUDA data path has no fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        vp, sz, i = c_void_p, c_size_t, c_int
        lib.b200dp_version.restype = i
        lib.b200dp_last_error.restype = c_char_p
        lib.b200dp_control_bytes.restype = sz
        lib.b200dp_push_max_bytes.restype = sz
        lib.b200dp_push_area_bytes.restype = sz
        sig = {
            "b200dp_init": [i, i, i, POINTER(vp)],
            "b200dp_finalize": [vp],
            "b200dp_query": [vp, i, POINTER(c_longlong)],
            "b200dp_set_timeout_ms": [vp, c_longlong],
            "b200dp_reset_status": [vp],
            "b200dp_seg_create": [vp, sz, i, POINTER(i), POINTER(vp), POINTER(sz), POINTER(i)],
            "b200dp_seg_export": [vp, i, c_char_p, POINTER(i)],
            "b200dp_seg_import": [vp, i, i, i, c_char_p, i],
            "b200dp_seg_peer_ptr": [vp, i, i, POINTER(vp)],
            "b200dp_seg_destroy": [vp, i],
            "b200dp_mc_create": [vp, i, POINTER(i)],
            "b200dp_mc_attach": [vp, i, i],
            "b200dp_mc_bind": [vp, i],
            "b200dp_seg_mc_ptr": [vp, i, POINTER(vp)],
            "b200dp_set_control": [vp, i],
            "b200dp_fdx_open": [vp, c_char_p],
            "b200dp_fdx_send": [vp, c_char_p, i, i],
            "b200dp_fdx_recv": [vp, POINTER(i), POINTER(i), i],
            "b200dp_fdx_close": [vp],
            "b200dp_stream_wait": [vp, vp, vp],
            "b200dp_fence": [vp, i, vp],
            "b200dp_allreduce_bucket": [vp, i, sz, sz, i, c_float, i, i, i, i, i, vp, vp],
            "b200dp_allreduce_oneshot": [vp, i, sz, sz, i, c_double, vp, i, i, vp],
            "b200dp_finalize_norm": [vp, i, i, i, c_double, c_float, vp, vp, vp],
            "b200dp_scale_flat": [vp, vp, sz, i, vp, i, vp],
            "b200dp_clamp_flat": [vp, vp, sz, i, c_float, i, vp],
            "b200dp_cast": [vp, vp, i, vp, i, sz, c_float, i, vp],
            "b200dp_multi_stats": [vp, POINTER(vp), POINTER(sz), i, i, i, vp],
            "b200dp_multi_scale": [vp, POINTER(vp), POINTER(sz), i, i, vp, vp],
            "b200dp_multi_adam": [vp, POINTER(vp), POINTER(vp), POINTER(sz), POINTER(sz), i, vp, vp, c_double, c_double,
                                  c_double, c_double, c_double, c_double, c_double, vp, vp, vp],
            "b200dp_pack": [vp, POINTER(vp), POINTER(vp), POINTER(sz), i, i, vp],
            "b200dp_pull": [vp, i, sz, sz, vp, sz, i, i, i, i, vp],
            "b200dp_gather": [vp, i, sz, vp, sz, vp, sz, i, i, vp],
            "b200dp_peer_copy": [vp, i, i, sz, vp, sz, vp],
            "b200dp_reduce": [vp, i, sz, vp, sz, i, c_double, vp, i, i, i, vp],
            "b200dp_push": [vp, i, i, vp, sz, i, c_double, vp, i, vp],
            "b200dp_exchange_small": [vp, vp, i, vp, vp],
            "b200dp_pad_copy": [vp, vp, vp, sz, sz, sz, sz, i, c_uint64, i, vp],
        }
        for name, argtypes in sig.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = i
        _lib = lib
    return _lib


def _check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.b200dp_last_error().decode(errors="replace")
        raise B200dpError(f"{what} failed (code {rc}): {msg}")


class Context:
    """One native context per process/GPU.  Thin, 1:1 with the C ABI."""

    def __init__(self, rank: int, world: int, device: int):
        self.lib = load()
        self.rank, self.world, self.device = rank, world, device
        h = c_void_p()
        _check(self.lib, self.lib.b200dp_init(rank, world, device, byref(h)), "b200dp_init")
        self.h = h
        self._closed = False

    # -- lifecycle --
    def close(self):
        if not self._closed and self.h:
            self.lib.b200dp_finalize(self.h)
            self._closed = True

    def query(self, key: int) -> int:
        out = c_longlong()
        _check(self.lib, self.lib.b200dp_query(self.h, key, byref(out)), "b200dp_query")
        return out.value

    @property
    def launches(self) -> int:
        return self.query(Q_LAUNCHES)

    @property
    def comm_stream(self) -> int:
        return self.query(Q_COMM_STREAM)

    def set_timeout_ms(self, ms: int):
        _check(self.lib, self.lib.b200dp_set_timeout_ms(self.h, int(ms)), "b200dp_set_timeout_ms")

    def reset_status(self):
        _check(self.lib, self.lib.b200dp_reset_status(self.h), "b200dp_reset_status")

    def check_status(self):
        st = self.query(Q_STATUS)
        if st != 0:
            raise B200dpError(
                f"device-side status {st}: a cross-GPU flag wait timed out (a peer rank hung, died, or issued "
                "collectives in a different order)"
            )

    # -- segments --
    def seg_create(self, nbytes: int, prefer_kind: int):
        seg, ptr, padded, kind = c_int(), c_void_p(), c_size_t(), c_int()
        _check(
            self.lib,
            self.lib.b200dp_seg_create(self.h, nbytes, prefer_kind, byref(seg), byref(ptr), byref(padded), byref(kind)),
            "b200dp_seg_create",
        )
        return seg.value, ptr.value, padded.value, kind.value

    def seg_export(self, seg: int):
        buf = ctypes.create_string_buffer(64)
        fd = c_int(-1)
        _check(self.lib, self.lib.b200dp_seg_export(self.h, seg, buf, byref(fd)), "b200dp_seg_export")
        return buf.raw, fd.value

    def seg_import(self, seg: int, peer: int, kind: int, handle: bytes, fd: int):
        _check(self.lib, self.lib.b200dp_seg_import(self.h, seg, peer, kind, handle, fd), "b200dp_seg_import")

    def seg_peer_ptr(self, seg: int, peer: int) -> int:
        p = c_void_p()
        _check(self.lib, self.lib.b200dp_seg_peer_ptr(self.h, seg, peer, byref(p)), "b200dp_seg_peer_ptr")
        return p.value

    def seg_destroy(self, seg: int):
        _check(self.lib, self.lib.b200dp_seg_destroy(self.h, seg), "b200dp_seg_destroy")

    def mc_create(self, seg: int) -> int:
        fd = c_int(-1)
        _check(self.lib, self.lib.b200dp_mc_create(self.h, seg, byref(fd)), "b200dp_mc_create")
        return fd.value

    def mc_attach(self, seg: int, fd: int):
        _check(self.lib, self.lib.b200dp_mc_attach(self.h, seg, fd), "b200dp_mc_attach")

    def mc_bind(self, seg: int):
        _check(self.lib, self.lib.b200dp_mc_bind(self.h, seg), "b200dp_mc_bind")

    def seg_mc_ptr(self, seg: int) -> int:
        p = c_void_p()
        _check(self.lib, self.lib.b200dp_seg_mc_ptr(self.h, seg, byref(p)), "b200dp_seg_mc_ptr")
        return p.value or 0

    def control_bytes(self) -> int:
        return self.lib.b200dp_control_bytes()

    def set_control(self, seg: int):
        _check(self.lib, self.lib.b200dp_set_control(self.h, seg), "b200dp_set_control")

    # -- fd passing --
    def fdx_open(self, name: str):
        _check(self.lib, self.lib.b200dp_fdx_open(self.h, name.encode()), "b200dp_fdx_open")

    def fdx_send(self, peer_name: str, fd: int, tag: int):
        _check(self.lib, self.lib.b200dp_fdx_send(self.h, peer_name.encode(), fd, tag), "b200dp_fdx_send")

    def fdx_recv(self, timeout_ms: int = 60000):
        fd, tag = c_int(-1), c_int(0)
        _check(self.lib, self.lib.b200dp_fdx_recv(self.h, byref(fd), byref(tag), timeout_ms), "b200dp_fdx_recv")
        return fd.value, tag.value

    def fdx_close(self):
        self.lib.b200dp_fdx_close(self.h)

    # -- data path --
    def stream_wait(self, waiter: int, signaller: int):
        _check(self.lib, self.lib.b200dp_stream_wait(self.h, waiter, signaller), "b200dp_stream_wait")

    def fence(self, channel: int, stream: int):
        _check(self.lib, self.lib.b200dp_fence(self.h, channel, stream), "b200dp_fence")

    def allreduce_bucket(self, seg, elem_off, count, dtype, prescale, slot, flags, algo, channel, max_blocks,
                         after_stream, stream):  # fmt: skip
        _check(
            self.lib,
            self.lib.b200dp_allreduce_bucket(self.h, seg, elem_off, count, dtype, prescale, slot, flags, algo, channel,
                                             max_blocks, after_stream, stream),
            "b200dp_allreduce_bucket",
        )  # fmt: skip

    def allreduce_oneshot(self, seg, byte_off, count, dtype, divisor, dst, channel, max_blocks, stream):
        _check(
            self.lib,
            self.lib.b200dp_allreduce_oneshot(self.h, seg, byte_off, count, dtype, divisor, dst, channel, max_blocks, stream),
            "b200dp_allreduce_oneshot",
        )

    def finalize_norm(self, slot0, nslots, all_ranks, norm_type, max_norm, scale_ptr, out_ptr, stream):
        _check(
            self.lib,
            self.lib.b200dp_finalize_norm(self.h, slot0, nslots, int(all_ranks), float(norm_type), float(max_norm),
                                          scale_ptr, out_ptr, stream),
            "b200dp_finalize_norm",
        )  # fmt: skip

    def scale_flat(self, ptr, count, dtype, mult_ptr, max_blocks, stream):
        _check(self.lib, self.lib.b200dp_scale_flat(self.h, ptr, count, dtype, mult_ptr, max_blocks, stream), "b200dp_scale_flat")

    def clamp_flat(self, ptr, count, dtype, v, max_blocks, stream):
        _check(self.lib, self.lib.b200dp_clamp_flat(self.h, ptr, count, dtype, float(v), max_blocks, stream), "b200dp_clamp_flat")

    def cast(self, src, src_dtype, dst, dst_dtype, count, divisor, max_blocks, stream):
        _check(self.lib, self.lib.b200dp_cast(self.h, src, src_dtype, dst, dst_dtype, count, float(divisor), max_blocks, stream),
               "b200dp_cast")  # fmt: skip

    def multi_stats(self, ptrs, numels, dtype, slot, stream):
        n = len(ptrs)
        pa = (c_void_p * n)(*ptrs)
        na = (c_size_t * n)(*numels)
        _check(self.lib, self.lib.b200dp_multi_stats(self.h, pa, na, n, dtype, slot, stream), "b200dp_multi_stats")
        return max(1, (n + MT_MAX - 1) // MT_MAX)

    def multi_scale(self, ptrs, numels, dtype, mult_ptr, stream):
        n = len(ptrs)
        pa = (c_void_p * n)(*ptrs)
        na = (c_size_t * n)(*numels)
        _check(self.lib, self.lib.b200dp_multi_scale(self.h, pa, na, n, dtype, mult_ptr, stream), "b200dp_multi_scale")

    def make_adam_table(self, params, numels, offs, grads=None):
        """ctypes arrays for multi_adam.  With ``grads`` (fusion-buffer views: stable pointers) the table is complete
        and can be cached; without, the gradient pointers are supplied per step."""
        n = len(params)
        ga = (c_void_p * n)(*grads) if grads is not None else None
        return (c_void_p * n)(*params), ga, (c_size_t * n)(*numels), (c_size_t * n)(*offs), n

    def multi_adam(self, table, exp_avg, exp_avg_sq, decay_mul, l2, beta1, beta2, neg_step_size, bc2_sqrt, eps,
                   mult_ptr, found_inf_ptr, stream, grads=None):  # fmt: skip
        pa, ga, na, oa, n = table
        if grads is not None:
            ga = (c_void_p * n)(*grads)
        _check(
            self.lib,
            self.lib.b200dp_multi_adam(self.h, pa, ga, na, oa, n, exp_avg, exp_avg_sq, decay_mul, l2, beta1, beta2,
                                       neg_step_size, bc2_sqrt, eps, mult_ptr, found_inf_ptr, stream),
            "b200dp_multi_adam",
        )  # fmt: skip

    def pack(self, srcs, dsts, nbytes, use_bulk, stream):
        """``srcs`` / ``dsts`` / ``nbytes``: Python sequences, or ctypes arrays built once (the reducer caches the
        destination pointers and sizes of every bucket: only the sources change from step to step)."""
        n = len(srcs)
        sa = srcs if isinstance(srcs, ctypes.Array) else (c_void_p * n)(*srcs)
        da = dsts if isinstance(dsts, ctypes.Array) else (c_void_p * n)(*dsts)
        ba = nbytes if isinstance(nbytes, ctypes.Array) else (c_size_t * n)(*nbytes)
        _check(self.lib, self.lib.b200dp_pack(self.h, sa, da, ba, n, int(use_bulk), stream), "b200dp_pack")

    def pull(self, seg, byte_off, nbytes, dst, dst_stride, first, nranks, channel, max_blocks, stream):
        _check(
            self.lib,
            self.lib.b200dp_pull(self.h, seg, byte_off, nbytes, dst, dst_stride, first, nranks, channel, max_blocks, stream),
            "b200dp_pull",
        )

    def peer_copy(self, seg, peer, byte_off, dst, nbytes, stream):
        _check(self.lib, self.lib.b200dp_peer_copy(self.h, seg, peer, byte_off, dst, nbytes, stream), "b200dp_peer_copy")

    def gather(self, seg, byte_off, src, nbytes, dst, dst_stride, channel, max_blocks, stream):
        _check(self.lib, self.lib.b200dp_gather(self.h, seg, byte_off, src, nbytes, dst, dst_stride, channel, max_blocks, stream),
               "b200dp_gather")  # fmt: skip

    def reduce(self, seg, byte_off, src, count, dtype, divisor, dst, algo, channel, max_blocks, stream):
        _check(self.lib, self.lib.b200dp_reduce(self.h, seg, byte_off, src, count, dtype, float(divisor), dst, algo, channel,
                                                max_blocks, stream), "b200dp_reduce")  # fmt: skip

    def push(self, seg, mode, src, nbytes, dtype, divisor, dst, channel, stream):
        _check(self.lib, self.lib.b200dp_push(self.h, seg, mode, src, nbytes, dtype, float(divisor), dst, channel, stream), "b200dp_push")

    def push_max_bytes(self) -> int:
        return self.lib.b200dp_push_max_bytes()

    def push_area_bytes(self) -> int:
        return self.lib.b200dp_push_area_bytes()

    def exchange_small(self, payload: bytes, stream) -> list[bytes]:
        n = len(payload)
        src = ctypes.create_string_buffer(payload, n)
        out = ctypes.create_string_buffer(n * self.world)
        _check(self.lib, self.lib.b200dp_exchange_small(self.h, src, n, out, stream), "b200dp_exchange_small")
        raw = out.raw
        return [raw[r * n : (r + 1) * n] for r in range(self.world)]

    def pad_copy(self, src, dst, outer, old_dim, new_dim, inner_bytes, elem_size, pad_bits, pad_first, stream):
        _check(
            self.lib,
            self.lib.b200dp_pad_copy(self.h, src, dst, outer, old_dim, new_dim, inner_bytes, elem_size, pad_bits,
                                     int(pad_first), stream),
            "b200dp_pad_copy",
        )  # fmt: skip
