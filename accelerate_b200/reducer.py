"""B200DataParallel: the gradient reducer that replaces ``torch.nn.parallel.DistributedDataParallel``
behind ``Accelerator.prepare`` (reference boundary: ``src/accelerate/accelerator.py:1877-1896``).

What torch's DDP + c10d::Reducer + NCCL do per sync-backward:

    per parameter : bucket_view <- grad * (1/N)                    (292 small kernels, GPT-2-medium)
    per bucket    : ncclAllReduce(SUM) on a side stream             (37 NCCL launches)
    end of bwd    : grad.copy_(bucket_view)                         (292 small kernels)
    later         : clip_grad_norm_ = foreach_norm + stack + norm + foreach_mul (reads grads twice)

What this reducer does instead:

* all gradients of one dtype live in ONE symmetric, peer-mapped *fusion buffer*; after the first
  reduction ``param.grad`` IS a view into it (no copy-back, ever);
* a gradient that autograd produced elsewhere (fresh ``.grad`` after ``zero_grad(set_to_none=True)``)
  is moved in by ONE multi-tensor pack launch per bucket (``b200dp_pack``, TMA bulk copies);
* ONE fused kernel per bucket (``b200dp_allreduce_bucket``) does pre-divide, the allreduce over
  NVLink by direct peer loads/stores (or NVLS multimem), and accumulates sum-of-squares / max-abs /
  non-finite statistics of the reduced values, so ``clip_grad_norm_`` and ``GradScaler.unscale_``
  need no read pass of their own;
* results are bit-identical on all ranks and equal to ``sum_r fl(g_r * 1/N)`` summed in rank order
  (the oracle's definition, ``oracle/ddp_oracle.py``; NVLS sums in the switch and scales after,
  which is the same value whenever N is a power of two).

Semantics mirrored from torch DDP because the reference's tests pin them: the
decision to reduce is latched at *forward* time from ``require_backward_grad_sync``; ``no_sync()``
only flips that flag; ``reducer.prepare_for_backward([])`` re-arms; buckets launch in index order;
parameters are broadcast from rank 0 at wrap time; buffers are re-broadcast each training forward
when ``broadcast_buffers=True``.

On a CPU device (``DistributedType.MULTI_CPU``, gloo) the same bucket/hook logic runs with
``torch.distributed.all_reduce`` moving the bytes — that is the reference's own CPU behaviour and
exists so the host logic is testable without a GPU.  It is selected by the *device type* of the
module, never as a fallback: a CUDA module without the native library raises.
"""

from __future__ import annotations

import os
from contextlib import contextmanager

import torch
import torch.distributed as dist
from torch.autograd import Variable
from torch.distributed.algorithms.join import Join, Joinable, JoinHook

import sys
import threading
import weakref

from . import lazy_grad, native
from .lazy_grad import raw_grad, raw_set_grad

_ALIGN = 32  # bytes: every parameter slice / bucket starts on a 32-byte boundary (16 B in the half-width wire copy)
_LIVE = weakref.WeakSet()
# debugging switches (diagnostics only; defaults are the product configuration)
_DEBUG = os.environ.get("B200DP_DEBUG", "0") == "1"
_KEEP_STAGED = os.environ.get("B200DP_KEEP_STAGED", "0") == "1"
_SAME_STREAM = os.environ.get("B200DP_SAME_STREAM", "0") == "1"
_PACK_ON_COMM = os.environ.get("B200DP_PACK_ON_COMM", "1") == "1"
_NVLS_MIN_WORLD = int(os.environ.get("B200DP_NVLS_MIN_WORLD", "4"))
# NVLS CTAs per bucket kernel, 0 = by world size: a rank streams S/N of every bucket through its SMs, so the CTA count
# that saturates the links shrinks with N.  8 GPUs: 8 CTAs give the same bus bandwidth as 16 on 25 MiB buckets (665 vs
# 657 GB/s) with half the SMs taken from the backward kernels (+0.55 % samples/s in r01 and again in r02:
# profiles/r02_bench_n8_nvls_16ctas.json vs _8ctas.json); 4 GPUs: 8 CTAs lose (441 vs 543 GB/s, 1.42 GB launch 529 vs
# 684: profiles/r02_bench_n4_nvls_8ctas.json vs r02_bench_n4.json), so 16 there.
_NVLS_BLOCKS = int(os.environ.get("B200DP_NVLS_BLOCKS", "0"))
_TWOSHOT_BLOCKS = int(os.environ.get("B200DP_TWOSHOT_BLOCKS", "32"))


def live_reducers():
    """Reducers currently alive in this process (the GradScaler and clip paths look them up)."""
    return list(_LIVE)


def compute_bucket_assignment(sizes_bytes, keys, first_cap: int, cap: int):
    """Bucket indices the way ``torch.distributed._compute_bucket_assignment_by_size`` does it.

    Walk tensors in the given order; tensors of one (dtype, device) key fill that key's open
    bucket; a bucket closes as soon as its bytes reach the key's current limit (``first_cap`` for a
    key's first bucket, then ``cap``); tensors are never split.  Buckets are returned ordered by
    their smallest member index.  (``tests/test_reducer_host.py`` checks this against torch.)
    """
    open_b: dict = {}
    limit_idx: dict = {}
    limits = [first_cap, cap]
    done = []
    for i, (nbytes, key) in enumerate(zip(sizes_bytes, keys)):
        members, total = open_b.get(key, ([], 0))
        members = members + [i]
        total += nbytes
        li = limit_idx.setdefault(key, 0)
        if total >= limits[li]:
            done.append(members)
            open_b[key] = ([], 0)
            if li + 1 < len(limits):
                limit_idx[key] = li + 1
        else:
            open_b[key] = (members, total)
    for members, _ in open_b.values():
        if members:
            done.append(members)
    done.sort(key=lambda m: min(m))
    return done


class GradBucket:
    __slots__ = (
        "index", "group", "elem_off", "count", "params", "offsets", "numels", "views", "ptrs", "nbytes",
        "pending", "seen", "staged", "slot", "flat", "c_ptrs", "c_nbytes", "fired", "trigger", "handles", "missed",
    )  # fmt: skip

    def __init__(self, index, group):
        self.index, self.group = index, group
        self.params, self.offsets, self.numels, self.views, self.ptrs, self.nbytes = [], [], [], [], [], []
        self.elem_off = self.count = 0
        self.pending = 0
        self.seen = []
        self.staged = []
        self.slot = -1
        self.flat = None
        self.c_ptrs = self.c_nbytes = None  # ctypes arrays of ptrs / nbytes, built once (CUDA)
        self.fired = []  # fast hooks: parameters whose gradient was accumulated in this backward (list.append IS the hook)
        self.trigger = -1  # index of the parameter expected to fire last (learned); its hook is the only Python one
        self.handles = []
        self.missed = False

    def reset(self):
        self.pending = len(self.params)
        for i in range(len(self.seen)):
            self.seen[i] = False
        self.fired.clear()  # never rebound: `fired.append` is registered as a hook


class HookBucket:
    """What a user communication hook receives: the methods of ``torch.distributed.GradBucket`` that Python hooks
    use (torch's own ``default_hooks`` / ``powerSGD_hook`` included), backed by a bucket of the fusion buffer.
    As in torch, ``buffer()`` is the flat bucket BEFORE any division: averaging is the hook's job."""

    __slots__ = ("_bucket", "_buffer", "_last")

    def __init__(self, bucket: GradBucket, last: bool):
        self._bucket, self._buffer, self._last = bucket, bucket.flat, last

    def index(self) -> int:
        return self._bucket.index

    def buffer(self) -> torch.Tensor:
        return self._buffer

    def set_buffer(self, tensor: torch.Tensor) -> None:
        self._buffer = tensor

    def gradients(self):
        return list(self._bucket.views)

    def parameters(self):
        return list(self._bucket.params)

    def is_last(self) -> bool:
        return self._last


class _DtypeGroup:
    """All gradients of one dtype: one flat fusion buffer (symmetric segment on CUDA)."""

    def __init__(self, dtype):
        self.dtype = dtype
        self.elem_size = torch.empty((), dtype=dtype).element_size()
        self.total_elems = 0
        self.segment = None  # runtime.Segment (CUDA)
        self.wire = None  # half-width symmetric copy used when a compress hook is registered
        self.wire_flat = None
        self.wire_native = None
        self.flat = None  # 1-D tensor over the whole buffer
        self.native_dtype = None


class GradReducer:
    """Bucket table + autograd hooks + launch ordering.  ``prepare_for_backward`` mirrors the one
    entry point the reference reaches into (``accelerator.py:1219-1222``)."""

    def __init__(self, owner, params, device, world, rank, bucket_cap_bytes, first_bucket_bytes,
                 find_unused_parameters, process_group):  # fmt: skip
        self.owner = owner
        self.device, self.world, self.rank = device, world, rank
        self.process_group = process_group
        self.find_unused_parameters = find_unused_parameters
        self.is_cuda = device.type == "cuda"
        self.params = params
        self._armed = False
        self._cb_queued = False
        self._next = 0
        self.stats_valid = False
        self.unscaled = False
        self.wire_dtype = None
        self.num_reductions = 0  # sync-backwards completed (tests / bench introspection)
        self.compute_stream = 0
        self._graveyard = []
        self._comm_dirty = False
        self._wire_dirty = False
        self._pending_mult = None  # lazy_grad.PendingScale of a deferred clip
        self._stats_versions = []
        self._views_epoch = -1
        self._closed = False
        self._locally_undefined = []  # (bucket, i) whose gradient was None when an unused parameter was filled in
        self._shadowing = False  # inside shadow_backward (a joined rank standing in for a backward it does not run)
        self.user_hook = None  # (state, hook): a user-supplied communication algorithm replaces the fused kernel
        self._hook_futures = []
        self.rt = None
        if self.is_cuda:
            from .runtime import SymmRuntime, native_dtype

            self.rt = SymmRuntime.get()  # raises if the native runtime is not up: no fallback
            if lazy_grad.enabled():
                lazy_grad.install()  # observe Python-level .grad accesses from the start (deferred clip, epoch fast paths)
            self._native_dtype = native_dtype
            self.comm_stream = self.rt.ctx.comm_stream
            # CTAs per bucket kernel (0 = pick by algorithm).  Measured on 8xB200 (profiles/r01_sweep_n8_u4.json,
            # r01_allreduce_n8.json): thanks to 4-16 x 16-byte requests in flight per thread, 16 CTAs (NVLS) /
            # 32 CTAs (two-shot) stream as fast as 148 with a LOWER fixed cost per bucket, and they leave >= 116
            # SMs to the backward kernels the reduction overlaps with.
            self.max_blocks = int(os.environ.get("B200DP_REDUCE_BLOCKS", "0"))
            algo = os.environ.get("B200DP_ALGO", "auto").lower()
            self.algo = {"auto": native.ALGO_AUTO, "twoshot": native.ALGO_TWOSHOT, "nvls": native.ALGO_NVLS}[algo]
        for p in params:
            if p.device != device:
                raise ValueError(f"parameter on {p.device}, expected {device}: all parameters must sit on the process's device")
            if p.is_sparse or getattr(p, "layout", torch.strided) != torch.strided:
                raise ValueError("sparse parameters are not supported by the B200 reducer")

        # ---- bucket table: reverse parameter order ~ gradient-ready order (what DDP starts from) ----
        order = list(range(len(params)))[::-1]
        sizes = [params[i].numel() * params[i].element_size() for i in order]
        keys = [params[i].dtype for i in order]
        assignment = compute_bucket_assignment(sizes, keys, first_bucket_bytes, bucket_cap_bytes)
        self.groups: dict = {}
        self.buckets: list[GradBucket] = []
        for members in assignment:
            dtype = keys[members[0]]
            group = self.groups.setdefault(dtype, _DtypeGroup(dtype))
            b = GradBucket(len(self.buckets), group)
            es = group.elem_size
            align_e = max(1, _ALIGN // es)
            b.elem_off = group.total_elems
            cursor = group.total_elems
            for m in members:
                p = params[order[m]]
                b.params.append(p)
                b.offsets.append(cursor)
                b.numels.append(p.numel())
                b.nbytes.append(p.numel() * es)
                cursor = (cursor + p.numel() + align_e - 1) // align_e * align_e
            b.count = cursor - b.elem_off
            group.total_elems = cursor
            b.seen = [False] * len(b.params)
            self.buckets.append(b)
        if len(self.buckets) > 512:
            raise ValueError(f"{len(self.buckets)} buckets: raise bucket_cap_mb (at most 512 buckets are supported)")

        # ---- fusion buffers ----
        for group in self.groups.values():
            nbytes = max(group.total_elems * group.elem_size, _ALIGN)
            if self.is_cuda:
                group.native_dtype = self._native_dtype(group.dtype)
                if group.native_dtype not in (native.F32, native.BF16, native.F16):
                    raise TypeError(f"gradient dtype {group.dtype} is not supported by the fused allreduce kernel")
                group.segment = self.rt.alloc_segment(nbytes, multicast=True)
                group.flat = group.segment.tensor(group.dtype, group.total_elems)
            else:
                group.flat = torch.zeros(group.total_elems, dtype=group.dtype, device=device)
        for b in self.buckets:
            b.flat = b.group.flat[b.elem_off : b.elem_off + b.count]
            for p, off, n in zip(b.params, b.offsets, b.numels):
                v = b.group.flat[off : off + n].view(p.shape)
                b.views.append(v)
                b.ptrs.append(v.data_ptr())
        if self.is_cuda:
            import ctypes

            self.slot_base = self.rt.reserve_slots(len(self.buckets))
            for b in self.buckets:
                b.slot = self.slot_base + b.index
                b.c_ptrs = (ctypes.c_void_p * len(b.ptrs))(*b.ptrs)
                b.c_nbytes = (ctypes.c_size_t * len(b.nbytes))(*b.nbytes)
        self._managed_ids = frozenset(id(p) for p in params)
        self._param_index = {id(p): k for k, p in enumerate(params)}
        # Hooks run on the autograd thread, and the backward of an eager transformer is host-bound (the engine enqueues
        # slower than the GPU executes: profiles/r02_timeline_n2.txt), so every microsecond here is a microsecond of step
        # time.  Fast hooks: the hook of all parameters but one per bucket is the C-level `list.append` of the bucket's
        # `fired` list; only the parameter expected to fire LAST (learned from the previous iteration) runs Python.
        self._fast = (not find_unused_parameters) and os.environ.get("B200DP_FAST_HOOKS", "1") != "0"
        self._relearn = []
        for b in self.buckets:
            if self._fast:
                b.trigger = len(b.params) - 1  # bucket order is reverse registration order ~ gradient-ready order
            self._install_bucket_hooks(b)
        self._reset_iteration()
        _LIVE.add(self)

    # ------------------------------------------------------------------ hooks
    def _make_hook(self, bucket: GradBucket, i: int):
        ptr, view, seen = bucket.ptrs[i], bucket.views[i], bucket.seen

        def hook(param):
            if not self._armed:
                # a gradient changed outside a reduction (second backward without a forward, no_sync micro-step):
                # the statistics of the last reduction no longer describe the buffer
                self.stats_valid = False
                return
            if not self._cb_queued:
                # end-of-backward callback of the graph task this hook runs in.  Under reentrant activation
                # checkpointing that may be a NESTED task; `_on_graph_task_end` then declines and the next hook
                # (in the enclosing task) queues again.
                self._cb_queued = True
                Variable._execution_engine.queue_callback(self._on_graph_task_end)
            if seen[i]:
                return  # same parameter fired twice in one backward (retain_graph replays): count once
            seen[i] = True
            g = raw_grad(param)
            if g is not view and g is not None and g.data_ptr() != ptr:
                # autograd put this gradient somewhere else: adopt it into the fusion buffer
                bucket.staged.append((i, g if g.is_contiguous() else g.contiguous()))
                raw_set_grad(param, view)
            bucket.pending -= 1
            if bucket.pending == 0:
                self._launch_ready()

        return hook

    def _install_bucket_hooks(self, b: GradBucket):
        for h in b.handles:
            h.remove()
        b.handles = []
        for i, p in enumerate(b.params):
            if not self._fast:
                fn = self._make_hook(b, i)
            elif i == b.trigger:
                fn = self._make_trigger(b)
            else:
                fn = b.fired.append  # returns None, takes the parameter: a valid post-accumulate-grad hook at C speed
            b.handles.append(p.register_post_accumulate_grad_hook(fn))

    def _make_trigger(self, bucket: GradBucket):
        fired = bucket.fired

        def trigger(param):
            fired.append(param)
            if not self._armed:
                self.stats_valid = False  # gradients changed outside a reduction (see `hook`)
                fired.clear()
                return
            if not self._cb_queued:
                self._cb_queued = True
                Variable._execution_engine.queue_callback(self._on_graph_task_end)
            self._launch_ready()
            if bucket.index >= self._next and not bucket.missed:
                # the expected-last parameter was not the last one this time: the bucket goes out with a later trigger
                # (or at the end of backward); learn the real order for the next iteration
                bucket.missed = True
                self._relearn.append(bucket)

        return trigger

    def _complete(self, b: GradBucket) -> bool:
        """Fast hooks: has every parameter of ``b`` fired in this backward?  If so adopt the gradients (what the
        per-parameter Python hook does one by one) and mark the bucket ready."""
        n = len(b.params)
        fired = b.fired
        if len(fired) < n or len({id(p) for p in fired}) < n:  # (a parameter may fire twice: count distinct ones)
            return False
        if b.missed:
            last = fired[-1]
            b.trigger = next(i for i, p in enumerate(b.params) if p is last)
        staged, ptrs, views = b.staged, b.ptrs, b.views
        for i, p in enumerate(b.params):
            g = raw_grad(p)
            if g is not views[i] and g is not None and g.data_ptr() != ptrs[i]:
                staged.append((i, g if g.is_contiguous() else g.contiguous()))
                raw_set_grad(p, views[i])
        seen = b.seen
        for i in range(n):
            seen[i] = True
        b.pending = 0
        return True

    def _launch_ready(self):
        buckets = self.buckets
        while self._next < len(buckets):  # strictly in index order
            b = buckets[self._next]
            if b.pending != 0 and not (self._fast and self._complete(b)):
                break
            self._reduce_bucket(b)
            self._next += 1

    # ------------------------------------------------------------------ the data path
    def _reduce_bucket(self, b: GradBucket):
        if self.is_cuda:
            ctx = self.rt.ctx
            cs = self.compute_stream
            comm = cs if _SAME_STREAM else self.comm_stream
            packed_on_comm = False
            if b.staged:
                srcs = [g.data_ptr() for _, g in b.staged]
                if len(b.staged) == len(b.ptrs) and all(i == k for k, (i, _) in enumerate(b.staged)):
                    dsts, sizes = b.c_ptrs, b.c_nbytes  # the usual case: every gradient of the bucket is fresh, in order
                else:
                    dsts = [b.ptrs[i] for i, _ in b.staged]
                    sizes = [b.nbytes[i] for i, _ in b.staged]
                if _PACK_ON_COMM and comm != cs and self.user_hook is None:
                    # adopt the gradients on the SIDE stream: the pack overlaps with the rest of backward instead
                    # of sitting on the compute stream.  The source tensors must then outlive the pack, so they
                    # are parked until the end-of-backward join (after which the compute stream is ordered behind
                    # the side stream and the allocator may hand the memory out again).
                    ctx.stream_wait(comm, cs)
                    ctx.pack(srcs, dsts, sizes, self.rt.bulk_for(b.count * b.group.elem_size), comm)
                    self._graveyard.extend(g for _, g in b.staged)
                    pack_stream = comm
                    packed_on_comm = True
                else:
                    ctx.pack(srcs, dsts, sizes, self.rt.bulk_for(b.count * b.group.elem_size), cs)
                    pack_stream = cs
                if _DEBUG:
                    print(f"[b200dp r{self.rank}] bucket {b.index}: pack {len(srcs)} spans on stream {pack_stream:#x}: "
                          f"{[(hex(s_), hex(d_), n_) for s_, d_, n_ in zip(srcs, dsts, sizes)]}", flush=True)
                if _KEEP_STAGED:
                    self._graveyard.extend(g for _, g in b.staged)
                b.staged.clear()  # same-stream pack: stream order protects the memory; side-stream pack: parked above
            if self.user_hook is not None:
                self._run_user_hook(b)
                return
            if self.wire_dtype is not None and b.group.wire is not None:
                self._reduce_bucket_wire(b, cs)
                return
            seg = b.group.segment
            # measured on B200 (profiles/): with 2 ranks the switch has nothing to save and plain peer
            # loads/stores stream faster; from 4 ranks on NVLS moves ~S instead of 2(N-1)/N*S per GPU
            algo = self._pick_algo(seg)
            # (a pack on the side stream already ordered that stream behind the compute stream)
            after = native.NO_STREAM if (packed_on_comm or comm == cs) else cs
            ctx.allreduce_bucket(seg.seg, b.elem_off, b.count, b.group.native_dtype, 1.0 / self.world, b.slot,
                                 native.F_STATS | native.F_NO_END_BARRIER, algo, native.CH_GRAD, self.blocks_for(algo), after,
                                 comm)  # fmt: skip
            self._comm_dirty = True
        else:
            for i, g in b.staged:
                b.views[i].copy_(g)
            b.staged.clear()
            if self.user_hook is not None:
                self._run_user_hook(b)
                return
            if self.wire_dtype is not None and b.flat.dtype != self.wire_dtype:
                # torch's compress hook, for every bucket whatever its dtype: buffer.to(dtype).div_(N) -> allreduce ->
                # copy back (default_hooks._compress_hook).  (The CUDA path compresses fp32 buckets only; half-precision
                # parameters under the OTHER half-precision hook keep their own dtype on the wire there.)
                narrow = b.flat.to(self.wire_dtype).div_(self.world)
                if self.world > 1:
                    dist.all_reduce(narrow, op=dist.ReduceOp.SUM, group=self.process_group)
                b.flat.copy_(narrow)
                return
            # same arithmetic as DDP: pre-divide, then SUM (reducer.hpp:276,524)
            b.flat.mul_(1.0 / self.world)
            if self.world > 1:
                dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.process_group)

    def set_user_hook(self, state, hook):
        """torch's operator plug point (``DistributedDataParallel.register_comm_hook``; reached from
        dataclasses.py:234-237): ``hook(state, bucket) -> Future[Tensor]`` is called once per bucket per sync
        backward, in bucket order on every rank, with the un-divided flat bucket; the future's value becomes the
        bucket's content.  The fused kernel (and its statistics) are out of the picture for this wrapper."""
        if self.user_hook is not None or self.wire_dtype is not None:
            raise RuntimeError("register_comm_hook can only be called once.")  # torch's rule and wording
        if not callable(hook):
            raise TypeError("Communication hook must be callable.")
        self.user_hook = (state, hook)

    def _run_user_hook(self, b: GradBucket):
        state, hook = self.user_hook
        fut = hook(state, HookBucket(b, b.index == len(self.buckets) - 1))
        if not isinstance(fut, torch._C.Future):  # c10d work futures are the base class of torch.futures.Future
            raise TypeError("Communication hook: return annotation should be torch.futures.Future.")
        self._hook_futures.append((b, fut))

    def _join_user_hooks(self):
        futures, self._hook_futures = self._hook_futures, []
        for b, fut in futures:
            result = fut.wait()
            if isinstance(result, (list, tuple)):  # c10d work futures resolve to a list of tensors
                result = result[0]
            if result.data_ptr() != b.flat.data_ptr() or result.dtype != b.flat.dtype:
                b.flat.copy_(result.reshape(-1))  # torch Reducer: copy the hook's result into the bucket views

    def _pick_algo(self, seg):
        algo = self.algo
        if algo == native.ALGO_AUTO:
            # the switch adds the N raw copies and the kernel scales the sum; that equals the oracle's
            # sum_r fl(g_r * 1/N) exactly only when 1/N is a power of two, so other world sizes keep the two-shot
            pow2 = self.world & (self.world - 1) == 0
            algo = native.ALGO_NVLS if (seg.multicast and pow2 and self.world >= _NVLS_MIN_WORLD) else native.ALGO_TWOSHOT
        return algo

    def blocks_for(self, algo: int) -> int:
        if self.max_blocks > 0:
            return self.max_blocks
        if algo == native.ALGO_NVLS:
            return _NVLS_BLOCKS if _NVLS_BLOCKS > 0 else (8 if self.world >= 8 else 16)
        return _TWOSHOT_BLOCKS

    def _reduce_bucket_wire(self, b: GradBucket, cs: int):
        """Half-precision wire (the reference's fp16/bf16 compress hooks): cast+pre-divide the bucket into the
        symmetric wire buffer, allreduce THERE (half the NVLink bytes, fp32 accumulation, one rounding), and
        convert the whole buffer back once, after the fence, in ``_finalize_backward``."""
        ctx, g = self.rt.ctx, b.group
        comm = cs if _SAME_STREAM else self.comm_stream
        ctx.stream_wait(comm, cs)
        ctx.cast(g.flat.data_ptr() + b.elem_off * g.elem_size, g.native_dtype, g.wire_flat.data_ptr() + b.elem_off * 2,
                 g.wire_native, b.count, float(self.world), 0, comm)  # fmt: skip
        algo = self._pick_algo(g.wire)
        ctx.allreduce_bucket(g.wire.seg, b.elem_off, b.count, g.wire_native, 1.0, b.slot,
                             native.F_STATS | native.F_NO_END_BARRIER, algo, native.CH_GRAD, self.blocks_for(algo),
                             native.NO_STREAM, comm)  # fmt: skip
        self._comm_dirty = True
        self._wire_dirty = True

    def set_wire_dtype(self, dtype):
        """Collective: allocate the half-width symmetric wire buffers (fp32 gradient groups only)."""
        if dtype not in (torch.bfloat16, torch.float16):
            raise ValueError(f"wire dtype must be torch.bfloat16 or torch.float16, got {dtype}")
        if self.user_hook is not None:
            raise RuntimeError("register_comm_hook can only be called once.")
        self.wire_dtype = dtype
        if not self.is_cuda:
            return
        for g in self.groups.values():
            if g.dtype != torch.float32 or g.wire is not None:
                continue
            g.wire = self.rt.alloc_segment(max(g.total_elems * 2, _ALIGN), multicast=True)
            g.wire_flat = g.wire.tensor(dtype, g.total_elems)
            g.wire_native = self._native_dtype(dtype)

    @staticmethod
    def _inside_nested_backward() -> bool:
        """True when the graph task whose final callbacks are running was started from inside another backward
        (``torch.utils.checkpoint(use_reentrant=True)`` re-runs autograd from a Function's backward).  The engine
        offers no query for that, but the Python stack does: a nested task is driven by a ``torch.autograd.backward``
        call, so an ``_engine_run_backward`` frame sits below this callback — one more than the outermost task has
        (none on the engine's own device threads, one on a Python thread that called ``backward`` itself)."""
        code = torch.autograd.graph._engine_run_backward.__code__
        depth, frame = 0, sys._getframe(1)
        while frame is not None:
            if frame.f_code is code:
                depth += 1
            frame = frame.f_back
        baseline = 0 if isinstance(threading.current_thread(), threading._DummyThread) else 1
        return depth > baseline

    def _on_graph_task_end(self):
        """Final callback of the graph task in which a gradient hook fired (torch's Reducer queues its finalizer when
        the last bucket is ready, reducer.cpp mark_variable_ready; we also need the end of a backward that left
        parameters unused, so the callback is queued early and decides here)."""
        if not self._armed:
            return
        if self._next < len(self.buckets) and self._inside_nested_backward():
            self._cb_queued = False  # the enclosing backward continues: its next hook queues the callback again
            return
        self._finalize_backward()

    def finish_open_backward(self):
        """Safety net for a backward whose last gradient hook ran in a nested graph task that could not know it was
        the last (reentrant checkpointing + parameters left unused): finish the reduction before anybody consumes
        the gradients.  A no-op in every ordinary run."""
        if self._armed and not self._cb_queued and any(any(b.seen) or b.fired for b in self.buckets):
            self._finalize_backward()

    def _finalize_backward(self):
        """Runs once per sync backward, after its last gradient."""
        try:
            if self._next < len(self.buckets) and self._fast:
                self._launch_ready()  # buckets whose learned trigger fired early are complete by now
            if self._next < len(self.buckets):
                self._handle_unready()
            if self._hook_futures:
                self._join_user_hooks()
            if self.is_cuda:
                # a device-side flag wait that timed out (dead or diverged peer) latches a status word in mapped
                # host memory; surface it here instead of training on with unreduced gradients
                self.rt.ctx.check_status()
                if self._comm_dirty:
                    # bucket kernels skip their closing barrier; ONE fence per backward proves that every
                    # peer's stores into this rank's fusion buffer (and statistics records) have landed
                    comm = self.compute_stream if _SAME_STREAM else self.comm_stream
                    self.rt.ctx.fence(native.CH_GRAD, comm)
                    self._comm_dirty = False
                    if self._wire_dirty:  # every peer's reduced halves have landed: widen back to fp32, one launch
                        for g in self.groups.values():
                            if g.wire is not None:
                                self.rt.ctx.cast(g.wire_flat.data_ptr(), g.wire_native, g.flat.data_ptr(), g.native_dtype,
                                                 g.total_elems, 1.0, 0, comm)  # fmt: skip
                        self._wire_dirty = False
                # the optimizer (compute stream) must see the reduced gradients
                if not _SAME_STREAM:
                    self.rt.ctx.stream_wait(self.compute_stream, self.comm_stream)
                self._graveyard.clear()
                self.stats_valid = self.user_hook is None  # a user hook bypasses the kernel that gathers them
                self.unscaled = False
                # the native kernels do not touch torch's version counters; anything that writes a gradient through
                # torch afterwards (p.grad.mul_(), a second backward accumulating into the views) does
                self._stats_versions = [g.flat._version for g in self.groups.values()]
                # every gradient is now its fusion-buffer view unless a parameter stayed undefined; until somebody assigns a
                # .grad from Python (observed through lazy_grad's property) the per-parameter pointer walk can be skipped
                self._views_epoch = lazy_grad.assign_epoch if not self._locally_undefined and not self.find_unused_parameters else -1
            if self.find_unused_parameters:
                self._restore_globally_unused()
            self.num_reductions += 1
        finally:
            self._armed = False
            self._hook_futures = []
            self._locally_undefined = []
            self._reset_iteration()

    def _handle_unready(self):
        if self._fast:  # the per-parameter bookkeeping lives in `fired`
            for b in self.buckets[self._next :]:
                ids = {id(p) for p in b.fired}
                for i, p in enumerate(b.params):
                    if id(p) in ids and not b.seen[i]:
                        g = raw_grad(p)
                        if g is not b.views[i] and g is not None and g.data_ptr() != b.ptrs[i]:
                            b.staged.append((i, g if g.is_contiguous() else g.contiguous()))
                            raw_set_grad(p, b.views[i])
                        b.seen[i] = True
                        b.pending -= 1
        missing = [(b, i) for b in self.buckets[self._next :] for i in range(len(b.params)) if not b.seen[i]]
        if missing and not self.find_unused_parameters:
            names = {id(p): n for n, p in self.owner.module.named_parameters()}
            some = ", ".join(names.get(id(b.params[i]), "?") for b, i in missing[:8])
            raise RuntimeError(
                f"{len(missing)} parameter(s) did not receive a gradient in this backward pass (e.g. {some}). "
                "Wrap with DistributedDataParallelKwargs(find_unused_parameters=True) if parts of the model are "
                "conditionally unused."
            )
        # torch Reducer (mark_variable_ready for an unused parameter): a defined gradient — e.g. one accumulated under
        # no_sync — goes into the bucket as it is; an undefined one contributes zeros, and stays undefined afterwards
        # if NO rank used the parameter (`_restore_globally_unused`)
        for b, i in missing:
            p, view = b.params[i], b.views[i]
            g = raw_grad(p)
            if g is None:
                view.zero_()
                self._locally_undefined.append((b, i))
                raw_set_grad(p, view)
            elif g.data_ptr() != b.ptrs[i]:
                b.staged.append((i, g if g.is_contiguous() else g.contiguous()))
                raw_set_grad(p, view)
            b.seen[i] = True
            b.pending -= 1
        self._launch_ready()

    def _restore_globally_unused(self):
        """``find_unused_parameters=True``: a parameter that no rank used in this backward keeps ``grad=None`` (torch
        Reducer::finalize_bucket_dense, "globally unused").  Every rank takes part in one small integer allreduce per
        sync backward, exactly like DDP's ``local_used_map`` reduction."""
        undefined, self._locally_undefined = self._locally_undefined, []
        if self.world > 1:
            flags = [0 if self._shadowing else 1] * len(self.params)  # a joined rank used nothing
            index = self._param_index
            for b, i in undefined:
                flags[index[id(b.params[i])]] = 0
            used = torch.tensor(flags, dtype=torch.int32, device=self.device)
            if self.is_cuda:
                from .collectives import all_reduce

                used = all_reduce(used)
            else:
                dist.all_reduce(used, op=dist.ReduceOp.SUM, group=self.process_group)
            used = used.tolist()
            undefined = [(b, i) for b, i in undefined if used[index[id(b.params[i])]] == 0]
        for b, i in undefined:
            raw_set_grad(b.params[i], None)

    def shadow_backward(self):
        """A rank that ran out of data (torch ``Join``) stands in for one sync backward of the others: it contributes
        zeros to every bucket, in bucket order, and takes part in every collective a real backward would issue
        (torch ``_match_all_reduce_for_bwd_pass`` / ``_match_unused_params_allreduce``).  Unlike torch, the in-place
        reduction leaves the OTHER ranks' mean gradient in this rank's fusion buffer."""
        self.prepare_for_backward([])
        self._shadowing = True
        try:
            with torch.no_grad():
                for group in self.groups.values():
                    group.flat.zero_()
            for b in self.buckets:
                b.pending = 0
                b.staged.clear()
                for i in range(len(b.seen)):
                    b.seen[i] = True
            self._launch_ready()
            self._finalize_backward()
        finally:
            self._shadowing = False

    def _reset_iteration(self):
        self._next = 0
        self._cb_queued = False
        for b in self.buckets:
            b.reset()
        if self._relearn:
            relearn, self._relearn = self._relearn, []
            for b in relearn:
                b.missed = False
                self._install_bucket_hooks(b)  # b.trigger was updated when the bucket completed

    # ------------------------------------------------------------------ API the wrapper / Accelerator use
    def prepare_for_backward(self, unused=None):
        """Arm the next backward (torch Reducer::prepare_for_backward; reached from accelerator.py:1222)."""
        self.finish_open_backward()
        if self._pending_mult is not None:
            # a deferred clip that no fused optimizer step consumed: the gradients may be kept (accumulation without
            # zero_grad), so they must carry the coefficient before new ones are added — what the reference's
            # in-place clip left behind
            self._pending_mult.apply()
            self._pending_mult = None
        self._armed = True
        self.stats_valid = False
        if self.is_cuda:
            self.compute_stream = torch.cuda.current_stream(self.device).cuda_stream
        self._reset_iteration()

    def grads_in_buffer(self) -> bool:
        """Every managed parameter's ``.grad`` is (still) its view into the fusion buffer."""
        if self._views_epoch == lazy_grad.assign_epoch and lazy_grad.installed():
            return True
        for b in self.buckets:
            for p, ptr in zip(b.params, b.ptrs):
                g = raw_grad(p)
                if g is None or g.data_ptr() != ptr:
                    return False
        return True

    def stats_fresh(self) -> bool:
        """The statistics of the last reduction still describe the fusion buffer: nothing wrote a gradient through
        torch since (the views share their version counter with the flat buffer)."""
        return (self.is_cuda and self.stats_valid
                and self._stats_versions == [g.flat._version for g in self.groups.values()])  # fmt: skip

    def effective_algo(self) -> str:
        """"nvls" or "twoshot": what the bucket kernels of this reducer actually run."""
        if not self.is_cuda:
            return "gloo"
        algo = self._pick_algo(next(iter(self.groups.values())).segment)
        return "nvls" if algo == native.ALGO_NVLS else "twoshot"

    def manages(self, parameters) -> bool:
        """True iff ``parameters`` (those with grads) are exactly this reducer's parameters and their
        gradients live in its buffers — the precondition for using the fused statistics."""
        if self._views_epoch == lazy_grad.assign_epoch and lazy_grad.installed():
            return {id(p) for p in parameters} == self._managed_ids  # all gradients known to be the buffer's views
        ids = {id(p) for p in parameters if raw_grad(p) is not None}
        return ids == self._managed_ids and self.grads_in_buffer()

    def fused_clip_(self, max_norm: float, norm_type: float = 2.0, scale: torch.Tensor | None = None, defer: bool = False):
        """clip_grad_norm_ and/or GradScaler.unscale_ from the statistics the allreduce kernels already
        produced: one 1-CTA finalize + one multiply pass over the fusion buffer that exits at once
        when the multiplier is 1.  Returns a 3-float device tensor ``[total_norm, multiplier,
        found_inf]``.

        * ``scale=None``: plain clip (bf16 / fp32 runs).
        * ``scale=S, max_norm<0``: unscale only (``g *= 1/S``); statistics stay usable for a later clip.
        * ``scale=S, max_norm>=0``: unscale and clip in one pass — or, if an unscale already ran, clip
          only (the norm is still derived from the scaled statistics: ``total = total_scaled / S``).
        """
        if not (self.is_cuda and self.stats_valid):
            raise RuntimeError("fused_clip_ needs fresh reduction statistics (call right after a sync backward)")
        ctx = self.rt.ctx
        st = torch.cuda.current_stream(self.device).cuda_stream
        out = torch.empty(3, dtype=torch.float32, device=self.device)
        scale_ptr = scale.data_ptr() if scale is not None else None
        ctx.finalize_norm(self.slot_base, len(self.buckets), True, norm_type, max_norm, scale_ptr, out.data_ptr(), st)
        # an earlier unscale-only multiplier that nobody materialised yet is superseded: `out[1]` = 1/S * coef
        earlier, self._pending_mult = self._pending_mult, None
        still_scaled = earlier is not None and earlier.consume()
        if scale is not None and self.unscaled and not still_scaled:
            # gradients already carry 1/S: keep only the clip coefficient (exact for power-of-two S)
            out[1].mul_(scale.to(torch.float32).reshape(()))
        if defer:
            # fused optimizer consumption: the multiplier stays pending; the fused Adam kernel
            # applies it on the fly, anything else that looks at a gradient first triggers the pass (lazy_grad.py)
            self._pending_mult = lazy_grad.defer(out, lambda: self._apply_multiplier(out), owner=self, param_ids=self._managed_ids)
        else:
            self._apply_multiplier(out)
        if scale is not None:
            self.unscaled = True
        if max_norm >= 0:
            self.stats_valid = False  # the buffer no longer holds the values the statistics describe
        return out

    def _apply_multiplier(self, out: torch.Tensor):
        """One multiply pass over the gradients in the fusion buffer by ``out[1]`` (device-side; exits at once when it
        is 1).  Multi-tensor over the parameter spans, NOT the flat buffer: the alignment padding between spans must
        stay zero whatever the multiplier is (0 * NaN after an fp16 overflow would poison every later bucket's
        statistics)."""
        if self._closed:
            return
        st = torch.cuda.current_stream(self.device).cuda_stream
        for group in self.groups.values():
            spans = self._spans_of(group)
            self.rt.ctx.multi_scale(spans[0], spans[1], group.native_dtype, out.data_ptr() + 4, st)

    def _spans_of(self, group):
        """(pointers, element counts) of every parameter's slice of ``group``'s fusion buffer (cached)."""
        cache = self.__dict__.setdefault("_span_cache", {})
        if group.dtype not in cache:
            ptrs, numels = [], []
            for b in self.buckets:
                if b.group is group:
                    ptrs.extend(b.ptrs)
                    numels.extend(b.numels)
            cache[group.dtype] = (ptrs, numels)
        return cache[group.dtype]

    def take_pending_multiplier(self):
        """The pending [total_norm, multiplier, found_inf] tensor of a deferred clip / unscale, handed to the fused
        optimizer step exactly once (None if there is none or it was already applied)."""
        pending, self._pending_mult = self._pending_mult, None
        if pending is None or not pending.consume():
            return None
        return pending.out

    def clamp_(self, clip_value: float):
        if self._pending_mult is not None:
            self._pending_mult.apply()
            self._pending_mult = None
        st = torch.cuda.current_stream(self.device).cuda_stream
        for group in self.groups.values():
            self.rt.ctx.clamp_flat(group.flat.data_ptr(), group.total_elems, group.native_dtype, clip_value, 0, st)
        self.stats_valid = False

    def can_average_in_buffer(self) -> bool:
        """The fusion buffer may be borrowed only while no live gradient is a view into it."""
        if not self.is_cuda or self._armed:
            return False
        for b in self.buckets:
            for p, ptr in zip(b.params, b.ptrs):
                g = raw_grad(p)
                if g is not None and g.data_ptr() == ptr:
                    return False
        return True

    def average_parameters_(self):
        """LocalSGD: params <- mean over ranks, through the fusion buffer (pack, ONE allreduce, unpack, divide)."""
        if self.world == 1:
            return
        ctx = self.rt.ctx
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.no_grad():
            for group in self.groups.values():
                mine = [(p, ptr, nb) for b in self.buckets if b.group is group for p, ptr, nb in zip(b.params, b.ptrs, b.nbytes)]
                srcs, slots, sizes = [p.data_ptr() for p, _, _ in mine], [ptr for _, ptr, _ in mine], [nb for _, _, nb in mine]
                bulk = self.rt.bulk_for(sum(sizes))
                ctx.pack(srcs, slots, sizes, bulk, st)
                algo = self._pick_algo(group.segment)
                ctx.allreduce_bucket(group.segment.seg, 0, group.total_elems, group.native_dtype, 1.0, -1, 0, algo,
                                     native.CH_USER, self.blocks_for(algo), native.NO_STREAM, st)  # fmt: skip
                ctx.pack(slots, srcs, sizes, bulk, st)
                torch._foreach_div_([p.data for p, _, _ in mine], float(self.world))  # the reference's `/= N`
                group.flat.zero_()

    def broadcast_parameters(self, module_params, src: int = 0):
        """Make every rank's parameters equal rank ``src``'s through the fusion buffers: pack, pull, unpack."""
        if self.world == 1:
            return
        if not self.is_cuda:
            for p in module_params:
                dist.broadcast(p.data, src=src, group=self.process_group)
            return
        ctx = self.rt.ctx
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.no_grad():
            for group in self.groups.values():
                mine = [(p, ptr, nb) for b in self.buckets if b.group is group for p, ptr, nb in zip(b.params, b.ptrs, b.nbytes)]
                bulk = self.rt.bulk_for(sum(nb for _, _, nb in mine))
                if self.rank == src:
                    ctx.pack([p.data_ptr() for p, _, _ in mine], [ptr for _, ptr, _ in mine], [nb for _, _, nb in mine],
                             bulk, st)  # fmt: skip
                nbytes = group.total_elems * group.elem_size
                ctx.pull(group.segment.seg, 0, nbytes, group.flat.data_ptr(), 0, src, 1, native.CH_USER, 0, st)
                for p, _, _ in mine:
                    if not p.is_contiguous():
                        raise ValueError("non-contiguous parameters are not supported by the B200 reducer")
                ctx.pack([ptr for _, ptr, _ in mine], [p.data_ptr() for p, _, _ in mine], [nb for _, _, nb in mine],
                         bulk, st)  # fmt: skip
                group.flat.zero_()

    def remove_hooks(self):
        for b in self.buckets:
            for h in b.handles:
                h.remove()
            b.handles = []

    def close(self):
        """Release everything the reducer took from the runtime: autograd hooks, the symmetric fusion / wire
        segments and the statistics slots (``Accelerator.free_memory``; the reference's DDP wrapper frees its buckets
        when it is garbage collected).  Gradients that were views of the fusion buffer are detached into ordinary
        tensors first.  Every rank is expected to close its reducers in the same order (slot numbering is
        positional), like every other collective-ordering rule of this class."""
        if self._closed:
            return
        self.remove_hooks()
        if self._pending_mult is not None:
            self._pending_mult.apply()
            self._pending_mult = None
        self._closed = True
        self._armed = False
        self.stats_valid = False
        _LIVE.discard(self)
        if not self.is_cuda:
            return
        with torch.no_grad():
            for b in self.buckets:
                for p, ptr in zip(b.params, b.ptrs):
                    g = raw_grad(p)
                    if g is not None and g.data_ptr() == ptr:
                        raw_set_grad(p, g.clone())
        torch.cuda.synchronize(self.device)
        for group in self.groups.values():
            group.flat = None
            group.wire_flat = None
            for seg_name in ("segment", "wire"):
                seg = getattr(group, seg_name)
                if seg is not None:
                    self.rt.free_segment(seg)
                    setattr(group, seg_name, None)
        for b in self.buckets:
            b.flat, b.views = None, []
        self.rt.release_slots(self.slot_base, len(self.buckets))


class _B200JoinHook(JoinHook):
    """What a rank does, once per iteration of the others, after it ran out of inputs inside ``torch``'s ``Join``
    context (reference: accelerator.py:1299-1372 ``join_uneven_inputs`` -> torch ``_DDPJoinHook``): shadow the buffer
    broadcast of the forward, learn whether the others reduce in this iteration, and if so shadow the reduction."""

    def __init__(self, ddp, divide_by_initial_world_size: bool):
        if not divide_by_initial_world_size:
            raise NotImplementedError(
                "divide_by_initial_world_size=False is not supported: the fused allreduce pre-divides by the world size"
            )
        self.ddp = ddp
        super().__init__()

    def main_hook(self):
        ddp = self.ddp
        if ddp._will_sync_buffers():
            ddp._broadcast_buffers(ddp._find_common_rank(False))
        should_sync = ddp._global_requires_backward_grad_sync(is_joined_rank=True)
        ddp.require_forward_param_sync = should_sync
        if should_sync:
            ddp.reducer.shadow_backward()

    def post_hook(self, is_last_joiner: bool):
        self.ddp._sync_final_model(is_last_joiner)


class B200DataParallel(torch.nn.Module, Joinable):
    """Drop-in for the ``DistributedDataParallel(model, device_ids=[local_rank], output_device=local_rank,
    **DistributedDataParallelKwargs.to_kwargs())`` call at ``accelerator.py:1892-1894``; exposes the
    attributes the reference and its tests read."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0, broadcast_buffers=True,
                 process_group=None, bucket_cap_mb=None, find_unused_parameters=False, check_reduction=False,
                 gradient_as_bucket_view=False, static_graph=False):  # fmt: skip
        super().__init__()
        Joinable.__init__(self)
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise RuntimeError("B200DataParallel is not needed when a module doesn't have any parameter that requires a gradient.")
        self.module = module
        device = params[0].device
        self.device_ids = [torch.device(d).index if not isinstance(d, int) else d for d in device_ids] if device_ids else None
        self.output_device = output_device
        self.dim = dim
        self.broadcast_buffers = broadcast_buffers
        self.find_unused_parameters = find_unused_parameters
        self.gradient_as_bucket_view = gradient_as_bucket_view
        self.static_graph = static_graph
        self.process_group = process_group
        self.require_backward_grad_sync = True
        self.require_forward_param_sync = True
        world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        # torch: bucket_cap_mb=None -> 25 MiB cap and a 1 MiB first bucket; explicit value -> both equal
        if bucket_cap_mb is None:
            self.bucket_bytes_cap_default = True
            self.bucket_bytes_cap = 25 * 1024 * 1024
            first = 1024 * 1024
        else:
            self.bucket_bytes_cap_default = False
            self.bucket_bytes_cap = int(bucket_cap_mb * 1024 * 1024)
            first = self.bucket_bytes_cap
        self.reducer = GradReducer(self, params, device, world, rank, self.bucket_bytes_cap, first,
                                   find_unused_parameters, process_group)  # fmt: skip
        # buffer list is fixed at wrap time, like torch DDP's `modules_buffers` (walking ~600 modules every
        # forward costs ~0.5 ms of host time in a launch-bound step)
        self._ddp_buffers = [b for b in module.buffers() if b is not None]
        # rank 0's weights (and buffers) become everyone's (torch: _sync_module_states)
        self.reducer.broadcast_parameters(params)
        self._broadcast_frozen(0)
        self._sync_buffers()

    # ---- DDP surface ----
    @contextmanager
    def no_sync(self):
        previous = self.require_backward_grad_sync
        self.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.require_backward_grad_sync = previous

    def register_comm_hook(self, state, hook):
        """``DistributedDataParallel.register_comm_hook`` (torch distributed.py; reached from dataclasses.py:234-237).

        * ``DDPCommunicationHookType.FP16/BF16`` — or torch's own ``fp16_compress_hook`` / ``bf16_compress_hook``
          functions — select the native half-precision wire (cast kernel + the fused allreduce on a half-width buffer);
        * torch's ``allreduce_hook`` is what the fused kernel already computes: nothing to change;
        * any other callable ``hook(state, bucket) -> Future[Tensor]`` (PowerSGD, user code) is run per bucket through
          the ``GradBucket`` protocol and owns the communication, typically over ``torch.distributed``.
        """
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks

        from .utils.dataclasses import DDPCommunicationHookType

        native_wire = {
            DDPCommunicationHookType.FP16: torch.float16, DDPCommunicationHookType.BF16: torch.bfloat16,
            default_hooks.fp16_compress_hook: torch.float16, default_hooks.bf16_compress_hook: torch.bfloat16,
        }  # fmt: skip
        if isinstance(hook, DDPCommunicationHookType) and hook not in native_wire:
            raise ValueError(f"{hook} names a hook family; pass the hook callable (DistributedDataParallelKwargs does)")
        if hook in native_wire:
            self.reducer.set_wire_dtype(native_wire[hook])
        elif hook is default_hooks.allreduce_hook:
            if self.reducer.user_hook is not None or self.reducer.wire_dtype is not None:
                raise RuntimeError("register_comm_hook can only be called once.")
        else:
            self.reducer.set_user_hook(state, hook)

    def _broadcast_frozen(self, src: int):
        """torch's ``_sync_module_states`` makes EVERY parameter equal to the source rank's, trainable or not; the
        frozen ones have no slot in the fusion buffer and go through the staged broadcast."""
        frozen = [p.data for p in self.module.parameters() if not p.requires_grad]
        if frozen and self.reducer.world > 1:
            from .utils.operations import broadcast

            with torch.no_grad():
                broadcast(frozen, from_process=src)

    def _will_sync_buffers(self) -> bool:
        return bool(self._ddp_buffers) and self.broadcast_buffers and self.require_forward_param_sync

    def _broadcast_buffers(self, src: int):
        if not self._ddp_buffers or self.reducer.world == 1:
            return
        from .utils.operations import broadcast

        with torch.no_grad():
            broadcast(self._ddp_buffers, from_process=src)

    def _sync_buffers(self):
        # inside torch's Join context rank 0 may already have run out of data: agree on a rank that is still
        # training (torch distributed.py:2214-2231)
        self._broadcast_buffers(self._find_common_rank(True) if self._join_config.enable else 0)

    # ---- uneven inputs: torch.distributed.algorithms.Join protocol ----
    def join_hook(self, **kwargs):
        return _B200JoinHook(self, divide_by_initial_world_size=kwargs.get("divide_by_initial_world_size", True))

    def join(self, divide_by_initial_world_size: bool = True, enable: bool = True, throw_on_early_termination: bool = False):
        """``with model.join():`` — torch DDP's shorthand for ``Join([model], ...)`` (distributed.py:1798-1870)."""
        return Join([self], enable, throw_on_early_termination, divide_by_initial_world_size=divide_by_initial_world_size)

    @property
    def join_device(self):
        return self.reducer.device

    @property
    def join_process_group(self):
        return self.process_group if self.process_group is not None else dist.group.WORLD

    def _find_common_rank(self, candidate: bool) -> int:
        """The largest rank for which ``candidate`` holds (one MAX allreduce, like torch's ``_find_common_rank``)."""
        t = torch.tensor([self.reducer.rank if candidate else -1], device=self.reducer.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.process_group)
        rank = int(t.item())
        if rank == -1:
            raise ValueError("no rank qualifies as the authoritative copy: every rank joined at once?")
        return rank

    def _global_requires_backward_grad_sync(self, is_joined_rank: bool):
        """Ranks still training announce whether their coming backward reduces; joined ranks listen."""
        flag = torch.ones(1, device=self.reducer.device) if (not is_joined_rank and self.require_backward_grad_sync) \
            else torch.zeros(1, device=self.reducer.device)  # fmt: skip
        work = dist.all_reduce(flag, group=self.process_group, async_op=True)
        if is_joined_rank:
            work.wait()
            return flag.item() != 0
        return None

    def _sync_final_model(self, is_last_joiner: bool):
        """Leaving the Join context: everyone adopts the model of the rank that trained longest."""
        src = self._find_common_rank(is_last_joiner)
        self.reducer.broadcast_parameters(self.reducer.params, src=src)
        self._broadcast_frozen(src)
        if self.broadcast_buffers:
            self._broadcast_buffers(src)

    def forward(self, *inputs, **kwargs):
        if self.device_ids:  # torch DDP moves inputs to its device (_pre_forward -> _to_kwargs)
            from .utils.operations import send_to_device

            dev = torch.device(self.reducer.device.type, self.device_ids[0])
            inputs = send_to_device(inputs, dev)
            kwargs = send_to_device(kwargs, dev)
        # torch DDP, _pre_forward / _post_forward (distributed.py:1553-1650): buffers are re-broadcast before a forward
        # iff the PREVIOUS forward was a gradient-syncing one (require_forward_param_sync), whatever the train/eval
        # mode; the sync decision for the coming backward is latched here, from require_backward_grad_sync
        # 
        joining = self._join_config.enable
        if joining:
            Join.notify_join_context(self)  # "this rank is still training" (one allreduce per iteration)
        if self._will_sync_buffers():
            self._sync_buffers()
        if joining:
            self._global_requires_backward_grad_sync(is_joined_rank=False)
        output = self.module(*inputs, **kwargs)
        will_sync = torch.is_grad_enabled() and self.require_backward_grad_sync
        self.require_forward_param_sync = will_sync
        if will_sync:
            self.reducer.prepare_for_backward([])
        # a non-syncing forward does not disarm: like torch's expect_autograd_hooks_, only the end of a reduction does
        return output


def find_reducers(model_or_models):
    models = model_or_models if isinstance(model_or_models, (list, tuple)) else [model_or_models]
    return [m.reducer for m in models if isinstance(m, B200DataParallel)]
