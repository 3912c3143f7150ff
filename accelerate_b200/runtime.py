"""Process-wide symmetric-memory runtime on top of the native context.

One :class:`SymmRuntime` per process (one process per GPU).  It owns

* the native context (``libb200dp.so``),
* the *control* segment (cross-GPU flag pads, norm/found-inf records, small-exchange slots),
* a lazily created *staging* segment used by ``gather`` / ``reduce`` / ``broadcast``,
* the rendezvous that maps every peer's copy of a segment into this process.

Rendezvous uses ``torch.distributed`` object collectives only (north star: NCCL stays for
init/barrier).  Segment memory is exchanged either as POSIX file descriptors of ``cuMemCreate``
allocations (passed with SCM_RIGHTS by the native library; required for NVLS multicast) or as
``cudaIpcMemHandle_t`` bytes.  Which one is used is agreed collectively so all ranks take the
same branch.

Reference call sites this replaces: the NCCL communicator that
``torch.nn.parallel.DistributedDataParallel`` (``src/accelerate/accelerator.py:1892``) and
``torch.distributed.all_gather_into_tensor`` / ``all_reduce`` / ``broadcast``
(``src/accelerate/utils/operations.py:342,871,522``) use under the hood.
"""

from __future__ import annotations

import os
import uuid
from dataclasses import dataclass, field

import torch
import torch.distributed as dist

from . import native

_TORCH_TO_NATIVE = {
    torch.float32: native.F32,
    torch.bfloat16: native.BF16,
    torch.float16: native.F16,
    torch.float64: native.F64,
    torch.int32: native.I32,
    torch.int64: native.I64,
    torch.uint8: native.U8,
    torch.bool: native.U8,
    torch.int8: native.U8,
}


def native_dtype(dtype: torch.dtype) -> int | None:
    return _TORCH_TO_NATIVE.get(dtype)


class _DeviceSpan:
    """Exposes raw device memory to torch through ``__cuda_array_interface__`` (no copy, no ownership)."""

    def __init__(self, ptr: int, nbytes: int, keepalive):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 3,
            "strides": None,
        }
        self._keepalive = keepalive


@dataclass
class Segment:
    """A symmetric allocation: same size on every rank, every peer's copy mapped locally."""

    seg: int
    ptr: int
    nbytes: int
    kind: int
    multicast: bool = False
    bytes_u8: torch.Tensor | None = field(default=None, repr=False)

    def tensor(self, dtype: torch.dtype, numel: int, byte_offset: int = 0) -> torch.Tensor:
        es = torch.empty((), dtype=dtype).element_size()
        return self.bytes_u8[byte_offset : byte_offset + numel * es].view(dtype)


class SymmRuntime:
    _instance: "SymmRuntime | None" = None

    def __init__(self, rank: int, world: int, device: torch.device, group=None):
        if device.type != "cuda":
            raise native.B200dpError("SymmRuntime needs a CUDA device: the b200dp data path has no CPU fallback")
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.ctx = native.Context(rank, world, device.index if device.index is not None else torch.cuda.current_device())
        t = os.environ.get("B200DP_TIMEOUT_MS")
        if t:
            self.ctx.set_timeout_ms(int(t))
        self.sm_count = self.ctx.query(native.Q_SM_COUNT)
        self._serial = 0
        self._pending_fds: dict[tuple[int, int], int] = {}
        self._slot_next = 0
        self.staging: Segment | None = None
        self.staging_bytes = max(int(os.environ.get("B200DP_STAGING_MB", "64")) << 20, 2 * self.ctx.push_area_bytes())
        self.push_max = self.ctx.push_max_bytes() if os.environ.get("B200DP_PUSH", "1") != "0" else 0
        # bucket pack variant: "auto" (default) = the TMA-engine kernel for launches of >= 128 MiB (6.4 TB/s at 1 GB vs 5.8
        # for LDG/STG), the LDG/STG kernel below (5.2 vs 4.5 TB/s on a 48 MB bucket, no shared memory: co-resides with the
        # backward GEMMs it overlaps) — profiles/r02_pack_sweep_n1.jsonl; "1" / "0" force one of them
        self.bulk_pack_mode = os.environ.get("B200DP_BULK_PACK", "auto").lower()
        self.bulk_pack_min_bytes = int(os.environ.get("B200DP_BULK_PACK_MIN_MB", "128")) << 20

        # ---- collective agreement on the handle kind / multicast ----
        want = os.environ.get("B200DP_SEG_KIND", "fd").lower()
        vmm_ok = bool(self.ctx.query(native.Q_VMM_OK)) and want != "ipc"
        mc_ok = bool(self.ctx.query(native.Q_MULTICAST_OK)) and os.environ.get("B200DP_NVLS", "1") != "0"
        dev_id = self._device_identity()
        infos = self._all_gather_object((vmm_ok, mc_ok, dev_id))
        self.kind = native.HANDLE_POSIXFD if all(i[0] for i in infos) else native.HANDLE_CUDAIPC
        distinct = len({i[2] for i in infos}) == len(infos)
        self.multicast_possible = (
            self.kind == native.HANDLE_POSIXFD and world > 1 and all(i[1] for i in infos) and distinct
        )
        self.oversubscribed = not distinct  # several ranks share one GPU (single-GPU test boxes)
        if self.kind == native.HANDLE_POSIXFD and world > 1:
            token = self._broadcast_object(uuid.uuid4().hex[:12] if rank == 0 else None)
            self._sock_names = [f"b200dp-{token}-{r}" for r in range(world)]
            self.ctx.fdx_open(self._sock_names[rank])
            self._barrier()

        # ---- control segment ----
        self.control = self.alloc_segment(self.ctx.control_bytes(), multicast=False)
        self.ctx.set_control(self.control.seg)
        # 3 floats of result per finalize call: total_norm, multiplier, found_inf
        self._barrier()

    # ------------------------------------------------------------------ singletons
    @classmethod
    def get(cls) -> "SymmRuntime":
        if cls._instance is None:
            raise native.B200dpError("SymmRuntime not initialised (PartialState creates it on CUDA devices)")
        return cls._instance

    @classmethod
    def initialise(cls, rank: int, world: int, device: torch.device, group=None) -> "SymmRuntime":
        if cls._instance is None:
            cls._instance = cls(rank, world, device, group)
        return cls._instance

    @classmethod
    def shutdown(cls):
        if cls._instance is not None:
            inst, cls._instance = cls._instance, None
            try:
                torch.cuda.synchronize(inst.device)
            except Exception:
                pass
            inst.ctx.fdx_close()
            inst.ctx.close()

    # ------------------------------------------------------------------ torch.distributed helpers
    def _all_gather_object(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def _broadcast_object(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=0, group=self.group)
        return box[0]

    def _barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)

    def _device_identity(self):
        try:
            props = torch.cuda.get_device_properties(self.device)
            return str(getattr(props, "uuid", self.ctx.device))
        except Exception:
            return str(self.ctx.device)

    def _all_ok(self, ok: bool) -> bool:
        return all(self._all_gather_object(bool(ok)))

    @property
    def use_bulk_pack(self) -> bool:
        """Variant for launches whose size is not known to the caller (kept for the staging helpers: small copies)."""
        return self.bulk_pack_mode == "1"

    def bulk_for(self, nbytes: int) -> bool:
        if self.bulk_pack_mode in ("0", "1"):
            return self.bulk_pack_mode == "1"
        return nbytes >= self.bulk_pack_min_bytes

    # ------------------------------------------------------------------ slots
    def reserve_slots(self, n: int) -> int:
        """``n`` consecutive statistics slots.  Released ranges are reused first-fit; every rank makes the same
        sequence of reserve / release calls, so slot numbers agree across ranks."""
        free = self.__dict__.setdefault("_slot_free", [])
        for k, (base, size) in enumerate(free):
            if size >= n:
                if size == n:
                    free.pop(k)
                else:
                    free[k] = (base + n, size - n)
                return base
        base = self._slot_next
        if base + n > native.MAX_SLOTS:
            raise native.B200dpError(f"out of stats slots ({native.MAX_SLOTS})")
        self._slot_next += n
        return base

    def release_slots(self, base: int, n: int):
        if n <= 0:
            return
        if base + n == self._slot_next:
            self._slot_next = base
            return
        free = self.__dict__.setdefault("_slot_free", [])
        free.append((base, n))
        free.sort()
        merged = []
        for b, s_ in free:  # coalesce neighbours
            if merged and merged[-1][0] + merged[-1][1] == b:
                merged[-1] = (merged[-1][0], merged[-1][1] + s_)
            else:
                merged.append((b, s_))
        self._slot_free = merged

    # ------------------------------------------------------------------ segments
    def _recv_fd(self, serial: int, src: int, mc: bool) -> int:
        key = (serial * 2 + int(mc), src)
        while key not in self._pending_fds:
            fd, tag = self.ctx.fdx_recv(60000)
            self._pending_fds[(tag // native.MAX_WORLD, tag % native.MAX_WORLD)] = fd
        return self._pending_fds.pop(key)

    def alloc_segment(self, nbytes: int, multicast: bool = True) -> Segment:
        """Collective: every rank must call with the same ``nbytes`` in the same order."""
        ctx = self.ctx
        serial = self._serial
        self._serial += 1
        seg, ptr, padded, kind = ctx.seg_create(nbytes, self.kind)
        if kind != self.kind:
            raise native.B200dpError(f"segment kind {kind} differs from the agreed kind {self.kind}")
        if self.world > 1:
            if kind == native.HANDLE_POSIXFD:
                _, fd = ctx.seg_export(seg)
                try:
                    for r in range(self.world):
                        if r != self.rank:
                            ctx.fdx_send(self._sock_names[r], fd, (serial * 2) * native.MAX_WORLD + self.rank)
                finally:
                    os.close(fd)
                for r in range(self.world):
                    if r != self.rank:
                        ctx.seg_import(seg, r, kind, b"\0" * 64, self._recv_fd(serial, r, False))
            else:
                handle, _ = ctx.seg_export(seg)
                handles = self._all_gather_object(handle)
                for r in range(self.world):
                    if r != self.rank:
                        ctx.seg_import(seg, r, kind, handles[r], -1)
            self._barrier()
        s = Segment(seg=seg, ptr=ptr, nbytes=padded, kind=kind)
        s.bytes_u8 = torch.as_tensor(_DeviceSpan(ptr, padded, self), device=self.device)
        if multicast and self.multicast_possible:
            s.multicast = self._bind_multicast(s, serial)
        return s

    def _bind_multicast(self, s: Segment, serial: int) -> bool:
        """NVLS: rank 0 creates the multicast object, everyone adds its device and binds its memory."""
        ctx = self.ctx
        ok = True
        fd = -1
        if self.rank == 0:
            try:
                fd = ctx.mc_create(s.seg)
                for r in range(1, self.world):
                    ctx.fdx_send(self._sock_names[r], fd, (serial * 2 + 1) * native.MAX_WORLD + 0)
            except native.B200dpError:
                ok = False
        if not self._all_ok(ok):  # creation failed: peers never got an fd, nobody waits for one
            return False
        try:
            if self.rank == 0:
                os.close(fd)
                ctx.mc_attach(s.seg, -1)
            else:
                ctx.mc_attach(s.seg, self._recv_fd(serial, 0, True))
        except native.B200dpError:
            ok = False
        if not self._all_ok(ok):
            return False
        try:
            ctx.mc_bind(s.seg)
        except native.B200dpError:
            ok = False
        return self._all_ok(ok)

    def free_segment(self, s: Segment):
        """Unmap and release this rank's part of a symmetric segment.  cuMem allocations are reference counted by
        their mappings, so a peer that still has this memory mapped keeps the physical pages alive until it frees
        its own side: no cross-rank barrier is needed (cudaIpc segments — the fallback kind — do need one)."""
        torch.cuda.synchronize(self.device)
        if self.kind != native.HANDLE_POSIXFD:
            self._barrier()
        s.bytes_u8 = None
        self.ctx.seg_destroy(s.seg)
        if s is self.staging:
            self.staging = None

    def ensure_staging(self) -> Segment:
        if self.staging is None:
            self.staging = self.alloc_segment(self.staging_bytes, multicast=True)
        return self.staging

    # ------------------------------------------------------------------ streams
    @staticmethod
    def current_stream() -> int:
        return torch.cuda.current_stream().cuda_stream
