"""Process / accelerator / gradient state singletons.

Behavioural mirror of the reference's ``src/accelerate/state.py`` for the pieces the data-parallel
path needs: ``PartialState`` (:177-328 process-group init and device selection, :377-414
``wait_for_everyone``), ``AcceleratorState`` (:871-1038 mixed precision) and ``GradientState``
(:1234-1374 ``sync_gradients`` + active dataloader stack).  All three share state between
instances (Borg), as in the reference.

B200 specifics: on a CUDA device the process group uses NCCL **only** for rendezvous and
barriers; the symmetric-memory runtime (``runtime.SymmRuntime``) that carries every byte of
gradient / gather / reduce / broadcast traffic is created right after ``init_process_group``.
"""

from __future__ import annotations

import os
import weakref
from contextlib import contextmanager
from functools import wraps

import torch
import torch.distributed as dist

from .utils.dataclasses import DistributedType, GradientAccumulationPlugin, PrecisionType
from .utils.environment import parse_choice_from_env, parse_flag_from_env


def is_initialized() -> bool:
    return AcceleratorState._shared != {}


class _Borg:
    """Instances of a subclass all alias one dict; ``_reset_state`` empties it."""

    _shared: dict

    def __init__(self):
        self.__dict__ = type(self)._shared

    @classmethod
    def _reset_state(cls):
        cls._shared.clear()


class PartialState(_Borg):
    _shared: dict = {}
    _shared_state = _shared  # the reference's name for the same dict (tests and user code peek at it)
    _known_attrs = [
        "_cpu", "backend", "device", "debug", "distributed_type", "fork_launched", "local_process_index",
        "num_processes", "process_index", "runtime",
    ]  # fmt: skip

    def __init__(self, cpu: bool = False, **kwargs):
        super().__init__()
        if self.initialized:
            return
        self._cpu = bool(cpu) or parse_flag_from_env("ACCELERATE_USE_CPU")
        self.debug = parse_flag_from_env("ACCELERATE_DEBUG_MODE")
        self.fork_launched = parse_flag_from_env("FORK_LAUNCHED", 0)
        self.runtime = None
        use_cuda = torch.cuda.is_available() and not self._cpu
        world = int(os.environ.get("WORLD_SIZE", "1"))
        launched = int(os.environ.get("LOCAL_RANK", "-1")) != -1 and world >= 1 and "RANK" in os.environ

        backend = kwargs.pop("backend", None)
        kwargs.pop("_from_accelerator", None)
        if launched and world > 1:
            local_rank = int(os.environ["LOCAL_RANK"])
            if use_cuda:
                ndev = torch.cuda.device_count()
                local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
                self.device = torch.device("cuda", local_rank % ndev)
                torch.cuda.set_device(self.device)
                self.distributed_type = DistributedType.MULTI_GPU
                if backend in (None, "nccl"):
                    # several ranks on one GPU (1-GPU CI boxes): NCCL refuses duplicate devices, so the
                    # rendezvous group falls back to gloo; the data path is unaffected (own P2P kernels).
                    backend = "gloo" if local_world > ndev else "nccl"
            else:
                self.device = torch.device("cpu")
                self.distributed_type = DistributedType.MULTI_CPU
                backend = backend if backend not in (None, "nccl") else "gloo"
            if not dist.is_initialized():
                init_kwargs = dict(kwargs)
                if backend == "nccl":
                    init_kwargs.setdefault("device_id", self.device)
                dist.init_process_group(backend=backend, **init_kwargs)
            self.backend = backend
            self.num_processes = dist.get_world_size()
            self.process_index = dist.get_rank()
            self.local_process_index = local_rank
        else:
            self.distributed_type = DistributedType.NO
            self.backend = None
            self.num_processes = 1
            self.process_index = self.local_process_index = 0
            env_device = os.environ.get("ACCELERATE_TORCH_DEVICE")  # explicit override (state.py:182-183)
            if env_device is not None:
                self.device = torch.device(env_device)
                use_cuda = self.device.type == "cuda"
            else:
                self.device = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
            if use_cuda:
                torch.cuda.set_device(self.device)

        if self.device.type == "cuda":
            from .runtime import SymmRuntime

            # fails loudly when libb200dp.so is missing: there is no CPU fallback for the CUDA path
            self.runtime = SymmRuntime.initialise(self.process_index, self.num_processes, self.device)

    # ---- introspection ----
    @property
    def initialized(self) -> bool:
        return self._shared != {}

    def __repr__(self):
        return (
            f"Distributed environment: {self.distributed_type}{('  Backend: ' + self.backend) if self.backend else ''}\n"
            f"Num processes: {self.num_processes}\nProcess index: {self.process_index}\n"
            f"Local process index: {self.local_process_index}\nDevice: {self.device}\n"
        )

    @property
    def use_distributed(self) -> bool:
        return self.distributed_type != DistributedType.NO and self.num_processes > 1

    @property
    def is_main_process(self) -> bool:
        return self.process_index == 0

    @property
    def is_local_main_process(self) -> bool:
        return self.local_process_index == 0

    @property
    def is_last_process(self) -> bool:
        return self.process_index == self.num_processes - 1

    @property
    def default_device(self) -> torch.device:
        return self.device

    # ---- synchronisation (stays on torch.distributed: north star keeps NCCL for init/barrier) ----
    def wait_for_everyone(self):
        if self.runtime is not None:
            self.runtime.ctx.check_status()  # surfaces a latched device-side timeout at every barrier point
        if self.use_distributed:
            if self.backend == "nccl":
                dist.barrier(device_ids=[self.device.index])
            else:
                dist.barrier()

    @contextmanager
    def _goes_first(self, is_first: bool):
        if not is_first:
            self.wait_for_everyone()
        yield
        if is_first:
            self.wait_for_everyone()

    @contextmanager
    def main_process_first(self):
        with self._goes_first(self.is_main_process):
            yield

    @contextmanager
    def local_main_process_first(self):
        with self._goes_first(self.is_local_main_process):
            yield

    @contextmanager
    def split_between_processes(self, inputs, apply_padding: bool = False):
        """Yield this process's contiguous share of ``inputs`` (list/tuple/dict of lists/tensor)."""
        if self.num_processes == 1:
            yield inputs
            return
        length = len(inputs) if not isinstance(inputs, dict) else len(next(iter(inputs.values())))
        per, extra = divmod(length, self.num_processes)
        start = self.process_index * per + min(self.process_index, extra)
        end = start + per + (1 if self.process_index < extra else 0)

        width = per + (1 if extra > 0 else 0)  # the longest share

        def take_dataset(data):
            # a Hugging Face `datasets.Dataset` is split by row selection (state.py:498-511); anything else passes through
            try:
                from datasets import Dataset
            except ImportError:
                return data
            if not isinstance(data, Dataset):
                return data
            lo, hi = min(start, len(data) - 1), min(end, len(data))
            rows = list(range(lo, hi))
            if apply_padding:
                rows += [hi - 1] * (width - len(rows))
            return data.select(rows)

        def take(seq):
            # semantics of state.py:477-492: a rank past the end still gets the last element; lists are padded with
            # the share's own last element, tensors with the input's last row through pad_across_processes (a collective)
            if not isinstance(seq, (list, tuple, torch.Tensor)):
                return take_dataset(seq)
            part = seq[-1:] if start >= len(seq) else seq[start:end]
            if apply_padding:
                if isinstance(part, torch.Tensor):
                    from .utils.operations import pad_across_processes, send_to_device

                    part = pad_across_processes(send_to_device(part, self.device), pad_index=seq[-1])
                else:
                    part += [part[-1]] * (width - len(part))
            return part

        if isinstance(inputs, dict):
            if not all(len(v) == length for v in inputs.values()):
                raise ValueError("All values in the dictionary must have the same length")
            for key in inputs.keys():  # the reference rewrites the caller's dict in place and yields it
                inputs[key] = take(inputs[key])
            yield inputs
        else:
            yield take(inputs)

    def on_main_process(self, function=None):
        return self._only_if(function, lambda: self.is_main_process)

    def on_local_main_process(self, function=None):
        return self._only_if(function, lambda: self.is_local_main_process)

    def on_last_process(self, function=None):
        return self._only_if(function, lambda: self.is_last_process)

    def on_process(self, function=None, process_index: int = None):
        if function is None:
            return lambda f: self.on_process(f, process_index)
        return self._only_if(function, lambda: self.process_index == process_index)

    def on_local_process(self, function=None, local_process_index: int = None):
        if function is None:
            return lambda f: self.on_local_process(f, local_process_index)
        return self._only_if(function, lambda: self.local_process_index == local_process_index)

    def _only_if(self, function, predicate):
        if function is None:
            return lambda f: self._only_if(f, predicate)
        if not self.initialized:
            raise ValueError("The `PartialState` or `Accelerator` must be initialized before calling this function.")

        @wraps(function)
        def guarded(*args, **kwargs):
            return function(*args, **kwargs) if (predicate() or not self.use_distributed) else None

        return guarded

    def print(self, *args, **kwargs):
        if self.is_local_main_process:
            print(*args, **kwargs)

    def destroy_process_group(self, group=None):
        if self.runtime is not None:
            from .runtime import SymmRuntime

            SymmRuntime.shutdown()
            self.runtime = None
        if dist.is_initialized():
            dist.destroy_process_group(group)

    def __getattr__(self, name: str):
        if name in PartialState._known_attrs:
            raise AttributeError(
                f"`PartialState` object has no attribute `{name}`. This happens if `PartialState._reset_state()` "
                "was called and an `Accelerator` or `PartialState` was not reinitialized."
            )
        raise AttributeError(f"'PartialState' object has no attribute '{name}'")

    @classmethod
    def _reset_state(cls):
        rt = cls._shared.get("runtime")
        if rt is not None:
            from .runtime import SymmRuntime

            SymmRuntime.shutdown()
        cls._shared.clear()


class AcceleratorState(_Borg):
    _shared: dict = {}
    _shared_state = _shared  # the reference's name for the same dict (tests and user code peek at it)

    def __init__(self, mixed_precision: str = None, cpu: bool = False, _from_accelerator: bool = False, **kwargs):
        super().__init__()
        if parse_flag_from_env("ACCELERATE_USE_CPU"):
            cpu = True
        if PartialState._shared == {}:
            PartialState(cpu, **kwargs)
        self.__dict__.update(PartialState._shared)
        self._check_initialized(mixed_precision, cpu)  # a second, differently configured Accelerator raises (state.py:913)
        if not self.initialized:
            if mixed_precision is None:
                mixed_precision = parse_choice_from_env("ACCELERATE_MIXED_PRECISION", "no")
            mixed_precision = str(mixed_precision).lower()
            if mixed_precision not in PrecisionType.list():
                raise ValueError(f"Unknown mixed_precision mode: {mixed_precision}. Choose between {PrecisionType.list()}")
            if mixed_precision == "fp8":
                raise NotImplementedError("fp8 recipes are out of scope for the B200 data-parallel build")
            self._mixed_precision = mixed_precision
            self._check_initialized_marker = True

    @property
    def initialized(self) -> bool:
        return "_mixed_precision" in self._shared

    @property
    def mixed_precision(self) -> str:
        return self._mixed_precision

    def _check_initialized(self, mixed_precision=None, cpu=None):
        if not self.initialized:
            return
        if cpu and self.device.type != "cpu":
            raise ValueError("AcceleratorState has already been initialized and cannot be changed, restart your runtime completely and pass `cpu=True` to `Accelerator()`.")
        if mixed_precision is not None and mixed_precision != self._mixed_precision:
            raise ValueError(f"AcceleratorState has already been initialized and cannot be changed, restart your runtime completely and pass `mixed_precision={mixed_precision!r}` to `Accelerator()`.")

    def __repr__(self):
        return PartialState().__repr__() + f"\nMixed precision type: {self.mixed_precision}\n"

    # delegate everything process-related to PartialState
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if not AcceleratorState._shared and name in PartialState._known_attrs:
            # the shared state was wiped under a live Accelerator (state.py:1221-1231)
            raise AttributeError(
                f"`AcceleratorState` object has no attribute `{name}`. This happens if `AcceleratorState._reset_state()` "
                "was called and an `Accelerator` or `PartialState` was not reinitialized."
            )
        partial = PartialState._shared
        if name in partial:
            return partial[name]
        try:
            return getattr(PartialState(), name)
        except AttributeError:
            raise AttributeError(f"'AcceleratorState' object has no attribute '{name}'") from None

    @classmethod
    def _reset_state(cls, reset_partial_state: bool = False):
        cls._shared.clear()
        if reset_partial_state:
            PartialState._reset_state()


class GradientState(_Borg):
    """``sync_gradients`` gates the reducer, the optimizer and the scheduler (state.py:1234-1374)."""

    _shared: dict = {}
    _shared_state = _shared  # the reference's name for the same dict (tests and user code peek at it)

    def __init__(self, gradient_accumulation_plugin: GradientAccumulationPlugin | None = None):
        super().__init__()
        if not self.initialized:
            self.sync_gradients = True
            self._loader_refs = [None]
            self.plugin_kwargs = (
                gradient_accumulation_plugin.to_kwargs() if gradient_accumulation_plugin is not None else {}
            )
        elif gradient_accumulation_plugin is not None:
            new = gradient_accumulation_plugin.to_kwargs()
            if new != self.plugin_kwargs:
                self.plugin_kwargs = new

    @property
    def initialized(self) -> bool:
        return self._shared != {}

    @property
    def num_steps(self) -> int:
        return self.plugin_kwargs.get("num_steps", 1)

    @property
    def adjust_scheduler(self) -> bool:
        return self.plugin_kwargs.get("adjust_scheduler", False)

    @property
    def sync_with_dataloader(self) -> bool:
        return self.plugin_kwargs.get("sync_with_dataloader", True)

    @property
    def sync_each_batch(self) -> bool:
        return self.plugin_kwargs.get("sync_each_batch", False)

    # ---- dataloader stack (weak references so loaders can be collected) ----
    @property
    def dataloader_references(self):
        return [ref() if ref is not None else None for ref in self._loader_refs]

    @dataloader_references.setter
    def dataloader_references(self, loaders):
        self._loader_refs = [weakref.ref(dl) if dl is not None else None for dl in loaders]

    @property
    def active_dataloader(self):
        return self.dataloader_references[-1]

    @property
    def in_dataloader(self) -> bool:
        return self.active_dataloader is not None

    @property
    def end_of_dataloader(self) -> bool:
        return self.active_dataloader.end_of_dataloader if self.in_dataloader else False

    @property
    def remainder(self) -> int:
        return self.active_dataloader.remainder if self.in_dataloader else -1

    def _set_sync_gradients(self, value: bool):
        self.sync_gradients = value

    def _add_dataloader(self, dataloader):
        self.dataloader_references = self.dataloader_references + [dataloader]

    def _remove_dataloader(self, dataloader):
        self.dataloader_references = [dl for dl in self.dataloader_references if dl is not dataloader]

    def __repr__(self):
        return (
            f"Sync Gradients: {self.sync_gradients}\nAt end of current dataloader: {self.end_of_dataloader}\n"
            f"Extra samples added: {self.remainder}\nGradient accumulation plugin: {self.plugin_kwargs}\n"
        )
