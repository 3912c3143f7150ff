"""Configuration surface of the data-parallel path: the plugin boundary.

Same class names, field names, defaults and ``to_kwargs()`` (diff-from-default) behaviour as the
reference's ``src/accelerate/utils/dataclasses.py`` for the handlers the hot path consumes:
``DistributedDataParallelKwargs`` (:154-237), ``DDPCommunicationHookType`` (:134-151),
``GradScalerKwargs`` (:240-269), ``AutocastKwargs`` (:112-131), ``InitProcessGroupKwargs``
(:272-299), ``GradientAccumulationPlugin`` (:976-1025), ``DataLoaderConfiguration`` (:818-910),
``DistributedType`` (:600-636), ``PrecisionType`` (:769-782).  Other backends' plugins (FSDP,
DeepSpeed, Megatron, TP/CP/SP, FP8 recipes) are out of scope for this build.
"""

from __future__ import annotations

import copy
import enum
from dataclasses import dataclass, field
from datetime import timedelta
from typing import Any, Optional


class KwargsHandler:
    """Mixin for config dataclasses: ``to_kwargs()`` returns only the fields that differ from a default instance."""

    def to_dict(self) -> dict:
        return copy.deepcopy(self.__dict__)

    def to_kwargs(self) -> dict:
        from .environment import clear_environment

        with clear_environment():  # defaults must not depend on the caller's env
            baseline = type(self)().to_dict()
        mine = self.to_dict()
        return {name: value for name, value in mine.items() if baseline[name] != value}


class _ContainsMeta(enum.EnumMeta):
    def __contains__(cls, item):
        try:
            cls(item)
            return True
        except ValueError:
            return False


class BaseEnum(enum.Enum, metaclass=_ContainsMeta):
    def __str__(self):
        return self.value

    @classmethod
    def list(cls):
        return [str(m) for m in cls]


class DistributedType(str, enum.Enum):
    """This build only ever takes NO, MULTI_GPU (CUDA) and MULTI_CPU (gloo; host-logic tests); the remaining members
    of the reference's enum (dataclasses.py:606-636) exist so that scripts comparing against them keep working."""

    NO = "NO"
    MULTI_CPU = "MULTI_CPU"
    MULTI_GPU = "MULTI_GPU"
    MULTI_NPU = "MULTI_NPU"
    MULTI_MLU = "MULTI_MLU"
    MULTI_SDAA = "MULTI_SDAA"
    MULTI_MUSA = "MULTI_MUSA"
    MULTI_XPU = "MULTI_XPU"
    DEEPSPEED = "DEEPSPEED"
    FSDP = "FSDP"
    XLA = "XLA"
    MEGATRON_LM = "MEGATRON_LM"
    MULTI_HPU = "MULTI_HPU"
    MULTI_NEURON = "MULTI_NEURON"


class PrecisionType(str, BaseEnum):
    NO = "no"
    FP16 = "fp16"
    BF16 = "bf16"
    FP8 = "fp8"  # recognised so that the error message is helpful; not supported by this build


class RNGType(BaseEnum):
    TORCH = "torch"
    CUDA = "cuda"
    GENERATOR = "generator"


@dataclass
class AutocastKwargs(KwargsHandler):
    enabled: bool = True
    cache_enabled: Optional[bool] = None


class DDPCommunicationHookType(BaseEnum):
    NO = "no"
    FP16 = "fp16"
    BF16 = "bf16"
    POWER_SGD = "power_sgd"
    BATCHED_POWER_SGD = "batched_power_sgd"


@dataclass
class DistributedDataParallelKwargs(KwargsHandler):
    """Knobs of the gradient reducer; field-compatible with the reference (dataclasses.py:181-195).

    ``comm_hook=FP16|BF16`` selects a half-precision *wire* format for the fused allreduce (the
    reference's compress hooks, dataclasses.py:206-216): gradients are cast and pre-divided into a
    half-width symmetric buffer, reduced there, and widened back once per step.  ``POWER_SGD`` /
    ``BATCHED_POWER_SGD`` and any ``comm_wrapper`` combination register torch's own hook functions,
    exactly as the reference does; the wrapper runs them per bucket (``GradReducer.set_user_hook``).
    As in the reference, a ``comm_wrapper`` without a ``comm_hook`` registers nothing.
    """

    dim: int = 0
    broadcast_buffers: bool = True
    bucket_cap_mb: int = 25
    find_unused_parameters: bool = False
    check_reduction: bool = False
    gradient_as_bucket_view: bool = False
    static_graph: bool = False

    comm_hook: DDPCommunicationHookType = DDPCommunicationHookType.NO
    comm_wrapper: DDPCommunicationHookType = DDPCommunicationHookType.NO
    comm_state_option: dict = field(default_factory=dict)

    def to_dict(self, ignore_keys=("comm_hook", "comm_wrapper", "comm_state_option")):
        return {k: v for k, v in super().to_dict().items() if k not in ignore_keys}

    def register_comm_hook(self, model):
        """Mirror of dataclasses.py:200-237: translate the enums into a hook on the wrapper.  A bare FP16/BF16
        compress hook maps to the reducer's native half-precision wire; PowerSGD and wrapped combinations are torch's
        own hook functions, which the wrapper runs per bucket through the ``GradBucket`` protocol."""
        hook_type, wrapper_type = self.comm_hook, self.comm_wrapper
        no, half = DDPCommunicationHookType.NO, (DDPCommunicationHookType.FP16, DDPCommunicationHookType.BF16)
        if hook_type == no:
            return  # the reference registers nothing without a hook, whatever the wrapper says (dataclasses.py:222)
        if hook_type in half and wrapper_type == no:
            model.register_comm_hook(None, hook_type)
            return
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks, powerSGD_hook

        hook = {
            DDPCommunicationHookType.FP16: default_hooks.fp16_compress_hook,
            DDPCommunicationHookType.BF16: default_hooks.bf16_compress_hook,
            DDPCommunicationHookType.POWER_SGD: powerSGD_hook.powerSGD_hook,
            DDPCommunicationHookType.BATCHED_POWER_SGD: powerSGD_hook.batched_powerSGD_hook,
        }[hook_type]
        if wrapper_type != no:
            hook = {
                DDPCommunicationHookType.FP16: default_hooks.fp16_compress_wrapper,
                DDPCommunicationHookType.BF16: default_hooks.bf16_compress_wrapper,
            }[wrapper_type](hook)
        state = None
        if hook_type in (DDPCommunicationHookType.POWER_SGD, DDPCommunicationHookType.BATCHED_POWER_SGD):
            state = powerSGD_hook.PowerSGDState(None, **self.comm_state_option)
        model.register_comm_hook(state, hook)


@dataclass
class GradScalerKwargs(KwargsHandler):
    init_scale: float = 65536.0
    growth_factor: float = 2.0
    backoff_factor: float = 0.5
    growth_interval: int = 2000
    enabled: bool = True


@dataclass
class InitProcessGroupKwargs(KwargsHandler):
    backend: Optional[str] = "nccl"
    init_method: Optional[str] = None
    timeout: Optional[timedelta] = None

    def __post_init__(self):
        if self.timeout is None:
            seconds = 1800 if self.backend != "nccl" else 600
            self.timeout = timedelta(seconds=seconds)


@dataclass
class GradientAccumulationPlugin(KwargsHandler):
    num_steps: int = None
    adjust_scheduler: bool = True
    sync_with_dataloader: bool = True
    sync_each_batch: bool = False


@dataclass
class DataLoaderConfiguration:
    split_batches: bool = False
    dispatch_batches: Optional[bool] = None
    even_batches: bool = True
    use_seedable_sampler: bool = False
    data_seed: Optional[int] = None
    non_blocking: bool = False
    use_stateful_dataloader: bool = False


@dataclass
class ProjectConfiguration:
    project_dir: Optional[str] = None
    logging_dir: Optional[str] = None
    automatic_checkpoint_naming: bool = False
    total_limit: Optional[int] = None
    iteration: int = 0
    save_on_each_node: bool = False

    def set_directories(self, project_dir: Optional[str] = None):
        self.project_dir = project_dir
        if self.logging_dir is None:
            self.logging_dir = project_dir

    def __post_init__(self):
        self.set_directories(self.project_dir)


@dataclass
class TensorInformation:
    shape: Any
    dtype: Any


__all__ = [n for n, v in list(globals().items()) if isinstance(v, type) and v.__module__ == __name__]
