"""Model unwrapping and the autocast context (reference: ``src/accelerate/utils/other.py:248-333``,
``src/accelerate/utils/modeling.py:2052-2093``)."""

from __future__ import annotations

import contextlib

import torch

from .dataclasses import AutocastKwargs


def PartialState():
    from ..state import PartialState as _PartialState

    return _PartialState()


def AcceleratorState():
    from ..state import AcceleratorState as _AcceleratorState

    return _AcceleratorState()


def extract_model_from_parallel(model, keep_fp32_wrapper: bool = True, keep_torch_compile: bool = True, recursive: bool = False):
    """Peel the data-parallel wrapper (ours, or torch's if a user wrapped by hand) and, unless
    ``keep_fp32_wrapper``, the autocast/fp32-output forward installed by ``prepare_model``."""
    from ..reducer import B200DataParallel

    wrappers = (B200DataParallel, torch.nn.parallel.DistributedDataParallel, torch.nn.DataParallel)
    compiled = None
    if hasattr(model, "_orig_mod"):  # torch.compile
        compiled, model = model, model._orig_mod
    while isinstance(model, wrappers):
        model = model.module
    if recursive:
        for name, child in list(model.named_children()):
            setattr(model, name, extract_model_from_parallel(child, keep_fp32_wrapper, keep_torch_compile, recursive))
    if not keep_fp32_wrapper:
        forward = model.forward
        original = model.__dict__.pop("_original_forward", None)
        if original is not None:
            while hasattr(forward, "__wrapped__"):
                forward = forward.__wrapped__
                if forward == original:
                    break
            model.forward = original
    if keep_torch_compile and compiled is not None:
        compiled._orig_mod = model
        return compiled
    return model


def get_mixed_precision_context_manager(native_amp: bool = False, autocast_kwargs: AutocastKwargs = None):
    """``torch.autocast`` for the configured precision, or a null context."""
    state = AcceleratorState()
    kwargs = autocast_kwargs.to_kwargs() if autocast_kwargs is not None else {}
    if not native_amp:
        return contextlib.nullcontext()
    device_type = state.device.type
    if state.mixed_precision == "fp16":
        return torch.autocast(device_type=device_type, dtype=torch.float16, **kwargs)
    if state.mixed_precision == "bf16":
        return torch.autocast(device_type=device_type, dtype=torch.bfloat16, **kwargs)
    return torch.autocast(device_type=device_type, **kwargs)


def wait_for_everyone():
    PartialState().wait_for_everyone()
