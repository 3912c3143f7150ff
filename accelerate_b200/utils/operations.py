"""Pytree-level collectives and helpers (API of the reference's ``src/accelerate/utils/operations.py``).

Dispatch is by the *process state*, never by try/except:

* ``DistributedType.NO``           -> identity semantics of the reference (reduce still clones),
* ``DistributedType.MULTI_GPU``    -> the hand-written P2P kernels in ``collectives.py``; a CUDA run
  without the native library fails loudly, it never falls back,
* ``DistributedType.MULTI_CPU``    -> ``torch.distributed`` over gloo, exactly what the reference does
  on CPU (``operations.py:345-350``); it exists so host logic is testable without a GPU.

Object collectives (``gather_object`` / ``broadcast_object_list``) move pickled bytes over the same P2P
kernels on CUDA and stay on ``torch.distributed`` for gloo.
"""

from __future__ import annotations

import pickle
import warnings
from collections.abc import Mapping
from functools import update_wrapper, wraps
from typing import Any

import torch
import torch.distributed as dist

from .dataclasses import DistributedType, TensorInformation


def PartialState():
    """Late import: ``state`` imports this package's dataclasses."""
    from ..state import PartialState as _PartialState

    return _PartialState()


class DistributedOperationException(Exception):
    """Raised in debug mode when a collective would be called with mismatching shapes."""


class CannotPadNestedTensorWarning(UserWarning):
    pass


def is_torch_tensor(x) -> bool:
    return isinstance(x, torch.Tensor)


def is_tensor_information(x) -> bool:
    return isinstance(x, TensorInformation)


def is_namedtuple(data) -> bool:
    return isinstance(data, tuple) and hasattr(data, "_asdict") and hasattr(data, "_fields")


def honor_type(obj, generator):
    """Rebuild a list/tuple/namedtuple of the same type as ``obj`` from ``generator``."""
    return type(obj)(*list(generator)) if is_namedtuple(obj) else type(obj)(generator)


def recursively_apply(func, data, *args, test_type=is_torch_tensor, error_on_other_type=False, **kwargs):
    """Apply ``func`` to every leaf of a nested list/tuple/dict for which ``test_type`` holds."""

    def walk(node):
        if isinstance(node, (tuple, list)):
            return honor_type(node, (walk(child) for child in node))
        if isinstance(node, Mapping):
            return type(node)({key: walk(child) for key, child in node.items()})
        if test_type(node):
            return func(node, *args, **kwargs)
        if error_on_other_type:
            raise TypeError(
                f"Unsupported types ({type(node)}) passed to `{func.__name__}`. Only nested list/tuple/dicts of "
                f"objects that are valid for `{test_type.__name__}` should be passed."
            )
        return node

    return walk(data)


def send_to_device(tensor, device, non_blocking=False, skip_keys=None):
    if is_torch_tensor(tensor) or hasattr(tensor, "to"):
        try:
            return tensor.to(device, non_blocking=non_blocking)
        except TypeError:  # objects whose .to() takes no non_blocking
            return tensor.to(device)
    if isinstance(tensor, (tuple, list)):
        return honor_type(tensor, (send_to_device(t, device, non_blocking, skip_keys) for t in tensor))
    if isinstance(tensor, Mapping):
        skip = [skip_keys] if isinstance(skip_keys, str) else (skip_keys or [])
        return type(tensor)(
            {k: (v if k in skip else send_to_device(v, device, non_blocking, skip_keys)) for k, v in tensor.items()}
        )
    return tensor


def get_data_structure(data):
    return recursively_apply(lambda t: TensorInformation(shape=t.shape, dtype=t.dtype), data)


def get_shape(data):
    return recursively_apply(lambda t: list(t.shape), data)


def initialize_tensors(data_structure):
    return recursively_apply(
        lambda info: torch.empty(*info.shape, dtype=info.dtype), data_structure, test_type=is_tensor_information
    )


def find_batch_size(data):
    if isinstance(data, (tuple, list, Mapping)) and len(data) == 0:
        raise ValueError(f"Cannot find the batch size from empty {type(data)}.")
    if isinstance(data, (tuple, list)):
        return find_batch_size(data[0])
    if isinstance(data, Mapping):
        return find_batch_size(next(iter(data.values())))
    if not isinstance(data, torch.Tensor):
        raise TypeError(f"Can only find the batch size of tensors but got {type(data)}.")
    return data.shape[0]


def ignorant_find_batch_size(data):
    try:
        return find_batch_size(data)
    except (ValueError, TypeError):
        return None


def listify(data):
    def to_list(t):
        t = t.detach().cpu()
        return (t.float() if t.dtype == torch.bfloat16 else t).tolist()

    return recursively_apply(to_list, data)


def find_device(data):
    if isinstance(data, Mapping):
        data = list(data.values())
    if isinstance(data, (tuple, list)):
        for item in data:
            dev = find_device(item)
            if dev is not None:
                return dev
        return None
    if isinstance(data, torch.Tensor):
        return data.device
    return None


def slice_tensors(data, tensor_slice, process_index=None, num_processes=None):
    return recursively_apply(lambda t: t[tensor_slice], data)


def concatenate(data, dim=0):
    head = data[0]
    if isinstance(head, (tuple, list)):
        return honor_type(head, (concatenate([d[i] for d in data], dim=dim) for i in range(len(head))))
    if isinstance(head, Mapping):
        return type(head)({k: concatenate([d[k] for d in data], dim=dim) for k in head.keys()})
    if isinstance(head, torch.Tensor):
        return torch.cat(data, dim=dim)
    if isinstance(data, (tuple, list)) and len(data) == 1:
        return head
    raise TypeError(f"Can only concatenate tensors but got {type(head)}")


# ---------------------------------------------------------------------------------------------
# debug-mode verification (ACCELERATE_DEBUG_MODE): all ranks must pass the same shapes
# ---------------------------------------------------------------------------------------------
def verify_operation(function):
    @wraps(function)
    def wrapper(*args, **kwargs):
        state = PartialState()
        if state.distributed_type == DistributedType.NO or not state.debug:
            return function(*args, **kwargs)
        operation = f"{function.__module__}.{function.__name__}"
        tensor = kwargs["tensor"] if "tensor" in kwargs else args[0]
        found = find_device(tensor)
        if found is not None and state.device.type != found.type:
            raise DistributedOperationException(
                f"One or more of the tensors passed to {operation} were not on the {found.type} while the "
                f"`Accelerator` is configured for {state.device.type}. "
                f"Please move it to the {state.device.type} before calling {operation}."
            )
        shapes = gather_object([get_shape(tensor)])
        if shapes[0] is not None and shapes.count(shapes[0]) != len(shapes):
            listing = "\n  - ".join(f"Process {i}: {shape}" for i, shape in enumerate(shapes))
            raise DistributedOperationException(
                "Cannot apply desired operation due to shape mismatches. All shapes across devices must be valid."
                f"\n\nOperation: `{operation}`\nInput shapes:\n  - {listing}"
            )
        return function(*args, **kwargs)

    return wrapper


def chained_operation(function):
    @wraps(function)
    def wrapper(*args, **kwargs):
        try:
            return function(*args, **kwargs)
        except DistributedOperationException as err:
            operation = f"{function.__module__}.{function.__name__}"
            raise DistributedOperationException(
                f"Error found while calling `{operation}`. Please see the earlier error for more details."
            ) from err

    return wrapper


# ---------------------------------------------------------------------------------------------
# collectives
# ---------------------------------------------------------------------------------------------
def _cpu_gather_one(t):
    state = PartialState()
    if t.ndim == 0:
        t = t.clone()[None]
    if not t.is_contiguous():
        t = t.contiguous()
    parts = [torch.empty_like(t) for _ in range(state.num_processes)]
    dist.all_gather(parts, t)
    return torch.cat(parts, dim=0)


@verify_operation
def gather(tensor):
    """All ranks' tensors concatenated on dim 0, for every leaf (operations.py:418-435)."""
    state = PartialState()
    if state.distributed_type == DistributedType.MULTI_GPU:
        from ..collectives import all_gather

        return recursively_apply(all_gather, tensor, error_on_other_type=True)
    if state.distributed_type == DistributedType.MULTI_CPU:
        return recursively_apply(_cpu_gather_one, tensor, error_on_other_type=True)
    return tensor


def gather_object(object: Any):
    """Gather picklable objects; like the reference, each rank passes a LIST and the lists are concatenated."""
    state = PartialState()
    if state.distributed_type == DistributedType.NO:
        return object
    if state.distributed_type == DistributedType.MULTI_GPU:
        from ..collectives import gather_object as p2p_gather_object

        parts = p2p_gather_object(object)
    else:
        parts = [None] * state.num_processes
        dist.all_gather_object(parts, object)
    return [item for part in parts for item in part]


# dtype codes of the shape message (operations.py:536-549)
TENSOR_TYPE_TO_INT = {
    torch.float: 1, torch.double: 2, torch.half: 3, torch.bfloat16: 4, torch.uint8: 5, torch.int8: 6, torch.int16: 7,
    torch.int32: 8, torch.int64: 9, torch.bool: 10,
}  # fmt: skip
TENSOR_INT_TO_DTYPE = {v: k for k, v in TENSOR_TYPE_TO_INT.items()}


def gather_tensor_shape(tensor):
    """Shape and dtype code of a tensor that exists on ONE process only (``None`` elsewhere), learnt by all through a
    sum-reduction of a small zero-initialised message (operations.py:552-574).  Returns ``(dims, dtype_code)``; as in
    the reference, dimensions of size 0 cannot be represented."""
    state = PartialState()
    message = torch.zeros(64, dtype=torch.int, device=state.device)
    if tensor is not None:
        dims = list(tensor.shape) + [TENSOR_TYPE_TO_INT[tensor.dtype]]
        message[: len(dims)] = torch.tensor(dims, dtype=torch.int)
    message = reduce(message, reduction="sum")
    message = message[message.nonzero()]
    return message[:-1], int(message[-1:][0])


def copy_tensor_to_devices(tensor=None) -> torch.Tensor:
    """Give every process a copy of a tensor only one of them holds, without the others knowing its shape
    (operations.py:577-591): shape message, then a sum-reduction against zeros."""
    state = PartialState()
    shape, dtype = gather_tensor_shape(tensor)
    if tensor is None:
        tensor = torch.zeros([int(d) for d in shape.flatten().tolist()], dtype=TENSOR_INT_TO_DTYPE[dtype]).to(state.device)
    return reduce(tensor, reduction="sum")


@verify_operation
def broadcast(tensor, from_process: int = 0):
    """In-place broadcast of every leaf from ``from_process`` (operations.py:594-613)."""
    state = PartialState()
    if state.distributed_type == DistributedType.MULTI_GPU:
        from ..collectives import broadcast_

        return recursively_apply(broadcast_, tensor, error_on_other_type=True, src=from_process)
    if state.distributed_type == DistributedType.MULTI_CPU:

        def one(t, src=0):
            dist.broadcast(t, src=src)
            return t

        return recursively_apply(one, tensor, error_on_other_type=True, src=from_process)
    return tensor


def broadcast_object_list(object_list, from_process: int = 0):
    """In place: every rank's list becomes ``from_process``'s (operations.py:669-690)."""
    state = PartialState()
    if state.distributed_type == DistributedType.MULTI_GPU:
        from ..collectives import broadcast_object

        received = broadcast_object(list(object_list), src=from_process)
        object_list[:] = received
    elif state.distributed_type != DistributedType.NO:
        dist.broadcast_object_list(object_list, src=from_process)
    return object_list


@chained_operation
def pad_across_processes(tensor, dim=0, pad_index=0, pad_first=False):
    """Pad every leaf along ``dim`` to the largest size any rank holds (operations.py:743-796)."""
    state = PartialState()

    def one(t, dim=0, pad_index=0, pad_first=False):
        if getattr(t, "is_nested", False):
            warnings.warn(
                "Cannot pad nested tensors without more information. Leaving unprocessed.",
                CannotPadNestedTensorWarning,
            )
            return t
        if dim >= t.ndim or dim < -t.ndim:
            return t
        if dim < 0:
            dim += t.ndim
        if state.distributed_type == DistributedType.MULTI_GPU:
            from ..collectives import pad_to_max

            return pad_to_max(t, dim, pad_index, pad_first)
        sizes = gather(torch.tensor(t.shape, device=t.device)[None]).cpu()
        max_size = max(int(s[dim]) for s in sizes)
        if max_size == t.shape[dim]:
            return t
        new_shape = list(t.shape)
        new_shape[dim] = max_size
        padded = t.new_zeros(tuple(new_shape)) + pad_index
        window = [slice(None)] * t.ndim
        window[dim] = slice(max_size - t.shape[dim], max_size) if pad_first else slice(0, t.shape[dim])
        padded[tuple(window)] = t
        return padded

    return recursively_apply(one, tensor, error_on_other_type=True, dim=dim, pad_index=pad_index, pad_first=pad_first)


def pad_input_tensors(tensor, batch_size, num_processes, dim=0):
    """Repeat-pad a batch so it splits evenly over ``num_processes`` (dispatcher helper)."""

    def one(t, batch_size, num_processes, dim=0):
        share = batch_size // num_processes
        leftover = batch_size - share * num_processes
        to_pad = (num_processes - batch_size) if share == 0 else (num_processes - share)
        if leftover > to_pad & to_pad < 1:
            to_pad = leftover - to_pad
        new_shape = list(t.shape)
        new_shape[0] = batch_size + to_pad
        out = t.new_zeros(tuple(new_shape))
        window = [slice(None)] * t.ndim
        window[dim] = slice(0, t.shape[dim])
        out[tuple(window)] = t
        return out

    return recursively_apply(one, tensor, error_on_other_type=True, batch_size=batch_size, num_processes=num_processes, dim=dim)


@verify_operation
def reduce(tensor, reduction="mean", scale=1.0):
    """Sum every leaf over ranks into a fresh tensor; divide by N only for "mean" (operations.py:839-878)."""
    state = PartialState()

    def one(t, reduction="mean", scale=1.0):
        if state.distributed_type == DistributedType.NO:
            return t.clone()
        if state.distributed_type == DistributedType.MULTI_GPU:
            from ..collectives import all_reduce

            return all_reduce(t, mean=(reduction == "mean"))
        out = t.clone()
        dist.all_reduce(out, dist.ReduceOp.SUM)
        if reduction == "mean":
            out /= state.num_processes
        return out

    return recursively_apply(one, tensor, error_on_other_type=True, reduction=reduction, scale=scale)


# ---------------------------------------------------------------------------------------------
# mixed-precision output conversion (operations.py:881-940): stays in Python by the north star
# ---------------------------------------------------------------------------------------------
def convert_to_fp32(tensor):
    def is_half(t):
        return (is_torch_tensor(t) or hasattr(t, "dtype")) and t.dtype in (torch.float16, torch.bfloat16)

    return recursively_apply(lambda t: t.float(), tensor, test_type=is_half)


class ConvertOutputsToFp32:
    def __init__(self, model_forward):
        self.model_forward = model_forward
        update_wrapper(self, model_forward)

    def __call__(self, *args, **kwargs):
        return convert_to_fp32(self.model_forward(*args, **kwargs))

    def __getstate__(self):
        raise pickle.PicklingError(
            "Cannot pickle a prepared model with automatic mixed precision, please unwrap the model with "
            "`Accelerator.unwrap_model(model)` before pickling it."
        )


def convert_outputs_to_fp32(model_forward):
    converter = ConvertOutputsToFp32(model_forward)

    def forward(*args, **kwargs):
        return converter(*args, **kwargs)

    forward.__wrapped__ = converter  # lets unwrap_model peel it off again
    return forward
