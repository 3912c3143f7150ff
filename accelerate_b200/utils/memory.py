"""Memory helpers training scripts written for the reference use (``src/accelerate/utils/memory.py``): drop
references and empty the CUDA cache, and retry a function with a smaller batch size after an out-of-memory error."""

import functools
import gc
import inspect

import torch


def clear_device_cache(garbage_collection: bool = False):
    """``torch.cuda.empty_cache()`` (optionally after a ``gc.collect()``, which is slow: use sparingly)."""
    if garbage_collection:
        gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def release_memory(*objects):
    """``a, b = release_memory(a, b)``: the names are rebound to ``None`` and the freed device memory is returned
    to the driver."""
    cleared = [None for _ in objects]
    clear_device_cache(garbage_collection=True)
    return cleared


_OOM_MARKERS = (
    " out of memory.",  # CUDA
    "cuDNN error: CUDNN_STATUS_NOT_SUPPORTED.",  # cuDNN refusing an oversized problem
    "DefaultCPUAllocator: can't allocate memory",  # host
)


def should_reduce_batch_size(exception: Exception) -> bool:
    """Is ``exception`` one of the out-of-memory errors a smaller batch would cure?"""
    if isinstance(exception, RuntimeError) and len(exception.args) == 1:
        return any(marker in exception.args[0] for marker in _OOM_MARKERS)
    return False


def find_executable_batch_size(function=None, starting_batch_size: int = 128, reduce_batch_size_fn=None):
    """Decorator: call ``function(batch_size, *args, **kwargs)``, and after an out-of-memory error again with the batch
    size multiplied by 0.9 (or whatever ``reduce_batch_size_fn()`` returns) until it fits; zero raises.  The decorated
    function is called WITHOUT the batch size (memory.py:118-184)."""
    if function is None:
        return functools.partial(find_executable_batch_size, starting_batch_size=starting_batch_size,
                                 reduce_batch_size_fn=reduce_batch_size_fn)  # fmt: skip
    batch_size = starting_batch_size

    def shrink():
        nonlocal batch_size
        batch_size = int(batch_size * 0.9) if reduce_batch_size_fn is None else reduce_batch_size_fn()
        return batch_size

    def attempt(*args, **kwargs):
        nonlocal batch_size
        clear_device_cache(garbage_collection=True)
        names = list(inspect.signature(function).parameters)
        if len(names) < len(args) + 1:  # the caller passed the batch size although the decorator supplies it
            shown = ", ".join(f"{n}={v}" for n, v in zip(names[1:], args[1:]))
            raise TypeError(
                f"Batch size was passed into `{function.__name__}` as the first argument when called."
                f"Remove this as the decorator already does so: `{function.__name__}({shown})`"
            )
        while True:
            if batch_size == 0:
                raise RuntimeError("No executable batch size found, reached zero.")
            try:
                return function(batch_size, *args, **kwargs)
            except Exception as err:
                if not should_reduce_batch_size(err):
                    raise
                clear_device_cache(garbage_collection=True)
                shrink()

    return attempt
