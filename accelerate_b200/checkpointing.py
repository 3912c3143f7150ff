"""save_state / load_state for the prepared objects, file-compatible with the reference's checkpoints
(``src/accelerate/checkpointing.py:63-338``, file names from ``utils/constants.py:20-36``):

    model.safetensors | pytorch_model.bin   (``_1``, ``_2`` ... for further models; keys are the UNWRAPPED module's,
                                             no ``module.`` prefix: like the reference, ``Accelerator._models`` keeps
                                             the module it was handed, not the data-parallel wrapper)
    optimizer.bin, scheduler.bin, sampler.bin, scaler.pt
    random_states_{process_index}.pkl       {"step", "random_state", "numpy_random_seed", "torch_manual_seed",
                                             "torch_cuda_manual_seed"}
    custom_checkpoint_{i}.pkl

so a run can be checkpointed under one package and resumed under the other.  Off the
step path: plain Python + torch.save / safetensors.
"""

from __future__ import annotations

import os
import random
import re
import shutil
from pathlib import Path

import numpy as np
import torch

MODEL_NAME, SAFE_MODEL_NAME = "pytorch_model", "model"
OPTIMIZER_NAME, SCHEDULER_NAME, SAMPLER_NAME = "optimizer", "scheduler", "sampler"
SCALER_NAME, RNG_STATE_NAME = "scaler.pt", "random_states"


def _numbered(stem: str, index: int, ext: str) -> str:
    return f"{stem}{'' if index == 0 else f'_{index}'}.{ext}"


def _write(obj, path: Path, is_writer: bool, safe: bool = False):
    if not is_writer:
        return
    if safe:
        from safetensors.torch import save_file

        # safetensors refuses aliased storage: keep one name per storage (tied weights), contiguous CPU copies
        seen, flat = {}, {}
        for name, t in obj.items():
            if not isinstance(t, torch.Tensor):
                continue
            key = (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape)) if t.numel() else name
            if key in seen:
                continue
            seen[key] = name
            flat[name] = t.detach().cpu().contiguous()
        save_file(flat, str(path), metadata={"format": "pt"})
    else:
        torch.save(obj, str(path))


def save_accelerator_state(output_dir, model_states, optimizers, schedulers, dataloaders, process_index: int, step: int,
                           scaler=None, save_on_each_node: bool = False, safe_serialization: bool = True,
                           is_main_process: bool = True, is_local_main_process: bool = True):  # fmt: skip
    out = Path(output_dir)
    writer = is_local_main_process if save_on_each_node else is_main_process
    for i, state in enumerate(model_states):
        name = _numbered(SAFE_MODEL_NAME, i, "safetensors") if safe_serialization else _numbered(MODEL_NAME, i, "bin")
        _write(state, out / name, writer, safe=safe_serialization)
    for i, opt in enumerate(optimizers):
        _write(opt.state_dict(), out / _numbered(OPTIMIZER_NAME, i, "bin"), writer)
    for i, sched in enumerate(schedulers):
        _write(sched.state_dict(), out / _numbered(SCHEDULER_NAME, i, "bin"), writer)
    from .data_loader import IterableDatasetShard, SeedableRandomSampler

    for i, dl in enumerate(dataloaders):
        if isinstance(dl.dataset, IterableDatasetShard):
            sampler = dl.get_sampler()
            if isinstance(sampler, SeedableRandomSampler):
                _write(sampler, out / _numbered(SAMPLER_NAME, i, "bin"), writer)
        if getattr(dl, "use_stateful_dataloader", False):
            # resume point inside the epoch; like the reference (checkpointing.py:141-145) EVERY process writes it
            _write(dl.state_dict(), out / _numbered("dl_state_dict", i, "bin"), True)
    if scaler is not None:
        _write(scaler.state_dict(), out / SCALER_NAME, writer)
    # every process keeps its own RNG streams
    states = {
        "step": step,
        "random_state": random.getstate(),
        "numpy_random_seed": np.random.get_state(),
        "torch_manual_seed": torch.get_rng_state(),
    }
    if torch.cuda.is_available():
        states["torch_cuda_manual_seed"] = torch.cuda.get_rng_state_all()
    torch.save(states, str(out / f"{RNG_STATE_NAME}_{process_index}.pkl"))
    return out


def _read(path: Path, map_location="cpu", weights_only: bool = True):
    try:
        return torch.load(str(path), map_location=map_location, weights_only=weights_only)
    except Exception:
        if weights_only:  # optimizer/sampler pickles written by older torch versions
            return torch.load(str(path), map_location=map_location, weights_only=False)
        raise


def load_accelerator_state(input_dir, models, optimizers, schedulers, dataloaders, process_index: int, scaler=None,
                           map_location=None, device=None, **load_model_kwargs):  # fmt: skip
    if map_location not in (None, "cpu", "on_device"):
        raise TypeError("Unsupported optimizer map location passed, please choose one of `None`, `'cpu'`, or `'on_device'`")
    where = "cpu" if map_location in (None, "cpu") else device
    src = Path(input_dir)
    for i, model in enumerate(models):
        safe = src / _numbered(SAFE_MODEL_NAME, i, "safetensors")
        if safe.exists():
            from safetensors.torch import load_file

            state = load_file(str(safe), device=str(where))
            # names dropped at save time because they aliased another tensor (tied weights) come back through
            # the module's own tying, so allow exactly those to be missing
            missing, unexpected = model.load_state_dict(state, strict=False, **load_model_kwargs)
            own = dict(model.state_dict())
            really_missing = [k for k in missing if not any(own[k].data_ptr() == own[o].data_ptr() for o in state if o in own)]
            if really_missing or unexpected:
                raise RuntimeError(f"Error(s) in loading state_dict: missing {really_missing}, unexpected {list(unexpected)}")
        else:
            model.load_state_dict(_read(src / _numbered(MODEL_NAME, i, "bin"), where), **load_model_kwargs)
    for i, opt in enumerate(optimizers):
        opt.load_state_dict(_read(src / _numbered(OPTIMIZER_NAME, i, "bin"), where))
    for i, sched in enumerate(schedulers):
        sched.load_state_dict(_read(src / _numbered(SCHEDULER_NAME, i, "bin")))
    from .data_loader import IterableDatasetShard, SeedableRandomSampler

    for i, dl in enumerate(dataloaders):
        if isinstance(dl.dataset, IterableDatasetShard) and isinstance(dl.get_sampler(), SeedableRandomSampler):
            dl.set_sampler(_read(src / _numbered(SAMPLER_NAME, i, "bin"), weights_only=False))
        if getattr(dl, "use_stateful_dataloader", False):
            state_file = src / _numbered("dl_state_dict", i, "bin")
            if state_file.exists():
                dl.load_state_dict(_read(state_file, weights_only=False))
    if scaler is not None:
        scaler.load_state_dict(_read(src / SCALER_NAME))
    overrides = {}
    try:
        states = _read(src / f"{RNG_STATE_NAME}_{process_index}.pkl", weights_only=False)
        if "step" in states:
            overrides["step"] = states["step"]
        random.setstate(states["random_state"])
        np.random.set_state(states["numpy_random_seed"])
        torch.set_rng_state(states["torch_manual_seed"])
        if torch.cuda.is_available() and "torch_cuda_manual_seed" in states:
            torch.cuda.set_rng_state_all(states["torch_cuda_manual_seed"])
    except Exception:
        pass  # same leniency as the reference: a checkpoint without RNG state still loads
    return overrides


def save_custom_state(obj, path, index: int = 0, is_writer: bool = True):
    _write(obj.state_dict(), Path(path) / f"custom_checkpoint_{index}.pkl", is_writer)


def load_custom_state(obj, path, index: int = 0):
    obj.load_state_dict(_read(Path(path) / f"custom_checkpoint_{index}.pkl", "cpu", weights_only=False))


def rotate_checkpoints(folder: str, total_limit: int | None):
    """Keep at most ``total_limit - 1`` existing ``checkpoint_<n>`` folders before a new one is written."""
    if total_limit is None or not os.path.isdir(folder):
        return
    def number(name):
        m = re.findall(r"[\\/]?([0-9]+)(?=[^\\/]*$)", name)
        return int(m[0]) if m else -1

    existing = sorted((os.path.join(folder, d) for d in os.listdir(folder)), key=number)
    for stale in existing[: max(0, len(existing) + 1 - total_limit)]:
        shutil.rmtree(stale, ignore_errors=True)
