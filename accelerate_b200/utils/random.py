"""Seeding and cross-rank RNG synchronisation (reference: ``src/accelerate/utils/random.py``).

``synchronize_rng_states`` runs once per dataloader epoch on a few KB: it stays on
``torch.distributed`` (SURVEY.md §2a row 10 — not hot)."""

from __future__ import annotations

import random

import numpy as np
import torch
import torch.distributed as dist

from .dataclasses import DistributedType, RNGType


def PartialState():
    from ..state import PartialState as _PartialState

    return _PartialState()


def set_seed(seed: int, device_specific: bool = False, deterministic: bool = False):
    """Seed python, numpy and torch (all devices); ``device_specific`` offsets by the process index."""
    if device_specific:
        seed += PartialState().process_index
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.use_deterministic_algorithms(True)


def synchronize_rng_state(rng_type: RNGType | None = None, generator: torch.Generator | None = None):
    rng_type = RNGType(rng_type) if rng_type is not None else None
    state = PartialState()
    if rng_type == RNGType.TORCH:
        rng_state = torch.get_rng_state()
    elif rng_type == RNGType.CUDA:
        rng_state = torch.cuda.get_rng_state()
    elif rng_type == RNGType.GENERATOR:
        if generator is None:
            raise ValueError("Need a generator to synchronize its seed.")
        rng_state = generator.get_state()
    else:
        raise ValueError(f"unsupported rng type {rng_type}")
    if state.distributed_type != DistributedType.NO:
        wire = rng_state.to(state.device) if state.backend == "nccl" else rng_state
        dist.broadcast(wire, 0)
        rng_state = wire.cpu()
    if rng_type == RNGType.TORCH:
        torch.set_rng_state(rng_state)
    elif rng_type == RNGType.CUDA:
        torch.cuda.set_rng_state(rng_state)
    else:
        generator.set_state(rng_state)


def synchronize_rng_states(rng_types, generator: torch.Generator | None = None):
    for rng_type in rng_types:
        synchronize_rng_state(RNGType(rng_type), generator=generator)
