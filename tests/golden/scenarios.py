"""Seeded scenario definitions shared by the golden generator (run against the reference) and the
parity tests (run against accelerate_b200).  Pure torch; imports neither package."""

import torch

# (num_processes, dataset length, batch size, drop_last, split_batches, even_batches)
SAMPLER_GRID = [
    (n, length, bs, drop_last, split, even)
    for n in (2, 3, 4)
    for length in (0, 1, 2, 3, 7, 8, 20, 21, 22, 24, 25, 26)
    for bs in (1, 3, 4)
    for drop_last in (False, True)
    for split in (False, True)
    for even in (True, False)
]


def make_model():
    """Small MLP with a tied-size mix of weight shapes; identical on every rank (fixed seed)."""
    g = torch.Generator().manual_seed(42)
    model = torch.nn.Sequential(torch.nn.Linear(16, 33), torch.nn.Tanh(), torch.nn.Linear(33, 8))
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.3)
    return model


def make_batch(rank: int, step: int):
    g = torch.Generator().manual_seed(1234 + 17 * rank + 1000 * step)
    x = torch.randn(6, 16, generator=g)
    y = torch.randn(6, 8, generator=g)
    return x, y


def ops_inputs(rank: int, world: int):
    g = torch.Generator().manual_seed(7 + rank)
    return {
        "arange": torch.arange(world, dtype=torch.int64) + 1 + world * rank,
        "scalar": torch.tensor(float(rank) + 0.5),
        "f32": torch.randn(5, 3, generator=g),
        "bf16": torch.randn(4, 4, generator=g).to(torch.bfloat16).to(torch.float32),  # bf16-representable values
        "i32": torch.randint(-50, 50, (7,), generator=g, dtype=torch.int32),
        "bool": torch.rand(6, generator=g) > 0.5,
        "noncontig": torch.randn(4, 6, generator=g).t(),
    }


# (dataset length, batch size, drop_last, split_batches, dispatch_batches, iterable dataset)
LOADER_GRID = [
    (length, bs, drop_last, split, dispatch, iterable)
    for length in (22, 24)
    for bs in (4,)
    for drop_last in (False, True)
    for split in (False, True)
    for dispatch, iterable in ((False, False), (True, False), (None, True), (False, True))
]


class _Stream(torch.utils.data.IterableDataset):
    def __init__(self, n):
        self.n = n

    def __iter__(self):
        return iter(torch.arange(self.n).tolist())

    def __len__(self):
        return self.n


def loader_dataset(length: int, iterable: bool):
    return _Stream(length) if iterable else torch.arange(length)
