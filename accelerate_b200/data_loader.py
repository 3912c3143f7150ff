"""Per-process sharding of batches (stays in Python by the north star).

API and index-level behaviour follow the reference's ``src/accelerate/data_loader.py``:
``SeedableRandomSampler`` (:73-107), ``BatchSamplerShard`` (:110-263), ``IterableDatasetShard``
(:266-362), ``DataLoaderShard`` (:502-658), ``DataLoaderDispatcher`` (:714-985),
``prepare_data_loader`` (:1006-1319), ``skip_first_batches`` (:1385).  The shard index tables are
pinned against the reference by ``tests/golden/sampler_shards.json``.

What touches the hot path: ``DataLoaderShard`` moves each batch to the device one step ahead and
tells ``GradientState`` when the epoch ends (that is what forces a gradient sync on the last,
possibly short, accumulation window); ``DataLoaderDispatcher`` broadcasts rank 0's batch with the
P2P broadcast kernel on CUDA.
"""

from __future__ import annotations

import math
from contextlib import suppress

import torch
from torch.utils.data import BatchSampler, DataLoader, IterableDataset, RandomSampler

from .state import GradientState, PartialState
from .utils.dataclasses import DistributedType
from .utils.operations import (
    broadcast,
    broadcast_object_list,
    concatenate,
    find_batch_size,
    get_data_structure,
    initialize_tensors,
    send_to_device,
    slice_tensors,
)
from .utils.random import synchronize_rng_states

# kwargs a DataLoader is rebuilt with (name -> default)
_LOADER_KWARGS = {
    "batch_size": 1, "shuffle": False, "sampler": None, "batch_sampler": None, "num_workers": 0,
    "collate_fn": None, "pin_memory": False, "drop_last": False, "timeout": 0, "worker_init_fn": None,
    "multiprocessing_context": None, "generator": None, "prefetch_factor": None, "persistent_workers": False,
    "pin_memory_device": "",
}  # fmt: skip
if "in_order" in __import__("inspect").signature(DataLoader.__init__).parameters:  # torch >= 2.6 (the reference gates on the version)
    _LOADER_KWARGS["in_order"] = True


class SeedableRandomSampler(RandomSampler):
    """RandomSampler whose permutation is a pure function of (initial seed, epoch): every rank
    draws the same order, every epoch a new one."""

    def __init__(self, *args, **kwargs):
        data_seed = kwargs.pop("data_seed", None)
        super().__init__(*args, **kwargs)
        self.initial_seed = data_seed if data_seed is not None else torch.random.initial_seed()
        self.epoch = 0

    def __iter__(self):
        if self.generator is None:
            self.generator = torch.Generator()
        self.generator.manual_seed(self.initial_seed + self.epoch)
        yield from super().__iter__()
        self.set_epoch(self.epoch + 1)

    def set_epoch(self, epoch: int):
        self.epoch = epoch


class BatchSamplerShard(BatchSampler):
    """Deal the batches of ``batch_sampler`` to ``num_processes`` ranks.

    Whole-batch mode: batch k goes to rank k mod N, released only once a full round of N batches
    exists.  Split mode: every batch is cut in N equal slices.  With ``even_batches`` the ragged
    tail is completed with indices recycled from the start of the epoch so all ranks run the same
    number of equally sized steps (which is what lets the gradient allreduce assume every rank
    shows up).
    """

    def __init__(self, batch_sampler, num_processes: int = 1, process_index: int = 0, split_batches: bool = False,
                 even_batches: bool = True):  # fmt: skip
        if split_batches and batch_sampler.batch_size % num_processes != 0:
            raise ValueError(
                f"To use `BatchSamplerShard` in `split_batches` mode, the batch size ({batch_sampler.batch_size}) "
                f"needs to be a round multiple of the number of processes ({num_processes})."
            )
        self.batch_sampler = batch_sampler
        self.num_processes, self.process_index = num_processes, process_index
        self.split_batches, self.even_batches = split_batches, even_batches
        self.batch_size = getattr(batch_sampler, "batch_size", None)
        self.drop_last = getattr(batch_sampler, "drop_last", False)
        if self.batch_size is None and self.even_batches:
            raise ValueError(
                "You need to use `even_batches=False` when the batch sampler has no batch size. If you "
                "are not calling this method directly, set `accelerator.even_batches=False` instead."
            )

    @property
    def total_length(self):
        return len(self.batch_sampler)

    def __len__(self):
        total, n = len(self.batch_sampler), self.num_processes
        if self.split_batches:
            return total
        whole, extra = divmod(total, n)
        if extra == 0 or self.drop_last:
            return whole
        if self.even_batches:
            return whole + 1
        return whole + (1 if self.process_index < extra else 0)

    def __iter__(self):
        return self._sliced() if self.split_batches else self._dealt()

    def _sliced(self):
        n, rank, bs = self.num_processes, self.process_index, self.batch_size
        width = bs // n
        lo, hi = width * rank, width * (rank + 1)
        first, last = [], None
        for k, last in enumerate(self.batch_sampler):
            if k == 0:
                first = last
            if len(last) == bs:
                yield last[lo:hi]
        if self.drop_last or not first or len(last) == bs:
            return
        if not self.even_batches:
            if len(last) > lo:
                yield last[lo:hi]
            return
        filler = list(first)
        while len(filler) < bs:
            filler += filler
        yield (list(last) + filler)[lo:hi]

    def _dealt(self):
        n, rank, bs = self.num_processes, self.process_index, self.batch_size
        head, mine, last, k = [], [], None, -1
        for k, last in enumerate(self.batch_sampler):
            if not self.drop_last and k < n:
                head += last  # recycling material: the first N batches
            seat = k % n
            if seat == rank:
                mine = last
            if seat == n - 1 and (bs is None or len(last) == bs):
                yield mine  # a complete round: everybody gets a full batch
                mine = []
        if self.drop_last or not head:
            return
        if not self.even_batches:
            if len(mine) > 0:
                yield mine
            return
        if len(mine) == bs:
            yield mine
        while len(head) < n * bs:
            head += head
        # finish the round: top up the short batch, then mint the missing ones, from `head`
        short = [] if len(last) == bs else list(last)
        pos = k + 1 if len(last) == bs else k
        cursor = 0
        while pos % n != 0 or short:
            take = bs - len(short)
            filled = short + head[cursor : cursor + take]
            if pos % n == rank:
                yield filled
            cursor += take
            short = []
            pos += 1


class IterableDatasetShard(IterableDataset):
    """Stream an IterableDataset in groups of one *global* batch and keep this rank's part.  A final
    partial group is dropped (``drop_last``) or completed with the first group's elements."""

    def __init__(self, dataset, batch_size: int = 1, drop_last: bool = False, num_processes: int = 1,
                 process_index: int = 0, split_batches: bool = False):  # fmt: skip
        if split_batches and batch_size > 1 and batch_size % num_processes != 0:
            raise ValueError(
                f"To use `IterableDatasetShard` in `split_batches` mode, the batch size ({batch_size}) "
                f"needs to be a round multiple of the number of processes ({num_processes})."
            )
        self.dataset = dataset
        self.batch_size, self.drop_last = batch_size, drop_last
        self.num_processes, self.process_index = num_processes, process_index
        self.split_batches = split_batches
        self.epoch = 0

    def set_epoch(self, epoch):
        self.epoch = epoch
        if hasattr(self.dataset, "set_epoch"):
            self.dataset.set_epoch(epoch)

    def __len__(self):
        per_group = self.batch_size * self.num_processes
        groups = len(self.dataset) // per_group if self.drop_last else math.ceil(len(self.dataset) / per_group)
        return groups * self.batch_size

    def __iter__(self):
        gen = getattr(self.dataset, "generator", None)
        if not hasattr(self.dataset, "set_epoch") and isinstance(gen, torch.Generator):
            gen.manual_seed(self.epoch)
        group_size = self.batch_size if self.split_batches else self.batch_size * self.num_processes
        share = self.batch_size // self.num_processes if self.split_batches else self.batch_size
        lo = self.process_index * share
        first_group, group = None, []
        for item in self.dataset:
            group.append(item)
            if len(group) == group_size:
                yield from group[lo : lo + share]
                if first_group is None:
                    first_group = list(group)
                group = []
        if self.drop_last or not group:
            return
        if first_group is None:
            first_group = list(group)
        while len(group) < group_size:
            group += first_group
        yield from group[lo : lo + share]


_STATEFUL_VARIANTS: dict = {}


def _stateful_base():
    """torchdata's drop-in ``StatefulDataLoader`` (mid-epoch ``state_dict`` / ``load_state_dict``)."""
    try:
        from torchdata.stateful_dataloader import StatefulDataLoader
    except ImportError as err:
        raise ImportError(
            "StatefulDataLoader is not available. Please install torchdata version 0.8.0 or higher to use it."
        ) from err
    return StatefulDataLoader


class _StatefulSupport:
    """``use_stateful_dataloader=True`` (data_loader.py:408-499): the loader is built on torchdata's
    ``StatefulDataLoader`` instead of ``DataLoader`` and carries a ``state_dict`` that resumes exactly after the last
    batch it has YIELDED — these loaders fetch one batch ahead of what they hand out, so the underlying iterator's
    state is snapshotted before each look-ahead fetch.  The stateful flavour of a class is a subclass created on first
    use (``isinstance`` holds for the class itself and for ``StatefulDataLoader``)."""

    use_stateful_dataloader = False

    def __new__(cls, *args, **kwargs):
        if kwargs.get("use_stateful_dataloader"):
            base = _stateful_base()
            if not issubclass(cls, base):
                if cls not in _STATEFUL_VARIANTS:
                    _STATEFUL_VARIANTS[cls] = type(cls.__name__, (cls, base), {"__module__": cls.__module__})
                cls = _STATEFUL_VARIANTS[cls]
        return super().__new__(cls)

    def _start_state_tracking(self, use_stateful_dataloader: bool):
        self.use_stateful_dataloader = bool(use_stateful_dataloader)
        if self.use_stateful_dataloader:
            self.dl_state_dict = super().state_dict()

    def state_dict(self):
        return self.dl_state_dict

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)

    def adjust_state_dict_for_prefetch(self):
        """In a distributed run every batch this process yields consumed ``num_processes`` batches of the underlying
        sampler, and the snapshot is taken one of OUR batches early: the counters are ``num_processes - 1`` ahead."""
        state = PartialState()
        if state.distributed_type == DistributedType.NO:
            return
        ahead = state.num_processes - 1
        sd = self.dl_state_dict
        for key in ("_sampler_iter_yielded", "_num_yielded"):  # (absent with worker processes)
            if sd.get(key, 0) > 0:
                sd[key] -= ahead
        inner = sd.get("_index_sampler_state")
        if inner is not None and inner.get("samples_yielded", 0) > 0:
            inner["samples_yielded"] -= self.batch_size * ahead

    def _update_state_dict(self):
        if self.use_stateful_dataloader:
            self.dl_state_dict = super().state_dict()
            self.adjust_state_dict_for_prefetch()
            self.dl_state_dict["_iterator_finished"] = self.end_of_dataloader


class _InstanceItself:
    """``loader.base_dataloader``: the reference wraps a ``DataLoader`` (``DataLoaderAdapter``); these classes ARE one.
    An instance-only attribute there, so class-level access raises."""

    def __get__(self, obj, owner=None):
        if obj is None:
            raise AttributeError("base_dataloader")
        return obj


class DataLoaderStateMixin:
    """end_of_dataloader / remainder tracking shared by the loader classes (data_loader.py:365-407)."""

    end_of_dataloader = False
    remainder = -1

    base_dataloader = _InstanceItself()

    def reset(self):
        self.end_of_dataloader = False
        self.remainder = -1

    def begin(self):
        self.reset()
        with suppress(Exception):
            if not self._drop_last:
                length = getattr(self.dataset, "total_dataset_length", len(self.dataset))
                self.remainder = length % self.total_batch_size
        self.gradient_state._add_dataloader(self)

    def end(self):
        self.gradient_state._remove_dataloader(self)


def _forward_set_epoch(loader, epoch: int):
    bs = loader.batch_sampler
    if hasattr(bs, "set_epoch"):
        bs.set_epoch(epoch)
    if hasattr(bs, "sampler") and hasattr(bs.sampler, "set_epoch"):
        bs.sampler.set_epoch(epoch)
    inner = getattr(bs, "batch_sampler", None)
    if inner is not None and hasattr(inner, "sampler") and hasattr(inner.sampler, "set_epoch"):
        inner.sampler.set_epoch(epoch)
    elif hasattr(loader.dataset, "set_epoch"):
        loader.dataset.set_epoch(epoch)


class DataLoaderShard(_StatefulSupport, DataLoader, DataLoaderStateMixin):
    """A DataLoader that (1) synchronises RNGs at epoch start, (2) puts batches on the device one
    batch ahead, (3) flags the last batch in ``GradientState`` *before* yielding it."""

    def __init__(self, dataset, device=None, rng_types=None, synchronized_generator=None, skip_batches=0,
                 _drop_last: bool = False, _non_blocking: bool = False, use_stateful_dataloader: bool = False,
                 **kwargs):  # fmt: skip
        kwargs.pop("torch_device_mesh", None)
        super().__init__(dataset, **kwargs)
        self.device = device
        self.rng_types = rng_types
        self.synchronized_generator = synchronized_generator
        self.skip_batches = skip_batches
        self.gradient_state = GradientState()
        self._drop_last = _drop_last
        self._non_blocking = _non_blocking
        self.iteration = 0
        self._start_state_tracking(use_stateful_dataloader)

    def adjust_state_dict_for_prefetch(self):
        """Nothing to correct here: every process iterates its own shard and the snapshot is taken before the
        look-ahead fetch, so the counters already equal the batches handed to the user (data_loader.py:560-566).
        The dispatcher, whose rank 0 draws ``num_processes`` batches per step, keeps the default correction."""

    def __iter__(self):
        if self.rng_types is not None:
            synchronize_rng_states(self.rng_types, self.synchronized_generator)
        self.begin()
        self.set_epoch(self.iteration)
        source = super().__iter__()
        try:
            held = next(source)
        except StopIteration:
            self.end()
            return
        index = 0
        while True:
            if self.device is not None:
                held = send_to_device(held, self.device, non_blocking=self._non_blocking)
            self._update_state_dict()  # resume point = right after `held`
            try:
                upcoming = next(source)
            except StopIteration:
                self.end_of_dataloader = True
                self._update_state_dict()
                if index >= self.skip_batches:
                    yield held
                break
            if index >= self.skip_batches:
                yield held
            index += 1
            held = upcoming
        self.iteration += 1
        self.end()

    def set_epoch(self, epoch: int):
        if self.iteration != epoch:
            self.iteration = epoch
        _forward_set_epoch(self, epoch)

    @property
    def total_batch_size(self):
        bs = self.sampler if isinstance(self.sampler, BatchSampler) else self.batch_sampler
        if getattr(bs, "split_batches", False):
            return bs.batch_size
        return bs.batch_size * getattr(bs, "num_processes", 1)

    @property
    def total_dataset_length(self):
        return self.dataset.total_length if hasattr(self.dataset, "total_length") else len(self.dataset)

    def get_sampler(self):
        return get_sampler(self)

    def set_sampler(self, sampler):
        if isinstance(self.sampler, BatchSampler):
            self.sampler.sampler = sampler
        else:
            self.batch_sampler.sampler = sampler
            if hasattr(self.batch_sampler, "batch_sampler"):
                self.batch_sampler.batch_sampler.sampler = sampler


class DataLoaderDispatcher(_StatefulSupport, DataLoader, DataLoaderStateMixin):
    """Rank 0 reads the data; every step it announces the batch's structure (object broadcast) and
    then its tensors (P2P broadcast kernel on CUDA); each rank keeps its slice."""

    def __init__(self, dataset, split_batches: bool = False, skip_batches=0, _drop_last: bool = False,
                 _non_blocking: bool = False, slice_fn=None, use_stateful_dataloader: bool = False, **kwargs):  # fmt: skip
        kwargs.pop("torch_device_mesh", None)
        shuffle = False
        super().__init__(dataset, **kwargs)
        self.split_batches = split_batches
        if shuffle:
            torch.utils.data.graph_settings.apply_shuffle_settings(dataset, shuffle=shuffle)
        self.gradient_state = GradientState()
        self.state = PartialState()
        self._drop_last = _drop_last
        self._non_blocking = _non_blocking
        self.skip_batches = skip_batches
        self.slice_fn = slice_tensors if slice_fn is None else slice_fn
        self.iteration = 0
        self._start_state_tracking(use_stateful_dataloader)

    def _fetch(self, source):
        """-> (batch, structure, exhausted).  Only rank 0 touches ``source``; one object broadcast tells
        everybody what (if anything) is coming.  ``exhausted`` with a structure = the ragged remainder."""
        batch, info = None, [None, True]
        if self.state.process_index == 0:
            pieces, exhausted = [], False
            try:
                if self.split_batches:
                    self._update_state_dict()
                    batch = next(source)
                else:
                    for _ in range(self.state.num_processes):
                        self._update_state_dict()
                        pieces.append(next(source))
            except StopIteration:
                exhausted = True
            if not self.split_batches and pieces and not (exhausted and self._drop_last):
                try:
                    batch = concatenate(pieces, dim=0)
                except RuntimeError as err:
                    raise RuntimeError(
                        "You can't use batches of different size with `dispatch_batches=True` or when using an "
                        "`IterableDataset`. Either pass `dispatch_batches=False` and have each process fetch its own "
                        "batch or pass `split_batches=True`."
                    ) from err
            info = [get_data_structure(batch) if batch is not None else None, exhausted]
        broadcast_object_list(info)
        return batch, info[0], info[1]

    def __iter__(self):
        self.begin()
        self.set_epoch(self.iteration)
        rank, world = self.state.process_index, self.state.num_processes
        source = super().__iter__() if rank == 0 else None
        first_batch = None
        batch, structure, exhausted = self._fetch(source)
        index = 0
        while structure is not None:
            if rank != 0:
                batch = initialize_tensors(structure)
            batch = send_to_device(batch, self.state.device, non_blocking=self._non_blocking)
            batch = broadcast(batch, from_process=0)
            if not self._drop_last and first_batch is None:
                # enough leading samples to top up a ragged final batch
                first_batch = self.slice_fn(batch, slice(0, world), process_index=rank, num_processes=world)
            observed = find_batch_size(batch)
            share = observed // world
            last, upcoming = exhausted, None
            if not exhausted:  # look ahead: the final batch must be flagged before it is yielded
                upcoming = self._fetch(source)
                last = upcoming[2] and upcoming[1] is None
            if last and not self._drop_last and observed % world != 0:
                batch = concatenate([batch, first_batch], dim=0)
                share += 1
            window = slice(rank * share, (rank + 1) * share)
            mine = self.slice_fn(batch, window, process_index=rank, num_processes=world)
            if last:
                self.end_of_dataloader = True
                self._update_state_dict()
                self.remainder = observed
            if index >= self.skip_batches:
                yield mine
            index += 1
            if last:
                break
            batch, structure, exhausted = upcoming
        self.iteration += 1
        self.end()

    def set_epoch(self, epoch: int):
        if self.iteration != epoch:
            self.iteration = epoch
        _forward_set_epoch(self, epoch)

    def __len__(self):
        whole = super().__len__()
        if self.split_batches:
            return whole
        if self._drop_last:
            return whole // self.state.num_processes
        return math.ceil(whole / self.state.num_processes)

    @property
    def total_batch_size(self):
        return self.dataset.batch_size if self.split_batches else self.dataset.batch_size * self.dataset.num_processes

    @property
    def total_dataset_length(self):
        return len(self.dataset)

    def get_sampler(self):
        return get_sampler(self)

    def set_sampler(self, sampler):
        if isinstance(self.sampler, BatchSampler):
            self.sampler.sampler = sampler
        elif hasattr(self.batch_sampler, "batch_sampler"):
            self.batch_sampler.batch_sampler.sampler = sampler
        else:
            self.batch_sampler.sampler = sampler


def _is_hf_iterable(dataset) -> bool:
    try:
        from datasets import IterableDataset as HFIterableDataset
    except ImportError:
        return False
    return isinstance(dataset, HFIterableDataset)


def get_sampler(dataloader):
    if isinstance(dataloader.sampler, BatchSampler):
        return getattr(dataloader.sampler, "sampler", None)
    return getattr(dataloader.batch_sampler, "sampler", None)


def prepare_data_loader(dataloader, device=None, num_processes=None, process_index=None, split_batches=False,
                        put_on_device=False, rng_types=None, dispatch_batches=None, even_batches=True,
                        slice_fn_for_dispatch=None, use_seedable_sampler=False, data_seed=None, non_blocking=False,
                        use_stateful_dataloader=False, torch_device_mesh=None):  # fmt: skip
    """Rebuild ``dataloader`` so that this process only sees its share (data_loader.py:1006-1319)."""
    if use_stateful_dataloader:
        _stateful_base()  # fail early, with the reference's message, when torchdata is missing
    if dispatch_batches is None:
        dispatch_batches = put_on_device and isinstance(dataloader.dataset, IterableDataset)
    if dispatch_batches and not put_on_device:
        raise ValueError("Using `dispatch_batches=True` requires `put_on_device=True`.")
    state = PartialState()
    num_processes = state.num_processes if num_processes is None else num_processes
    process_index = state.process_index if process_index is None else process_index

    if split_batches:
        bs = dataloader.batch_size if dataloader.batch_size is not None else getattr(dataloader.batch_sampler, "batch_size", None)
        if bs is None:
            raise ValueError(
                "In order to use `split_batches==True` you must have a `batch_size` attribute either in the passed "
                "`dataloader` or `dataloader.batch_sampler` objects, and it has to return a natural number."
            )
        if bs > 1 and bs % num_processes != 0:
            raise ValueError(
                f"To use a `DataLoader` in `split_batches` mode, the batch size ({dataloader.batch_size}) "
                f"needs to be a round multiple of the number of processes ({num_processes})."
            )

    dataset = dataloader.dataset
    iterable = isinstance(dataset, IterableDataset)
    new_batch_sampler = None if iterable else dataloader.batch_sampler
    sampler_is_batch_sampler = isinstance(dataloader.sampler, BatchSampler)
    synchronized_generator = None
    sampler = get_sampler(dataloader)
    if isinstance(sampler, RandomSampler) and use_seedable_sampler:
        sampler = SeedableRandomSampler(
            data_source=sampler.data_source,
            replacement=sampler.replacement,
            num_samples=sampler._num_samples,
            generator=getattr(sampler, "generator", torch.Generator()),
            data_seed=data_seed,
        )

    if num_processes != 1 and not dispatch_batches:
        if _is_hf_iterable(dataset) and not split_batches and dataset.n_shards >= num_processes:
            # a Hugging Face streaming dataset with enough shards is split natively, whole shards per process
            # (data_loader.py:1214-1221); fewer shards fall back to striding through the stream below
            dataset = dataset.shard(num_shards=num_processes, index=process_index)
        elif iterable:
            if getattr(dataset, "generator", None) is not None:
                synchronized_generator = dataset.generator
            dataset = IterableDatasetShard(
                dataset, batch_size=dataloader.batch_size, drop_last=dataloader.drop_last, num_processes=num_processes,
                process_index=process_index, split_batches=split_batches,
            )  # fmt: skip
        else:
            if not use_seedable_sampler and hasattr(sampler, "generator"):
                if sampler.generator is None:
                    sampler.generator = torch.Generator()
                    sampler.generator.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
                synchronized_generator = sampler.generator
            inner = dataloader.sampler if sampler_is_batch_sampler else dataloader.batch_sampler
            new_batch_sampler = BatchSamplerShard(
                inner, num_processes=num_processes, process_index=process_index, split_batches=split_batches,
                even_batches=even_batches,
            )  # fmt: skip

    if rng_types is not None and synchronized_generator is None and "generator" in rng_types:
        rng_types.remove("generator")

    handled = {"batch_size", "shuffle", "sampler", "batch_sampler", "drop_last"}
    kwargs = {k: getattr(dataloader, k, default) for k, default in _LOADER_KWARGS.items() if k not in handled}
    if new_batch_sampler is None:  # iterable datasets batch inside the DataLoader
        kwargs["drop_last"] = dataloader.drop_last
        kwargs["batch_size"] = (
            dataloader.batch_size // num_processes if split_batches and not dispatch_batches else dataloader.batch_size
        )
    target = device if put_on_device else None
    if dispatch_batches:
        kwargs.pop("generator")
        new = DataLoaderDispatcher(
            dataset, split_batches=split_batches, batch_sampler=new_batch_sampler, _drop_last=dataloader.drop_last,
            _non_blocking=non_blocking, slice_fn=slice_fn_for_dispatch, use_stateful_dataloader=use_stateful_dataloader,
            **kwargs,
        )  # fmt: skip
    elif sampler_is_batch_sampler:
        new = DataLoaderShard(
            dataset, device=target, sampler=new_batch_sampler, batch_size=dataloader.batch_size, rng_types=rng_types,
            synchronized_generator=synchronized_generator, _drop_last=dataloader.drop_last, _non_blocking=non_blocking,
            use_stateful_dataloader=use_stateful_dataloader, **kwargs,
        )  # fmt: skip
    else:
        new = DataLoaderShard(
            dataset, device=target, batch_sampler=new_batch_sampler, rng_types=rng_types,
            synchronized_generator=synchronized_generator, _drop_last=dataloader.drop_last, _non_blocking=non_blocking,
            use_stateful_dataloader=use_stateful_dataloader, **kwargs,
        )  # fmt: skip
    if isinstance(sampler, SeedableRandomSampler) and use_seedable_sampler:
        new.set_sampler(sampler)
    return new


class SkipBatchSampler(BatchSampler):
    """Skips the first ``skip_batches`` batches of ``batch_sampler``."""

    def __init__(self, batch_sampler, skip_batches=0):
        self.batch_sampler = batch_sampler
        self.skip_batches = skip_batches

    def __iter__(self):
        for k, batch in enumerate(self.batch_sampler):
            if k >= self.skip_batches:
                yield batch

    @property
    def total_length(self):
        return len(self.batch_sampler)

    def __len__(self):
        return len(self.batch_sampler) - self.skip_batches


class SkipDataLoader(_StatefulSupport, DataLoader, DataLoaderStateMixin):
    def __init__(self, dataset, skip_batches=0, use_stateful_dataloader: bool = False, **kwargs):
        super().__init__(dataset, **kwargs)
        self.skip_batches = skip_batches
        self.gradient_state = GradientState()
        self._drop_last = kwargs.get("drop_last", False)
        self._start_state_tracking(use_stateful_dataloader)

    def __iter__(self):
        self.begin()
        for k, batch in enumerate(super().__iter__()):
            if k >= self.skip_batches:
                self._update_state_dict()
                yield batch
        self.end()

    def __len__(self):
        return super().__len__() - self.skip_batches

    @property
    def total_batch_size(self):
        return self.batch_size


def skip_first_batches(dataloader, num_batches=0):
    """A dataloader that resumes ``num_batches`` into the epoch (mid-epoch checkpoint resume)."""
    dataset = dataloader.dataset
    sampler_is_batch_sampler = isinstance(dataloader.sampler, BatchSampler)
    if isinstance(dataset, IterableDataset):
        new_batch_sampler = None
    else:
        inner = dataloader.sampler if sampler_is_batch_sampler else dataloader.batch_sampler
        new_batch_sampler = SkipBatchSampler(inner, skip_batches=num_batches)
    handled = {"batch_size", "shuffle", "sampler", "batch_sampler", "drop_last"}
    kwargs = {k: getattr(dataloader, k, default) for k, default in _LOADER_KWARGS.items() if k not in handled}
    if new_batch_sampler is None:
        kwargs["drop_last"] = dataloader.drop_last
        kwargs["batch_size"] = dataloader.batch_size
    if isinstance(dataloader, DataLoaderDispatcher):
        if new_batch_sampler is None:
            kwargs["skip_batches"] = num_batches
        kwargs.pop("generator", None)
        return DataLoaderDispatcher(dataset, split_batches=dataloader.split_batches, batch_sampler=new_batch_sampler,
                                    _drop_last=dataloader._drop_last, **kwargs)  # fmt: skip
    if isinstance(dataloader, DataLoaderShard):
        if new_batch_sampler is None:
            kwargs["skip_batches"] = num_batches
        extra = dict(device=dataloader.device, rng_types=dataloader.rng_types,
                     synchronized_generator=dataloader.synchronized_generator, _drop_last=dataloader._drop_last)  # fmt: skip
        if sampler_is_batch_sampler:
            return DataLoaderShard(dataset, sampler=new_batch_sampler, batch_size=dataloader.batch_size, **extra, **kwargs)
        return DataLoaderShard(dataset, batch_sampler=new_batch_sampler, **extra, **kwargs)
    if new_batch_sampler is None:
        return SkipDataLoader(dataset, skip_batches=num_batches, **kwargs)
    return DataLoader(dataset, batch_sampler=new_batch_sampler, **kwargs)
