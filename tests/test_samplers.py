"""Shard index tables of BatchSamplerShard / IterableDatasetShard: exact equality with the tables the
reference produced (tests/golden/sampler_shards.json; reference tests: tests/test_data_loader.py:109-400)."""

import json
import os

import pytest
import torch
from torch.utils.data import BatchSampler, SequentialSampler

from accelerate_b200.data_loader import BatchSamplerShard, IterableDatasetShard

TABLE = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_shards.json")))


def test_batch_sampler_shard_tables():
    assert len(TABLE["batch_sampler"]) > 500
    for case in TABLE["batch_sampler"]:
        n = case["n"]
        inner = BatchSampler(SequentialSampler(range(case["length"])), batch_size=case["bs"], drop_last=case["drop_last"])
        for r in range(n):
            shard = BatchSamplerShard(inner, n, r, split_batches=case["split"], even_batches=case["even"])
            got = [list(b) for b in shard]
            assert got == case["shards"][r], (case, r, got)
            assert len(shard) == case["lens"][r], (case, r)


def test_iterable_dataset_shard_tables():
    assert len(TABLE["iterable"]) > 200
    for case in TABLE["iterable"]:
        length = case["length"]

        class Stream(torch.utils.data.IterableDataset):
            def __iter__(self):
                return iter(range(length))

        for r in range(case["n"]):
            got = list(IterableDatasetShard(Stream(), batch_size=case["bs"], drop_last=case["drop_last"],
                                            num_processes=case["n"], process_index=r, split_batches=case["split"]))  # fmt: skip
            assert got == case["shards"][r], (case, r)


def test_argument_validation():
    inner = BatchSampler(SequentialSampler(range(10)), batch_size=3, drop_last=False)
    with pytest.raises(ValueError, match="round multiple"):
        BatchSamplerShard(inner, 2, 0, split_batches=True)

    class NoSize:
        def __iter__(self):
            return iter([[0, 1], [2]])

        def __len__(self):
            return 2

    with pytest.raises(ValueError, match="even_batches=False"):
        BatchSamplerShard(NoSize(), 2, 0)
    got = [list(BatchSamplerShard(NoSize(), 2, r, even_batches=False)) for r in range(2)]
    assert got == [[[0, 1]], [[2]]]
