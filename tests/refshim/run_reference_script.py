"""Run one of the REFERENCE'S OWN distributed test scripts (``/root/reference/src/accelerate/test_utils/scripts``)
against this package: ``import accelerate`` (and every ``accelerate.x.y``) is aliased to ``accelerate_b200``; the two
helper modules the scripts use — ``accelerate.test_utils`` (RegressionModel/RegressionDataset: the reference's
training.py is executed from where it lies, nothing is copied) and ``accelerate.test_utils.testing`` (two tiny
assertion helpers, restated here) — are provided by this shim.  Build-container only: the scripts are read from the
reference tree at run time.

    torchrun --nproc-per-node 2 tests/refshim/run_reference_script.py test_sync.py
"""
import importlib
import importlib.abc
import importlib.util
import os
import runpy
import sys
import types
from contextlib import contextmanager

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/src/accelerate"
sys.path.insert(0, ROOT)


class AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """accelerate[.sub] -> the very module object of accelerate_b200[.sub] (one copy: singletons stay shared)."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "accelerate" and not fullname.startswith("accelerate."):
            return None
        if fullname.startswith("accelerate.test_utils"):
            return None  # served from sys.modules (installed below)
        return importlib.util.spec_from_loader(fullname, self)

    def create_module(self, spec):
        return importlib.import_module("accelerate_b200" + spec.name[len("accelerate") :])

    def exec_module(self, module):
        pass


def install():
    import torch

    sys.meta_path.insert(0, AliasFinder())
    import accelerate  # noqa: F401  (now accelerate_b200)
    from accelerate_b200.state import AcceleratorState
    from accelerate_b200.utils import gather

    test_utils = types.ModuleType("accelerate.test_utils")
    test_utils.__path__ = []
    sys.modules["accelerate.test_utils"] = test_utils
    spec = importlib.util.spec_from_file_location("accelerate.test_utils.training", os.path.join(REF, "test_utils", "training.py"))
    training = importlib.util.module_from_spec(spec)
    sys.modules["accelerate.test_utils.training"] = training
    spec.loader.exec_module(training)
    for name in ("RegressionDataset", "RegressionModel"):
        setattr(test_utils, name, getattr(training, name))

    def are_the_same_tensors(tensor):
        state = AcceleratorState()
        mine = tensor[None].clone().to(state.device)
        everyone = gather(mine).cpu()
        return all(torch.equal(everyone[i], mine[0].cpu()) for i in range(everyone.shape[0]))

    @contextmanager
    def assert_exception(exception_class, msg=None):
        ran = False
        try:
            yield
            ran = True
        except Exception as err:
            assert isinstance(err, exception_class), f"Expected exception of type {exception_class} but got {type(err)}"
            if msg is not None:
                assert msg in str(err), f"Expected message '{msg}' to be in exception but got '{str(err)}'"
        if ran:
            raise AssertionError(f"Expected exception of type {exception_class} but ran without issue.")

    testing = types.ModuleType("accelerate.test_utils.testing")
    testing.are_the_same_tensors = are_the_same_tensors
    testing.assert_exception = assert_exception
    sys.modules["accelerate.test_utils.testing"] = testing
    test_utils.testing = testing
    test_utils.training = training
    test_utils.are_the_same_tensors = are_the_same_tensors
    test_utils.assert_exception = assert_exception


def data_loop_on_cpu(path):
    """test_distributed_data_loop.py minus its join_uneven_inputs cases — those need one GPU per process (on MULTI_CPU the
    reference's join is a documented no-op and uneven inputs would simply hang DDP) — and minus the torchdata
    StatefulDataLoader cases; everything else runs as written."""
    import torch
    from torch.utils.data import BatchSampler, DataLoader, RandomSampler, default_collate

    G = runpy.run_path(path, run_name="reference_data_loop")
    accelerator = G["create_accelerator"]()
    torch.manual_seed(accelerator.process_index)
    G["test_default_ensures_even_batch_sizes"]()
    G["test_can_disable_even_batches"]()
    G["test_pickle_accelerator"]()
    dataset = G["DummyDataset"]()
    bs, workers = G["BATCH_SIZE"], G["NUM_WORKERS"]
    G["test_data_loader"](DataLoader(dataset, shuffle=False, batch_size=bs, num_workers=workers), accelerator)
    G["test_data_loader"](DataLoader(dataset, shuffle=True, batch_size=bs, num_workers=workers), accelerator)
    sampler = BatchSampler(RandomSampler(dataset), batch_size=bs, drop_last=False)
    G["test_data_loader"](DataLoader(dataset, batch_sampler=sampler, num_workers=workers), accelerator)
    sampler = BatchSampler(RandomSampler(dataset), batch_size=bs, drop_last=False)
    G["test_data_loader"](DataLoader(dataset, sampler=sampler, batch_size=None, collate_fn=default_collate, num_workers=workers),
                          accelerator)  # fmt: skip
    # (test_stateful_dataloader*: torchdata's StatefulDataLoader is not supported by this build)
    accelerator.end_training()


def main():
    script = sys.argv[1]
    sys.argv = [script] + sys.argv[2:]
    install()
    path = os.path.join(REF, "test_utils", "scripts", script)
    if script == "test_distributed_data_loop.py":
        data_loop_on_cpu(path)
    else:
        runpy.run_path(path, run_name="__main__")
    print(f"REFERENCE_SCRIPT_OK {script} rank {os.environ.get('RANK', '0')}", flush=True)


if __name__ == "__main__":
    main()
