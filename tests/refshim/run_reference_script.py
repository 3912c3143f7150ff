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


class OutOfScope:
    """Stands in for a name of the reference this build deliberately does not provide (big-model inference, hooks,
    FP8, profiling ...): importing it works, USING it skips the test."""

    def __init__(self, name):
        self._name = name

    def _skip(self, *args, **kwargs):
        import inspect
        import unittest

        why = f"`{self._name}` is outside the data-parallel scope of this build"
        if len(args) == 1 and not kwargs and (inspect.isfunction(args[0]) or inspect.isclass(args[0])):
            return unittest.skip(why)(args[0])  # used as a decorator: the decorated test skips when it runs
        raise unittest.SkipTest(why)

    __call__ = __enter__ = _skip

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return OutOfScope(f"{self._name}.{item}")


class StubModule(types.ModuleType):
    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return OutOfScope(f"{self.__name__}.{item}")


OUT_OF_SCOPE_MODULES = ("accelerate.big_modeling", "accelerate.hooks", "accelerate.utils.modeling",
                        "accelerate.utils.offload", "accelerate.tracking", "accelerate.inference")  # fmt: skip
OUT_OF_SCOPE_NAMES = {
    "accelerate_b200": ["infer_auto_device_map", "init_empty_weights", "load_checkpoint_and_dispatch", "cpu_offload",
                        "dispatch_model", "notebook_launcher", "find_executable_batch_size"],
    "accelerate_b200.utils": ["FP8RecipeKwargs", "ProfileKwargs", "TorchDynamoPlugin", "check_os_kernel",
                              "convert_dict_to_env_variables", "has_offloaded_params", "purge_accelerate_environment",
                              "save", "DeepSpeedPlugin", "FullyShardedDataParallelPlugin"],
}  # fmt: skip


class AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """accelerate[.sub] -> the very module object of accelerate_b200[.sub] (one copy: singletons stay shared)."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "accelerate" and not fullname.startswith("accelerate."):
            return None
        if fullname.startswith("accelerate.test_utils") or fullname in OUT_OF_SCOPE_MODULES:
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

    for modname in OUT_OF_SCOPE_MODULES:
        sys.modules[modname] = StubModule(modname)
    for modname, names in OUT_OF_SCOPE_NAMES.items():
        mod = importlib.import_module(modname)
        for name in names:
            if not hasattr(mod, name):
                setattr(mod, name, OutOfScope(name))
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

    import unittest

    class AccelerateTestCase(unittest.TestCase):
        """Resets the process-global state after every test (testing.py:667-677)."""

        def tearDown(self):
            super().tearDown()
            AcceleratorState._reset_state(True)
            from accelerate_b200.state import GradientState

            GradientState._reset_state()

    def skip_unless(condition, why):
        return unittest.skipUnless(condition, why)

    cuda = torch.cuda.is_available()

    class Testing(types.ModuleType):
        """``require_*`` decorators resolve lazily: hardware / extras this container lacks skip the test."""

        def __getattr__(self, name):
            if name.startswith("require_"):
                yes = {
                    "require_cpu": not cuda, "require_datasets": importlib.util.find_spec("datasets") is not None,
                    "require_non_torch_xla": True, "require_non_hpu": True, "require_fp16": True,
                    "require_non_cpu": cuda, "require_cuda": cuda, "require_cuda_or_xpu": cuda, "require_multi_device": False,
                    "require_torchdata_stateful_dataloader": importlib.util.find_spec("torchdata") is not None,
                }.get(name, False)  # fmt: skip
                if name == "require_torch_min_version":
                    return lambda version=None: (lambda test: test)
                return skip_unless(yes, f"{name} is not satisfied in this container")
            raise AttributeError(name)

    testing = Testing("accelerate.test_utils.testing")
    testing.AccelerateTestCase = AccelerateTestCase
    testing.slow = unittest.skip("slow test")
    testing.run_first = lambda test: test
    testing.torch_device = "cuda" if cuda else "cpu"
    testing.device_count = torch.cuda.device_count() if cuda else 1
    testing.DEFAULT_LAUNCH_COMMAND = ["torchrun", "--nproc_per_node=1"]
    testing.are_the_same_tensors = are_the_same_tensors
    testing.assert_exception = assert_exception
    testing.memory_allocated_func = torch.cuda.memory_allocated if cuda else (lambda *a: 0)
    testing.path_in_accelerate_package = lambda *parts: os.path.join(REF, *parts)
    testing.get_launch_command = lambda **kw: ["torchrun"] + [f"--{k}={v}" for k, v in kw.items()]

    def execute_subprocess_async(cmd, env=None, **kw):  # only reached by multi-device tests, which are skipped here
        raise unittest.SkipTest("subprocess launches through the reference's CLI are out of scope")

    testing.execute_subprocess_async = execute_subprocess_async
    # the two script modules the reference's tests import by name
    for name in ("test_sync", "test_ops", "test_script"):
        spec_s = importlib.util.spec_from_file_location(f"accelerate.test_utils.scripts.{name}",
                                                        os.path.join(REF, "test_utils", "scripts", f"{name}.py"))  # fmt: skip
        lazy = importlib.util.LazyLoader(spec_s.loader)
        spec_s.loader = lazy
        mod = importlib.util.module_from_spec(spec_s)
        sys.modules[spec_s.name] = mod
        lazy.exec_module(mod)
        setattr(testing, name, mod)
    sys.modules["accelerate.test_utils.testing"] = testing
    test_utils.testing = testing
    test_utils.training = training
    test_utils.__getattr__ = lambda name: getattr(testing, name)  # `from accelerate.test_utils import require_cpu, ...`


def data_loop_on_cpu(path):
    """test_distributed_data_loop.py minus its join_uneven_inputs cases — those need one GPU per process (on MULTI_CPU the
    reference's join is a documented no-op and uneven inputs would simply hang DDP); everything else runs as written."""
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
    G["test_stateful_dataloader"](accelerator)
    G["test_stateful_dataloader_save_state"](accelerator)
    accelerator.end_training()


def run_pytest(args):
    """``--pytest tests/test_data_loader.py [-k expr ...]``: a unit-test FILE of the reference, under pytest."""
    import pytest

    install()
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))  # the `parameterized` stand-in
    target = os.path.join("/root/reference", args[0])
    # nothing may be written next to the read-only reference sources
    sys.exit(pytest.main([target, "-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", "-c", os.devnull, *args[1:]]))


def main():
    if sys.argv[1] == "--pytest":
        run_pytest(sys.argv[2:])
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
