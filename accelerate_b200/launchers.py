"""``debug_launcher``: run a function in a few forked CPU processes that form a gloo group — what the reference
offers for debugging and what its own CPU tests are built on (``src/accelerate/launchers.py:276-309``).  The CLI and
notebook launchers of the reference are out of scope: real runs start one process per GPU with ``torchrun``."""

import os
import socket

from .utils.environment import patch_environment


class _ForkedWorker:
    def __init__(self, function):
        self.function = function

    def __call__(self, index, *args):
        os.environ["RANK"] = os.environ["LOCAL_RANK"] = str(index)
        self.function(*args)


def debug_launcher(function, args=(), num_processes: int = 2):
    """Call ``function(*args)`` in ``num_processes`` forked processes on the CPU (``DistributedType.MULTI_CPU``, gloo).
    Meant for tests and debugging only."""
    from torch.multiprocessing import start_processes

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with patch_environment(
        world_size=num_processes, local_world_size=num_processes, master_addr="127.0.0.1", master_port=str(port),
        accelerate_mixed_precision="no", accelerate_use_cpu="yes", fork_launched="1", omp_num_threads="1",
    ):  # fmt: skip
        start_processes(_ForkedWorker(function), args=args, nprocs=num_processes, start_method="fork")
