"""accelerate_b200 — a Blackwell-native data-parallel training wrapper with the Accelerate API.

``Accelerator`` / ``prepare`` / ``backward`` / ``clip_grad_norm_`` / ``gather`` / ``reduce`` /
``pad_across_processes`` behave like huggingface/accelerate's; the per-step gradient reduction
underneath is hand-written sm_100a CUDA (``csrc/``) reached through a C ABI (``include/b200dp.h``).
"""

__version__ = "0.1.0"

from .accelerator import Accelerator
from .data_loader import skip_first_batches
from .launchers import debug_launcher
from .local_sgd import LocalSGD
from .optimizer import AcceleratedOptimizer
from .reducer import B200DataParallel
from .scheduler import AcceleratedScheduler
from .state import AcceleratorState, GradientState, PartialState
from .utils import (
    find_executable_batch_size,
    synchronize_rng_states,
    AutocastKwargs,
    DataLoaderConfiguration,
    DDPCommunicationHookType,
    DistributedDataParallelKwargs,
    DistributedType,
    GradientAccumulationPlugin,
    GradScalerKwargs,
    InitProcessGroupKwargs,
    ProjectConfiguration,
)
