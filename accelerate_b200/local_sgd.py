"""LocalSGD: run ``local_sgd_steps`` optimizer steps without gradient synchronisation, then average the
model parameters across ranks (reference: ``src/accelerate/local_sgd.py:20-107``).

The reference averages with one ``accelerator.reduce(param, "mean")`` per parameter (clone + NCCL allreduce +
divide, 3 x 292 launches for GPT-2-medium).  On CUDA the wrapped model already owns a symmetric, peer-mapped
buffer with one slot per parameter, so the average is: ONE multi-tensor pack, ONE fused allreduce over the whole
buffer, ONE multi-tensor unpack, one foreach divide — same values (sum in rank order, then true division by N).
"""

from __future__ import annotations

import torch

from .reducer import B200DataParallel
from .utils.dataclasses import DistributedType


class LocalSGD:
    def __init__(self, accelerator, model: torch.nn.Module, local_sgd_steps: int, enabled: bool = True):
        if accelerator.distributed_type not in (DistributedType.NO, DistributedType.MULTI_CPU, DistributedType.MULTI_GPU):
            raise NotImplementedError("LocalSGD is supported only for CPUs and GPUs (no DeepSpeed or MegatronLM)")
        self.enabled = enabled and accelerator.distributed_type != DistributedType.NO
        self.num_steps = 0
        if self.enabled:
            self.accelerator = accelerator
            self.model = model
            self.local_sgd_steps = local_sgd_steps

    def __enter__(self):
        if self.enabled:
            self._no_sync = self.model.no_sync()
            self._no_sync.__enter__()
        return self

    def __exit__(self, exc_type, exc, tb):
        if self.enabled:
            self._sync_and_avg_model_params()  # average once more on exit
            self._no_sync.__exit__(exc_type, exc, tb)

    def step(self):
        self.num_steps += 1
        if self.enabled and self.num_steps % self.local_sgd_steps == 0:
            self._sync_and_avg_model_params()

    def _sync_and_avg_model_params(self):
        self.accelerator.wait_for_everyone()
        model = self.model
        if isinstance(model, B200DataParallel) and model.reducer.is_cuda and model.reducer.can_average_in_buffer():
            model.reducer.average_parameters_()
            # the reference averages EVERY parameter (local_sgd.py:99-107); frozen ones have no slot in the fusion buffer
            for param in model.parameters():
                if not param.requires_grad:
                    param.data = self.accelerator.reduce(param.data, reduction="mean")
            return
        with self.accelerator.autocast():
            for param in model.parameters():
                param.data = self.accelerator.reduce(param.data, reduction="mean")
