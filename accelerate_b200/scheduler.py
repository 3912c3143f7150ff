"""AcceleratedScheduler (reference: ``src/accelerate/scheduler.py:25-98``): an LR scheduler that only
advances when the optimizer really stepped, and advances ``num_processes`` times per step unless the
dataloader splits one global batch across ranks."""

from __future__ import annotations

from .state import AcceleratorState, GradientState


class AcceleratedScheduler:
    def __init__(self, scheduler, optimizers, step_with_optimizer: bool = True, split_batches: bool = False):
        self.scheduler = scheduler
        self.optimizers = list(optimizers) if isinstance(optimizers, (list, tuple)) else [optimizers]
        self.split_batches = split_batches
        self.step_with_optimizer = step_with_optimizer
        self.gradient_state = GradientState()

    def step(self, *args, **kwargs):
        if not self.step_with_optimizer:
            self.scheduler.step(*args, **kwargs)
            return
        if not self.gradient_state.sync_gradients:
            if self.gradient_state.adjust_scheduler:
                self.scheduler._step_count += 1  # keep warm-up lengths honest under accumulation
            return
        if any(opt.step_was_skipped for opt in self.optimizers):
            return  # fp16 overflow: the optimizer did nothing, neither do we
        repeats = 1 if self.split_batches else AcceleratorState().num_processes
        total = getattr(self.scheduler, "total_steps", None)
        for _ in range(repeats):
            if total is None or self.scheduler._step_count <= total:  # OneCycle-style schedulers must not overrun
                self.scheduler.step(*args, **kwargs)

    def get_last_lr(self):
        return self.scheduler.get_last_lr()

    def state_dict(self):
        return self.scheduler.state_dict()

    def load_state_dict(self, state_dict):
        self.scheduler.load_state_dict(state_dict)

    def get_lr(self):
        return self.scheduler.get_lr()

    def print_lr(self, *args, **kwargs):
        return self.scheduler.print_lr(*args, **kwargs)
