"""AcceleratedScheduler (reference: ``src/accelerate/scheduler.py:25-98``): an LR scheduler that only
advances when the optimizer really stepped, and advances ``num_processes`` times per step unless the
dataloader splits one global batch across ranks."""

from __future__ import annotations

from .state import AcceleratorState, GradientState


class AcceleratedScheduler:
    # everything that is not about WHEN to step (get_last_lr, get_lr, state_dict, load_state_dict, print_lr,
    # last_epoch, base_lrs, ...) is the wrapped scheduler's own business: see __getattr__

    def __init__(self, scheduler, optimizers, step_with_optimizer: bool = True, split_batches: bool = False):
        self.scheduler = scheduler
        self.optimizers = list(optimizers) if isinstance(optimizers, (list, tuple)) else [optimizers]
        self.step_with_optimizer, self.split_batches = step_with_optimizer, split_batches
        self.gradient_state = GradientState()

    def __getattr__(self, name):
        if name in ("scheduler", "__setstate__", "__getstate__") or name.startswith("__"):
            raise AttributeError(name)  # (unpickling probes these before __init__ ran)
        return getattr(self.scheduler, name)

    def _times_to_advance(self) -> int:
        """0 = hold still this call."""
        state = self.gradient_state
        if not state.sync_gradients:
            if state.adjust_scheduler:
                self.scheduler._step_count += 1  # keep warm-up lengths honest under accumulation
            return 0
        if any(opt.step_was_skipped for opt in self.optimizers):
            return 0  # fp16 overflow: the optimizer did nothing, neither do we
        # one global batch per step when batches are split; otherwise each process consumed its own batch
        return 1 if self.split_batches else AcceleratorState().num_processes

    def step(self, *args, **kwargs):
        inner = self.scheduler
        if not self.step_with_optimizer:
            inner.step(*args, **kwargs)
            return
        # OneCycle-style schedulers must not overrun their length when each step advances num_processes times; with
        # split batches the reference steps unconditionally (scheduler.py:70-82)
        budget = None if self.split_batches else getattr(inner, "total_steps", None)
        for _ in range(self._times_to_advance()):
            if budget is None or inner._step_count <= budget:
                inner.step(*args, **kwargs)
