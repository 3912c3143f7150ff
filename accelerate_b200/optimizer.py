"""AcceleratedOptimizer: gates ``step`` / ``zero_grad`` on ``GradientState.sync_gradients`` and
reports skipped steps under fp16 loss scaling (reference: ``src/accelerate/optimizer.py:38-213``).

The fused reducer feeds this class in two ways: the gradients it steps on are views into the fusion
buffer, already averaged (and, after ``clip_grad_norm_`` / ``unscale_``, already scaled) by the
device kernels; and the found-inf flag the ``GradScaler`` consults comes from the allreduce
kernel's statistics rather than from a separate pass over the gradients.
"""

from __future__ import annotations

import inspect

import torch

from .state import AcceleratorState, GradientState


def _to_device(obj, device):
    if isinstance(obj, torch.Tensor):
        return obj.to(device)
    if isinstance(obj, dict):
        return type(obj)({k: _to_device(v, device) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device) for v in obj)
    return obj


class AcceleratedOptimizer(torch.optim.Optimizer):
    def __init__(self, optimizer, device_placement: bool = True, scaler=None):
        self.optimizer = optimizer
        self.scaler = scaler
        self.accelerator_state = AcceleratorState()
        self.gradient_state = GradientState()
        self.device_placement = device_placement
        self._is_overflow = False
        self._stepped = False  # set by the instrumented inner step
        if scaler is not None:
            self._arm_step_probe()
        if device_placement:
            self.optimizer.load_state_dict(_to_device(self.optimizer.state_dict(), self.accelerator_state.device))

    # the GradScaler calls optimizer.step only when no inf/NaN was found: observing that call is
    # how a skipped step is detected (optimizer.py:163-177,208-213)
    def _arm_step_probe(self):
        self._inner_step = self.optimizer.step

        def probed(*args, **kwargs):
            self._stepped = True
            return self._inner_step(*args, **kwargs)

        self._probed_step = probed

    # ---- torch.optim.Optimizer surface, forwarded ----
    @property
    def state(self):
        return self.optimizer.state

    @state.setter
    def state(self, value):
        self.optimizer.state = value

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    @param_groups.setter
    def param_groups(self, value):
        self.optimizer.param_groups = value

    @property
    def defaults(self):
        return self.optimizer.defaults

    @defaults.setter
    def defaults(self, value):
        self.optimizer.defaults = value

    def add_param_group(self, param_group):
        self.optimizer.add_param_group(param_group)

    def load_state_dict(self, state_dict):
        self.optimizer.load_state_dict(state_dict)

    def state_dict(self):
        return self.optimizer.state_dict()

    def train(self):
        if callable(getattr(self.optimizer, "train", None)):
            self.optimizer.train()

    def eval(self):
        if callable(getattr(self.optimizer, "eval", None)):
            self.optimizer.eval()

    # ---- gated operations ----
    def zero_grad(self, set_to_none=None):
        if not self.gradient_state.sync_gradients:
            return  # still accumulating
        if "set_to_none" in inspect.signature(self.optimizer.zero_grad).parameters:
            self.optimizer.zero_grad(set_to_none=True if set_to_none is None else set_to_none)
        else:
            if set_to_none is not None:
                raise ValueError("`set_to_none` for Optimizer.zero_grad` is not supported by this optimizer.")
            self.optimizer.zero_grad()

    def step(self, closure=None):
        if not self.gradient_state.sync_gradients:
            return
        if self.scaler is None:
            self.optimizer.step(closure)
            return
        self._stepped = False
        self.optimizer.step = self._probed_step
        try:
            self.scaler.step(self.optimizer, closure)
            self.scaler.update()
        finally:
            self.optimizer.step = self._inner_step
        self._is_overflow = not self._stepped

    @property
    def step_was_skipped(self) -> bool:
        return self._is_overflow

    def _switch_parameters(self, parameters_map):
        for group in self.optimizer.param_groups:
            group["params"] = [parameters_map.get(p, p) for p in group["params"]]

    def __getstate__(self):
        drop = {"_inner_step", "_probed_step"}
        return {k: v for k, v in self.__dict__.items() if k not in drop}

    def __setstate__(self, state):
        self.__dict__.update(state)
        if self.scaler is not None:
            self._arm_step_probe()
