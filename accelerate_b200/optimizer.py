"""AcceleratedOptimizer: gates ``step`` / ``zero_grad`` on ``GradientState.sync_gradients`` and
reports skipped steps under fp16 loss scaling (reference: ``src/accelerate/optimizer.py:38-213``).

The fused reducer feeds this class in two ways: the gradients it steps on are views into the fusion
buffer, already averaged (and, after ``clip_grad_norm_`` / ``unscale_``, already scaled) by the
device kernels; and the found-inf flag the ``GradScaler`` consults comes from the allreduce
kernel's statistics rather than from a separate pass over the gradients.
"""

from __future__ import annotations

import inspect
import os

import torch

from . import lazy_grad
from .lazy_grad import raw_grad
from .state import AcceleratorState, GradientState


def _to_device(obj, device):
    if isinstance(obj, torch.Tensor):
        return obj.to(device)
    if isinstance(obj, dict):
        return type(obj)({k: _to_device(v, device) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device) for v in obj)
    return obj


def _has_step_hooks(optimizer) -> bool:
    """User-registered optimizer step hooks must fire around ``optimizer.step``: such optimizers keep torch's step."""
    import torch.optim.optimizer as O

    return bool(
        getattr(optimizer, "_optimizer_step_pre_hooks", None)
        or getattr(optimizer, "_optimizer_step_post_hooks", None)
        or getattr(O, "_global_optimizer_pre_hooks", None)
        or getattr(O, "_global_optimizer_post_hooks", None)
    )


class FusedAdamPlan:
    """Fused optimizer consumption; ``B200DP_FUSED_ADAM=0`` turns it off.

    When the wrapped optimizer is a plain ``torch.optim.Adam`` / ``AdamW`` over fp32 CUDA parameters, its
    ~10 foreach launches are replaced by ONE multi-tensor kernel (``b200dp_multi_adam``) that reads p, g, m, v
    and writes p, m, v.  Two layouts:

    * *buffer* (multi-GPU): the gradients are views of one reducer's fusion buffer; the moments live in two flat
      buffers with the same layout, the pointer table is built once.  The pending clip / unscale multiplier
      (``lazy_grad.py``, on by default) is applied on the fly, so the separate multiply pass over the gradients
      disappears; under a ``GradScaler`` the kernel also reads the found-inf flag on the device.
    * *loose* (one process, no reducer): the gradients are whatever tensors autograd left in ``param.grad``; their
      pointers are collected per step, the moments still live in two flat buffers.

    The optimizer object stays the user's: ``param_groups`` (schedulers), ``state`` (``step`` / ``exp_avg`` /
    ``exp_avg_sq``, now views of the flat buffers) and ``state_dict()`` keep their torch layout, and any step that
    is not eligible falls through to ``optimizer.step()`` on that very state.
    """

    def __init__(self, optimizer, ctx, device, reducer=None, group=None):
        self.optimizer, self.ctx, self.device, self.reducer, self.group = optimizer, ctx, device, reducer, group
        self.decoupled_default = isinstance(optimizer, torch.optim.AdamW)
        where = {}
        if reducer is not None:
            for b in reducer.buckets:
                if b.group is group:
                    for p, off, n in zip(b.params, b.offsets, b.numels):
                        where[id(p)] = (off, n)
            total = group.total_elems
        else:
            total = 0
            for pg in optimizer.param_groups:
                for p in pg["params"]:
                    if p.requires_grad and id(p) not in where:
                        where[id(p)] = (total, p.numel())
                        total += (p.numel() + 3) & ~3  # keep every moment span 16-byte aligned
        self.m_flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.v_flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.tables = []
        self.views = {}  # id(param) -> (exp_avg view, exp_avg_sq view, step view) installed in optimizer.state
        self.step_flats = []  # one CPU float32 vector per parameter group: state[p]["step"] are 0-dim views of it, so
        #                       the per-step `step += 1` is ONE host op instead of a foreach over ~300 scalars
        self.signature = self._signature(optimizer)
        for pg in optimizer.param_groups:
            params = [p for p in pg["params"] if p.requires_grad]
            ptrs, numels, offs, gptrs = [], [], [], []
            step_flat = torch.zeros(len(params), dtype=torch.float32)
            self.step_flats.append(step_flat)
            for k, p in enumerate(params):
                off, n = where[id(p)]
                self.views[id(p)] = (self.m_flat[off : off + n].view(p.shape), self.v_flat[off : off + n].view(p.shape), step_flat[k])
                ptrs.append(p.data_ptr())
                numels.append(n)
                offs.append(off)
                if reducer is not None:
                    gptrs.append(group.flat.data_ptr() + 4 * off)
            self.tables.append((params, ctx.make_adam_table(ptrs, numels, offs, gptrs if reducer is not None else None)))
        self._adopt_state()

    @staticmethod
    def _signature(optimizer):
        return tuple((len(pg["params"]), id(pg["params"][0]) if pg["params"] else 0) for pg in optimizer.param_groups)

    def _adopt_state(self):
        """Make ``optimizer.state[p]['exp_avg'/'exp_avg_sq']`` the views of the flat moment buffers.  Values that
        are already there (a resumed run, or ``load_state_dict`` replacing the tensors) are copied in first."""
        state, views = self.optimizer.state, self.views
        for params, _ in self.tables:
            for p in params:
                st = state[p]
                m_view, v_view, s_view = views[id(p)]
                if st.get("exp_avg") is not m_view:
                    if "exp_avg" in st:
                        m_view.copy_(st["exp_avg"])
                        v_view.copy_(st["exp_avg_sq"])
                    st["exp_avg"], st["exp_avg_sq"] = m_view, v_view
                if st.get("step") is not s_view:
                    if "step" in st:
                        s_view.fill_(float(st["step"]))
                    st["step"] = s_view

    @staticmethod
    def try_build(optimizer):
        if os.environ.get("B200DP_FUSED_ADAM", "1") == "0" or type(optimizer) not in (torch.optim.Adam, torch.optim.AdamW):
            return None
        from .reducer import live_reducers

        wanted = [p for pg in optimizer.param_groups for p in pg["params"] if p.requires_grad]
        if not wanted or any(p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() for p in wanted):
            return None
        if len({p.device for p in wanted}) != 1 or len({id(p) for p in wanted}) != len(wanted):
            return None
        for pg in optimizer.param_groups:
            if pg.get("amsgrad") or pg.get("maximize") or pg.get("capturable") or pg.get("differentiable") or pg.get("fused"):
                return None
            if isinstance(pg["lr"], torch.Tensor) or any(isinstance(b, torch.Tensor) for b in pg["betas"]):
                return None
        ids = {id(p) for p in wanted}
        touched = False
        for red in live_reducers():
            if red.is_cuda and ids <= red._managed_ids and torch.float32 in red.groups:
                return FusedAdamPlan(optimizer, red.rt.ctx, red.device, red, red.groups[torch.float32])
            touched = touched or bool(ids & red._managed_ids)
        if touched:
            return None  # parameters split over reducers / partly managed: leave it to torch
        from .runtime import SymmRuntime

        rt = SymmRuntime._instance
        if rt is None or rt.device != wanted[0].device:
            return None
        return FusedAdamPlan(optimizer, rt.ctx, rt.device)

    def _loose_grads(self, params):
        """Gradient pointers for a loose-layout step, or None when this step is not eligible."""
        out = []
        for p in params:
            g = raw_grad(p)
            if g is None or g.dtype != torch.float32 or g.device != p.device or g.is_sparse or not g.is_contiguous():
                return None
            out.append(g.data_ptr())
        return out

    def step(self, found_inf: torch.Tensor | None = None) -> bool:
        """Run the fused update; False = not applicable right now (caller falls back to optimizer.step()).

        ``found_inf`` (GradScaler runs): 1-element float32 device tensor; the kernels return at once when it is
        non-zero, so an overflowed step is skipped without the host ever reading the flag."""
        red = self.reducer
        pending = None
        if red is not None:
            if not red.grads_in_buffer():
                return False
            grads_per_group = [None] * len(self.tables)
        else:
            grads_per_group = [self._loose_grads(params) for params, _ in self.tables]
            if any(g is None for g in grads_per_group):
                return False
        self._adopt_state()  # cheap identity checks; heals optimizer.load_state_dict()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if red is not None:
            pending = red.take_pending_multiplier()  # deferred clip / unscale: applied on the fly by the kernel
        elif lazy_grad.has_pending():
            pending = lazy_grad.take_for(frozenset(ptr for grads in grads_per_group for ptr in grads))
            pending = pending.out if pending is not None else None
        if lazy_grad.has_pending():  # a multiplier computed for a different selection of these gradients: materialise it
            lazy_grad.flush_params({id(p) for params, _ in self.tables for p in params})
        mult_ptr = pending.data_ptr() + 4 if pending is not None else None
        # without a GradScaler torch applies the update even to non-finite gradients: keep that (found_ptr None)
        found_ptr = found_inf.data_ptr() if found_inf is not None else None
        self.last_steps = []
        for pg, (params, table), grads, step_flat in zip(self.optimizer.param_groups, self.tables, grads_per_group, self.step_flats):
            if not params:
                continue
            step_flat.add_(1.0)
            self.last_steps.append(step_flat)
            t = float(step_flat[0])
            lr, (beta1, beta2), eps, wd = float(pg["lr"]), pg["betas"], float(pg["eps"]), float(pg["weight_decay"])
            decoupled = pg.get("decoupled_weight_decay", self.decoupled_default)
            decay_mul = (1 - lr * wd) if (decoupled and wd != 0) else 1.0
            l2 = wd if (not decoupled and wd != 0) else 0.0
            bc1 = 1 - beta1**t
            bc2_sqrt = (1 - beta2**t) ** 0.5
            self.ctx.multi_adam(table, self.m_flat.data_ptr(), self.v_flat.data_ptr(), decay_mul, l2, beta1, beta2,
                                -(lr / bc1), bc2_sqrt, eps, mult_ptr, found_ptr, stream, grads=grads)  # fmt: skip
        self._keepalive = (pending, found_inf)  # the kernels read them asynchronously
        return True


class AcceleratedOptimizer(torch.optim.Optimizer):
    def __init__(self, optimizer, device_placement: bool = True, scaler=None):
        self.optimizer = optimizer
        self.scaler = scaler
        self.accelerator_state = AcceleratorState()
        self.gradient_state = GradientState()
        self.device_placement = device_placement
        self._is_overflow = False
        self._stepped = False  # set by the instrumented inner step
        self._fused_plan = None  # built lazily at the first step (parameters must be wrapped by then)
        self._fused_plan_tried = False
        self._skip_pending = None  # (event, step counters) of a fused step under a GradScaler, not yet settled
        self._skip_host = None
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
        self._settle_skip()  # `step` counters must be final
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
            if closure is None and self._fused_step():
                return
            self.optimizer.step(closure)
            return
        self._settle_skip()
        if closure is None and self._fused_scaled_step():
            return
        self._stepped = False
        self.optimizer.step = self._probed_step
        try:
            self.scaler.step(self.optimizer, closure)
            self.scaler.update()
        finally:
            self.optimizer.step = self._inner_step
        self._is_overflow = not self._stepped

    def _plan(self):
        if not self._fused_plan_tried:
            self._fused_plan_tried = True
            if self.accelerator_state.device.type == "cuda":
                self._fused_plan = FusedAdamPlan.try_build(self.optimizer)
        plan = self._fused_plan
        if plan is not None and plan.signature != FusedAdamPlan._signature(self.optimizer):
            self._fused_plan = plan = FusedAdamPlan.try_build(self.optimizer)  # add_param_group() happened
        if plan is None or _has_step_hooks(self.optimizer):
            return None
        return plan

    def _fused_step(self) -> bool:
        plan = self._plan()
        if plan is None or not plan.step():
            return False
        # torch's LRScheduler patches optimizer.step to learn that a step happened before scheduler.step()
        self.optimizer._opt_called = True
        return True

    def _fused_scaled_step(self) -> bool:
        """fp16 (reference optimizer.py:163-177: ``scaler.step(optimizer)`` + ``scaler.update()``): the unscale comes
        from the allreduce statistics, the Adam kernel applies ``1/scale`` (and a pending clip coefficient) on the
        fly and returns at once when the device-side found-inf flag is set, and ``scaler.update()`` adjusts the scale
        on the device.  torch's ``scaler.step`` reads the flag on the host (``.item()``) to decide whether to call
        ``optimizer.step``; here nothing waits for the GPU.  Whether the step was skipped is settled lazily — when
        ``step_was_skipped`` is read (schedulers do) or at the next step — from a pinned copy of the flag."""
        from torch.amp.grad_scaler import OptState

        scaler = self.scaler
        plan = self._plan() if scaler.is_enabled() else None
        if plan is None:
            return False
        state = scaler._per_optimizer_states[id(self.optimizer)]
        if state["stage"] is OptState.STEPPED:
            raise RuntimeError("step() has already been called since the last update().")
        if state["stage"] is OptState.READY:
            scaler.unscale_(self.optimizer)  # fused: statistics -> found_inf; the multiply stays pending
        found = list(state["found_inf_per_device"].values())
        if len(found) != 1 or found[0].device != plan.device:
            return False  # scaler.step() below finds the gradients already unscaled and takes torch's route
        found_inf = found[0]
        if not plan.step(found_inf=found_inf):
            return False
        state["stage"] = OptState.STEPPED
        # async copy of the flag; the event completes long before anybody asks
        if self._skip_host is None:
            self._skip_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self._skip_host.copy_(found_inf.reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._skip_pending = (ev, plan.last_steps)
        scaler.update()
        self.optimizer._opt_called = True
        return True

    def _settle_skip(self):
        """Resolve the previous fused step: if the device skipped it, take the optimistic `step` increment back."""
        pending, self._skip_pending = self._skip_pending, None
        if pending is None:
            return
        ev, steps = pending
        ev.synchronize()
        skipped = bool(self._skip_host.item() != 0.0)
        if skipped and steps:
            torch._foreach_sub_(steps, 1)
        self._is_overflow = skipped

    @property
    def step_was_skipped(self) -> bool:
        self._settle_skip()
        return self._is_overflow

    def _switch_parameters(self, parameters_map):
        for group in self.optimizer.param_groups:
            group["params"] = [parameters_map.get(p, p) for p in group["params"]]

    def __getstate__(self):
        self._settle_skip()
        drop = {"_inner_step", "_probed_step", "_fused_plan", "_skip_pending", "_skip_host"}
        return {k: v for k, v in self.__dict__.items() if k not in drop}

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._skip_pending = self._skip_host = None
        self._fused_plan = None
        if self.scaler is not None:
            self._arm_step_probe()
