"""Deferred gradient multiplier: the clip coefficient (and 1/loss-scale) that ``clip_grad_norm_`` /
``GradScaler.unscale_`` computed stays a 3-float device tensor until somebody needs the scaled gradients.

Reference flow (``src/accelerate/accelerator.py:2946-3007`` -> ``torch.nn.utils.clip_grad_norm_``): the clip
ends with ``_foreach_mul_(grads, coef)`` — a read+write pass over every gradient (2.84 GB per GPT-2-medium
step) whose only consumer, in the usual loop, is the optimizer step right behind it.  SURVEY.md §8f row 1 folds
that multiply into the optimizer kernel (``b200dp_multi_adam`` reads ``g * coef`` on the fly).  To do that
WITHOUT changing what a user can observe, the multiplier is *pending*, and is applied in place
(``b200dp_scale_flat`` / ``b200dp_multi_scale``) the moment anything reads or writes a ``Parameter.grad``
from Python before the fused optimizer step consumed it:

* ``AcceleratedOptimizer.step`` with the fused Adam kernel  -> consumes the multiplier (no extra pass);
* any other optimizer, a user printing / editing ``p.grad``, a second ``clip_grad_norm_``, ``zero_grad`` ->
  the first ``.grad`` access applies it, then proceeds: they see exactly the reference's values;
* the next forward of the wrapped model (gradients kept for accumulation) -> applied before new gradients
  are added.

Observation works through a property installed on ``torch.nn.Parameter`` (``grad`` is a plain C getset
descriptor of ``TensorBase``; the property forwards to it).  It is installed only when the first multiplier is
deferred and costs one list-emptiness test per access.  ``B200DP_DEFER_CLIP=0`` turns the whole mechanism off
(the multiply pass then runs inside ``clip_grad_norm_`` as in round 1).  The library's own code reads
gradients through :func:`raw_grad` so that bookkeeping never triggers the pass.
"""

from __future__ import annotations

import os
import threading

import torch

_DESC = torch._C.TensorBase.grad
raw_grad = _DESC.__get__  # raw_grad(p): p.grad without observing it
raw_set_grad = _DESC.__set__

_pending: list = []  # PendingScale objects neither consumed nor applied yet
assign_epoch = 0  # bumped by every Python-level assignment to a Parameter's .grad (once the observer is installed)
_lock = threading.RLock()
_installed = False


def enabled() -> bool:
    return os.environ.get("B200DP_DEFER_CLIP", "1") != "0"


class PendingScale:
    """``out`` = device tensor [total_norm, multiplier, found_inf]; ``apply()`` multiplies the gradients it was
    computed for by ``out[1]`` in place (stream-ordered, exits at once on the device when the multiplier is 1)."""

    __slots__ = ("out", "_apply", "owner", "param_ids", "done")

    def __init__(self, out: torch.Tensor, apply, owner=None, param_ids=None):
        self.out, self._apply, self.owner, self.done = out, apply, owner, False
        self.param_ids = param_ids  # ids of the parameters whose gradients wait for the multiplier (None: any)

    @property
    def mult_ptr(self) -> int:
        return self.out.data_ptr() + 4

    def apply(self):
        """Materialise: run the multiply pass now."""
        with _lock:
            if self.done:
                return
            self.done = True
            if self in _pending:
                _pending.remove(self)
        self._apply()

    def consume(self):
        """The fused optimizer kernel applies the multiplier on the fly: nothing else to do."""
        with _lock:
            if self.done:
                return False
            self.done = True
            if self in _pending:
                _pending.remove(self)
        return True

    def discard(self):
        with _lock:
            self.done = True
            if self in _pending:
                _pending.remove(self)


def defer(out: torch.Tensor, apply, owner=None, param_ids=None) -> PendingScale:
    install()
    p = PendingScale(out, apply, owner, param_ids)
    with _lock:
        _pending.append(p)
    return p


def take_for(owner):
    """The pending multiplier registered for ``owner`` (compared with ==), consumed; None if there is none."""
    with _lock:
        for p in _pending:
            if p.owner == owner:
                break
        else:
            return None
    return p if p.consume() else None


def flush_all():
    """Apply every pending multiplier."""
    while _pending:
        with _lock:
            if not _pending:
                break
            p = _pending[0]
        p.apply()


def flush_for(param):
    """Apply the pending multipliers that cover ``param`` (called on the first observation of its ``.grad``)."""
    pid = id(param)
    for p in list(_pending):
        if p.param_ids is None or pid in p.param_ids:
            p.apply()


def flush_params(param_ids):
    """Apply the pending multipliers that cover any of ``param_ids``."""
    for p in list(_pending):
        if p.param_ids is None or not p.param_ids.isdisjoint(param_ids):
            p.apply()


def has_pending() -> bool:
    return bool(_pending)


def _get(self):
    if _pending:
        flush_for(self)
    return raw_grad(self)


def _set(self, value):
    global assign_epoch
    if _pending:
        flush_for(self)
    assign_epoch += 1
    raw_set_grad(self, value)


def _del(self):
    raise RuntimeError("cannot delete a tensor's .grad; set it to None")


def installed() -> bool:
    return _installed


def install():
    """Route Python-level ``Parameter.grad`` accesses through the observer (idempotent)."""
    global _installed
    if _installed:
        return
    with _lock:
        if _installed:
            return
        torch.nn.Parameter.grad = property(_get, _set, _del, doc=_DESC.__doc__)
        _installed = True


def uninstall():
    """Tests only: restore the plain descriptor."""
    global _installed
    with _lock:
        flush_all()
        if _installed:
            try:
                del torch.nn.Parameter.grad
            except AttributeError:
                pass
            _installed = False
