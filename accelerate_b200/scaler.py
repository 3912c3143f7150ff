"""GradScaler whose unscale / found-inf step consumes the allreduce kernel's statistics.

Reference flow (fp16 only; ``src/accelerate/accelerator.py:2911-2944``, ``optimizer.py:163-177``):
``scaler.unscale_(opt)`` runs ``_amp_foreach_non_finite_check_and_unscale_`` over every gradient
(read + write of all gradients) to multiply by ``1/scale`` and OR-reduce a found-inf flag.  Here the
non-finite check already happened inside the fused allreduce (on the reduced values, which is what
the reference checks too), so unscaling is one multiply pass over the fusion buffer and found-inf is
read from the statistics records.  Everything else (``step`` skipping, ``update`` growth/backoff)
is torch's own ``GradScaler`` logic, untouched.
"""

from __future__ import annotations

import types

import torch

from . import lazy_grad
from .lazy_grad import raw_grad
from .reducer import live_reducers


class FusedGradScaler(torch.amp.GradScaler):
    def _unscale_grads_(self, optimizer, inv_scale, found_inf, allow_fp16):
        params = [p for group in optimizer.param_groups for p in group["params"] if raw_grad(p) is not None]
        ids = {id(p) for p in params}
        fused = []
        for red in live_reducers():
            red.finish_open_backward()
            if red.is_cuda and red.stats_fresh() and not red.unscaled and red._managed_ids <= ids and red.grads_in_buffer():
                fused.append(red)
        result = {}
        covered = set()
        for red in fused:
            # 1/scale becomes a pending multiplier (the fused Adam step applies it on the fly; any other reader of a
            # gradient triggers the pass first), found_inf comes from the allreduce statistics
            out = red.fused_clip_(-1.0, 2.0, scale=self._scale, defer=lazy_grad.enabled())
            found_inf.copy_(torch.maximum(found_inf, out[2].to(found_inf.device)))
            covered |= red._managed_ids
            result[found_inf.device] = found_inf
        rest = [p for p in params if id(p) not in covered]
        if rest:
            shim = types.SimpleNamespace(param_groups=[{"params": rest}])
            for dev, flag in super()._unscale_grads_(shim, inv_scale, found_inf, allow_fp16).items():
                result[dev] = flag
        if not result:
            result[found_inf.device] = found_inf
        return result
