"""``Accelerator``: the user-facing object, API-compatible with the reference's
``src/accelerate/accelerator.py`` for the data-parallel training loop
(``prepare`` :1414, ``backward`` :2818, ``accumulate`` :1254, ``no_sync`` :1132,
``trigger_sync_in_backward`` :1181, ``unscale_gradients`` :2911, ``clip_grad_norm_`` :2946,
``clip_grad_value_`` :3009, ``gather`` :3036, ``gather_for_metrics`` :3068, ``reduce`` :3141,
``pad_across_processes`` :3177, ``unwrap_model`` :3212, ``set_trigger/check_trigger`` :2852-2909,
``autocast`` :4172), so loops written for the reference run unmodified.

The difference is underneath: ``prepare_model`` wraps with :class:`~accelerate_b200.reducer.B200DataParallel`
(hand-written P2P allreduce kernels, gradients as views of a symmetric fusion buffer) instead of
``torch.nn.parallel.DistributedDataParallel``; ``clip_grad_norm_`` / ``unscale_gradients`` consume the
statistics the allreduce kernel already produced; ``gather`` / ``reduce`` / ``pad_across_processes`` run
on the same P2P kernels.  Other parallelism backends of the reference (FSDP, DeepSpeed, Megatron,
TP/CP/SP, XLA, FP8) are out of scope and raise ``NotImplementedError`` when requested.
"""

from __future__ import annotations

import contextlib
import functools
from collections import OrderedDict
import math
import os
from contextlib import contextmanager
from types import MethodType

import torch

from . import lazy_grad
from .data_loader import prepare_data_loader, skip_first_batches
from .lazy_grad import raw_grad
from .optimizer import AcceleratedOptimizer
from .reducer import B200DataParallel, live_reducers
from .scheduler import AcceleratedScheduler
from .state import AcceleratorState, GradientState, PartialState
from .utils.dataclasses import (
    AutocastKwargs,
    DataLoaderConfiguration,
    DistributedDataParallelKwargs,
    DistributedType,
    GradientAccumulationPlugin,
    GradScalerKwargs,
    InitProcessGroupKwargs,
    KwargsHandler,
    PrecisionType,
    ProjectConfiguration,
)
from .utils.operations import (
    convert_outputs_to_fp32,
    gather,
    gather_object,
    pad_across_processes,
    recursively_apply,
    reduce,
)
from .utils.other import extract_model_from_parallel, get_mixed_precision_context_manager

_UNSUPPORTED_PLUGINS = (
    "deepspeed_plugin", "deepspeed_plugins", "fsdp_plugin", "megatron_lm_plugin", "parallelism_config",
    "torch_tp_plugin", "dynamo_plugin", "dynamo_backend",
)  # fmt: skip


class Accelerator:
    def __init__(self, device_placement: bool = True, split_batches: bool = None, mixed_precision=None,
                 gradient_accumulation_steps: int = 1, cpu: bool = False, dataloader_config=None, log_with=None,
                 project_dir=None, project_config=None, gradient_accumulation_plugin=None,
                 step_scheduler_with_optimizer: bool = True, kwargs_handlers=None, rng_types=None, **unsupported):  # fmt: skip
        for key, value in unsupported.items():
            if key in _UNSUPPORTED_PLUGINS and value is not None:
                raise NotImplementedError(f"`{key}` belongs to a parallelism backend this data-parallel build does not provide")
            if key not in _UNSUPPORTED_PLUGINS:
                raise TypeError(f"Accelerator.__init__() got an unexpected keyword argument '{key}'")
        if log_with is not None:
            raise NotImplementedError("experiment trackers are out of scope for this build")
        self.project_configuration = project_config or ProjectConfiguration(project_dir=project_dir)
        if project_dir is not None and self.project_configuration.project_dir is None:
            self.project_configuration.set_directories(project_dir)
        if mixed_precision is not None:
            mixed_precision = str(mixed_precision)
            if mixed_precision not in PrecisionType:
                raise ValueError(f"Unknown mixed_precision mode: {mixed_precision}. Choose between {PrecisionType.list()}")

        # ---- kwargs handlers: at most one of each kind (accelerator.py:416-452) ----
        self.ddp_handler = self.scaler_handler = self.init_handler = self.autocast_handler = None
        slots = {
            DistributedDataParallelKwargs: "ddp_handler", GradScalerKwargs: "scaler_handler",
            InitProcessGroupKwargs: "init_handler", AutocastKwargs: "autocast_handler",
        }  # fmt: skip
        for handler in kwargs_handlers or []:
            if not isinstance(handler, KwargsHandler):
                raise TypeError(f"Unsupported kwargs handler passed: {handler}, must be one that inherits `accelerate.utils.KwargsHandler`.")
            attr = next((name for cls, name in slots.items() if isinstance(handler, cls)), None)
            if attr is None:
                raise NotImplementedError(f"kwargs handler {type(handler).__name__} is not supported by this build")
            if getattr(self, attr) is not None:
                raise ValueError(f"You can only pass one {type(handler).__name__} in `kwargs_handlers`.")
            setattr(self, attr, handler)

        init_kwargs = self.init_handler.to_kwargs() if self.init_handler else {}
        self.state = AcceleratorState(mixed_precision=mixed_precision, cpu=cpu, _from_accelerator=True, **init_kwargs)
        self.device_placement = device_placement

        if dataloader_config is None:
            dataloader_config = DataLoaderConfiguration()
        if split_batches is not None:
            dataloader_config.split_batches = split_batches
        self.dataloader_config = dataloader_config
        self.step_scheduler_with_optimizer = step_scheduler_with_optimizer

        # ---- gradient accumulation (accelerator.py:539-553) ----
        if gradient_accumulation_plugin is not None:
            if gradient_accumulation_steps != 1:
                raise ValueError("You can only pass one of `gradient_accumulation_steps` and `gradient_accumulation_plugin`. Please only pass in the created `GradientAccumulationPlugin` object.")
        else:
            steps = int(os.environ.get("ACCELERATE_GRADIENT_ACCUMULATION_STEPS", gradient_accumulation_steps))
            gradient_accumulation_plugin = GradientAccumulationPlugin(num_steps=steps)
        self.gradient_state = GradientState(gradient_accumulation_plugin=gradient_accumulation_plugin)

        # ---- mixed precision: fp16 gets a (fused) GradScaler, bf16 none (accelerator.py:564-596) ----
        self.scaler = None
        self.native_amp = False
        if self.state.mixed_precision == "fp16" and self.device.type != "cpu":
            # (on a CPU device the reference silently trains in full precision: accelerator.py:564-567)
            self.native_amp = True
            if self.device.type != "cuda":
                raise ValueError(f"fp16 mixed precision requires a device in ('cuda',) (not {self.device.type!r}).")
            from .scaler import FusedGradScaler

            self.scaler = FusedGradScaler("cuda", **(self.scaler_handler.to_kwargs() if self.scaler_handler else {}))
        elif self.state.mixed_precision == "bf16":
            self.native_amp = True

        self.step = 0
        self._optimizers, self._models, self._schedulers, self._dataloaders = [], [], [], []
        self._wrappers = []  # B200DataParallel instances handed to the user (their hooks are removed by free_memory)
        self._save_model_state_pre_hook, self._load_model_state_pre_hook = OrderedDict(), OrderedDict()
        self._custom_objects = []
        self.flag_tensor = None
        self.rng_types = rng_types if rng_types is not None else ["generator"]
        self.last_clip_mode = None  # "fused" | "multi_tensor" | "torch" (introspection for tests/bench)

    # ------------------------------------------------------------------ simple properties
    @property
    def distributed_type(self):
        return self.state.distributed_type

    @property
    def num_processes(self):
        return self.state.num_processes

    @property
    def process_index(self):
        return self.state.process_index

    @property
    def local_process_index(self):
        return self.state.local_process_index

    @property
    def device(self):
        return self.state.device

    @property
    def use_distributed(self):
        return self.state.use_distributed

    @property
    def is_main_process(self):
        return self.state.is_main_process

    @property
    def is_local_main_process(self):
        return self.state.is_local_main_process

    @property
    def is_last_process(self):
        return self.process_index == self.num_processes - 1

    @property
    def mixed_precision(self):
        return self.state.mixed_precision

    @property
    def split_batches(self):
        return self.dataloader_config.split_batches

    @property
    def dispatch_batches(self):
        return self.dataloader_config.dispatch_batches

    @property
    def even_batches(self):
        return self.dataloader_config.even_batches

    @even_batches.setter
    def even_batches(self, value: bool):
        self.dataloader_config.even_batches = value

    @property
    def use_seedable_sampler(self):
        return self.dataloader_config.use_seedable_sampler

    @property
    def non_blocking(self):
        return self.dataloader_config.non_blocking

    @property
    def project_dir(self):
        return self.project_configuration.project_dir

    @property
    def logging_dir(self):
        return self.project_configuration.logging_dir

    @property
    def save_iteration(self):
        return self.project_configuration.iteration

    @property
    def use_stateful_dataloader(self):
        return bool(getattr(self.dataloader_config, "use_stateful_dataloader", False))

    @property
    def multi_device(self):
        """True when every process drives its own accelerator device (accelerator.py:657-668)."""
        return self.use_distributed and self.distributed_type == DistributedType.MULTI_GPU

    @property
    def sync_gradients(self):
        return self.gradient_state.sync_gradients

    @sync_gradients.setter
    def sync_gradients(self, value):
        self.gradient_state.sync_gradients = value

    @property
    def gradient_accumulation_steps(self):
        return self.gradient_state.num_steps

    @gradient_accumulation_steps.setter
    def gradient_accumulation_steps(self, value):
        self.gradient_state.plugin_kwargs.update({"num_steps": value})

    @property
    def optimizer_step_was_skipped(self):
        return any(opt.step_was_skipped for opt in self._optimizers)

    # ------------------------------------------------------------------ process control passthroughs
    def on_main_process(self, function=None):
        return self._process_decorator("on_main_process", function)

    def on_local_main_process(self, function=None):
        return self._process_decorator("on_local_main_process", function)

    def on_last_process(self, function=None):
        return self._process_decorator("on_last_process", function)

    def _process_decorator(self, name, function):
        if function is None:
            return functools.partial(self._process_decorator, name)
        return getattr(PartialState(), name)(function)

    def on_process(self, function=None, process_index=None):
        if function is None:
            return functools.partial(self.on_process, process_index=process_index)
        return PartialState().on_process(function, process_index)

    def on_local_process(self, function=None, local_process_index=None):
        if function is None:
            return functools.partial(self.on_local_process, local_process_index=local_process_index)
        return PartialState().on_local_process(function, local_process_index)

    @contextmanager
    def main_process_first(self):
        with self.state.main_process_first():
            yield

    @contextmanager
    def local_main_process_first(self):
        with self.state.local_main_process_first():
            yield

    @contextmanager
    def split_between_processes(self, inputs, apply_padding: bool = False):
        with PartialState().split_between_processes(inputs, apply_padding=apply_padding) as part:
            yield part

    def wait_for_everyone(self):
        PartialState().wait_for_everyone()

    def print(self, *args, **kwargs):
        self.state.print(*args, **kwargs)

    # ------------------------------------------------------------------ gradient synchronisation control
    @contextmanager
    def no_sync(self, model):
        """Skip the gradient allreduce for backward passes whose *forward* ran inside (accelerator.py:1132-1178)."""
        ctx = contextlib.nullcontext
        if self.use_distributed:
            ctx = getattr(model, "no_sync", ctx)
        with ctx():
            yield

    @staticmethod
    @contextmanager
    def trigger_sync_in_backward(model):
        """Force the next backward to reduce even though its forward ran under ``no_sync`` (accelerator.py:1181-1227)."""
        if not isinstance(model, B200DataParallel):
            yield
            return
        old_bwd, old_fwd = model.require_backward_grad_sync, model.require_forward_param_sync
        model.require_backward_grad_sync = True
        model.require_forward_param_sync = True
        model.reducer.prepare_for_backward([])
        try:
            yield
        finally:
            model.require_backward_grad_sync = old_bwd
            model.require_forward_param_sync = old_fwd

    def _do_sync(self):
        if self.gradient_state.sync_with_dataloader and self.gradient_state.end_of_dataloader:
            self.step = 0
            self.gradient_state._set_sync_gradients(True)
        else:
            self.step += 1
            self.gradient_state._set_sync_gradients((self.step % self.gradient_state.num_steps) == 0)

    @contextmanager
    def accumulate(self, *models):
        self._do_sync()
        allow_sync = self.sync_gradients or (self.use_distributed and self.gradient_state.sync_each_batch)
        with contextlib.ExitStack() as stack:
            for m in models:
                stack.enter_context(contextlib.nullcontext() if allow_sync else self.no_sync(m))
            yield

    @contextmanager
    def join_uneven_inputs(self, joinables, even_batches=None):
        """Train on uneven per-rank inputs (accelerator.py:1299-1372): inside the context a rank that runs out of
        batches shadows the others' collectives through torch's ``Join`` (the prepared model is a ``Joinable``).
        ``even_batches`` temporarily overrides the setting of the prepared map-style dataloaders.  Like the reference,
        only the one-device-per-process GPU configuration joins; elsewhere the context has no effect."""
        import warnings

        if self.multi_device:
            from torch.distributed.algorithms.join import Join

            from .data_loader import DataLoaderDispatcher

            overridden = []
            if even_batches is not None:
                iterable_seen = False
                for idx, dl in enumerate(self._dataloaders):
                    if isinstance(dl, DataLoaderDispatcher):
                        iterable_seen = True
                        continue
                    overridden.append((idx, dl.batch_sampler.even_batches))
                    dl.batch_sampler.even_batches = even_batches
                if iterable_seen:
                    warnings.warn(
                        "Overriding even_batches is only supported for map-style datasets, yet some dataloaders given were iterable"
                    )
            else:
                even_batches = self.even_batches
            try:
                with Join(joinables, enable=not even_batches, throw_on_early_termination=False):
                    yield
            finally:
                for idx, value in overridden:
                    self._dataloaders[idx].batch_sampler.even_batches = value
        else:
            if self.distributed_type != DistributedType.NO:
                warnings.warn(
                    "Joining uneven inputs is only supported for multi-GPU training, as a result `join_uneven_inputs` will have no effect."
                )
            with contextlib.nullcontext(joinables):
                yield

    # ------------------------------------------------------------------ prepare
    def prepare(self, *args, device_placement=None):
        if device_placement is None:
            device_placement = [None] * len(args)
        elif len(device_placement) != len(args):
            raise ValueError(f"`device_placement` should be a list with {len(args)} elements (the number of objects passed).")
        first = [self._prepare_one(o, first_pass=True, device_placement=d) for o, d in zip(args, device_placement)]
        result = tuple(self._prepare_one(o, device_placement=d) for o, d in zip(first, device_placement))
        for item in result:
            if any(item in bag for bag in (self._dataloaders, self._models, self._optimizers, self._schedulers)):
                item._is_accelerate_prepared = True
        return result if len(result) > 1 else result[0]

    def _prepare_one(self, obj, first_pass=False, device_placement=None):
        if first_pass:
            if isinstance(obj, torch.utils.data.DataLoader):
                return self.prepare_data_loader(obj, device_placement=device_placement)
            if isinstance(obj, torch.nn.Module):
                return self.prepare_model(obj, device_placement=device_placement)
            if isinstance(obj, torch.optim.Optimizer):
                return self.prepare_optimizer(obj, device_placement=device_placement)
        elif isinstance(obj, torch.optim.lr_scheduler.LRScheduler):
            return self.prepare_scheduler(obj)
        return obj

    def prepare_model(self, model: torch.nn.Module, device_placement: bool = None, evaluation_mode: bool = False):
        if device_placement is None:
            device_placement = self.device_placement
        if getattr(model, "_is_accelerate_prepared", False):
            if model not in self._models:
                self._models.append(model)
            return model
        if getattr(model, "hf_device_map", None) and len(set(model.hf_device_map.values())) > 1:
            raise ValueError("You can't train a model that has been loaded with `device_map='auto'` in any distributed mode.")
        self._models.append(model)

        if self.native_amp:  # autocast the forward and hand back fp32 outputs (accelerator.py:1818-1829)
            model._original_forward = model.forward
            ctx = get_mixed_precision_context_manager(self.native_amp, self.autocast_handler)
            inner = model.forward.__func__ if hasattr(model.forward, "__func__") else model.forward
            wrapped = ctx(inner)
            if hasattr(model.forward, "__func__"):
                model.forward = MethodType(wrapped, model)
                model.forward = MethodType(convert_outputs_to_fp32(model.forward.__func__), model)
            else:
                model.forward = convert_outputs_to_fp32(wrapped)

        if device_placement:
            model = model.to(self.device)
        if not evaluation_mode and self.use_distributed and any(p.requires_grad for p in model.parameters()):
            kwargs = self.ddp_handler.to_kwargs() if self.ddp_handler is not None else {}
            if self.distributed_type == DistributedType.MULTI_GPU:
                index = self.device.index
                wrapped_model = B200DataParallel(model, device_ids=[index], output_device=index, **kwargs)
            else:
                wrapped_model = B200DataParallel(model, **kwargs)
            if self.ddp_handler is not None:
                self.ddp_handler.register_comm_hook(wrapped_model)
            # like the reference (accelerator.py:1805,1892-1896), `_models` keeps the UNWRAPPED module: save_state /
            # load_state therefore read and write parameter names without the wrapper's "module." prefix, which is what
            # makes checkpoints interchangeable between the two libraries in distributed runs
            self._wrappers.append(wrapped_model)
            model = wrapped_model
        model._is_accelerate_prepared = True
        return model

    def prepare_data_loader(self, data_loader, device_placement=None, slice_fn_for_dispatch=None):
        if getattr(data_loader, "_is_accelerate_prepared", False):
            if data_loader not in self._dataloaders:
                self._dataloaders.append(data_loader)
            return data_loader
        if device_placement is None:
            device_placement = self.device_placement
        prepared = prepare_data_loader(
            data_loader, self.device, num_processes=self.num_processes, process_index=self.process_index,
            split_batches=self.split_batches, put_on_device=device_placement, rng_types=list(self.rng_types),
            dispatch_batches=self.dispatch_batches, even_batches=self.even_batches,
            slice_fn_for_dispatch=slice_fn_for_dispatch, use_seedable_sampler=self.use_seedable_sampler,
            data_seed=self.dataloader_config.data_seed, non_blocking=self.non_blocking,
            use_stateful_dataloader=self.use_stateful_dataloader,
        )  # fmt: skip
        self._dataloaders.append(prepared)
        return prepared

    def prepare_optimizer(self, optimizer, device_placement=None):
        if getattr(optimizer, "_is_accelerate_prepared", False):
            if optimizer not in self._optimizers:
                self._optimizers.append(optimizer)
            return optimizer
        if device_placement is None:
            device_placement = self.device_placement
        wrapped = AcceleratedOptimizer(optimizer, device_placement=device_placement, scaler=self.scaler)
        self._optimizers.append(wrapped)
        return wrapped

    def prepare_scheduler(self, scheduler):
        if getattr(scheduler, "_is_accelerate_prepared", False):
            if scheduler not in self._schedulers:
                self._schedulers.append(scheduler)
            return scheduler
        # the scheduler follows the optimizer(s) it was built on
        target = getattr(scheduler, "optimizer", None)
        owners = [opt for opt in self._optimizers if target is opt.optimizer] or list(self._optimizers)
        wrapped = AcceleratedScheduler(scheduler, owners, step_with_optimizer=self.step_scheduler_with_optimizer,
                                       split_batches=self.split_batches)  # fmt: skip
        self._schedulers.append(wrapped)
        return wrapped

    # ------------------------------------------------------------------ the step
    def backward(self, loss, **kwargs):
        """Scale for accumulation (and fp16), then backward; the reducer's hooks run the fused
        bucket allreduces while autograd is still producing later gradients (accelerator.py:2818-2850)."""
        loss = loss / self.gradient_accumulation_steps
        if self.scaler is not None:
            self.scaler.scale(loss).backward(**kwargs)
        else:
            loss.backward(**kwargs)

    def unscale_gradients(self, optimizer=None):
        """fp16 only: divide gradients by the loss scale and record found-inf (accelerator.py:2911-2944)."""
        if self.scaler is None:
            return
        optimizers = self._optimizers if optimizer is None else (optimizer if isinstance(optimizer, (list, tuple)) else [optimizer])
        for opt in optimizers:
            while isinstance(opt, AcceleratedOptimizer):
                opt = opt.optimizer
            self.scaler.unscale_(opt)

    def _fused_reducer_for(self, params):
        for red in live_reducers():
            red.finish_open_backward()
            if red.is_cuda and red.stats_fresh() and red.manages(params):
                return red
        return None

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        """Clip by global norm; returns the total norm (0-dim tensor).  On CUDA, when ``parameters`` are
        exactly a prepared model's parameters and this follows a sync backward, the norm comes from the
        statistics the allreduce kernels produced; the multiply by the coefficient stays *pending*
        (``lazy_grad.py``): the fused Adam step applies it on the fly, anything else that looks at a gradient
        first triggers the pass, so observable values are the reference's (accelerator.py:2946-3007)."""
        params = [parameters] if isinstance(parameters, torch.Tensor) else list(parameters)
        norm_type = float(norm_type)
        if self.device.type != "cuda":
            self.unscale_gradients()
            self.last_clip_mode = "torch"
            return torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type=norm_type)
        red = self._fused_reducer_for(params) if norm_type in (2.0, math.inf) else None
        defer = lazy_grad.enabled()
        if red is not None:
            scale = self.scaler._scale if (self.scaler is not None and self._scaler_pending(red)) else None
            if scale is not None and not red.unscaled:
                # unscale and clip in ONE pass; tell the scaler its work is done
                out = red.fused_clip_(float(max_norm), norm_type, scale=scale, defer=defer)
                self._mark_unscaled(red, out)
            else:
                self.unscale_gradients()
                out = red.fused_clip_(float(max_norm), norm_type, scale=self.scaler._scale if red.unscaled else None,
                                      defer=defer)
            self.last_clip_mode = "fused"
            return out[0]
        self.unscale_gradients()
        self.last_clip_mode = "multi_tensor"
        return self._multi_tensor_clip(params, float(max_norm), norm_type, defer=defer)

    def _scaler_pending(self, red) -> bool:
        """True when every optimizer holding this reducer's parameters still has scaled gradients."""
        from torch.amp.grad_scaler import OptState

        owners = self._owning_optimizers(red)
        if not owners:
            return False
        return all(self.scaler._per_optimizer_states[id(o)]["stage"] == OptState.READY for o in owners) and all(
            {id(p) for g in o.param_groups for p in g["params"] if raw_grad(p) is not None} <= red._managed_ids for o in owners
        )

    def _owning_optimizers(self, red):
        owners = []
        for opt in self._optimizers:
            inner = opt
            while isinstance(inner, AcceleratedOptimizer):
                inner = inner.optimizer
            ids = {id(p) for g in inner.param_groups for p in g["params"]}
            if ids & red._managed_ids:
                owners.append(inner)
        return owners

    def _mark_unscaled(self, red, out):
        from torch.amp.grad_scaler import OptState

        for inner in self._owning_optimizers(red):
            st = self.scaler._per_optimizer_states[id(inner)]
            st["found_inf_per_device"] = {self.device: out[2].clone()}
            st["stage"] = OptState.UNSCALED

    def _multi_tensor_clip(self, params, max_norm: float, norm_type: float, defer: bool = False):
        """Gradients outside a fusion buffer (world size 1, parameter subsets): two multi-tensor launches
        per dtype — statistics, then scale — instead of torch's foreach_norm/stack/norm/foreach_mul."""
        if norm_type not in (2.0, math.inf):
            return torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type=norm_type)
        from .runtime import SymmRuntime, native_dtype

        rt = SymmRuntime.get()
        st = torch.cuda.current_stream(self.device).cuda_stream
        if lazy_grad.has_pending():  # an earlier pending multiplier belongs to these values
            lazy_grad.flush_params({id(p) for p in params})
        by_dtype = {}
        for p in params:
            g = raw_grad(p)
            if g is None:
                continue
            if g.device != self.device or not g.is_contiguous() or native_dtype(g.dtype) not in (0, 1, 2):
                return torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type=norm_type)
            by_dtype.setdefault(g.dtype, []).append(g)
        out = torch.empty(3, dtype=torch.float32, device=self.device)
        if not by_dtype:
            return out[0].zero_()
        if not hasattr(rt, "_mt_slot_base"):
            rt._mt_slot_base = rt.reserve_slots(64)  # one reservation per process: the launches are stream-ordered
        slot = rt._mt_slot_base
        for dtype, grads in by_dtype.items():
            used = rt.ctx.multi_stats([g.data_ptr() for g in grads], [g.numel() for g in grads], native_dtype(dtype), slot, st)
            slot += used
            if slot - rt._mt_slot_base > 64:
                raise RuntimeError("too many gradient tensors for the multi-tensor clip")
        rt.ctx.finalize_norm(rt._mt_slot_base, slot - rt._mt_slot_base, False, norm_type, max_norm, None, out.data_ptr(), st)

        def apply():  # holds the gradient tensors alive until the multiplier is applied or consumed
            stream = torch.cuda.current_stream(self.device).cuda_stream
            for dtype, grads in by_dtype.items():
                rt.ctx.multi_scale([g.data_ptr() for g in grads], [g.numel() for g in grads], native_dtype(dtype),
                                   out.data_ptr() + 4, stream)  # fmt: skip

        if defer:
            lazy_grad.defer(out, apply, owner=frozenset(g.data_ptr() for grads in by_dtype.values() for g in grads),
                            param_ids=frozenset(id(p) for p in params))
        else:
            apply()
        return out[0]

    def clip_grad_value_(self, parameters, clip_value):
        params = [parameters] if isinstance(parameters, torch.Tensor) else list(parameters)
        self.unscale_gradients()
        if self.device.type == "cuda":
            for red in live_reducers():
                if red.is_cuda and red.manages(params):
                    red.clamp_(float(clip_value))
                    return
        torch.nn.utils.clip_grad_value_(params, clip_value)

    # ------------------------------------------------------------------ collectives
    def gather(self, tensor):
        return gather(tensor)

    def gather_for_metrics(self, input_data, use_gather_object: bool = False):
        """Gather, then drop the samples that ``even_batches`` duplicated at the end of the epoch (accelerator.py:3068-3139)."""
        try:
            recursively_apply(lambda x: x, input_data, error_on_other_type=True)
            all_tensors = True
        except TypeError:
            all_tensors = False
        use_gather_object = use_gather_object or not all_tensors
        data = gather_object(input_data) if use_gather_object else self.gather(input_data)
        try:
            if self.gradient_state.end_of_dataloader and self.gradient_state.remainder > 0:
                keep = self.gradient_state.remainder
                if use_gather_object:
                    return data[:keep]
                return recursively_apply(lambda t: t[:keep], data)
        except Exception:
            pass
        return data

    def reduce(self, tensor, reduction="sum", scale=1.0):
        return reduce(tensor, reduction, scale)

    def pad_across_processes(self, tensor, dim=0, pad_index=0, pad_first=False):
        return pad_across_processes(tensor, dim=dim, pad_index=pad_index, pad_first=pad_first)

    def set_trigger(self):
        self.flag_tensor = torch.tensor(1, device=self.device)

    def check_trigger(self) -> bool:
        if self.flag_tensor is None:
            self.flag_tensor = torch.tensor(0, device=self.device)
        total = self.reduce(self.flag_tensor)
        if total.item() >= 1:
            self.flag_tensor = torch.tensor(0, device=self.device)
            return True
        return False

    # ------------------------------------------------------------------ misc
    def unwrap_model(self, model, keep_fp32_wrapper: bool = True, keep_torch_compile: bool = True):
        return extract_model_from_parallel(model, keep_fp32_wrapper, keep_torch_compile)

    @contextmanager
    def autocast(self, autocast_handler: AutocastKwargs = None):
        handler = autocast_handler or self.autocast_handler
        ctx = get_mixed_precision_context_manager(self.native_amp, handler)
        ctx.__enter__()
        try:
            yield
        finally:
            ctx.__exit__(None, None, None)

    def skip_first_batches(self, dataloader, num_batches: int = 0):
        return skip_first_batches(dataloader, num_batches=num_batches)

    def free_memory(self, *objects):
        for m in self._wrappers:
            m.reducer.close()  # hooks, symmetric fusion/wire segments, statistics slots
        self._wrappers = []
        self._schedulers, self._optimizers, self._models, self._dataloaders = [], [], [], []
        self.step = 0
        import gc

        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        return [None for _ in objects]  # `a, b = accelerator.free_memory(a, b)` (a list, as in the reference)

    clear = free_memory

    def get_state_dict(self, model, unwrap: bool = True):
        if unwrap:
            model = self.unwrap_model(model)
        return model.state_dict()

    def save(self, obj, f, safe_serialization: bool = False):
        if self.is_main_process:
            torch.save(obj, f)

    # ------------------------------------------------------------------ checkpoint / resume
    def register_for_checkpointing(self, *objects):
        """Objects with ``state_dict`` / ``load_state_dict`` that save_state/load_state should carry along."""
        bad = [f"\t- Item at index {i}, `{type(o).__name__}`" for i, o in enumerate(objects)
               if not (hasattr(o, "state_dict") and hasattr(o, "load_state_dict"))]  # fmt: skip
        if bad:
            raise ValueError("All `objects` must include a `state_dict` and `load_state_dict` function to be stored. "
                             "The following inputs are invalid:\n" + "\n".join(bad))  # fmt: skip
        self._custom_objects.extend(objects)

    def register_save_state_pre_hook(self, hook):
        """``hook(models, weights, output_dir)`` runs inside save_state before anything is written
        (accelerator.py:3547-3577); entries removed from ``weights`` are not saved.  Returns a removable handle."""
        from torch.utils import hooks

        handle = hooks.RemovableHandle(self._save_model_state_pre_hook)
        self._save_model_state_pre_hook[handle.id] = hook
        return handle

    def register_load_state_pre_hook(self, hook):
        """``hook(models, input_dir)`` runs inside load_state before anything is read (accelerator.py:3714-3743);
        models removed from the list are not loaded.  Returns a removable handle."""
        from torch.utils import hooks

        handle = hooks.RemovableHandle(self._load_model_state_pre_hook)
        self._load_model_state_pre_hook[handle.id] = hook
        return handle

    def save_state(self, output_dir: str = None, safe_serialization: bool = True, **save_model_func_kwargs):
        """Write models, optimizers, schedulers, scaler, RNG states (every rank) and registered objects in the
        reference's on-disk layout (accelerator.py:3579-3712)."""
        from .checkpointing import rotate_checkpoints, save_accelerator_state, save_custom_state

        cfg = self.project_configuration
        if cfg.automatic_checkpoint_naming:
            output_dir = os.path.join(self.project_dir, "checkpoints")
        if output_dir is None:
            raise ValueError("`output_dir` is required unless `automatic_checkpoint_naming` is enabled")
        os.makedirs(output_dir, exist_ok=True)
        if cfg.automatic_checkpoint_naming:
            if self.is_main_process:
                rotate_checkpoints(output_dir, cfg.total_limit)
            output_dir = os.path.join(output_dir, f"checkpoint_{cfg.iteration}")
            if os.path.exists(output_dir):
                raise ValueError(f"Checkpoint directory {output_dir} ({cfg.iteration}) already exists. Please manually "
                                 "override `self.save_iteration` with what iteration to start with.")  # fmt: skip
            self.wait_for_everyone()
        os.makedirs(output_dir, exist_ok=True)
        weights = [self.get_state_dict(m, unwrap=False) for m in self._models]
        for hook in self._save_model_state_pre_hook.values():
            hook(self._models, weights, output_dir)
        save_accelerator_state(
            output_dir, weights, [o for o in self._optimizers], [s for s in self._schedulers], self._dataloaders,
            self.process_index, self.step, scaler=self.scaler, save_on_each_node=cfg.save_on_each_node,
            safe_serialization=safe_serialization, is_main_process=self.is_main_process,
            is_local_main_process=self.is_local_main_process,
        )  # fmt: skip
        writer = self.is_local_main_process if cfg.save_on_each_node else self.is_main_process
        for i, obj in enumerate(self._custom_objects):
            save_custom_state(obj, output_dir, i, is_writer=writer)
        cfg.iteration += 1
        self.wait_for_everyone()
        return output_dir

    def load_state(self, input_dir: str = None, load_kwargs: dict = None, **load_model_func_kwargs):
        """Restore what save_state wrote (also checkpoints written by the reference; accelerator.py:3745-3895)."""
        from .checkpointing import load_accelerator_state, load_custom_state

        if input_dir is not None:
            input_dir = os.path.expanduser(input_dir)
            if not os.path.isdir(input_dir):
                raise ValueError(f"Tried to find {input_dir} but folder does not exist")
        elif self.project_configuration.automatic_checkpoint_naming:
            folder = os.path.join(self.project_dir, "checkpoints")
            names = sorted(os.listdir(folder), key=lambda n: int(n.rsplit("_", 1)[-1]))
            input_dir = os.path.join(folder, names[-1])
        else:
            raise ValueError("No input_dir provided and automatic checkpoint naming is disabled.")
        models = list(self._models)
        for hook in self._load_model_state_pre_hook.values():
            hook(models, input_dir)
        map_location = load_model_func_kwargs.pop("map_location", None)
        if map_location is None:
            map_location = "on_device" if self.num_processes > 1 and self.device.type == "cuda" else "cpu"
        overrides = load_accelerator_state(
            input_dir, models, self._optimizers, self._schedulers, self._dataloaders, self.process_index,
            scaler=self.scaler, map_location=map_location, device=self.device, **load_model_func_kwargs,
        )  # fmt: skip
        if "step" in overrides:
            self.step = overrides["step"]
        for i, obj in enumerate(self._custom_objects):
            load_custom_state(obj, input_dir, i)
        self.wait_for_everyone()

    def save_model(self, model, save_directory, max_shard_size="10GB", safe_serialization: bool = True):
        """Unwrapped weights for re-loading outside a training run (accelerator.py:3433-3546): ``model.safetensors`` /
        ``pytorch_model.bin``, or — above ``max_shard_size`` — numbered shard files plus an ``*.index.json`` weight map,
        the layout ``transformers`` / ``load_checkpoint_in_model`` read.  Main process writes."""
        import json
        import re
        from pathlib import Path

        from huggingface_hub import split_torch_state_dict_into_shards

        from .checkpointing import MODEL_NAME, SAFE_MODEL_NAME, _write

        if os.path.isfile(save_directory):
            raise ValueError(f"Provided path ({save_directory}) should be a directory, not a file")
        if any(p.device == torch.device("meta") for p in model.parameters()):
            raise RuntimeError("You can't save the model since some parameters are on the meta device.")
        state_dict = self.get_state_dict(model)
        os.makedirs(save_directory, exist_ok=True)
        if safe_serialization:  # one name per storage (tied weights), as safetensors requires
            seen, unique = set(), {}
            for name, t in state_dict.items():
                key = (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape)) if t.numel() else name
                if key not in seen:
                    seen.add(key)
                    unique[name] = t
            state_dict = unique
        stem, ext = (SAFE_MODEL_NAME, "safetensors") if safe_serialization else (MODEL_NAME, "bin")
        split = split_torch_state_dict_into_shards(
            state_dict, filename_pattern=stem + "{suffix}." + ext, max_shard_size=max_shard_size
        )
        # shard files of an earlier save that this one does not overwrite would poison a later load
        shard_name = re.compile(re.escape(stem) + r"-\d{5}-of-\d{5}\." + ext)
        if self.is_main_process:
            for filename in os.listdir(save_directory):
                if shard_name.fullmatch(filename) and filename not in split.filename_to_tensors:
                    os.remove(os.path.join(save_directory, filename))
        for filename, names in split.filename_to_tensors.items():
            shard = {name: state_dict[name] for name in names}
            _write(shard, Path(save_directory) / filename, self.is_main_process, safe=safe_serialization)
        if split.is_sharded and self.is_main_process:
            index = {"metadata": split.metadata, "weight_map": split.tensor_to_filename}
            with open(os.path.join(save_directory, f"{stem}.{ext}.index.json"), "w", encoding="utf-8") as f:
                f.write(json.dumps(index, indent=2, sort_keys=True) + "\n")
        self.wait_for_everyone()

    def end_training(self):
        self.state.destroy_process_group()

    def __deepcopy__(self, memo):
        return self
