"""Host-side behaviour that the reference's own CPU tests pin, re-stated against
accelerate_b200: kwargs handlers, state singletons, optimizer/scheduler gating, the prepared dataloader,
and BASELINE.json config 1 (the nlp_example loop on CPU, world size 1, synthetic MRPC-shaped batches)."""

import math

import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from accelerate_b200 import (
    Accelerator,
    AcceleratedOptimizer,
    AcceleratedScheduler,
    DistributedDataParallelKwargs,
    GradientAccumulationPlugin,
    GradScalerKwargs,
)
from accelerate_b200.data_loader import DataLoaderShard, skip_first_batches
from accelerate_b200.state import AcceleratorState, GradientState, PartialState
from accelerate_b200.utils import (
    AutocastKwargs,
    DDPCommunicationHookType,
    find_batch_size,
    pad_across_processes,
    recursively_apply,
    send_to_device,
)


def test_kwargs_handlers_report_only_non_defaults():
    # reference tests/test_kwargs_handlers.py:60-74,173-202
    assert DistributedDataParallelKwargs().to_kwargs() == {}
    assert DistributedDataParallelKwargs(bucket_cap_mb=15, find_unused_parameters=True).to_kwargs() == {
        "bucket_cap_mb": 15, "find_unused_parameters": True}  # fmt: skip
    # bucket_cap_mb=25 equals the default -> not forwarded -> wrapper keeps the 1 MiB first bucket
    assert "bucket_cap_mb" not in DistributedDataParallelKwargs(bucket_cap_mb=25).to_kwargs()
    assert "comm_hook" not in DistributedDataParallelKwargs(comm_hook=DDPCommunicationHookType.BF16).to_kwargs()
    assert GradScalerKwargs(init_scale=1024, growth_factor=2).to_kwargs() == {"init_scale": 1024}
    assert AutocastKwargs(enabled=False).to_kwargs() == {"enabled": False}
    assert GradientAccumulationPlugin(num_steps=4).to_kwargs() == {"num_steps": 4}


def test_accelerator_rejects_duplicate_and_foreign_handlers():
    with pytest.raises(ValueError, match="only pass one"):
        Accelerator(cpu=True, kwargs_handlers=[GradScalerKwargs(), GradScalerKwargs()])
    with pytest.raises(NotImplementedError):
        Accelerator(cpu=True, fsdp_plugin=object())
    with pytest.raises(ValueError, match="Unknown mixed_precision"):
        Accelerator(cpu=True, mixed_precision="int3")


def test_state_singletons_share_and_reset():
    a = Accelerator(cpu=True, gradient_accumulation_steps=3)
    assert PartialState().device.type == "cpu" and PartialState().num_processes == 1
    assert AcceleratorState().mixed_precision == "no"
    assert GradientState().num_steps == 3 and a.gradient_accumulation_steps == 3
    GradientState._reset_state()
    assert GradientState().num_steps == 1
    with pytest.raises(ValueError):
        AcceleratorState()._check_initialized(mixed_precision="bf16")


def _regression(n=16, bs=4):
    x = torch.linspace(-1, 1, n).unsqueeze(1)
    y = 2 * x + 3
    return DataLoader(TensorDataset(x, y), batch_size=bs)


def test_optimizer_and_scheduler_are_gated_by_accumulation():
    # reference tests/test_scheduler.py:105-142, optimizer.py:112-122,145-181
    acc = Accelerator(cpu=True, gradient_accumulation_steps=2)
    model = torch.nn.Linear(1, 1)
    opt = torch.optim.SGD(model.parameters(), lr=1.0)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda step: 1.0 / (step + 1))
    model, opt, sched, dl = acc.prepare(model, opt, sched, _regression())
    assert isinstance(opt, AcceleratedOptimizer) and isinstance(sched, AcceleratedScheduler) and isinstance(dl, DataLoaderShard)
    seen = []
    for i, (x, y) in enumerate(dl):
        before = model.weight.detach().clone()
        with acc.accumulate(model):
            acc.backward(torch.nn.functional.mse_loss(model(x), y))
            opt.step()
            sched.step()
            had_grad = model.weight.grad is not None
            opt.zero_grad()
        stepped = not torch.equal(before, model.weight.detach())
        seen.append((acc.sync_gradients, stepped))
        if not acc.sync_gradients:
            assert had_grad and model.weight.grad is not None, "zero_grad must be a no-op while accumulating"
        else:
            assert model.weight.grad is None  # set_to_none=True default (optimizer.py:116-118)
    assert seen == [(False, False), (True, True), (False, False), (True, True)]
    assert sched.scheduler.last_epoch == 2  # one scheduler step per optimizer step (1 process)


def test_sync_forced_at_end_of_dataloader():
    acc = Accelerator(cpu=True, gradient_accumulation_steps=3)
    model, dl = acc.prepare(torch.nn.Linear(1, 1), _regression(n=16, bs=4))  # 4 batches, 3-step windows
    flags = []
    for x, y in dl:
        with acc.accumulate(model):
            flags.append((acc.sync_gradients, acc.gradient_state.end_of_dataloader))
    assert flags == [(False, False), (False, False), (True, False), (True, True)]
    assert acc.gradient_state.in_dataloader is False


def test_skipped_step_is_reported_to_scheduler():
    acc = Accelerator(cpu=True)
    model = torch.nn.Linear(1, 1)
    opt = acc.prepare(torch.optim.SGD(model.parameters(), lr=0.1))
    sched = AcceleratedScheduler(torch.optim.lr_scheduler.StepLR(opt.optimizer, 1), opt)
    opt._is_overflow = True
    sched.step()
    assert sched.scheduler.last_epoch == 0
    opt._is_overflow = False
    sched.step()
    assert sched.scheduler.last_epoch == 1


def test_prepare_is_idempotent_and_unwrap_restores_forward():
    # reference tests/test_accelerator.py:438-478
    acc = Accelerator(cpu=True, mixed_precision="bf16")
    model = torch.nn.Linear(2, 2)
    original_forward = model.forward
    prepared = acc.prepare(model)
    assert prepared is model and model._is_accelerate_prepared
    assert acc.prepare(prepared) is prepared and len(acc._models) == 1
    out = prepared(torch.randn(3, 2))
    assert out.dtype == torch.float32, "autocast outputs come back as fp32 (operations.py:881-940)"
    plain = acc.unwrap_model(prepared, keep_fp32_wrapper=False)
    assert plain is model and plain.forward == original_forward


def test_pytree_helpers():
    data = {"a": [torch.ones(2, 3), (torch.zeros(2), 5)], "b": "text"}
    doubled = recursively_apply(lambda t: t * 2, data)
    assert doubled["a"][0].sum() == 12 and doubled["a"][1][1] == 5 and doubled["b"] == "text"
    with pytest.raises(TypeError):
        recursively_apply(lambda t: t, data, error_on_other_type=True)
    assert find_batch_size(data) == 2
    moved = send_to_device(data, "cpu", skip_keys="b")
    assert moved["b"] == "text"
    t = torch.ones(2, 2)
    assert pad_across_processes(t) is t  # DistributedType.NO... no padding needed -> same object


def test_skip_first_batches():
    acc = Accelerator(cpu=True)
    dl = acc.prepare(_regression(n=16, bs=4))
    rest = skip_first_batches(dl, 2)
    assert len(rest) == 2
    assert [b[0][0].item() for b in rest] == [b[0][0].item() for b in list(dl)[2:]]


def test_config1_nlp_example_loop_on_cpu():
    """BASELINE.json config 1: examples/nlp_example.py:149-189 loop shape, BERT-shaped seq-cls model,
    synthetic MRPC-shaped batches, world size 1, CPU."""
    from transformers import BertConfig, BertForSequenceClassification, get_linear_schedule_with_warmup

    torch.manual_seed(42)
    cfg = BertConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, vocab_size=1000)
    model = BertForSequenceClassification(cfg)
    g = torch.Generator().manual_seed(0)

    def batch(n, length):
        return {"input_ids": torch.randint(0, 1000, (n, length), generator=g),
                "token_type_ids": torch.randint(0, 2, (n, length), generator=g),
                "attention_mask": torch.ones(n, length, dtype=torch.long),
                "labels": torch.randint(0, 2, (n,), generator=g)}  # fmt: skip

    train = [batch(16, 40 + 8 * i) for i in range(4)]
    evals = [batch(16, 64), batch(7, 50)]
    acc = Accelerator(cpu=True, gradient_accumulation_steps=2)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-5)
    train_dl = DataLoader(train, batch_size=None)
    eval_dl = DataLoader(evals, batch_size=None)
    sched = get_linear_schedule_with_warmup(opt, num_warmup_steps=1, num_training_steps=2)
    model, opt, train_dl, eval_dl, sched = acc.prepare(model, opt, train_dl, eval_dl, sched)
    model.train()
    losses = []
    for step, b in enumerate(train_dl):
        out = model(**b)
        loss = out.loss / 2
        acc.backward(loss)
        if step % 2 == 1:
            acc.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            sched.step()
            opt.zero_grad()
        losses.append(out.loss.item())
    assert all(math.isfinite(v) for v in losses)
    model.eval()
    total = 0
    for b in eval_dl:
        with torch.no_grad():
            preds = model(**b).logits.argmax(dim=-1)
        preds, refs = acc.gather_for_metrics((preds, b["labels"]))
        assert preds.shape == refs.shape
        total += preds.numel()
    assert total == 23
