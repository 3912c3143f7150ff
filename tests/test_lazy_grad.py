"""Host logic of the pending gradient multiplier (accelerate_b200/lazy_grad.py; reference semantics it must preserve:
after ``accelerator.clip_grad_norm_`` — src/accelerate/accelerator.py:2946-3007 — every reader of ``p.grad`` sees the
clipped values).  CPU only: the multiply is a plain torch op here; the CUDA wiring is covered by -m gpu tests."""

import torch

from accelerate_b200 import lazy_grad


def params(n=3):
    ps = [torch.nn.Parameter(torch.full((4,), float(i + 1))) for i in range(n)]
    for p in ps:
        lazy_grad.raw_set_grad(p, torch.full((4,), 2.0))
    return ps


def pending_for(ps, factor, log):
    out = torch.tensor([0.0, factor, 0.0])

    def apply():
        log.append("applied")
        for p in ps:
            lazy_grad.raw_grad(p).mul_(out[1])

    return lazy_grad.defer(out, apply, owner="test", param_ids=frozenset(id(p) for p in ps))


def test_read_triggers_the_pass_exactly_once_and_only_for_covered_parameters():
    ps, other, log = params(), params(1)[0], []
    pend = pending_for(ps, 0.25, log)
    assert lazy_grad.installed() and lazy_grad.has_pending()
    assert float(lazy_grad.raw_grad(ps[0])[0]) == 2.0, "raw access must not observe"
    assert float(other.grad[0]) == 2.0 and not log, "a parameter outside the pending set must not trigger the pass"
    assert float(ps[1].grad[0]) == 0.5 and log == ["applied"], "first observation applies the multiplier"
    assert float(ps[0].grad[0]) == 0.5 and float(ps[2].grad[0]) == 0.5 and log == ["applied"]
    assert pend.done and not lazy_grad.has_pending()
    assert pend.consume() is False, "nothing left for the optimizer kernel to fold in"


def test_consume_hands_the_multiplier_to_the_optimizer_and_skips_the_pass():
    ps, log = params(), []
    pend = pending_for(ps, 0.25, log)
    taken = lazy_grad.take_for("test")
    assert taken is pend and pend.done and not lazy_grad.has_pending()
    assert float(ps[0].grad[0]) == 2.0 and not log, "consumed: gradients stay unscaled, the kernel multiplies on the fly"


def test_assignment_applies_first_and_bumps_the_epoch():
    ps, log = params(), []
    pending_for(ps, 0.5, log)
    epoch = lazy_grad.assign_epoch
    ps[0].grad = None  # zero_grad(set_to_none=True) on a covered parameter
    assert log == ["applied"] and lazy_grad.assign_epoch == epoch + 1
    assert ps[0].grad is None and float(ps[1].grad[0]) == 1.0


def test_torch_optimizers_and_clip_see_scaled_gradients():
    ps, log = params(), []
    pending_for(ps, 0.5, log)
    opt = torch.optim.SGD(ps, lr=1.0)
    opt.step()  # reads p.grad from Python -> pass runs first
    assert log == ["applied"] and float(ps[0].detach()[0]) == 1.0 - 1.0 and float(ps[2].detach()[0]) == 3.0 - 1.0
    pending_for(ps, 0.5, log)
    total = torch.nn.utils.clip_grad_norm_(ps, 1e9)
    assert log == ["applied", "applied"] and abs(float(total) - (12 * 0.25) ** 0.5) < 1e-6
    opt.zero_grad()
    assert all(p.grad is None for p in ps)


def test_flush_params_and_discard():
    a, b, log_a, log_b = params(2), params(2), [], []
    pa, pb = pending_for(a, 0.5, log_a), pending_for(b, 0.5, log_b)
    lazy_grad.flush_params({id(b[1])})
    assert log_b == ["applied"] and not log_a and pb.done and not pa.done
    pa.discard()
    assert not lazy_grad.has_pending() and float(a[0].grad[0]) == 2.0
