"""save_state / load_state: round trip through our own files, the reference's file
names, and — when the reference tree is present (build container only) — a checkpoint written BY the reference
resumed by us and vice versa."""

import os
import subprocess
import sys

import pytest
import torch

from accelerate_b200 import Accelerator, ProjectConfiguration

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make(seed):
    torch.manual_seed(seed)
    model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    sched = torch.optim.lr_scheduler.StepLR(opt, 1, gamma=0.5)
    return model, opt, sched


def train(acc, model, opt, sched, steps, seed=0):
    g = torch.Generator().manual_seed(seed)
    for _ in range(steps):
        x = torch.randn(8, 4, generator=g)
        acc.backward(model(x).square().mean())
        opt.step()
        sched.step()
        opt.zero_grad()


def test_round_trip_and_file_names(tmp_path):
    acc = Accelerator(cpu=True)
    model, opt, sched = acc.prepare(*make(0))
    train(acc, model, opt, sched, 3)
    acc.step = 7
    torch.manual_seed(123)
    out = acc.save_state(str(tmp_path / "ckpt"))
    expected_next = torch.rand(3)
    assert sorted(os.listdir(out)) == ["model.safetensors", "optimizer.bin", "random_states_0.pkl", "scheduler.bin"]
    want = {k: v.clone() for k, v in model.state_dict().items()}
    want_lr = sched.get_last_lr()
    train(acc, model, opt, sched, 2, seed=5)  # drift away
    acc.step = 0
    acc.load_state(out)
    assert all(torch.equal(v, want[k]) for k, v in model.state_dict().items())
    assert sched.get_last_lr() == want_lr and acc.step == 7
    assert torch.equal(torch.rand(3), expected_next), "torch RNG state must be restored"
    # pickle format + automatic naming + rotation
    cfg = ProjectConfiguration(project_dir=str(tmp_path / "proj"), automatic_checkpoint_naming=True, total_limit=2)
    acc.project_configuration = cfg
    for _ in range(3):
        acc.save_state(safe_serialization=False)
    assert sorted(os.listdir(tmp_path / "proj" / "checkpoints")) == ["checkpoint_1", "checkpoint_2"]
    assert "pytorch_model.bin" in os.listdir(tmp_path / "proj" / "checkpoints" / "checkpoint_2")
    acc.load_state()  # newest
    with pytest.raises(ValueError, match="state_dict"):
        acc.register_for_checkpointing(object())


def test_custom_objects_and_tied_weights(tmp_path):
    class Counter:
        def __init__(self):
            self.n = 0

        def state_dict(self):
            return {"n": self.n}

        def load_state_dict(self, s):
            self.n = s["n"]

    class Tied(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(10, 4)
            self.head = torch.nn.Linear(4, 10, bias=False)
            self.head.weight = self.emb.weight

        def forward(self, ids):
            return self.head(self.emb(ids))

    acc = Accelerator(cpu=True)
    model = acc.prepare(Tied())
    counter = Counter()
    counter.n = 41
    acc.register_for_checkpointing(counter)
    out = acc.save_state(str(tmp_path / "c"))
    counter.n = 0
    with torch.no_grad():
        model.emb.weight.add_(1.0)
    before = model.emb.weight.clone()
    acc.load_state(out)
    assert counter.n == 41 and not torch.equal(model.emb.weight, before)
    assert model.head.weight.data_ptr() == model.emb.weight.data_ptr()


REF = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_checkpoints_interchange_with_the_reference(tmp_path):
    script = r"""
import sys, torch
mode, path = sys.argv[1], sys.argv[2]
from accelerate import Accelerator
acc = Accelerator(cpu=True)
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
sched = torch.optim.lr_scheduler.StepLR(opt, 1, gamma=0.5)
model, opt, sched = acc.prepare(model, opt, sched)
if mode == "save":
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        acc.backward(model(torch.randn(8, 4, generator=g)).square().mean()); opt.step(); sched.step(); opt.zero_grad()
    acc.save_state(path)
    torch.save({k: v for k, v in model.state_dict().items()}, path + "/expected.pt")
else:
    acc.load_state(path)
    exp = torch.load(path + "/expected.pt")
    assert all(torch.equal(v, exp[k]) for k, v in model.state_dict().items())
    assert abs(sched.get_last_lr()[0] - 1e-2 * 0.5 ** 3) < 1e-12, sched.get_last_lr()
    print("REF_LOADED_OK")
"""
    env = dict(os.environ, PYTHONPATH=REF, PYTHONDONTWRITEBYTECODE="1", ACCELERATE_USE_CPU="yes")
    # 1) the reference writes, we read
    ref_dir = str(tmp_path / "by_reference")
    res = subprocess.run([sys.executable, "-c", script, "save", ref_dir], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    acc = Accelerator(cpu=True)
    model, opt, sched = acc.prepare(*make(1))
    acc.load_state(ref_dir)
    exp = torch.load(ref_dir + "/expected.pt")
    assert all(torch.equal(v, exp[k]) for k, v in model.state_dict().items())
    assert abs(sched.get_last_lr()[0] - 1e-2 * 0.5**3) < 1e-12
    assert len(opt.state_dict()["state"]) == 4
    # 2) we write, the reference reads
    ours = str(tmp_path / "by_us")
    acc.save_state(ours)
    torch.save({k: v for k, v in model.state_dict().items()}, ours + "/expected.pt")
    res = subprocess.run([sys.executable, "-c", script, "load", ours], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "REF_LOADED_OK" in res.stdout, res.stderr[-2000:]
