"""Diagnostic stress for the accumulate -> sync path on CUDA (not part of the test suite)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from scenarios import make_batch, make_model
from oracle import ddp_oracle as oracle

def main():
    from accelerate_b200 import Accelerator
    from accelerate_b200.state import AcceleratorState, GradientState, PartialState
    acc = Accelerator(gradient_accumulation_steps=2)
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    model = acc.prepare(make_model())
    twin = make_model().to(dev)
    bad = 0
    names = ["W1", "b1", "W2", "b2"]
    for it in range(int(os.environ.get("DIAG_ITERS", "30"))):
        model.zero_grad(set_to_none=(it % 3 != 2)); twin.zero_grad(set_to_none=True)
        micro_grads = []
        for micro in range(2):
            x, y = (t.to(dev) for t in make_batch(rank, 10 * it + micro))
            (torch.nn.functional.mse_loss(twin(x), y) / 2).backward()
            if micro == 0:
                micro_grads = [p.grad.detach().clone() for p in twin.parameters()]
            with acc.accumulate(model):
                acc.backward(torch.nn.functional.mse_loss(model(x), y))
        torch.cuda.synchronize()
        for i, (p, tp) in enumerate(zip(model.parameters(), twin.parameters())):
            loc = acc.gather(tp.grad.reshape(1, -1)).cpu().numpy()
            m0 = acc.gather(micro_grads[i].reshape(1, -1)).cpu().numpy()
            want = oracle.ddp_mean_reduce([loc[r] for r in range(world)])
            want_m0 = oracle.ddp_mean_reduce([m0[r] for r in range(world)])
            got = p.grad.reshape(-1).cpu().numpy()
            if got.tobytes() != want.tobytes():
                bad += 1
                idx = np.nonzero(got != want)[0]
                kind = "other"
                if np.allclose(got[idx], want_m0[idx], rtol=1e-6): kind = "micro0-only"
                elif np.allclose(got[idx], (want - want_m0)[idx], rtol=1e-4, atol=1e-8): kind = "micro1-only"
                elif np.allclose(got[idx], loc[rank][idx], rtol=1e-6): kind = "local-unreduced"
                print(f"[rank {rank}] it {it} {names[i]}: {idx.size}/{got.size} wrong, first idx {idx[:6]}, kind={kind}, "
                      f"got {got[idx[:3]]} want {want[idx[:3]]} m0mean {want_m0[idx[:3]]} local {loc[rank][idx[:3]]}", flush=True)
    PartialState().runtime.ctx.check_status()
    acc.wait_for_everyone()
    print(f"DIAG rank {rank} bad={bad} bulk={os.environ.get('B200DP_BULK_PACK','1')}", flush=True)
    PartialState().destroy_process_group()

if __name__ == "__main__":
    main()
