"""Any world size: the fused bucket allreduce (two-shot and NVLS) against the oracle on seeded data, through
the C ABI, for several sizes/dtypes; results must be identical on all ranks, two-shot bit-exact vs the
oracle's rank-order sum, NVLS within float32 rounding of it (the switch adds in its own fixed order; bit-exact at N=2)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ddp_oracle as oracle


def main():
    from accelerate_b200 import Accelerator, native
    from accelerate_b200.state import PartialState
    acc = Accelerator()
    rank, world, dev = acc.process_index, acc.num_processes, acc.device
    rt = PartialState().runtime
    st = torch.cuda.current_stream().cuda_stream
    max_elems = (1 << 22) + 4096
    seg = rt.alloc_segment(max_elems * 4, multicast=True)
    algos = [("twoshot", native.ALGO_TWOSHOT)] + ([("nvls", native.ALGO_NVLS)] if seg.multicast else [])
    slot = rt.reserve_slots(1)
    for dtype, dt in ((torch.float32, native.F32), (torch.bfloat16, native.BF16), (torch.float16, native.F16)):
        es = 2 if dtype != torch.float32 else 4
        for numel in (16 // es * 1, 16 // es * 513, 1 << 16, (1 << 22) + 16 // es * 77):
            g = torch.Generator().manual_seed(1000 * rank + numel % 997)
            x = (torch.randn(numel, generator=g) * 3).to(dtype).to(dev)
            everyone = acc.gather(x.reshape(1, -1)).float().cpu().numpy()
            want = oracle.ddp_mean_reduce([everyone[r] for r in range(world)])
            if dtype != torch.float32:
                want = torch.from_numpy(want).to(dtype).float().numpy()
            for name, algo in algos:
                view = seg.tensor(dtype, numel)
                view.copy_(x)
                acc.wait_for_everyone()
                rt.ctx.allreduce_bucket(seg.seg, 0, numel, dt, 1.0 / world, slot, native.F_STATS, algo, native.CH_USER, 0,
                                        native.NO_STREAM, st)
                got = view.float().cpu().numpy()
                allranks = acc.gather(view.reshape(1, -1).clone()).float().cpu().numpy()
                assert all(allranks[0].tobytes() == allranks[r].tobytes() for r in range(world)), (name, dtype, numel, "ranks differ")
                if name == "twoshot" or world == 2:  # NVLS adds in the switch's own (fixed, but not rank) order
                    assert got.tobytes() == want.tobytes(), (name, dtype, numel, np.abs(got - want).max())
                else:
                    tol = 1e-6 if dtype == torch.float32 else 2e-2
                    assert np.allclose(got, want, rtol=tol, atol=tol), (name, dtype, numel, np.abs(got - want).max())
                out = torch.empty(3, device=dev)
                rt.ctx.finalize_norm(slot, 1, True, 2.0, -1.0, None, out.data_ptr(), st)
                ref = float(np.sqrt((got.astype(np.float64) ** 2).sum()))
                assert abs(out[0].item() - ref) <= 1e-5 * ref + 1e-12, (name, dtype, numel, out[0].item(), ref)
    rt.ctx.check_status()
    acc.wait_for_everyone()
    print(f"ALLREDUCE_CHECK_OK rank {rank}/{world} algos {[a for a, _ in algos]}", flush=True)
    PartialState().destroy_process_group()


if __name__ == "__main__":
    main()
