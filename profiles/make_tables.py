#!/usr/bin/env python
"""Render the markdown tables in BASELINE.md / DESIGN.md from the JSON files under profiles/."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        return None
    text = open(path).read().strip().splitlines()[-1] if name.startswith("r01_bench") else open(path).read()
    return json.loads(text)


def bench_table(files):
    print("| GPUs | accelerate_b200 samples/s (ms/step) | torch DDP+NCCL samples/s (ms/step) | ratio | weak-scaling eff. | e2e samples/s |")
    print("|---|---|---|---|---|---|")
    base = None
    for n, f in files:
        d = load(f)
        if d is None:
            continue
        if n == 1:
            base = d["value"]
        ddp = d.get("torch_ddp_nccl") or {}
        eff = f"{100 * d['value'] / (n * base):.1f} %" if base else "-"
        ratio = f"{d['value'] / ddp['value']:.3f}" if ddp.get("value") else "-"
        ddp_s = f"{ddp['value']:.1f} ({ddp['ms_per_step']:.2f})" if ddp.get("value") else "-"
        print(f"| {n} | {d['value']:.1f} ({d['ms_per_step']:.2f}) | {ddp_s} | {ratio} | {eff} | {d['e2e']['value']:.1f} |")


def allreduce_table(name, cols):
    d = load(name)
    if d is None:
        return
    print("| bytes | " + " | ".join(c for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in d["rows"]:
        cells = []
        for c in cols:
            us, bw = r.get(c + "_us"), r.get(c + "_busbw")
            cells.append(f"{us:.0f} us / {bw:.0f}" if us else "-")
        print(f"| {r['bytes']} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    print("## training step\n")
    bench_table([(1, "r01_bench_n1.json"), (2, "r01_bench_n2.json"), (4, "r01_bench_n4.json"), (8, "r01_bench_n8.json")])
    print("\n## allreduce, 8 GPUs (us / bus GB/s)\n")
    allreduce_table("r01_allreduce_n8.json", ["twoshot_b32", "nvls_b32", "api_reduce", "nccl_allreduce", "api_gather", "nccl_allgather"])
    print("\n## allreduce, 2 GPUs (us / bus GB/s)\n")
    allreduce_table("r01_allreduce_n2.json", ["twoshot_b32", "nvls_b32", "api_reduce", "nccl_allreduce"])
