"""The collective / process-control API surface driven by ONE script under the reference and under this package
(CPU/gloo); everything returned is serialised (dtype, shape, values) to JSON per rank and compared by
tests/test_reference_loop_parity.py.  Complements the golden fixtures with nested containers, odd dtypes and the
process-control helpers."""
import argparse
import json
import os
import sys

import torch


def ser(obj):
    if isinstance(obj, torch.Tensor):
        return {"__tensor__": str(obj.dtype), "shape": list(obj.shape), "values": obj.detach().cpu().double().flatten().tolist()}
    if isinstance(obj, dict):
        return {"__dict__": type(obj).__name__, "items": {str(k): ser(v) for k, v in obj.items()}}
    if isinstance(obj, (list, tuple)):
        return {"__seq__": type(obj).__name__, "items": [ser(v) for v in obj]}
    if isinstance(obj, (int, float, str, bool)) or obj is None:
        return obj
    return {"__repr__": repr(obj)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ours", "reference"], required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--project", required=True, help="scratch directory for the checkpointing part")
    args = ap.parse_args()
    os.environ["ACCELERATE_USE_CPU"] = "yes"
    torch.set_num_threads(1)
    if args.impl == "reference":
        sys.path.insert(0, "/root/reference/src")
        import accelerate as lib
        from accelerate import utils
        from accelerate.local_sgd import LocalSGD
    else:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        import accelerate_b200 as lib
        from accelerate_b200 import utils
        from accelerate_b200.local_sgd import LocalSGD

    project = utils.ProjectConfiguration(project_dir=args.project, automatic_checkpoint_naming=True, total_limit=2)
    acc = lib.Accelerator(cpu=True, project_config=project)
    rank, world = acc.process_index, acc.num_processes
    g = torch.Generator().manual_seed(11 + rank)
    log = {}

    nested = {
        "a": torch.randn(3, 2, generator=g),
        "b": [torch.arange(4, dtype=torch.int64) + 10 * rank, (torch.tensor(rank == 0), torch.randn(2, 2, 2, generator=g).double())],
        "c": torch.randint(0, 255, (5,), generator=g, dtype=torch.uint8),
        "h": torch.randn(2, 3, generator=g).to(torch.float16),
    }
    log["gather_nested"] = ser(acc.gather(nested))
    log["gather_scalar"] = ser(acc.gather(torch.tensor(3.5 * (rank + 1))))
    log["gather_noncontig"] = ser(acc.gather(torch.randn(4, 6, generator=g).t()))
    log["reduce_sum"] = ser(acc.reduce({"x": torch.randn(4, generator=g), "n": torch.tensor([rank + 1, 7])}, "sum"))
    log["reduce_mean"] = ser(acc.reduce([torch.randn(3, 3, generator=g), torch.randn(2, generator=g).double()], "mean"))
    ragged = torch.randn(2 + rank, 3, generator=g)
    for kw in ({}, {"dim": 0, "pad_index": -7}, {"dim": 0, "pad_first": True}, {"dim": 1}, {"dim": -2, "pad_index": 2}, {"dim": 5}):
        log[f"pad_{sorted(kw.items())}"] = ser(acc.pad_across_processes(ragged, **kw))
    log["pad_nested"] = ser(acc.pad_across_processes({"r": ragged, "same": torch.ones(2, 2)}, dim=0))
    log["pad_int"] = ser(acc.pad_across_processes(torch.arange(3 + 2 * rank), pad_index=-1))
    # degenerate shapes
    log["gather_empty"] = ser(acc.gather(torch.empty(0, 3)))
    log["gather_zero_dim_inside"] = ser(acc.gather(torch.zeros(2, 0, 4)))
    log["reduce_empty"] = ser(acc.reduce(torch.empty(0), "sum"))
    log["pad_from_empty"] = ser(acc.pad_across_processes(torch.ones(rank, 2), dim=0, pad_index=9))  # rank 0 holds nothing
    log["pad_bool"] = ser(acc.pad_across_processes(torch.ones(1 + rank, dtype=torch.bool)))
    log["pad_float_index_on_int"] = ser(acc.pad_across_processes(torch.arange(2 + rank), pad_index=0.5))
    payload = {"rank": rank, "words": ["x"] * (rank + 1), "t": [1.5, None]}
    log["gather_object"] = ser(utils.gather_object([payload, rank]))
    objs = [payload if rank == 0 else None, "keep" if rank == 0 else "drop"]
    utils.broadcast_object_list(objs, from_process=0)
    log["broadcast_object_list"] = ser(objs)
    bt = {"w": torch.full((2, 3), float(rank)), "i": torch.arange(3) * (rank + 1)}
    log["broadcast_from_1"] = ser(utils.broadcast(bt, from_process=world - 1))
    log["bt_mutated_in_place"] = ser(bt)
    log["concatenate"] = ser(utils.concatenate([{"k": torch.ones(1, 2) * rank}, {"k": torch.zeros(2, 2)}]))
    log["find_batch_size"] = utils.find_batch_size({"z": [torch.zeros(5, 2)], "y": torch.zeros(5)})
    log["slice_tensors"] = ser(utils.slice_tensors({"q": torch.arange(10).reshape(5, 2)}, slice(1, 3)))

    items = list(range(7))
    with acc.split_between_processes(items) as part:
        log["split_list"] = ser(part)
    with acc.split_between_processes(items, apply_padding=True) as part:
        log["split_list_padded"] = ser(part)
    with acc.split_between_processes({"a": list(range(5)), "b": torch.arange(10).reshape(5, 2)}) as part:
        log["split_dict"] = ser(part)
    with acc.split_between_processes(torch.arange(9).reshape(3, 3), apply_padding=True) as part:
        log["split_tensor_padded"] = ser(part)

    @acc.on_main_process
    def only_main():
        return "ran"

    @acc.on_last_process
    def only_last():
        return "ran"

    @acc.on_process(process_index=world - 1)
    def only_idx():
        return "ran"

    log["on_main"], log["on_last"], log["on_idx"] = only_main(), only_last(), only_idx()
    log["flags"] = [acc.is_main_process, acc.is_local_main_process, acc.is_last_process, acc.use_distributed, str(acc.distributed_type).split(".")[-1]]
    log["trigger_before"] = bool(acc.check_trigger())
    if rank == world - 1:
        acc.set_trigger()
    log["trigger_after"] = bool(acc.check_trigger())
    log["trigger_reset"] = bool(acc.check_trigger())

    # LocalSGD: parameters are averaged every 2 steps; no gradient synchronisation in between
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 2)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    model, opt = acc.prepare(model, opt)
    with LocalSGD(accelerator=acc, model=model, local_sgd_steps=2, enabled=True) as local_sgd:
        for step in range(5):
            x = torch.randn(3, 4, generator=g)
            with acc.accumulate(model):
                acc.backward(model(x).square().mean())
                opt.step()
                opt.zero_grad()
                local_sgd.step()
            log[f"local_sgd_{step}"] = ser(acc.unwrap_model(model).weight)

    import warnings

    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        with acc.join_uneven_inputs([model], even_batches=False):
            pass
    log["join_uneven_inputs_on_cpu"] = sorted({str(w.message) for w in caught if "Joining uneven" in str(w.message)})

    @acc.on_local_process(local_process_index=0)
    def only_local0():
        return "ran"

    log["on_local"] = only_local0()

    # checkpoint bookkeeping: automatic naming, rotation with total_limit=2, pre-hooks, registered objects
    class Counter:
        def __init__(self):
            self.n = 0

        def state_dict(self):
            return {"n": self.n}

        def load_state_dict(self, sd):
            self.n = sd["n"]

    counter = Counter()
    acc.register_for_checkpointing(counter)
    seen = []
    h_save = acc.register_save_state_pre_hook(lambda models, weights, out: seen.append(("save", len(models), sorted(weights[0]), os.path.basename(out))))
    h_load = acc.register_load_state_pre_hook(lambda models, d: seen.append(("load", len(models), os.path.basename(d))))
    for k in range(3):
        counter.n = 10 + k
        acc.save_state()
        log[f"save_iteration_{k}"] = acc.save_iteration
    acc.wait_for_everyone()
    log["checkpoint_dirs"] = sorted(os.listdir(os.path.join(args.project, "checkpoints")))
    log["checkpoint_files"] = sorted(os.listdir(os.path.join(args.project, "checkpoints", "checkpoint_2")))
    counter.n = -1
    acc.load_state()
    log["counter_after_load"] = counter.n
    h_save.remove()
    h_load.remove()
    log["hook_calls"] = ser(seen)
    log["project"] = [os.path.basename(str(acc.project_dir)), acc.logging_dir is None or os.path.basename(str(acc.logging_dir))]

    class Words(torch.utils.data.Dataset):
        def __len__(self):
            return 11

        def __getitem__(self, i):
            return {"id": i, "word": f"w{i}"}

    wdl = acc.prepare(torch.utils.data.DataLoader(Words(), batch_size=2, collate_fn=lambda b: b))
    collected = []
    for b in wdl:
        collected.extend(acc.gather_for_metrics(b, use_gather_object=True))
    log["gather_objects_for_metrics"] = ser(collected)

    # model-facing helpers: value clipping, unwrap / state-dict names, sharded save_model layout, free_memory
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(6, 40), torch.nn.BatchNorm1d(40), torch.nn.Linear(40, 3))
    net_opt = torch.optim.SGD(net.parameters(), lr=0.1)
    net, net_opt = acc.prepare(net, net_opt)
    acc.backward(net(torch.randn(5, 6, generator=g)).square().sum() * 100)
    acc.clip_grad_value_(net.parameters(), 0.25)
    log["clip_grad_value"] = ser([p.grad for p in net.parameters()])
    log["wrapped_state_dict_keys"] = sorted(net.state_dict())
    log["get_state_dict_keys"] = sorted(acc.get_state_dict(net))
    log["get_state_dict_wrapped_keys"] = sorted(acc.get_state_dict(net, unwrap=False))
    log["unwrap_type"] = type(acc.unwrap_model(net)).__name__
    model_dir = os.path.join(args.project, "saved_model")
    acc.save_model(net, model_dir, max_shard_size=400)
    acc.wait_for_everyone()
    log["save_model_files"] = sorted(os.listdir(model_dir))
    with open(os.path.join(model_dir, "model.safetensors.index.json")) as f:
        log["save_model_index"] = json.load(f)
    log["free_memory"] = ser(acc.free_memory(net_opt))
    log["after_free_memory"] = [len(acc._models), len(acc._optimizers), len(acc._dataloaders)]

    dl = acc.prepare(torch.utils.data.DataLoader(list(range(22)), batch_size=3))
    log["skip_first"] = ser([b for b in acc.skip_first_batches(dl, 2)])
    log["dl_len"], log["dl_total"] = len(dl), getattr(dl, "total_dataset_length", None)
    acc.wait_for_everyone()
    with open(f"{args.out}.rank{rank}.json", "w") as f:
        json.dump(log, f)
    print(f"API_DONE {args.impl} rank {rank}", flush=True)
    acc.end_training()


if __name__ == "__main__":
    main()
