"""The same user training script under the reference and under this package (2 ranks, CPU/gloo): every value the loop
can observe must be identical — at two ranks the gradient mean is exact, so even the final weights match bit for bit.
Runs only where the reference tree exists (the build container); its outputs for the GPU box are tests/golden/."""

import json
import os

import pytest

from conftest import run_concurrently, run_distributed

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")


def check_training_loop(tmp_path):
    outs = {}
    for impl in ("reference", "ours"):
        base = str(tmp_path / impl)
        env = {"PYTHONDONTWRITEBYTECODE": "1"}
        out = run_distributed("dist_loop_diff.py", 2, "--impl", impl, "--out", base, "--variant", "all", env=env, timeout=900)
        assert out.count("LOOP_DONE") == 2, out[-3000:]
        outs[impl] = [json.load(open(f"{base}.rank{r}.json")) for r in range(2)]
    for r in range(2):
        assert outs["reference"][r].keys() == outs["ours"][r].keys() and len(outs["ours"][r]) == 7
        for variant in outs["reference"][r]:
            ref, ours = outs["reference"][r][variant], outs["ours"][r][variant]
            assert len(ref["steps"]) == len(ours["steps"]) > 0
            for k, (a, b) in enumerate(zip(ref["steps"], ours["steps"])):
                assert a == b, (variant, r, k, a, b)
            assert ref["eval"] == ours["eval"], (variant, r)
            assert ref["reduced_probe"] == ours["reduced_probe"]
            assert ref["final_state_sha256"] == ours["final_state_sha256"], (variant, r)


def check_api(tmp_path, world=3):
    outs = {}
    for impl in ("reference", "ours"):
        base = str(tmp_path / impl)
        project = tmp_path / f"scratch_{impl}" / "project"
        project.mkdir(parents=True)
        out = run_distributed("dist_api_diff.py", world, "--impl", impl, "--out", base, "--project", str(project),
                              env={"PYTHONDONTWRITEBYTECODE": "1"}, timeout=600)  # fmt: skip
        assert out.count("API_DONE") == world, out[-3000:]
        outs[impl] = [json.load(open(f"{base}.rank{r}.json")) for r in range(world)]
    for r in range(world):
        ref, ours = outs["reference"][r], outs["ours"][r]
        assert ref.keys() == ours.keys()
        for key in ref:
            assert ref[key] == ours[key], (world, r, key, ref[key], ours[key])


def check_dataloaders(tmp_path, world=3):
    """~130 (length, batch size, drop_last, split_batches, shard/dispatch/iterable, shuffle, even_batches) cases: every
    batch on every rank over two epochs, end_of_dataloader/remainder, gather_for_metrics, len, skip_first_batches."""
    outs = {}
    for impl in ("reference", "ours"):
        base = str(tmp_path / impl)
        out = run_distributed("dist_loader_diff.py", world, "--impl", impl, "--out", base, env={"PYTHONDONTWRITEBYTECODE": "1"}, timeout=900)
        assert out.count("LOADER_DONE") == world, out[-3000:]
        outs[impl] = [json.load(open(f"{base}.rank{r}.json")) for r in range(world)]
    for r in range(world):
        ref, ours = outs["reference"][r]["runs"], outs["ours"][r]["runs"]
        assert len(ref) == len(ours) > 50
        for a, b in zip(ref, ours):
            assert a == b, (world, r, a["cfg"])


def check_resume(tmp_path):
    """save_state in the middle of an epoch with either library, resume (load_state + skip_first_batches) with either
    library, 2 ranks: all four combinations end with the weights of an uninterrupted run; same files on disk."""
    out = run_distributed("dist_resume_diff.py", 2, "--dir", str(tmp_path), env={"PYTHONDONTWRITEBYTECODE": "1"}, timeout=600)
    assert out.count("RESUME_OK") == 2, out[-3000:]


def test_same_script_under_both_libraries(tmp_path):
    """The four differential checks above, run concurrently (each launches its own torchrun groups):
    training loop (7 variants, 2 ranks), collective / process-control / checkpoint API (3 ranks), prepared dataloaders
    (~70 configurations, 3 ranks), mid-epoch save/resume across the libraries (2 ranks)."""
    checks = {"training_loop": check_training_loop, "api": check_api, "dataloaders": check_dataloaders, "resume": check_resume}
    attempt = {name: 0 for name in checks}

    def fresh_dir(name):  # a retry must not trip over the files of the first attempt
        attempt[name] += 1
        d = tmp_path / f"{name}_{attempt[name]}"
        d.mkdir()
        return d

    run_concurrently({name: (lambda n=name, fn=fn: fn(fresh_dir(n))) for name, fn in checks.items()}, max_workers=4)
