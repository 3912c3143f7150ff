import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(autouse=True)
def _fresh_state():
    """Every test starts with pristine Borg singletons (the reference's AccelerateTestCase.tearDown)."""
    yield
    from accelerate_b200.state import AcceleratorState, GradientState

    AcceleratorState._reset_state()
    GradientState._reset_state()
    # PartialState (process group / native runtime) survives: one per test process
    from accelerate_b200 import lazy_grad

    for pending in list(lazy_grad._pending):  # a deferred clip nobody consumed must not leak into the next test
        pending.discard()


import threading  # noqa: E402

_PORT_LOCK = threading.Lock()
_PORTS_HANDED_OUT = set()


def run_distributed(script, nproc, *args, env=None, timeout=600):
    """torchrun a script under tests/scripts with `nproc` ranks on this node; returns combined output."""
    import socket
    import subprocess

    with _PORT_LOCK:  # concurrent launches (thread pools in the differential tests) must not be handed the same port
        while True:
            s = socket.socket()
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
            s.close()
            if port not in _PORTS_HANDED_OUT:
                _PORTS_HANDED_OUT.add(port)
                break
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "scripts", script), *args]  # fmt: skip
    full_env = dict(os.environ)
    full_env.setdefault("OMP_NUM_THREADS", "1")
    full_env.update(env or {})
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=full_env, cwd=ROOT)
    out = res.stdout + "\n" + res.stderr
    assert res.returncode == 0, f"{' '.join(cmd)} failed ({res.returncode}):\n{out[-6000:]}"
    return out


def run_concurrently(jobs, max_workers=3):
    """jobs: {label: thunk}.  Run them in a thread pool (each thunk launches its own subprocesses); a thunk that
    raised is retried ONCE, alone — a loaded machine can time a rendezvous out — and only a second failure counts.
    Returns {label: result}; raises AssertionError listing every label that failed twice."""
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=max_workers) as pool:
        futures = {label: pool.submit(thunk) for label, thunk in jobs.items()}
    results, failures = {}, {}
    for label, fut in futures.items():
        first = fut.exception()
        if first is None:
            results[label] = fut.result()
            continue
        import warnings

        warnings.warn(f"{label} needed a retry after: {str(first)[-1500:]}")
        try:
            results[label] = jobs[label]()
        except BaseException as second:  # noqa: BLE001 - reported below
            failures[label] = f"first attempt: {first!r}\nretry: {second!r}"
    assert not failures, "\n\n".join(f"{label}:\n{why}" for label, why in failures.items())
    return results
