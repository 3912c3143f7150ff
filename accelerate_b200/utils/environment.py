"""Environment helpers used by the state singletons (reference: src/accelerate/utils/environment.py)."""

from __future__ import annotations

import os
from contextlib import contextmanager

_TRUE = {"1", "true", "t", "yes", "y", "on"}
_FALSE = {"0", "false", "f", "no", "n", "off"}


def str_to_bool(value) -> int:
    text = str(value).strip().lower()
    if text in _TRUE:
        return 1
    if text in _FALSE:
        return 0
    raise ValueError(f"invalid truth value {value!r}")


def parse_flag_from_env(key: str, default: bool = False) -> bool:
    raw = os.environ.get(key)
    return bool(default) if raw is None else bool(str_to_bool(raw))


def parse_choice_from_env(key: str, default: str = "no") -> str:
    return os.environ.get(key, str(default))


def get_int_from_env(keys, default: int) -> int:
    for key in keys:
        raw = os.environ.get(key)
        if raw is not None and int(raw) >= 0:
            return int(raw)
    return default


@contextmanager
def clear_environment():
    """Temporarily empty ``os.environ`` (restored even on error)."""
    saved = dict(os.environ)
    os.environ.clear()
    try:
        yield
    finally:
        os.environ.clear()
        os.environ.update(saved)


@contextmanager
def patch_environment(**kwargs):
    """Temporarily set upper-cased env vars; previous values come back afterwards."""
    previous = {}
    for key, value in kwargs.items():
        key = key.upper()
        previous[key] = os.environ.get(key)
        os.environ[key] = str(value)
    try:
        yield
    finally:
        for key, old in previous.items():
            if old is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = old


def purge_accelerate_environment(func_or_cls):
    """Decorator for tests (functions, ``unittest.TestCase`` or plain pytest classes, subclasses included): whatever
    ``ACCELERATE_*`` variables a test sets are removed — and pre-existing ones restored — when it returns, so tests do
    not leak configuration into each other (``utils/environment.py:405-474``)."""
    import functools

    def guarded(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            before = {k: v for k, v in os.environ.items() if k.startswith("ACCELERATE_")}
            try:
                return func(*args, **kwargs)
            finally:
                for key in [k for k in os.environ if k.startswith("ACCELERATE_")]:
                    if key in before:
                        os.environ[key] = before[key]
                    else:
                        del os.environ[key]

        wrapper._accelerate_is_purged_environment_wrapped = True
        return wrapper

    if not isinstance(func_or_cls, type):
        return guarded(func_or_cls)

    def wrap_tests(cls):
        for name in dir(cls):
            method = getattr(cls, name)
            if name.startswith("test") and callable(method) and not hasattr(method, "_accelerate_is_purged_environment_wrapped"):
                setattr(cls, name, guarded(method))
        return cls

    wrap_tests(func_or_cls)
    func_or_cls.__init_subclass__ = classmethod(lambda cls, **kw: wrap_tests(cls))
    return func_or_cls
