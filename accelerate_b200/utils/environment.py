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
