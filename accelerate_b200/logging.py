"""``get_logger``: a ``logging`` adapter that knows about ranks, as training scripts written for the reference use
it (``src/accelerate/logging.py``): records carry ``[RANK r]``, by default only the main process emits them
(``main_process_only=False`` for everyone, ``in_order=True`` for everyone in rank order), ``warning_once`` de-duplicates."""

from __future__ import annotations

import functools
import logging
import os

from .state import PartialState


class MultiProcessAdapter(logging.LoggerAdapter):
    def process(self, msg, kwargs):
        msg, kwargs = super().process(msg, kwargs)
        kwargs.setdefault("stacklevel", 2)  # report the caller's frame, not this adapter's
        return f"[RANK {PartialState().process_index}] {msg}", kwargs

    def log(self, level, msg, *args, **kwargs):
        if PartialState._shared_state == {}:
            raise RuntimeError(
                "You must initialize the accelerate state by calling either `PartialState()` or `Accelerator()` before "
                "using the logging utility."
            )
        main_process_only = kwargs.pop("main_process_only", True)
        in_order = kwargs.pop("in_order", False)
        if not self.isEnabledFor(level):
            return
        msg, kwargs = self.process(msg, kwargs)
        state = PartialState()
        if in_order:  # one process at a time, everybody takes part
            for turn in range(state.num_processes):
                if turn == state.process_index:
                    self.logger.log(level, msg, *args, **kwargs)
                state.wait_for_everyone()
        elif not main_process_only or state.is_main_process:
            self.logger.log(level, msg, *args, **kwargs)

    @functools.lru_cache(None)
    def warning_once(self, *args, **kwargs):
        """``warning`` that fires once per distinct argument list."""
        self.warning(*args, **kwargs)


def get_logger(name: str, log_level: str | None = None):
    """A rank-aware logger for ``name``; ``log_level`` defaults to ``ACCELERATE_LOG_LEVEL`` (unset: leave as is)."""
    if log_level is None:
        log_level = os.environ.get("ACCELERATE_LOG_LEVEL")
    logger = logging.getLogger(name)
    if log_level is not None:
        logger.setLevel(log_level.upper())
        logger.root.setLevel(log_level.upper())
    return MultiProcessAdapter(logger, {})
