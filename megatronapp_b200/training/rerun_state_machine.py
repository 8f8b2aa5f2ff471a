"""Result validation with re-execution to classify faults, plus numerical fault injection.

Parity: reference ``core/rerun_state_machine.py:131-990`` (``RerunStateMachine`` with modes
``disabled`` / ``validate_results`` / ``report_stats``; ``should_run_forward_backward`` loop
protocol; ``validate_result(result, rejection_func, ...)``; re-running the same step on the same
data through the rewindable ``RerunDataIterator`` to tell *transient* (first rerun differs) from
*persistent* (reproducible on this GPU) from *correct* (reproducible => result is legitimate);
exit codes for checkpoint-and-exit) and ``RerunErrorInjector`` (:1147-1240,
``--error-injection-rate/-type``).
"""
from __future__ import annotations

import math
import random
from enum import Enum
from typing import Any, Callable, Iterator, List, Optional

import torch

EXIT_CODE_RESUME_TO_DISAMBIGUATE = 16
EXIT_CODE_FAILED_ON_RESULT_VALIDATION = 17


class RerunMode(str, Enum):
    DISABLED = "disabled"
    VALIDATE_RESULTS = "validate_results"
    REPORT_DETERMINISM_STATS = "report_stats"


class RerunState(Enum):
    NOT_RUNNING_YET = 0
    INITIAL_RUN = 1
    RERUNNING_IN_PLACE = 2
    WILL_RERUN_FROM_CHECKPOINT = 3
    RERUNNING_FROM_CHECKPOINT = 4


class RerunDiagnostic(str, Enum):
    CORRECT_RESULT = "correct_result"
    TRANSIENT_ERROR = "transient_error"
    PERSISTENT_ERROR = "persistent_error"


class RerunDataIterator:
    """Records the batches of the current step so the step can be replayed on identical data."""

    def __init__(self, iterable: Iterator[Any]):
        self.iterable = iterable
        self.saved: List[Any] = []
        self.replaying = False
        self.pos = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self.replaying:
            if self.pos < len(self.saved):
                v = self.saved[self.pos]
                self.pos += 1
                return v
            self.replaying = False
        v = next(self.iterable)
        self.saved.append(v)
        return v

    def rewind(self):
        self.replaying, self.pos = True, 0

    def advance(self):
        self.saved, self.replaying, self.pos = [], False, 0

    def state_dict(self):
        return {"saved": self.saved, "replaying": self.replaying, "pos": self.pos}

    def load_state_dict(self, sd):
        self.saved, self.replaying, self.pos = sd["saved"], sd["replaying"], sd["pos"]


class RerunErrorInjector:
    """Randomly corrupts validated results so the recovery path can be exercised."""

    def __init__(self, error_injection_rate: int = 0, error_injection_type: str = "transient_error", seed: int = 0):
        self.rate = error_injection_rate
        self.kind = RerunDiagnostic(error_injection_type)
        self.rng = random.Random(seed)
        self.armed_iteration: Optional[int] = None
        self.injected = 0

    def maybe_inject(self, iteration: int, state: RerunState) -> bool:
        if self.rate <= 0:
            return False
        if state == RerunState.INITIAL_RUN:
            if self.rng.randrange(self.rate) == 0:
                self.armed_iteration = iteration
                self.injected += 1
                return True
            return False
        if self.armed_iteration == iteration:
            # transient: only the first run is wrong; persistent / "correct": reproducible
            return self.kind in (RerunDiagnostic.PERSISTENT_ERROR, RerunDiagnostic.CORRECT_RESULT)
        return False


class RerunStateMachine:
    def __init__(self, mode: RerunMode = RerunMode.DISABLED, error_injector: Optional[RerunErrorInjector] = None,
                 result_rejected_tracker_filename: Optional[str] = None):
        self.mode = RerunMode(mode)
        self.state = RerunState.NOT_RUNNING_YET
        self.injector = error_injector or RerunErrorInjector()
        self.current_iteration = 0
        self.rerun_requested = False
        self.first_results: List[float] = []
        self.call_idx = 0
        self.diagnostics: List[tuple] = []
        self.checkpoint_requested = False
        self.exit_requested = False
        self.exit_code = 0
        self.data_iterators: List[RerunDataIterator] = []
        self.stats = {"validations": 0, "rejections": 0, "max_rel_diff": 0.0}
        self.max_values: dict = {}
        self.large_value_counts: dict = {}

    # -- loop protocol ------------------------------------------------------------------
    def should_run_forward_backward(self, data_iterator) -> bool:
        its = data_iterator if isinstance(data_iterator, (list, tuple)) else [data_iterator]
        self.data_iterators = [it for it in its if isinstance(it, RerunDataIterator)]
        self.call_idx = 0
        if self.state == RerunState.NOT_RUNNING_YET:
            self.state = RerunState.INITIAL_RUN
            self.rerun_requested = False
            self.first_results = []
            return True
        if self.mode == RerunMode.DISABLED:
            self._finish_step()
            return False
        if self.state == RerunState.INITIAL_RUN and (self.rerun_requested or self.mode == RerunMode.REPORT_DETERMINISM_STATS):
            self.state = RerunState.RERUNNING_IN_PLACE
            for it in self.data_iterators:
                it.rewind()
            self.rerun_requested = False
            return True
        self._finish_step()
        return False

    def _finish_step(self):
        self.state = RerunState.NOT_RUNNING_YET
        self.first_results = []
        for it in self.data_iterators:
            it.advance()
        self.current_iteration += 1

    # -- validation -----------------------------------------------------------------------
    def is_unexpectedly_large(self, result, threshold: float, context: str, num_samples: int = 100, resample: bool = False) -> bool:
        """Rejection function for spiky losses / gradients (reference rerun_state_machine.py:655): the first ``num_samples``
        values of a ``context`` establish the largest magnitude seen, later values are flagged at ``threshold`` times that."""
        value = abs(float(result.item() if torch.is_tensor(result) else result))
        if math.isnan(value) or math.isinf(value):
            return False                                  # checked separately
        if resample or context not in self.large_value_counts:
            self.large_value_counts[context] = 0
        if self.large_value_counts[context] < num_samples:
            self.large_value_counts[context] += 1
            self.max_values[context] = max(self.max_values.get(context, 0.0), value)
            return False
        return value >= self.max_values[context] * threshold

    def validate_result(self, result, rejection_func: Callable[[Any], bool], message: str = "unexpected result",
                        tolerance: float = 0.0, fatal: bool = True) -> None:
        if self.mode == RerunMode.DISABLED:
            # reruns off: still validate, a rejected fatal result stops the job (reference rerun_state_machine.py:491-508)
            if rejection_func(float(result.item() if torch.is_tensor(result) else result)):
                self.stats["rejections"] += 1
                text = f"iteration {self.current_iteration}: unexpected result {float(result)} ({message})"
                if fatal:
                    raise RuntimeError(text)
                print(f"WARNING: {text}", flush=True)
            return
        value = float(result.item() if torch.is_tensor(result) else result)
        if self.injector.maybe_inject(self.current_iteration, self.state):
            value = float("nan")
        self.stats["validations"] += 1
        if self.state == RerunState.INITIAL_RUN:
            self.first_results.append(value)
            if self.mode == RerunMode.VALIDATE_RESULTS and rejection_func(value):
                self.stats["rejections"] += 1
                self.rerun_requested = True
            return
        # second execution of the same step on the same data
        first = self.first_results[self.call_idx] if self.call_idx < len(self.first_results) else value
        self.call_idx += 1
        same = (math.isnan(first) and math.isnan(value)) or abs(first - value) <= tolerance * max(1.0, abs(first))
        if first is not None and not (math.isnan(first) or math.isnan(value)) and first != 0:
            self.stats["max_rel_diff"] = max(self.stats["max_rel_diff"], abs(first - value) / abs(first))
        if self.mode == RerunMode.REPORT_DETERMINISM_STATS:
            return
        if not same:
            diag = RerunDiagnostic.TRANSIENT_ERROR
        elif rejection_func(value):
            diag = RerunDiagnostic.PERSISTENT_ERROR
        else:
            diag = RerunDiagnostic.CORRECT_RESULT
        self.diagnostics.append((self.current_iteration, diag.value, message, first, value))
        if diag == RerunDiagnostic.PERSISTENT_ERROR and fatal:
            # reproducible bad value on this GPU: checkpoint and resume elsewhere to disambiguate
            self.checkpoint_requested, self.exit_requested = True, True
            self.exit_code = EXIT_CODE_RESUME_TO_DISAMBIGUATE
        elif diag == RerunDiagnostic.TRANSIENT_ERROR and fatal:
            self.checkpoint_requested, self.exit_requested = False, True
            self.exit_code = EXIT_CODE_FAILED_ON_RESULT_VALIDATION

    def should_checkpoint_and_exit(self):
        return self.checkpoint_requested, self.exit_requested, self.exit_code

    def state_dict(self, data_iterator=None, ckpt_format="torch"):
        return {"mode": self.mode.value, "current_iteration": self.current_iteration, "diagnostics": self.diagnostics,
                "stats": self.stats}

    def load_state_dict(self, sd):
        if sd:
            self.current_iteration = sd.get("current_iteration", 0)
            self.diagnostics = sd.get("diagnostics", [])


_RSM: Optional[RerunStateMachine] = None


def initialize_rerun_state_machine(**kw) -> RerunStateMachine:
    global _RSM
    _RSM = RerunStateMachine(**kw)
    return _RSM


def get_rerun_state_machine() -> RerunStateMachine:
    global _RSM
    if _RSM is None:
        _RSM = RerunStateMachine()
    return _RSM


def destroy_rerun_state_machine():
    global _RSM
    _RSM = None


# -- bookkeeping types of the reference's ``rerun_state_machine.py`` ----------------------------------------------------------------------
from collections import namedtuple as _namedtuple  # noqa: E402
from enum import Enum as _Enum  # noqa: E402

Caller = _namedtuple("Caller", ["filename", "lineno", "rank"])       # where a validate_result call sits
Call = _namedtuple("Call", ["caller", "sequence"])                   # ... and which call of that site within the step


class RerunValidationStatus(str, _Enum):
    """Outcome recorded for a rejected result."""
    RERUN_DISABLED = "rerun_disabled"
    INITIAL_RUN = "initial_run"
    FIRST_RERUN_NOT_REPRODUCIBLE = "first_rerun_not_reproducible"
    FIRST_RERUN_REPRODUCIBLE = "first_rerun_reproducible"
    SECOND_RERUN_NOT_REPRODUCIBLE = "second_rerun_not_reproducible"
    SECOND_RERUN_REPRODUCIBLE = "second_rerun_reproducible"


class QuickStats:
    """Distribution of a statistic (relative differences between an execution and its re-execution in ``report_stats`` mode):
    up to ``max_size`` non-zero samples (a ring once full), the count of exact zeros, the maximum; ``print_stats`` -> one line."""

    def __init__(self, max_size: int = 100000) -> None:
        self.samples: List[float] = []
        self.pos = 0
        self.zero_cnt = 0
        self.max = 0.0
        self.max_size = max_size

    def record(self, data: float) -> None:
        if data == 0.0:
            self.zero_cnt += 1
            return
        if self.pos < self.max_size:
            self.samples.append(data)
        else:
            self.samples[self.pos % self.max_size] = data
        self.pos += 1
        self.max = max(self.max, data)

    def combine(self, others: List["QuickStats"]) -> None:
        if not others:
            return
        n = len(self.samples) + sum(len(o.samples) for o in others)
        if n <= self.max_size:
            for o in others:
                self.samples.extend(o.samples)
            self.pos = n
        self.zero_cnt += sum(o.zero_cnt for o in others)
        self.max = max(self.max, max(o.max for o in others))

    def reset(self) -> None:
        self.samples, self.pos, self.zero_cnt, self.max = [], 0, 0, 0.0

    def print_stats(self) -> str:
        z, n = self.zero_cnt, len(self.samples)
        if n == 0:
            return f"{z:,} samples, all identical"
        xs = sorted(self.samples)
        t = z + n

        def pct(p):
            i = int(t * p) - z
            return f"{xs[min(i, n - 1)]:.3E}" if i >= 0 else "0.0"

        return (f"{t:,}/{z:,} total/identical samples, rel. variability: avg= {sum(xs) / t:.3E}, p50= {pct(0.5)}, p90= {pct(0.9)}, "
                f"p99= {pct(0.99)}, p99.9= {pct(0.999)}, max: {self.max:.3E}")
