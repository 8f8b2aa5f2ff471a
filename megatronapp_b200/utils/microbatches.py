"""Number-of-microbatches calculators: constant, and global-batch ramp-up.

Parity: reference ``core/num_microbatches_calculator.py`` (``ConstantNumMicroBatchesCalculator``,
``RampupBatchsizeNumMicroBatchesCalculator``, module-level get/update/init functions).
"""
from __future__ import annotations

from typing import List, Optional

_CALC = None


class ConstantNumMicroBatchesCalculator:
    def __init__(self, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed=False):
        mbs_dp = micro_batch_size * data_parallel_size
        if decrease_batch_size_if_needed and global_batch_size % mbs_dp != 0:
            global_batch_size = (global_batch_size // mbs_dp) * mbs_dp
        if global_batch_size % mbs_dp != 0:
            raise ValueError(f"global batch size ({global_batch_size}) is not divisible by micro batch size "
                             f"({micro_batch_size}) times data parallel size ({data_parallel_size})")
        self.num_micro_batches = global_batch_size // mbs_dp
        self.current_global_batch_size = global_batch_size
        self.current_running_global_batch_size = global_batch_size
        self.micro_batch_size = micro_batch_size

    def get(self):
        return self.num_micro_batches

    def get_current_global_batch_size(self):
        return self.current_global_batch_size

    def get_micro_batch_size(self):
        return self.micro_batch_size

    def get_current_running_global_batch_size(self):
        return self.current_running_global_batch_size

    def update(self, consumed_samples, consistency_check, verbose=False):
        pass


class RampupBatchsizeNumMicroBatchesCalculator(ConstantNumMicroBatchesCalculator):
    def __init__(self, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed,
                 start_global_batch_size, batch_size_increment, ramup_samples):
        self.micro_batch_size = micro_batch_size
        self.data_parallel_size = data_parallel_size
        self.decrease = decrease_batch_size_if_needed
        self.global_batch_size = global_batch_size
        self.start = start_global_batch_size
        self.incr = batch_size_increment
        self.ramup_samples = ramup_samples
        if start_global_batch_size <= 0 or batch_size_increment <= 0 or ramup_samples < 0:
            raise ValueError("bad ramp-up parameters")
        diff = global_batch_size - start_global_batch_size
        if diff < 0 or diff % batch_size_increment != 0:
            raise ValueError("global batch size - start batch size must be a non-negative multiple of the increment")
        n_inc = diff // batch_size_increment
        self.rampup_samples_per_increment = ramup_samples / n_inc if n_inc else 0
        self.update(0, False)

    def update(self, consumed_samples, consistency_check, verbose=False):
        if consumed_samples > self.ramup_samples or self.rampup_samples_per_increment == 0:
            cur = self.global_batch_size
        else:
            steps = int(consumed_samples / self.rampup_samples_per_increment)
            cur = min(self.global_batch_size, self.start + steps * self.incr)
        self.current_global_batch_size = cur
        mbs_dp = self.micro_batch_size * self.data_parallel_size
        if cur % mbs_dp != 0:
            if consistency_check and not self.decrease:
                raise ValueError(f"current global batch size ({cur}) is not divisible by mbs*dp ({mbs_dp})")
            self.current_running_global_batch_size = (cur // mbs_dp) * mbs_dp
        else:
            self.current_running_global_batch_size = cur
        self.num_micro_batches = self.current_running_global_batch_size // mbs_dp


def init_num_microbatches_calculator(rank, rampup_batch_size, global_batch_size, micro_batch_size,
                                     data_parallel_size, decrease_batch_size_if_needed=False):
    global _CALC
    if rampup_batch_size is None:
        _CALC = ConstantNumMicroBatchesCalculator(global_batch_size, micro_batch_size, data_parallel_size,
                                                  decrease_batch_size_if_needed)
    else:
        if len(rampup_batch_size) != 3:
            raise ValueError("--rampup-batch-size takes <start> <increment> <ramp-up samples>")
        _CALC = RampupBatchsizeNumMicroBatchesCalculator(global_batch_size, micro_batch_size, data_parallel_size,
                                                         decrease_batch_size_if_needed, int(rampup_batch_size[0]),
                                                         int(rampup_batch_size[1]), int(rampup_batch_size[2]))
    return _CALC


def destroy_num_microbatches_calculator():
    global _CALC
    _CALC = None


def get_num_microbatches():
    return _CALC.get()


def get_current_global_batch_size():
    return _CALC.get_current_global_batch_size()


def get_micro_batch_size():
    return _CALC.get_micro_batch_size()


def get_current_running_global_batch_size():
    return _CALC.get_current_running_global_batch_size()


def update_num_microbatches(consumed_samples, consistency_check=True, verbose=False):
    _CALC.update(consumed_samples, consistency_check, verbose)


NumMicroBatchesCalculator = ConstantNumMicroBatchesCalculator       # base type name of the reference


def unset_num_microbatches_calculator():
    destroy_num_microbatches_calculator()


def reconfigure_num_microbatches_calculator(rank, rampup_batch_size, global_batch_size, micro_batch_size, data_parallel_size,
                                            decrease_batch_size_if_needed=False):
    """Replace the calculator (e.g. evaluation with another batch size) without tripping the "already initialised" guard."""
    return init_num_microbatches_calculator(rank, rampup_batch_size, global_batch_size, micro_batch_size, data_parallel_size,
                                            decrease_batch_size_if_needed)
