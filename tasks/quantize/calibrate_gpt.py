#!/usr/bin/env python
"""Calibrate a GPT checkpoint for post-training quantisation (reference ``tasks/quantize/calibrate_gpt.py``): runs the
calibration batches through the model with quantizers inserted, records amax / pre-quant scales, and saves the folded
weights + quantizer state.  Same engine and flags as ``examples/post_training/quantize.py`` (``--export-quant-cfg``,
``--calib-size``, ``--calib-batch-size``, ``--export-real-weights``); kept under ``tasks/`` for users of the reference layout.

    torchrun --nproc-per-node 1 tasks/quantize/calibrate_gpt.py <model args> --load ckpt --save ckpt_fp8 --export-quant-cfg fp8
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    runpy.run_path(os.path.join(ROOT, "examples", "post_training", "quantize.py"), run_name="__main__")
