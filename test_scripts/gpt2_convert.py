#!/usr/bin/env python
"""HF GPT-2 -> this framework (reference ``test_scripts/gpt2_convert.py``): thin front of ``tools/checkpoint/convert.py`` with
the GPT-2 loader, plus a logit check of the converted model against the HF forward pass when ``transformers`` can load the
source directory.

    python test_scripts/gpt2_convert.py --hf-dir gpt2 --save-dir ckpt/gpt2 [--tp 2] [--verify]
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hf-dir", required=True)
    ap.add_argument("--save-dir", required=True)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--verify", action="store_true")
    a = ap.parse_args()
    cmd = [sys.executable, os.path.join(ROOT, "tools", "checkpoint", "convert.py"), "--model-type", "GPT", "--loader", "hf_gpt2", "--saver", "core",
           "--load-dir", a.hf_dir, "--save-dir", a.save_dir, "--target-tensor-parallel-size", str(a.tp), "--target-pipeline-parallel-size", str(a.pp)]
    subprocess.check_call(cmd)
    if a.verify:
        sys.path.insert(0, ROOT)
        import torch
        from transformers import AutoModelForCausalLM

        from tools.checkpoint import convert as conv

        ck = conv.load_hf_gpt2(a.hf_dir)
        hf = AutoModelForCausalLM.from_pretrained(a.hf_dir).eval()
        ids = torch.randint(0, ck["config"]["vocab_size"], (1, 16))
        with torch.no_grad():
            ref = hf(ids).logits
        model = conv.build_core_model(ck) if hasattr(conv, "build_core_model") else None
        if model is None:
            print("verification needs tools.checkpoint.convert.build_core_model; skipped")
            return
        with torch.no_grad():
            got = model(ids, torch.arange(16).unsqueeze(0), None)
        print("max |logit difference| vs HF:", float((got.float() - ref.float()).abs().max()))


if __name__ == "__main__":
    main()
