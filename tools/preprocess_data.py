#!/usr/bin/env python
"""jsonl -> Megatron ``.bin/.idx`` (CLI parity with reference ``tools/preprocess_data.py``:
``--input --json-keys --tokenizer-type --vocab-file --merge-file --append-eod --output-prefix
--workers --split-sentences``)."""
import argparse
import json
import multiprocessing
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from megatronapp_b200.data.indexed_dataset import DType, IndexedDatasetBuilder  # noqa: E402
from megatronapp_b200.data.tokenizer import build_tokenizer  # noqa: E402

_TOK = None


def _init(args):
    global _TOK
    _TOK = build_tokenizer(args)


def _encode(line_and_args):
    line, keys, append_eod, split_sentences = line_and_args
    data = json.loads(line)
    out = {}
    for k in keys:
        text = data[k]
        sents = [s for s in text.replace("\n", " ").split(". ") if s] if split_sentences else [text]
        ids = [_TOK.tokenize(s) for s in sents]
        ids = [i for i in ids if i]
        if ids and append_eod:
            ids[-1].append(_TOK.eod)
        out[k] = ids
    return out, len(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True)
    ap.add_argument("--json-keys", nargs="+", default=["text"])
    ap.add_argument("--split-sentences", action="store_true")
    ap.add_argument("--tokenizer-type", required=True)
    ap.add_argument("--tokenizer-model", default=None)
    ap.add_argument("--vocab-file", default=None)
    ap.add_argument("--merge-file", default=None)
    ap.add_argument("--vocab-size", type=int, default=None)
    ap.add_argument("--vocab-extra-ids", type=int, default=0)
    ap.add_argument("--append-eod", action="store_true")
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--make-vocab-size-divisible-by", type=int, default=128)
    ap.add_argument("--tensor-model-parallel-size", type=int, default=1)
    args = ap.parse_args()
    args.rank, args.padded_vocab_size, args.mock_data = 1, None, False
    _init(args)
    dtype = DType.optimal_dtype(_TOK.vocab_size)
    level = "sentence" if args.split_sentences else "document"
    builders = {k: IndexedDatasetBuilder(f"{args.output_prefix}_{k}_{level}.bin", dtype=dtype) for k in args.json_keys}
    with open(args.input, encoding="utf-8") as f:
        work = ((line, args.json_keys, args.append_eod, args.split_sentences) for line in f if line.strip())
        if args.workers > 1:
            with multiprocessing.Pool(args.workers, initializer=_init, initargs=(args,)) as pool:
                results = pool.imap(_encode, work, 32)
                _consume(results, builders)
        else:
            _consume(map(_encode, work), builders)
    for k, b in builders.items():
        b.finalize(f"{args.output_prefix}_{k}_{level}.idx")
    print(f"wrote {args.output_prefix}_<key>_{level}.bin/.idx")


def _consume(results, builders):
    for doc, _ in results:
        for k, sents in doc.items():
            if not sents:
                continue
            flat = np.concatenate([np.array(s) for s in sents])
            builders[k].add_document(flat, [len(s) for s in sents])


if __name__ == "__main__":
    main()
