#!/usr/bin/env python
"""(caption, image) pairs -> one multimodal ``.bin/.idx`` (reference ``tools/preprocess_mmdata.py``).

Every document holds two sequences: mode 0 = the caption token ids padded / truncated to ``--pad-length`` with
``eod``, mode 1 = the raw image file bytes viewed as int32 words (zero-padded to a multiple of 4, the pad count
stored as the first word so the reader can strip it).  Captions come from ``--input`` (jsonl, key ``text``), images
from ``--input-image`` in sorted file order (line *i* pairs with file *i*).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from megatronapp_b200.data.indexed_dataset import IndexedDatasetBuilder, get_bin_path, get_idx_path  # noqa: E402
from megatronapp_b200.data.tokenizer import build_tokenizer  # noqa: E402


def image_words(path: str) -> np.ndarray:
    raw = bytearray(open(path, "rb").read())
    pad = (4 - len(raw) % 4) % 4
    raw.extend(b"\0" * pad)
    return np.concatenate([[pad], np.frombuffer(bytes(raw), dtype=np.int32)]).astype(np.int32)


def words_to_image_bytes(words: np.ndarray) -> bytes:
    pad = int(words[0])
    b = np.asarray(words[1:], dtype=np.int32).tobytes()
    return b[: len(b) - pad] if pad else b


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True)
    ap.add_argument("--input-image", required=True)
    ap.add_argument("--pad-length", type=int, required=True)
    ap.add_argument("--json-key", default="text")
    ap.add_argument("--tokenizer-type", required=True)
    ap.add_argument("--tokenizer-model", default=None)
    ap.add_argument("--vocab-file", default=None)
    ap.add_argument("--merge-file", default=None)
    ap.add_argument("--vocab-size", type=int, default=None)
    ap.add_argument("--vocab-extra-ids", type=int, default=0)
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--log-interval", type=int, default=100)
    ap.add_argument("--make-vocab-size-divisible-by", type=int, default=128)
    ap.add_argument("--tensor-model-parallel-size", type=int, default=1)
    args = ap.parse_args(argv)
    args.rank, args.padded_vocab_size, args.mock_data = 1, None, False
    tok = build_tokenizer(args)
    images = sorted(os.path.join(args.input_image, f) for f in os.listdir(args.input_image))
    builder = IndexedDatasetBuilder(get_bin_path(args.output_prefix), dtype=np.int32, multimodal=True)
    n = 0
    with open(args.input, encoding="utf-8") as f:
        for line, img in zip((ln for ln in f if ln.strip()), images):
            ids = tok.tokenize(json.loads(line)[args.json_key])[: args.pad_length]
            ids = ids + [tok.eod] * (args.pad_length - len(ids))
            builder.add_item(np.asarray(ids, dtype=np.int32), 0)
            builder.add_item(image_words(img), 1)
            builder.end_document()
            n += 1
            if n % args.log_interval == 0:
                print(f"processed {n} documents", flush=True)
    builder.finalize(get_idx_path(args.output_prefix))
    print(f"wrote {n} (caption, image) documents to {args.output_prefix}.bin/.idx")


if __name__ == "__main__":
    main()
