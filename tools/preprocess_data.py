#!/usr/bin/env python
"""jsonl -> Megatron ``.bin/.idx`` (CLI parity with reference ``tools/preprocess_data.py``:
``--input --json-keys --tokenizer-type --vocab-file --merge-file --append-eod --output-prefix
--workers --split-sentences --keep-newlines --lang --partitions --keep-sequential-samples --log-interval``).

``--split-sentences`` uses nltk's punkt model for ``--lang`` when nltk is installed and a rule-based splitter otherwise
(sentence-final punctuation followed by white space and an opening character, with a short list of abbreviations that do not
end a sentence).  ``--partitions N`` cuts the input (one file, or a glob of files) into N parts that are encoded one after the
other by the worker pool and merged at the end -- bounded memory for very large corpora; parts are contiguous line ranges with
``--keep-sequential-samples`` and round-robin otherwise, as in the reference."""
import argparse
import glob
import json
import multiprocessing
import os
import re
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from megatronapp_b200.data.indexed_dataset import DType, IndexedDatasetBuilder  # noqa: E402
from megatronapp_b200.data.tokenizer import build_tokenizer  # noqa: E402

_TOK = None
_SPLIT = None
_ABBREV = {"mr", "mrs", "ms", "dr", "prof", "sr", "jr", "st", "vs", "etc", "inc", "ltd", "co", "corp", "no", "e.g", "i.e", "u.s",
           "jan", "feb", "mar", "apr", "jun", "jul", "aug", "sep", "sept", "oct", "nov", "dec", "fig", "gen", "col", "lt", "mt"}
_BOUNDARY = re.compile(r"""([.!?]+["')\]]*)(\s+)(?=["'(\[]?[A-Z0-9])""")


def rule_based_sentences(text: str, keep_newlines: bool = False):
    """Split after ., ! or ? (plus closing quotes / brackets) when white space and an upper-case letter, digit or opening quote
    follow, unless the word before the full stop is a known abbreviation or a single initial.  With ``keep_newlines`` the white
    space between sentences stays attached to the sentence it follows (newlines survive tokenisation), otherwise newlines are
    treated as spaces."""
    if not keep_newlines:
        text = text.replace("\n", " ")
    out, start = [], 0
    for m in _BOUNDARY.finditer(text):
        head = text[start:m.start()]
        last = head.rsplit(None, 1)[-1].lower().rstrip(".") if head.strip() else ""
        if m.group(1).startswith(".") and (last in _ABBREV or (len(last) == 1 and last.isalpha())):
            continue
        end = m.end() if keep_newlines else m.end(1)
        out.append(text[start:end])
        start = m.end()
    if text[start:].strip():
        out.append(text[start:])
    return [x for x in out if x.strip()]


def _init(args):
    global _TOK, _SPLIT
    _TOK = build_tokenizer(args)
    _SPLIT = None
    if getattr(args, "split_sentences", False):
        keep = bool(getattr(args, "keep_newlines", False))
        try:
            import nltk

            punkt = nltk.data.load(f"tokenizers/punkt/{getattr(args, 'lang', 'english')}.pickle")
            _SPLIT = lambda t: punkt.tokenize(t if keep else t.replace("\n", " "))  # noqa: E731
        except Exception:  # noqa: BLE001  (no nltk / no punkt model in this environment)
            _SPLIT = lambda t: rule_based_sentences(t, keep)  # noqa: E731


def _encode(line_and_args):
    line, keys, append_eod, split_sentences = line_and_args
    data = json.loads(line)
    out = {}
    for k in keys:
        text = data[k]
        sents = _SPLIT(text) if split_sentences else [text]
        ids = [_TOK.tokenize(s) for s in sents]
        ids = [i for i in ids if i]
        if ids and append_eod:
            ids[-1].append(_TOK.eod)
        out[k] = ids
    return out, len(line)


def _parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True, help="jsonl file (or a glob of files with --partitions > 1)")
    ap.add_argument("--json-keys", nargs="+", default=["text"])
    ap.add_argument("--split-sentences", action="store_true")
    ap.add_argument("--keep-newlines", action="store_true", help="keep newlines between sentences when splitting")
    ap.add_argument("--lang", default="english", help="language of the nltk punkt model (when nltk is installed)")
    ap.add_argument("--tokenizer-type", required=True)
    ap.add_argument("--tokenizer-model", default=None)
    ap.add_argument("--vocab-file", default=None)
    ap.add_argument("--merge-file", default=None)
    ap.add_argument("--vocab-size", type=int, default=None)
    ap.add_argument("--vocab-extra-ids", type=int, default=0)
    ap.add_argument("--append-eod", action="store_true")
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--partitions", type=int, default=1, help="encode the input in this many parts and merge them")
    ap.add_argument("--keep-sequential-samples", action="store_true",
                    help="partitions are contiguous line ranges (document order is kept) instead of round-robin")
    ap.add_argument("--log-interval", type=int, default=0)
    ap.add_argument("--make-vocab-size-divisible-by", type=int, default=128)
    ap.add_argument("--tensor-model-parallel-size", type=int, default=1)
    args = ap.parse_args()
    args.rank, args.padded_vocab_size, args.mock_data = 1, None, False
    return args


def _encode_lines(args, lines, prefix, level, dtype, pool):
    builders = {k: IndexedDatasetBuilder(f"{prefix}_{k}_{level}.bin", dtype=dtype) for k in args.json_keys}
    work = ((line, args.json_keys, args.append_eod, args.split_sentences) for line in lines if line.strip())
    results = pool.imap(_encode, work, 32) if pool is not None else map(_encode, work)
    _consume(results, builders, args.log_interval)
    for k, b in builders.items():
        b.finalize(f"{prefix}_{k}_{level}.idx")


def _partition_lines(files, n_parts, part, sequential):
    if sequential:
        total = 0
        for fn in files:
            with open(fn, encoding="utf-8") as f:
                total += sum(1 for _ in f)
        lo, hi = part * total // n_parts, (part + 1) * total // n_parts
    i = 0
    for fn in files:
        with open(fn, encoding="utf-8") as f:
            for line in f:
                if (lo <= i < hi) if sequential else (i % n_parts == part):
                    yield line
                i += 1


def main():
    args = _parse()
    _init(args)
    dtype = DType.optimal_dtype(_TOK.vocab_size)
    level = "sentence" if args.split_sentences else "document"
    files = sorted(glob.glob(args.input)) or [args.input]
    pool = multiprocessing.Pool(args.workers, initializer=_init, initargs=(args,)) if args.workers > 1 else None
    try:
        if args.partitions <= 1:
            if len(files) != 1:
                raise SystemExit("--input matches several files: use --partitions > 1")
            with open(files[0], encoding="utf-8") as f:
                _encode_lines(args, f, args.output_prefix, level, dtype, pool)
        else:
            parts = [f"{args.output_prefix}_part{p:03d}" for p in range(args.partitions)]
            for p, prefix in enumerate(parts):
                _encode_lines(args, _partition_lines(files, args.partitions, p, args.keep_sequential_samples), prefix, level, dtype, pool)
            for k in args.json_keys:                      # merge the parts in order, then drop them
                merged = IndexedDatasetBuilder(f"{args.output_prefix}_{k}_{level}.bin", dtype=dtype)
                for prefix in parts:
                    merged.add_index(f"{prefix}_{k}_{level}")
                merged.finalize(f"{args.output_prefix}_{k}_{level}.idx")
                for prefix in parts:
                    for ext in (".bin", ".idx"):
                        os.remove(f"{prefix}_{k}_{level}{ext}")
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    print(f"wrote {args.output_prefix}_<key>_{level}.bin/.idx")


def _consume(results, builders, log_interval=0):
    t0, nbytes = time.time(), 0
    for n, (doc, size) in enumerate(results, 1):
        nbytes += size
        if log_interval and n % log_interval == 0:
            dt = max(time.time() - t0, 1e-9)
            print(f"Processed {n} documents ({n / dt:.1f} docs/s, {nbytes / dt / 1e6:.2f} MB/s).", file=sys.stderr)
        for k, sents in doc.items():
            if not sents:
                continue
            flat = np.concatenate([np.array(s) for s in sents])
            builders[k].add_document(flat, [len(s) for s in sents])


if __name__ == "__main__":
    main()
