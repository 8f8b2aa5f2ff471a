#!/usr/bin/env python
"""Plain-text parallel corpora -> Megatron ``.bin/.idx`` (reference ``tools/preprocess_data_nmt.py``): every line of the
input file is one sentence = one document; no json, no sentence splitting.  Run once per side of the parallel corpus
(``--input train.en --output-prefix train_en`` / ``--input train.de --output-prefix train_de``): line i of both outputs are
translations of each other, which is what a T5 / encoder-decoder translation dataset indexes by position.

    python tools/preprocess_data_nmt.py --input train.en --output-prefix train_en --tokenizer-type SentencePieceTokenizer \\
        --tokenizer-model spm.model --workers 8
"""
import argparse
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


def _encode(line):
    ids = _TOK.tokenize(line.rstrip("\n"))
    return ids, len(line)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True, help="text file, one sentence per line")
    ap.add_argument("--tokenizer-type", required=True)
    ap.add_argument("--tokenizer-model", default=None)
    ap.add_argument("--vocab-file", default=None)
    ap.add_argument("--merge-file", default=None)
    ap.add_argument("--vocab-size", type=int, default=None)
    ap.add_argument("--vocab-extra-ids", type=int, default=0)
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--log-interval", type=int, default=100000)
    ap.add_argument("--keep-empty", action="store_true", help="keep empty lines as 1-token (eod) documents so both sides stay aligned")
    ap.add_argument("--make-vocab-size-divisible-by", type=int, default=128)
    ap.add_argument("--tensor-model-parallel-size", type=int, default=1)
    args = ap.parse_args(argv)
    args.rank, args.padded_vocab_size, args.mock_data = 1, None, False
    _init(args)
    builder = IndexedDatasetBuilder(f"{args.output_prefix}.bin", dtype=DType.optimal_dtype(_TOK.vocab_size))
    n = nbytes = 0
    with open(args.input, encoding="utf-8") as f:
        if args.workers > 1:
            pool = multiprocessing.Pool(args.workers, initializer=_init, initargs=(args,))
            results = pool.imap(_encode, f, 64)
        else:
            pool, results = None, map(_encode, f)
        for ids, ln in results:
            if not ids:
                if not args.keep_empty:
                    continue
                ids = [_TOK.eod]
            builder.add_document(np.asarray(ids), [len(ids)])
            n += 1
            nbytes += ln
            if n % args.log_interval == 0:
                print(f"processed {n} sentences ({nbytes / 1e6:.1f} MB)", file=sys.stderr)
        if pool is not None:
            pool.close()
    builder.finalize(f"{args.output_prefix}.idx")
    print(f"wrote {args.output_prefix}.bin/.idx: {n} sentences")
    return n


if __name__ == "__main__":
    main()
