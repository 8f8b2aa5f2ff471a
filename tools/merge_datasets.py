"""Merge several ``.bin/.idx`` indexed datasets into one (reference ``tools/merge_datasets.py``)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatronapp_b200.data.indexed_dataset import (IndexedDataset, IndexedDatasetBuilder, get_bin_path,  # noqa: E402
                                                   get_idx_path)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True, help="directory with the prefixes to merge")
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--multimodal", action="store_true", help="the datasets carry a per-sequence mode (text / image) column")
    a = ap.parse_args()
    prefixes = sorted({os.path.join(a.input, f[:-4]) for f in os.listdir(a.input) if f.endswith(".idx")})
    assert prefixes, "no .idx files found"
    first = IndexedDataset(prefixes[0], multimodal=a.multimodal)
    b = IndexedDatasetBuilder(get_bin_path(a.output_prefix), dtype=first.index.dtype, multimodal=a.multimodal)
    for p in prefixes:
        assert os.path.exists(get_bin_path(p)), f"missing {get_bin_path(p)}"
        b.add_index(p)
    b.finalize(get_idx_path(a.output_prefix))
    print(f"merged {len(prefixes)} datasets into {a.output_prefix}")
