"""Retro preprocessing driver (reference ``tools/retro/preprocess_data.py`` with ``--retro-tasks``):

    db-build         chunk database over the training blend                     -> <project>/db/
    index-train      train the ANN index on a sample of chunk embeddings        -> <project>/index/
    index-add        embed and add every chunk                                  -> <project>/index/
    query-neighbors  neighbours of every chunk of every pre-training sample     -> <project>/query/
    build            all of the above

    python tools/retro/preprocess_data.py --retro-project-dir work --retro-tasks build --data-path corpus_text_document \\
        --seq-length 2048 --retro-gpt-chunk-length 64 --train-samples 100000 --retro-index-str IVF4096,PQ32

then ``pretrain_retro.py --retro-project-dir work --data-path corpus_text_document ...`` (same data / split / seed /
sequence length) trains on the retrieved neighbours."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from megatronapp_b200.data.retro import ChunkDB, RetroPreprocessingConfig, build_db, index_factory, query_neighbors  # noqa: E402
from megatronapp_b200.data.retro.index import load_index, save_index  # noqa: E402
from megatronapp_b200.data.retro.query import build_train_dataset, get_embedder  # noqa: E402

TASKS = ("db-build", "index-train", "index-add", "query-neighbors")


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--retro-project-dir", required=True)
    ap.add_argument("--retro-tasks", default="build")
    ap.add_argument("--data-path", nargs="*", default=None)
    ap.add_argument("--split", default="969,30,1")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--seq-length", type=int, default=2048)
    ap.add_argument("--train-samples", type=int, default=0)
    ap.add_argument("--vocab-size", type=int, default=50304)
    ap.add_argument("--retro-gpt-chunk-length", type=int, default=64)
    ap.add_argument("--retro-index-str", default="IVF256,PQ16")
    ap.add_argument("--retro-index-ntrain", type=int, default=100_000)
    ap.add_argument("--retro-query-nprobe", type=int, default=16)
    ap.add_argument("--retro-query-num-neighbors-query", type=int, default=20)
    ap.add_argument("--retro-query-num-neighbors-save", type=int, default=2)
    ap.add_argument("--retro-retrieved-length", type=int, default=128)
    ap.add_argument("--retro-block-size", type=int, default=10_000)
    ap.add_argument("--retro-embed-dim", type=int, default=256)
    return ap.parse_args(argv)


def main(argv=None):
    a = parse(argv)
    tasks = list(TASKS) if a.retro_tasks == "build" else a.retro_tasks.split(",")
    cfg_path = os.path.join(a.retro_project_dir, "config.json")
    if "db-build" in tasks or not os.path.exists(cfg_path):
        if not a.data_path:
            raise SystemExit("--data-path is required to start a project")
        cfg = RetroPreprocessingConfig(
            retro_project_dir=a.retro_project_dir, retro_gpt_data_path=list(a.data_path), retro_gpt_seq_length=a.seq_length,
            retro_gpt_chunk_length=a.retro_gpt_chunk_length, retro_gpt_split=a.split, retro_gpt_seed=a.seed,
            retro_gpt_train_samples=a.train_samples, retro_vocab_size=a.vocab_size, retro_index_str=a.retro_index_str,
            retro_index_ntrain=a.retro_index_ntrain, retro_query_nprobe=a.retro_query_nprobe,
            retro_query_num_neighbors_query=a.retro_query_num_neighbors_query,
            retro_query_num_neighbors_save=a.retro_query_num_neighbors_save, retro_retrieved_length=a.retro_retrieved_length,
            retro_block_size=a.retro_block_size, retro_embed_dim=a.retro_embed_dim)
        cfg.save()
    else:
        cfg = RetroPreprocessingConfig.load(a.retro_project_dir)
    embed = get_embedder(cfg)
    index_path = cfg.path("index", "index.npz")
    for task in tasks:
        t0 = time.time()
        if task == "db-build":
            db = build_db(cfg)
            print(f"[retro] db-build: {len(db)} chunks of {cfg.retro_gpt_chunk_length} tokens ({time.time() - t0:.1f}s)")
        elif task == "index-train":
            db = ChunkDB(cfg)
            ix = index_factory(cfg.retro_index_str, cfg.retro_embed_dim, seed=cfg.retro_gpt_seed)
            ix.train(embed(db.batch_tokens(db.sampled)))
            os.makedirs(cfg.path("index"), exist_ok=True)
            save_index(ix, index_path)
            print(f"[retro] index-train: {cfg.retro_index_str} on {len(db.sampled)} chunks ({time.time() - t0:.1f}s)")
        elif task == "index-add":
            db = ChunkDB(cfg)
            ix = load_index(index_path)
            for b0 in range(ix.ntotal, len(db), cfg.retro_doc_block_size):
                ix.add(embed(db.batch_tokens(range(b0, min(len(db), b0 + cfg.retro_doc_block_size)))))
            save_index(ix, index_path)
            print(f"[retro] index-add: {ix.ntotal} vectors ({time.time() - t0:.1f}s)")
        elif task == "query-neighbors":
            db = ChunkDB(cfg)
            ix = load_index(index_path)
            ds = build_train_dataset(cfg)
            nb = query_neighbors(cfg, db, ix, ds, "train", embed)
            found = float((np.asarray(nb) >= 0).mean())
            print(f"[retro] query-neighbors: {nb.shape[0]} samples x {nb.shape[1]} chunks x {nb.shape[2]} neighbours, "
                  f"{100 * found:.1f}% filled ({time.time() - t0:.1f}s)")
        else:
            raise SystemExit(f"unknown task {task!r}; choose from {TASKS} or 'build'")
    return cfg


if __name__ == "__main__":
    main()
