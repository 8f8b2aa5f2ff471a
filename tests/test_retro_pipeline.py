"""Retro preprocessing: chunk DB, IVF-PQ index, neighbour query, RetroDataset, training from a project directory."""
import json
import os
import re
import socket
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from megatronapp_b200.data.retro.index import FlatIndex, IVFPQIndex, index_factory, load_index, save_index  # noqa: E402


def _clustered(n, dim, n_centers=32, seed=0):
    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(n_centers, dim, generator=g)
    x = centers[torch.randint(0, n_centers, (n,), generator=g)] + 0.35 * torch.randn(n, dim, generator=g)
    return torch.nn.functional.normalize(x, dim=-1)


def test_ivfpq_recall_against_exact_search_and_persistence(tmp_path):
    x, q = _clustered(6000, 64), _clustered(200, 64, seed=1)
    flat = FlatIndex(64, "cpu")
    flat.add(x)
    _, truth = flat.search(q, 10)
    ivf = index_factory("OPQ16_64,IVF64_HNSW8,PQ16", 64, "cpu")
    assert isinstance(ivf, IVFPQIndex) and ivf.m == 16
    ivf.train(x[:3000])
    ivf.add(x[:2500])
    ivf.add(x[2500:])                                              # incremental adds land in the same inverted lists
    assert ivf.ntotal == 6000
    _, got = ivf.search(q, 10, nprobe=16)
    recall = np.mean([len(set(got[i]) & set(truth[i])) / 10 for i in range(len(q))])
    r1 = np.mean([truth[i][0] in got[i] for i in range(len(q))])
    assert recall > 0.6 and r1 > 0.85, (recall, r1)
    _, narrow = ivf.search(q, 10, nprobe=1)
    assert np.mean([len(set(narrow[i]) & set(truth[i])) / 10 for i in range(len(q))]) <= recall + 1e-9   # more probes never hurt
    exact_lists = index_factory("IVF32,Flat", 64, "cpu")
    exact_lists.train(x)
    exact_lists.add(x)
    _, e = exact_lists.search(q, 10, nprobe=32)                    # every list probed + exact residuals == brute force
    assert (np.sort(e, 1) == np.sort(truth, 1)).mean() > 0.999
    save_index(ivf, str(tmp_path / "ix.npz"))
    again = load_index(str(tmp_path / "ix.npz"), "cpu")
    _, got2 = again.search(q, 10, nprobe=16)
    assert np.array_equal(got, got2)


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_retro_project_from_sample_corpus_to_training(tmp_path):
    """shipped corpus -> preprocess_data.py -> tools/retro/preprocess_data.py (db, index, query) -> pretrain_retro.py"""
    small = tmp_path / "c.jsonl"
    with open(os.path.join(ROOT, "datasets_gpt", "dataset.json")) as f, open(small, "w") as g:
        for _, line in zip(range(120), f):
            g.write(line)
    tok = ["--tokenizer-type", "GPT2BPETokenizer", "--vocab-file", os.path.join(ROOT, "datasets_gpt", "vocab.json"),
           "--merge-file", os.path.join(ROOT, "datasets_gpt", "merges.txt")]
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "preprocess_data.py"), "--input", str(small), "--output-prefix",
                        str(tmp_path / "c"), "--workers", "1", "--append-eod", *tok], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    proj = str(tmp_path / "retro")
    from tools.retro import preprocess_data as rp

    cfg = rp.main(["--retro-project-dir", proj, "--data-path", str(tmp_path / "c_text_document"), "--seq-length", "64",
                   "--retro-gpt-chunk-length", "16", "--retro-retrieved-length", "32", "--train-samples", "24", "--vocab-size", "2048",
                   "--retro-index-str", "IVF8,PQ8", "--retro-embed-dim", "64", "--retro-query-nprobe", "8", "--split", "100,0,0",
                   "--retro-block-size", "10"])
    from megatronapp_b200.data.retro import ChunkDB, RetroDataset
    from megatronapp_b200.data.retro.query import _sample_doc_keys, build_train_dataset

    db = ChunkDB(cfg)
    assert len(db) > 200 and db.tokens(0).shape == (16,) and db.tokens(0, 32).shape == (32,)
    assert np.array_equal(db.tokens(0, 32)[:16], db.tokens(0))                       # continuation extends the chunk
    ds = build_train_dataset(cfg)
    rds = RetroDataset(cfg, ds, db)
    s = rds[3]
    assert s["neighbor_tokens"].shape == (4, 2, 32) and s["tokens"].shape == (64,)
    nb = np.load(os.path.join(proj, "query", "train_neighbors.npy"))
    assert nb.shape == (len(ds), 4, 2) and (nb >= 0).mean() > 0.95
    for i in range(len(ds)):                                                         # never retrieve from the sample's own documents
        own = _sample_doc_keys(ds, i)
        assert all(db.doc_key(int(j)) not in own for j in nb[i].ravel() if j >= 0)
    assert json.load(open(os.path.join(proj, "query", "train_progress.json")))["done"] == len(ds)
    # train on it through the CLI
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    args = ("--num-layers 2 --hidden-size 32 --num-attention-heads 4 --seq-length 64 --max-position-embeddings 64 --micro-batch-size 2 "
            "--global-batch-size 4 --train-iters 3 --lr 1e-3 --log-interval 1 --eval-iters 1 --eval-interval 100 --distributed-backend gloo "
            "--no-gradient-accumulation-fusion --retro-chunk-length 16 --retro-retrieved-length 32 --retro-num-neighbors 2 --split 100,0,0").split()
    p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_retro.py"), *args, *tok, "--data-path", str(tmp_path / "c_text_document"),
                        "--retro-project-dir", proj], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert re.search(r"iteration\s+3/\s+3 \|.*lm loss: [0-9.E+\-]+", out)
