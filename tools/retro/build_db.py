"""Retro retrieval preprocessing without faiss: chunk a tokenised corpus, embed the chunks, and find each training
chunk's nearest-neighbour chunks (+ their continuations) by exact batched inner-product search.

Parity: reference ``tools/retro/`` + ``core/datasets/retro`` (chunk database ``db/``, index ``index/`` (faiss
IVF-PQ there, exact search here -- sufficient up to a few million chunks on one GPU / CPU node), neighbour query
``query/`` excluding chunks from the same document, output consumed by ``pretrain_retro.py`` as ``neighbor_tokens``).

    python tools/retro/build_db.py --data-prefix corpus_text_document --out retro_work --chunk-length 64 \
        --num-neighbors 2 --retrieved-length 128
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def build_chunk_db(indexed, chunk_length: int):
    """Returns (chunk_doc_ids [n], chunk_starts [n] as (doc seq index, offset)) over all documents."""
    doc_ids, seq_ids, offsets = [], [], []
    docs = np.asarray(indexed.document_indices)
    for d in range(len(docs) - 1):
        for s in range(int(docs[d]), int(docs[d + 1])):
            n = int(indexed.sequence_lengths[s])
            for off in range(0, max(n - 1, 1), chunk_length):
                doc_ids.append(d); seq_ids.append(s); offsets.append(off)
    return np.asarray(doc_ids), np.asarray(seq_ids), np.asarray(offsets)


def chunk_tokens(indexed, seq_id: int, offset: int, length: int, pad: int = 0) -> np.ndarray:
    t = np.asarray(indexed.get(int(seq_id), int(offset), min(length, int(indexed.sequence_lengths[seq_id]) - int(offset))))
    if len(t) < length:
        t = np.concatenate([t, np.full(length - len(t), pad, dtype=t.dtype)])
    return t


def embed_chunks(tokens: np.ndarray, vocab_size: int, dim: int = 256, seed: int = 0, embedder=None) -> torch.Tensor:
    """Chunk embeddings: a supplied encoder (e.g. BERT mean-pool via ``tools/bert_embedding``) or, by default, a
    fixed random-projection bag-of-tokens (hashing) embedding -- deterministic and corpus-independent."""
    if embedder is not None:
        return torch.nn.functional.normalize(embedder(torch.from_numpy(tokens.astype(np.int64))).float(), dim=-1)
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(vocab_size, dim, generator=g)
    e = table[torch.from_numpy(tokens.astype(np.int64)).clamp_(0, vocab_size - 1)].mean(dim=1)
    return torch.nn.functional.normalize(e, dim=-1)


def query_neighbors(emb: torch.Tensor, doc_ids: np.ndarray, k: int, block: int = 4096) -> np.ndarray:
    """Exact top-k by inner product, never returning a chunk of the querying chunk's own document."""
    n = emb.shape[0]
    d = torch.from_numpy(doc_ids)
    out = np.zeros((n, k), dtype=np.int64)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    E = emb.to(dev)
    for i in range(0, n, block):
        sim = E[i:i + block] @ E.t()
        same = d[i:i + block].unsqueeze(1).to(dev) == d.unsqueeze(0).to(dev)
        sim.masked_fill_(same, -2.0)
        out[i:i + block] = sim.topk(min(k, n), dim=-1).indices.cpu().numpy()[:, :k]
    return out


def main(argv=None):
    from megatronapp_b200.data.indexed_dataset import IndexedDataset

    ap = argparse.ArgumentParser()
    ap.add_argument("--data-prefix", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--chunk-length", type=int, default=64)
    ap.add_argument("--num-neighbors", type=int, default=2)
    ap.add_argument("--retrieved-length", type=int, default=128)
    ap.add_argument("--vocab-size", type=int, default=50304)
    a = ap.parse_args(argv)
    ds = IndexedDataset(a.data_prefix)
    doc_ids, seq_ids, offs = build_chunk_db(ds, a.chunk_length)
    toks = np.stack([chunk_tokens(ds, s, o, a.chunk_length) for s, o in zip(seq_ids, offs)])
    emb = embed_chunks(toks, a.vocab_size)
    nbr = query_neighbors(emb, doc_ids, a.num_neighbors)
    # retrieved text = neighbour chunk + its continuation (retrieved_length tokens from the neighbour's start)
    retrieved = np.stack([[chunk_tokens(ds, seq_ids[j], offs[j], a.retrieved_length) for j in row] for row in nbr])
    os.makedirs(a.out, exist_ok=True)
    np.save(os.path.join(a.out, "chunk_doc_ids.npy"), doc_ids)
    np.save(os.path.join(a.out, "chunk_seq_ids.npy"), seq_ids)
    np.save(os.path.join(a.out, "chunk_offsets.npy"), offs)
    np.save(os.path.join(a.out, "neighbor_chunk_ids.npy"), nbr)
    np.save(os.path.join(a.out, "neighbor_tokens.npy"), retrieved)
    print(f"retro db: {len(doc_ids)} chunks, {a.num_neighbors} neighbours x {a.retrieved_length} tokens -> {a.out}")
    return nbr, retrieved


if __name__ == "__main__":
    main()
