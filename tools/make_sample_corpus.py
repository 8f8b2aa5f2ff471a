"""Generate the small sample corpora shipped in ``datasets_gpt/`` and ``datasets_bert/`` (the role of the reference's
``datasets_gpt/{dataset.json,vocab.json,merges.txt}`` and ``datasets_bert/{dataset.json,dataset_ss.json,vocab.txt}``).

The text is synthetic -- short financial-newswire style briefs produced from a seeded grammar -- so the repository carries
no third-party data; the byte-level BPE vocabulary / merges (GPT-2 file format) and the WordPiece vocabulary are *trained on
that text* by the two small trainers below, so ``tools/preprocess_data.py`` + ``pretrain_gpt.py`` / ``pretrain_bert.py`` run
end to end on files in this tree.

    python tools/make_sample_corpus.py --docs 1500 --bpe-merges 1500 --wordpiece-size 2000
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import random
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COMPANIES = ["Northwind Holdings", "Alder Creek Energy", "Blue Harbor Logistics", "Cobalt Ridge Mining", "Delta Quill Software",
             "Evergreen Foods", "Falcon Peak Aerospace", "Granite Bay Bancorp", "Helix Orchard Pharma", "Ironwood Rail",
             "Juniper Lane Retail", "Keystone Tidal Power", "Lumen Field Semiconductors", "Marigold Insurance Group",
             "Nimbus Air Freight", "Orchid Valley Textiles", "Pinecrest Robotics", "Quarry Hill Cement", "Redwater Telecom",
             "Silverline Motors", "Tamarack Paper", "Umber Coast Shipping", "Vireo Medical Devices", "Westmark Hotels",
             "Yarrow Street Media", "Zenith Loop Networks"]
CITIES = ["Lisbon", "Osaka", "Toronto", "Nairobi", "Santiago", "Helsinki", "Mumbai", "Denver", "Leeds", "Perth", "Oslo", "Austin"]
MONTHS = ["Jan", "Feb", "March", "April", "May", "June", "July", "Aug", "Sept", "Oct", "Nov", "Dec"]
VERBS = ["reports", "posts", "forecasts", "confirms", "sees", "announces", "reaffirms", "expects"]
ITEMS = ["revenue", "net income", "operating profit", "free cash flow", "adjusted EBITDA", "order backlog", "same-store sales"]
DIRS = [("rose", "up"), ("fell", "down"), ("climbed", "higher"), ("slipped", "lower"), ("was little changed", "flat")]
ACTIONS = ["launches tender offer to acquire", "agrees to sell its stake in", "signs supply agreement with", "completes merger with",
           "opens new distribution centre with", "extends credit facility arranged by", "forms joint venture with",
           "settles patent dispute with"]
QUOTES = ["Demand in the quarter was ahead of our plan", "We continue to manage costs carefully", "The integration is on schedule",
          "Input prices remain a headwind", "Our balance sheet gives us room to invest", "We are encouraged by early orders",
          "Currency effects reduced reported growth", "The board approved a higher dividend"]
TITLES = ["chief executive", "chief financial officer", "chairman", "head of strategy", "treasurer"]
NAMES = ["A. Moreau", "B. Lindqvist", "C. Okafor", "D. Tanaka", "E. Castillo", "F. Brennan", "G. Novak", "H. Adeyemi"]


def _money(r):
    v = r.choice([r.uniform(1, 99), r.uniform(100, 999), r.uniform(1, 40)])
    return f"${v:.1f} {'million' if v > 40 or r.random() < 0.5 else 'billion'}"


def _sentences(r):
    a, b = r.sample(COMPANIES, 2)
    head = r.choice(["BRIEF-", "UPDATE 1-", "", "", "CORRECTED-"])
    kind = r.random()
    out = []
    if kind < 0.4:
        it, (vb, adj) = r.choice(ITEMS), r.choice(DIRS)
        q = r.choice(["first", "second", "third", "fourth"])
        out.append(f"{head}{a} {r.choice(VERBS)} {q}-quarter {it} of {_money(r)}.")
        out.append(f"{r.choice(MONTHS)} {r.randint(1, 28)} ({r.choice(CITIES)}) - {a} said {q}-quarter {it} {vb} {r.uniform(0.3, 24):.1f} percent "
                   f"from a year earlier, {adj} from the {_money(r)} analysts had expected.")
        out.append(f"{r.choice(ITEMS).capitalize()} for the full year is now seen at {_money(r)} to {_money(r)}.")
    elif kind < 0.75:
        act = r.choice(ACTIONS)
        out.append(f"{head}{a} {act} {b} for {_money(r)}.")
        out.append(f"{r.choice(MONTHS)} {r.randint(1, 28)} ({r.choice(CITIES)}) - {a}: * {a} {act} {b} * Deal valued at ${r.uniform(2, 80):.2f} per share in cash "
                   f"* Transaction expected to close in the {r.choice(['first', 'second'])} half of {r.randint(2017, 2024)}, subject to regulatory approval.")
        out.append(f"{b} shares {r.choice(DIRS)[0]} {r.uniform(0.2, 9):.1f} percent in {r.choice(CITIES)} trading after the announcement.")
    else:
        out.append(f"{head}{a} names {r.choice(NAMES)} as {r.choice(TITLES)}.")
        out.append(f"{r.choice(MONTHS)} {r.randint(1, 28)} ({r.choice(CITIES)}) - {a} said on {r.choice(['Monday', 'Tuesday', 'Wednesday', 'Thursday', 'Friday'])} that "
                   f"{r.choice(NAMES)} will step down at the end of {r.choice(MONTHS)} after {r.randint(2, 19)} years with the company.")
    for _ in range(r.randint(1, 4)):
        out.append(f'"{r.choice(QUOTES)}," {r.choice(NAMES)}, the {r.choice(TITLES)} of {a}, said in a statement.')
    if r.random() < 0.5:
        out.append(f"The company employs about {r.randint(3, 90) * 100} people and has operations in {r.choice(CITIES)} and {r.choice(CITIES)}.")
    return out


def make_docs(n, seed=1234):
    r = random.Random(seed)
    return [_sentences(r) for _ in range(n)]


# ---- byte-level BPE trainer (GPT-2 vocab.json / merges.txt format) -----------------------------------------------------------

def train_bpe(texts, n_merges):
    from megatronapp_b200.data.tokenizer import bytes_to_unicode

    b2u = bytes_to_unicode()
    pat = re.compile(r"'s|'t|'re|'ve|'m|'ll|'d| ?[A-Za-z]+| ?[0-9]+| ?[^\sA-Za-z0-9]+|\s+(?!\S)|\s+")
    words = collections.Counter()
    for t in texts:
        for w in pat.findall(t):
            words[tuple(b2u[b] for b in w.encode("utf-8"))] += 1
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), c = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        if c < 2:
            break
        merges.append((a, b))
        new = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and w[i] == a and w[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        words = new
    vocab = {u: i for i, u in enumerate(b2u[b] for b in range(256))}
    for a, b in merges:
        vocab.setdefault(a + b, len(vocab))
    vocab["<|endoftext|>"] = len(vocab)
    return vocab, merges


# ---- WordPiece trainer (BERT vocab.txt) ----------------------------------------------------------------------------------

def train_wordpiece(texts, size):
    counts = collections.Counter()
    for t in texts:
        for w in re.findall(r"[a-z0-9]+|[^\sa-z0-9]", t.lower()):
            counts[w] += 1
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    chars = sorted({c for w in counts for c in w})
    vocab += chars + ["##" + c for c in chars if c.isalnum()]
    seen = set(vocab)
    # whole words by frequency, then their most frequent suffix pieces
    for w, _ in counts.most_common():
        if len(vocab) >= size * 3 // 4:
            break
        if w not in seen:
            vocab.append(w)
            seen.add(w)
    suffix = collections.Counter()
    for w, c in counts.items():
        for k in range(2, min(6, len(w))):
            suffix["##" + w[-k:]] += c
    for s, _ in suffix.most_common():
        if len(vocab) >= size:
            break
        if s not in seen:
            vocab.append(s)
            seen.add(s)
    return vocab


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1500)
    ap.add_argument("--bpe-merges", type=int, default=1500)
    ap.add_argument("--wordpiece-size", type=int, default=2000)
    ap.add_argument("--out", default=ROOT)
    a = ap.parse_args()
    docs = make_docs(a.docs)
    texts = [" ".join(d) for d in docs]
    for sub in ("datasets_gpt", "datasets_bert"):
        os.makedirs(os.path.join(a.out, sub), exist_ok=True)
        with open(os.path.join(a.out, sub, "dataset.json"), "w") as f:
            for t in texts:
                f.write(json.dumps({"src": "synthetic-newswire", "text": t}) + "\n")
    with open(os.path.join(a.out, "datasets_bert", "dataset_ss.json"), "w") as f:       # pre-split sentences
        for d in docs:
            f.write(json.dumps({"text": d}) + "\n")
    vocab, merges = train_bpe(texts, a.bpe_merges)
    with open(os.path.join(a.out, "datasets_gpt", "vocab.json"), "w") as f:
        json.dump(vocab, f, ensure_ascii=False)
    with open(os.path.join(a.out, "datasets_gpt", "merges.txt"), "w") as f:
        f.write("#version: 0.2\n" + "".join(f"{x} {y}\n" for x, y in merges))
    with open(os.path.join(a.out, "datasets_bert", "vocab.txt"), "w") as f:
        f.write("\n".join(train_wordpiece(texts, a.wordpiece_size)) + "\n")
    print(f"{len(texts)} documents, {len(vocab)} BPE tokens, {len(merges)} merges")


if __name__ == "__main__":
    main()
