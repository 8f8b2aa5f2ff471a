"""Score generated answers against references (reference ``tools/retro/text_generation/evaluate.py``).

    python tools/retro/text_generation/evaluate.py --generations out.txt --references test.json

``--generations``: one answer per line (everything after the first ``Answer:`` marker, up to the end of the line / the first
``<|endoftext|>``, is the answer).  ``--references``: json list with ``answers`` (list) or ``answer`` (str) per item, or a
text file with one reference per line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))

from tools.retro.text_generation.metrics import corpus_scores  # noqa: E402


def read_generations(path):
    out = []
    with open(path, encoding="utf-8") as f:
        for line in f:
            t = line.rstrip("\n")
            if "Answer:" in t:
                t = t.split("Answer:", 1)[1]
            out.append(t.split("<|endoftext|>")[0].strip())
    return out


def read_references(path):
    with open(path, encoding="utf-8") as f:
        head = f.read(1)
        f.seek(0)
        if head == "[":
            rows = json.load(f)
            return [r.get("answers") or [r.get("answer", "")] for r in rows]
        return [line.rstrip("\n") for line in f]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--generations", required=True)
    ap.add_argument("--references", required=True)
    a = ap.parse_args(argv)
    f1, em = corpus_scores(read_generations(a.generations), read_references(a.references))
    print(f"F1: {f1:.2f}  EM: {em:.2f}")
    return f1, em


if __name__ == "__main__":
    main()
