"""Token-level precision / recall / F1 for dialogue generation (reference ``tasks/msdp/metrics.py``)."""
import re
from collections import Counter
from typing import List, Optional, Tuple

_ARTICLES = re.compile(r"\b(a|an|the)\b")
_PUNCT = re.compile(r"[!\"#$%&()*+,\-./:;<=>?@\[\]\\^`{|}~_']")


def normalize_answer(s: str) -> str:
    return " ".join(_ARTICLES.sub(" ", _PUNCT.sub(" ", s.lower())).split())


def prf1(guess: str, answer: str) -> Optional[Tuple[float, float, float]]:
    """``None`` when there is no gold answer (the pair is skipped); zeros for an empty guess."""
    if answer == "":
        return None
    if guess == "":
        return 0.0, 0.0, 0.0
    g, a = normalize_answer(guess).split(), normalize_answer(answer).split()
    same = sum((Counter(g) & Counter(a)).values())
    if same == 0:
        return 0.0, 0.0, 0.0
    p, r = same / len(g), same / len(a)
    return p, r, 2 * p * r / (p + r)


def corpus_prf1(guesses: List[str], answers: List[str]) -> Tuple[float, float, float]:
    if len(guesses) != len(answers):
        raise ValueError("guess / answer count mismatch")
    rows = [x for x in (prf1(g, a) for g, a in zip(guesses, answers)) if x is not None]
    n = max(1, len(rows))
    return tuple(sum(r[i] for r in rows) / n for i in range(3))
