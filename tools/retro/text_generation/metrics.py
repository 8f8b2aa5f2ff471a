"""Generation metrics for retrieval-augmented QA (reference ``tools/retro/text_generation/metrics.py``): token-level F1 and
exact match after the SQuAD normalisation (lower-case, strip punctuation / articles / extra white space); with several
reference answers the best score counts."""
import os
import sys
from typing import List, Sequence, Tuple, Union

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))

from tasks.msdp.metrics import normalize_answer, prf1  # noqa: E402

Answers = Union[str, Sequence[str]]


def f1_score(guess: str, answers: Answers) -> float:
    answers = [answers] if isinstance(answers, str) else list(answers)
    best = 0.0
    for a in answers:
        r = prf1(guess, a)
        if r is not None:
            best = max(best, r[2])
    return best


def exact_match(guess: str, answers: Answers) -> float:
    answers = [answers] if isinstance(answers, str) else list(answers)
    return float(any(normalize_answer(guess) == normalize_answer(a) for a in answers))


def corpus_scores(guesses: List[str], answers: List[Answers]) -> Tuple[float, float]:
    if len(guesses) != len(answers):
        raise ValueError(f"{len(guesses)} generations vs {len(answers)} references")
    n = max(1, len(guesses))
    return (100.0 * sum(f1_score(g, a) for g, a in zip(guesses, answers)) / n,
            100.0 * sum(exact_match(g, a) for g, a in zip(guesses, answers)) / n)
