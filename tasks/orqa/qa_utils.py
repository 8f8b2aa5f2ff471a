"""Answer matching for open-retrieval QA (reference ``tasks/orqa/unsupervised/qa_utils.py``): a retrieved passage is
a hit when it contains one of the gold answers as a token sub-sequence (``string``) or matches it as a regular
expression (``regex``), after unicode NFD normalisation, lower-casing and simple word tokenisation."""
import re
import unicodedata
from typing import Dict, List, Sequence, Tuple

_WORD = re.compile(r"[\w]+|[^\s\w]", re.UNICODE)


def normalize(text: str) -> str:
    return unicodedata.normalize("NFD", text)


def words(text: str) -> List[str]:
    return [w.lower() for w in _WORD.findall(normalize(text))]


def has_answer(answers: Sequence[str], text: str, match_type: str = "string") -> bool:
    if match_type == "regex":
        for a in answers:
            try:
                if re.compile(normalize(a), re.IGNORECASE | re.UNICODE | re.MULTILINE).search(normalize(text)):
                    return True
            except re.error:
                continue
        return False
    tw = words(text)
    for a in answers:
        aw = words(a)
        n = len(aw)
        if n and any(tw[i:i + n] == aw for i in range(len(tw) - n + 1)):
            return True
    return False


def calculate_matches(all_docs: Dict[object, Tuple[str, str]], answers: List[List[str]],
                      closest_docs: List[Tuple[List[object], List[float]]], match_type: str = "string"):
    """``closest_docs[i] = (doc ids by decreasing score, scores)``.  Returns ``(top_k_hits, per-question hit lists)``
    where ``top_k_hits[k-1]`` counts the questions answered by one of their first ``k`` passages."""
    n_docs = len(closest_docs[0][0]) if closest_docs else 0
    top_k = [0] * n_docs
    per_q = []
    for ans, (ids, _) in zip(answers, closest_docs):
        hits = [has_answer(ans, all_docs[d][0], match_type) for d in ids]
        per_q.append(hits)
        first = next((i for i, h in enumerate(hits) if h), None)
        if first is not None:
            for k in range(first, n_docs):
                top_k[k] += 1
    return top_k, per_q


def exact_match_score(prediction: str, truth: str) -> bool:
    from tasks.msdp.metrics import normalize_answer

    return normalize_answer(prediction) == normalize_answer(truth)
