"""F1 between generated and reference responses / knowledge (reference ``tasks/msdp/evaluate.py``)."""
from .metrics import corpus_prf1


def _clean(line: str) -> str:
    line = line.strip()
    return line.replace("<|endoftext|>", "") if "<|endoftext|>" in line else line


def evaluate_f1(guess_file: str, answer_file: str):
    with open(guess_file) as f:
        guesses = [_clean(ln) for ln in f]
    with open(answer_file) as f:
        answers = [("" if ln.strip() == "no_passages_used" else ln.strip()) for ln in f]
    p, r, f1 = corpus_prf1(guesses, answers)
    print(f"Precision: {p:.4f}; recall: {r:.4f}; f1: {f1:.4f}")
    return p, r, f1


def main():
    from megatronapp_b200.training.global_vars import get_args

    a = get_args()
    return evaluate_f1(a.guess_file, a.answer_file)
