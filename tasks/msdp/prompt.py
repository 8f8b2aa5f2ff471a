"""Multi-stage dialogue prompting (reference ``tasks/msdp/prompt.py``, Liu et al. 2022).

Stage 1 (``knowledge``) prompts the LM with few-shot examples *selected per topic / last turn* to generate the
knowledge sentence; stage 2 (``response``) prompts it with fixed examples plus that knowledge to generate the reply.
Input lines are ``topic \\t dialogue context (turns joined by [SEP]) [\\t knowledge]``; the generation stops at the
first newline.  Generation goes through this framework's inference engine (or an HTTP ``/api`` endpoint).
"""
import json
from typing import Callable, Dict, List, Optional


def read_prompts(prompt_path: str, prompt_type: str, n_example: int):
    """knowledge: jsonl of ``{"<topic> <last turn>": [examples...]}`` -> dict of joined few-shot blocks;
    response: plain lines -> one fixed block of the first ``n_example`` lines."""
    if prompt_type == "knowledge":
        out: Dict[str, str] = {}
        with open(prompt_path) as f:
            for line in f:
                if line.strip():
                    d = json.loads(line)
                    for key, examples in d.items():
                        out.setdefault(key, "".join(e.strip() + " \n" for e in examples[:n_example]))
        return out
    with open(prompt_path) as f:
        lines = [ln.strip() for ln in f if ln.strip()][:n_example]
    return "".join(ln + " \n" for ln in lines)


def build_input(line: str, prompts, prompt_type: str) -> str:
    parts = line.rstrip("\n").split("\t")
    topic, turns = parts[0], parts[1].split(" [SEP] ")
    last = turns[-1]
    if prompt_type == "knowledge":
        shots = prompts.get(f"{topic} {last}", "")
        return f"{shots}( {last} ) {topic} =>"
    knowledge = parts[2] if len(parts) > 2 else ""
    return f"{prompts}Topic: {topic}. User says: {last} We know that: {knowledge} System replies:"


def first_line(text: str) -> str:
    return text.strip().split("\n")[0].strip()


def generate_samples(generate: Callable[[List[str], int], List[str]], sample_input_file: str, sample_output_file: str,
                     prompt_file: str, prompt_type: str, num_prompt_examples: int = 10, out_seq_length: int = 64,
                     batch_size: int = 8) -> int:
    """``generate(prompts, tokens_to_generate) -> continuations`` (without the prompt)."""
    assert prompt_type in ("knowledge", "response")
    prompts = read_prompts(prompt_file, prompt_type, num_prompt_examples)
    with open(sample_input_file) as f:
        lines = [ln for ln in f if ln.strip()]
    n = 0
    with open(sample_output_file, "w") as out:
        for i in range(0, len(lines), batch_size):
            inputs = [build_input(ln, prompts, prompt_type) for ln in lines[i:i + batch_size]]
            for text in generate(inputs, out_seq_length):
                out.write(first_line(text) + "\n")
                n += 1
    return n


def engine_generate_fn(engine) -> Callable[[List[str], int], List[str]]:
    from megatronapp_b200.inference import SamplingParams

    def fn(prompts: List[str], n_tokens: int) -> List[str]:
        res = engine.generate(prompts, SamplingParams(top_k=1, num_tokens_to_generate=n_tokens))
        return [r.generated_text for r in res]

    return fn


def api_generate_fn(url: str) -> Callable[[List[str], int], List[str]]:
    import urllib.request

    def fn(prompts: List[str], n_tokens: int) -> List[str]:
        body = json.dumps({"prompts": prompts, "tokens_to_generate": n_tokens, "top_k": 1}).encode()
        req = urllib.request.Request(url, data=body, method="PUT", headers={"Content-Type": "application/json"})
        out = json.loads(urllib.request.urlopen(req).read())
        return [t[len(p):] for t, p in zip(out["text"], prompts)]

    return fn
