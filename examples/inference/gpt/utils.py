"""Shared pieces of the GPT inference examples (reference ``examples/inference/gpt/utils.py``): request construction with
optional arrival times, common CLI flags, a results table."""
import json
import os
import random
import sys
import time
from dataclasses import dataclass, field
from typing import List, Optional

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))


def add_common_inference_args(parser):
    g = parser.add_argument_group("inference example")
    g.add_argument("--prompts", nargs="+", default=None, help="prompt strings")
    g.add_argument("--prompt-file", default=None, help="one prompt per line, or jsonl with a 'text' field")
    g.add_argument("--num-tokens-to-generate", type=int, default=30)
    g.add_argument("--top_k", "--top-k", dest="top_k", type=int, default=1)
    g.add_argument("--top_p", "--top-p", dest="top_p", type=float, default=0.0)
    g.add_argument("--temperature", type=float, default=1.0)
    g.add_argument("--return-log-probs", action="store_true")
    g.add_argument("--incoming-requests-per-sec", type=float, default=0.0,
                   help="> 0: requests arrive as a Poisson process at this rate (continuous batching demo); 0: all at time 0")
    g.add_argument("--num-synthetic-prompts", type=int, default=0, help="generate this many random-token prompts instead")
    g.add_argument("--synthetic-prompt-length", type=int, default=32)
    g.add_argument("--output-path", default=None, help="write results as json")
    return parser


@dataclass
class Request:
    prompt: str
    time_offset: float = 0.0
    tokens: Optional[List[int]] = None
    time_arrival: Optional[float] = None
    time_end: Optional[float] = None
    output: str = ""
    output_tokens: List[int] = field(default_factory=list)


def build_requests(args, tokenizer) -> List[Request]:
    prompts: List[str] = list(args.prompts or [])
    if args.prompt_file:
        with open(args.prompt_file, encoding="utf-8") as f:
            for line in f:
                line = line.rstrip("\n")
                if line:
                    prompts.append(json.loads(line)["text"] if line.startswith("{") else line)
    rng = random.Random(args.seed)
    reqs = [Request(p) for p in prompts]
    for _ in range(args.num_synthetic_prompts):
        toks = [rng.randrange(1, max(2, tokenizer.vocab_size - 1)) for _ in range(args.synthetic_prompt_length)]
        reqs.append(Request(tokenizer.detokenize(toks), tokens=toks))
    if not reqs:
        raise SystemExit("give --prompts, --prompt-file or --num-synthetic-prompts")
    t = 0.0
    for r in reqs:
        if args.incoming_requests_per_sec > 0:
            t += rng.expovariate(args.incoming_requests_per_sec)
        r.time_offset = t
    return reqs


def get_curr_time() -> float:
    return time.time()


def report(reqs: List[Request], wall: float, path: Optional[str] = None):
    n_out = sum(len(r.output_tokens) for r in reqs)
    lat = [r.time_end - r.time_arrival for r in reqs if r.time_end and r.time_arrival]
    print(f"{len(reqs)} requests, {n_out} generated tokens in {wall:.2f} s -> {n_out / max(wall, 1e-9):.1f} tokens/s"
          + (f", mean latency {sum(lat) / len(lat):.2f} s, max {max(lat):.2f} s" if lat else ""))
    for i, r in enumerate(reqs[:8]):
        print(f"  [{i}] {r.prompt[:60]!r} -> {r.output[:80]!r}")
    if path:
        with open(path, "w") as f:
            json.dump([{"prompt": r.prompt, "output": r.output, "tokens": r.output_tokens,
                        "latency": (r.time_end - r.time_arrival) if r.time_end and r.time_arrival else None} for r in reqs], f, indent=1)
