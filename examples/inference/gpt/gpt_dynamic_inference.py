"""Continuous-batching generation with the dynamic engine -- paged KV cache, chunked prefill, requests joining between any
two steps (reference ``examples/inference/gpt/gpt_dynamic_inference.py``).  With ``--incoming-requests-per-sec R`` the
requests arrive as a Poisson process and are admitted while earlier ones are still decoding.

    python examples/inference/gpt/gpt_dynamic_inference.py --load ckpt ... --num-synthetic-prompts 64 --incoming-requests-per-sec 8 \\
        --inference-dynamic-batching-buffer-size-gb 20 --num-tokens-to-generate 64
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from utils import add_common_inference_args, build_requests, get_curr_time, report  # noqa: E402

from megatronapp_b200.inference import DynamicInferenceEngine, SamplingParams  # noqa: E402
from megatronapp_b200.training import checkpointing  # noqa: E402
from megatronapp_b200.training.global_vars import get_args, get_tokenizer  # noqa: E402
from megatronapp_b200.training.initialize import initialize_megatron  # noqa: E402
from megatronapp_b200.training.training import get_model  # noqa: E402
from pretrain_gpt import model_provider  # noqa: E402


def main():
    initialize_megatron(extra_args_provider=add_common_inference_args, args_defaults={"micro_batch_size": 1})
    args = get_args()
    model = get_model(model_provider, wrap_with_ddp=False)
    if args.load:
        checkpointing.load_checkpoint(model, None, None)
    tok = get_tokenizer()
    cfg = model[0].config if hasattr(model[0], "config") else None
    engine = DynamicInferenceEngine.from_args(model[0], tok, args, cfg)
    engine.padded_vocab_size = args.padded_vocab_size
    reqs = build_requests(args, tok)
    sp = SamplingParams(num_tokens_to_generate=args.num_tokens_to_generate, top_k=args.top_k, top_p=args.top_p,
                        temperature=args.temperature, return_log_probs=args.return_log_probs)
    t0 = get_curr_time()
    pending = sorted(range(len(reqs)), key=lambda i: reqs[i].time_offset)
    ids = {}
    while pending or engine.has_unfinished_requests():
        now = get_curr_time() - t0
        while pending and reqs[pending[0]].time_offset <= now:                 # admit everything that has "arrived"
            i = pending.pop(0)
            r = reqs[i]
            r.time_arrival = get_curr_time()
            ids[engine.add_request(r.prompt, sp, prompt_tokens=r.tokens)] = i
        if not engine.has_unfinished_requests():
            continue                                                            # idle until the next arrival
        for done in engine.step():
            r = reqs[ids[done.request_id]]
            r.output, r.output_tokens, r.time_end = done.generated_text, [int(x) for x in done.generated_tokens], get_curr_time()
    if args.rank == 0:
        print(f"engine steps: {engine.steps}, preemptions: {engine.num_preemptions}, paged-attention kernel calls: {engine.context.kernel_calls}")
        report(reqs, get_curr_time() - t0, args.output_path)


if __name__ == "__main__":
    main()
