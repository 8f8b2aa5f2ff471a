"""Static-batch generation with a GPT checkpoint (reference ``examples/inference/gpt/gpt_static_inference.py``).

    torchrun --nproc-per-node 1 examples/inference/gpt/gpt_static_inference.py --load ckpt --tokenizer-type GPT2BPETokenizer \\
        --vocab-file vocab.json --merge-file merges.txt --num-layers 24 --hidden-size 1024 --num-attention-heads 16 \\
        --seq-length 1024 --max-position-embeddings 1024 --micro-batch-size 1 --prompts "Hello" --num-tokens-to-generate 30
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from utils import add_common_inference_args, build_requests, get_curr_time, report  # noqa: E402

from megatronapp_b200.inference import GPTInferenceWrapper, SamplingParams, StaticInferenceEngine, TextGenerationController  # noqa: E402
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
    wrapper = GPTInferenceWrapper(model[0], args.inference_max_requests, args.inference_max_seq_length, args.padded_vocab_size)
    engine = StaticInferenceEngine(TextGenerationController(wrapper, tok), args.inference_max_requests)
    reqs = build_requests(args, tok)
    sp = SamplingParams(num_tokens_to_generate=args.num_tokens_to_generate, top_k=args.top_k, top_p=args.top_p,
                        temperature=args.temperature, return_log_probs=args.return_log_probs)
    t0 = get_curr_time()
    for r in reqs:
        r.time_arrival = t0
    out = engine.generate([r.prompt for r in reqs], sp)
    t1 = get_curr_time()
    for r, o in zip(reqs, out):
        r.output, r.output_tokens, r.time_end = o.generated_text, [int(x) for x in o.generated_tokens], t1
    if args.rank == 0:
        report(reqs, t1 - t0, args.output_path)


if __name__ == "__main__":
    main()
