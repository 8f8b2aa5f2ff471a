"""Latency / throughput test of the inference engines (reference ``tools/run_inference_performance_test.py``).

Takes the regular model arguments (``--num-layers ... --load ckpt --tokenizer-type ...``; ``--model-provider gpt|mamba``) plus

    --prompts "..." "..." | --num-input-tokens N      prompts as text, or random prompts of N tokens
    --max-batch-size B                                 prompts processed at once (and number of random prompts)
    --num-tokens-to-generate N --temperature --top_k --top_p --return-log-probs
    --stream                                           token streaming through the dynamic engine: time to first token + rate
    --inference-dynamic-batching                       continuous-batching engine instead of the static one

and prints one JSON line: prefill + decode seconds (CUDA events on a GPU), generated tokens / s, latency per request; with
``--stream`` also the time to the first token.  Without any model arguments a small random-init GPT is used (``--num-layers 4
--hidden-size 512 --num-attention-heads 8 --vocab-size 32000``; the older ``--prompt-length`` spelling of
this tool is kept)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def add_text_generate_args(parser):
    g = parser.add_argument_group("text generation")
    g.add_argument("--temperature", type=float, default=1.0)
    g.add_argument("--top_k", "--top-k", dest="top_k", type=int, default=1)
    g.add_argument("--top_p", "--top-p", dest="top_p", type=float, default=0.0)
    g.add_argument("--return-log-probs", action="store_true")
    g.add_argument("--num-tokens-to-generate", type=int, default=30)
    g.add_argument("--prompts", type=str, default=None, nargs="+")
    g.add_argument("--num-input-tokens", "--prompt-length", dest="num_input_tokens", type=int, default=None)
    g.add_argument("--max-batch-size", dest="inference_max_requests", type=int, default=8)
    g.add_argument("--stream", action="store_true")
    g.add_argument("--model-provider", choices=["mamba", "gpt"], default="gpt")
    g.add_argument("--num-timed-runs", type=int, default=1)
    return parser


_DEFAULTS = {"num_layers": 4, "hidden_size": 512, "num_attention_heads": 8, "seq_length": 512, "max_position_embeddings": 512,
             "micro_batch_size": 1, "tokenizer_type": "NullTokenizer", "vocab_size": 31999, "position_embedding_type": "rope"}


def main():
    import torch

    from megatronapp_b200.inference import (DynamicInferenceEngine, GPTInferenceWrapper, SamplingParams, StaticInferenceEngine,
                                             TextGenerationController)
    from megatronapp_b200.training import checkpointing
    from megatronapp_b200.training.global_vars import get_args, get_tokenizer
    from megatronapp_b200.training.initialize import initialize_megatron
    from megatronapp_b200.training.training import get_model

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    initialize_megatron(extra_args_provider=add_text_generate_args, args_defaults=_DEFAULTS)
    args = get_args()
    if args.model_provider == "mamba":
        from pretrain_mamba import model_provider
    else:
        from pretrain_gpt import model_provider
    model = get_model(model_provider, wrap_with_ddp=False)
    if args.load:
        checkpointing.load_checkpoint(model, None, None)
    for m in model:
        m.eval()
    tok = get_tokenizer()
    cuda = torch.cuda.is_available()
    n_new, batch = args.num_tokens_to_generate, args.inference_max_requests
    if args.prompts:
        prompts = list(args.prompts)
    else:
        g = torch.Generator().manual_seed(args.seed)
        n_in = args.num_input_tokens or 128
        hi = max(3, min(tok.vocab_size, args.padded_vocab_size) - 2)
        prompts = [tok.detokenize([int(x) for x in torch.randint(1, hi, (n_in,), generator=g)]) for _ in range(batch)]
    max_len = max(len(tok.tokenize(p)) for p in prompts) + n_new + 1
    sp = SamplingParams(temperature=args.temperature, top_k=args.top_k, top_p=args.top_p, return_log_probs=args.return_log_probs,
                        num_tokens_to_generate=n_new)
    dynamic = args.stream or args.inference_dynamic_batching
    if dynamic:
        engine = DynamicInferenceEngine.from_args(model[0], tok, args, model[0].config) if args.inference_dynamic_batching else \
            DynamicInferenceEngine(model[0], tok, max_sequence_length=max(max_len, 16), max_requests=batch)
    else:
        wrapper = GPTInferenceWrapper(model[0], batch, max(max_len, args.inference_max_seq_length or 0), args.padded_vocab_size)
        engine = StaticInferenceEngine(TextGenerationController(wrapper, tok), max_batch_size=batch)

    def sync():
        if cuda:
            torch.cuda.synchronize()

    def timed_run():
        first = None
        sync()
        t0 = time.perf_counter()
        if args.stream:
            # one streamed request with the rest of the batch sharing its steps: time to first token and inter-token rate
            others = [engine.add_request(p, sp) for p in prompts[1:]]
            n = 0
            for _chunk in engine.generate_stream(prompts[0], sp):
                if first is None:
                    sync()
                    first = time.perf_counter() - t0
                n += 1
            while engine.has_unfinished_requests():
                engine.step()
            generated = n + sum(n_new for _ in others)
        else:
            out = engine.generate(prompts, sp)
            generated = sum(int(r.generated_length) for r in out)
        sync()
        return time.perf_counter() - t0, generated, first

    warm = SamplingParams(top_k=1, num_tokens_to_generate=2)
    if dynamic:
        engine.generate(prompts[:1], warm)
    else:
        engine.generate(prompts[:1], warm)
    best = None
    for _ in range(max(1, args.num_timed_runs)):
        run = timed_run()
        best = run if best is None or run[0] < best[0] else best
    seconds, generated, first = best
    if args.rank == 0:
        line = {"engine": "dynamic" if dynamic else "static", "model_provider": args.model_provider, "batch": len(prompts),
                "prompt_tokens": [len(tok.tokenize(p)) for p in prompts][:4], "generated_tokens": generated, "seconds": seconds,
                "tokens_per_s": generated / max(seconds, 1e-9), "latency_per_request_s": seconds,
                "tp": args.tensor_model_parallel_size, "pp": args.pipeline_model_parallel_size}
        if first is not None:
            line["time_to_first_token_s"] = first
        print(json.dumps(line))


if __name__ == "__main__":
    main()
