#!/usr/bin/env python
"""Text generation server (reference ``tools/run_text_generation_server.py``): REST on --port, or
the MegaScope WebSocket protocol with --enable-ws-server.  Loads ``--load`` if given."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatronapp_b200.inference import GPTInferenceWrapper, StaticInferenceEngine, TextGenerationController  # noqa: E402
from megatronapp_b200.inference.server import InferenceWSServer, MegatronServer  # noqa: E402
from megatronapp_b200.training import checkpointing  # noqa: E402
from megatronapp_b200.training.global_vars import get_args, get_tokenizer  # noqa: E402
from megatronapp_b200.training.initialize import initialize_megatron  # noqa: E402
from megatronapp_b200.training.training import get_model  # noqa: E402
from pretrain_gpt import model_provider  # noqa: E402


def add_text_generate_args(parser):
    """The reference tool's flag group; sampling options are defaults a client request may override."""
    g = parser.add_argument_group("text generation")
    g.add_argument("--temperature", type=float, default=1.0)
    g.add_argument("--top_k", "--top-k", dest="top_k", type=int, default=1)
    g.add_argument("--top_p", "--top-p", dest="top_p", type=float, default=0.0)
    g.add_argument("--return-log-probs", action="store_true")
    g.add_argument("--num-tokens-to-generate", type=int, default=30)
    g.add_argument("--prompts", type=str, default=None, nargs="+", help="generate for these once at start-up (printed by rank 0)")
    g.add_argument("--max-batch-size", type=int, default=None, dest="max_batch_size_override",
                   help="prompts processed at once (same as --inference-max-requests)")
    return parser


def main(provider=None, extra_args_provider=None):
    def extra(parser):
        parser = add_text_generate_args(parser)
        return extra_args_provider(parser) if extra_args_provider else parser

    initialize_megatron(extra_args_provider=extra, args_defaults={"tokenizer_type": "NullTokenizer", "micro_batch_size": 1})
    args = get_args()
    if args.max_batch_size_override:
        args.inference_max_requests = args.max_batch_size_override
    model = get_model(provider or model_provider, wrap_with_ddp=False)
    if args.load:
        checkpointing.load_checkpoint(model, None, None)
    tok = get_tokenizer()
    wrapper = GPTInferenceWrapper(model[0], args.inference_max_requests, args.inference_max_seq_length, args.padded_vocab_size)
    if args.inference_dynamic_batching:
        from megatronapp_b200.inference import DynamicInferenceEngine
        from megatronapp_b200.training.arguments import core_transformer_config_from_args

        engine = DynamicInferenceEngine.from_args(model[0], tok, args, core_transformer_config_from_args(args))
    else:
        engine = StaticInferenceEngine(TextGenerationController(wrapper, tok), args.inference_max_requests)
    if args.enable_cuda_graph or args.prompts:
        # every rank takes part: CUDA-graph warm-up (reference "Running warmup for CUDA graphs...") and / or the start-up prompts
        from megatronapp_b200.inference import SamplingParams

        sp = SamplingParams(temperature=args.temperature, top_k=args.top_k, top_p=args.top_p, return_log_probs=args.return_log_probs,
                            num_tokens_to_generate=args.num_tokens_to_generate if args.prompts else 10)
        warmup = "1 2 3" if type(tok).__name__ == "NullTokenizer" else "Test prompt"
        for r in engine.generate(list(args.prompts or [warmup]), sp):
            if args.rank == 0 and args.prompts:
                print(f"{r.prompt!r} -> {r.generated_text!r}", flush=True)
    if args.inference_dynamic_batching:
        # continuous batching: every rank runs the engine's serving loop in lockstep (rank 0 in a thread next to HTTP)
        if args.rank != 0:
            return engine.run_loop()
        engine.start_background_loop()
    if args.rank != 0:                       # model-parallel peers follow rank 0's announcements
        from megatronapp_b200.inference.server import serve_worker_ranks

        return serve_worker_ranks(engine)
    if args.enable_ws_server:
        InferenceWSServer(engine, args.num_layers, tok, port=args.ws_port).run()
    else:
        MegatronServer(engine).run("0.0.0.0" if os.environ.get("B200_BIND_ALL") else "127.0.0.1", args.port)


if __name__ == "__main__":
    main()
