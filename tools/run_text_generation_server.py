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


def main():
    initialize_megatron(args_defaults={"tokenizer_type": "NullTokenizer", "micro_batch_size": 1})
    args = get_args()
    model = get_model(model_provider, wrap_with_ddp=False)
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
