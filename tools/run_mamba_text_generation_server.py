#!/usr/bin/env python
"""Text generation server for Mamba / hybrid models (reference ``tools/run_mamba_text_generation_server.py``).
Mamba layers decode recurrently (conv + SSM state in the inference context), attention layers use the KV cache."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatronapp_b200.inference import GPTInferenceWrapper, StaticInferenceEngine, TextGenerationController  # noqa: E402
from megatronapp_b200.inference.server import InferenceWSServer, MegatronServer  # noqa: E402
from megatronapp_b200.training import checkpointing  # noqa: E402
from megatronapp_b200.training.global_vars import get_args, get_tokenizer  # noqa: E402
from megatronapp_b200.training.initialize import initialize_megatron  # noqa: E402
from megatronapp_b200.training.training import get_model  # noqa: E402
from pretrain_mamba import model_provider  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from run_text_generation_server import add_text_generate_args  # noqa: E402


def main():
    initialize_megatron(extra_args_provider=add_text_generate_args,
                        args_defaults={"tokenizer_type": "NullTokenizer", "micro_batch_size": 1})
    args = get_args()
    if args.max_batch_size_override:
        args.inference_max_requests = args.max_batch_size_override
    model = get_model(model_provider, wrap_with_ddp=False)
    if args.load:
        checkpointing.load_checkpoint(model, None, None)
    tok = get_tokenizer()
    wrapper = GPTInferenceWrapper(model[0], args.inference_max_requests, args.inference_max_seq_length, args.padded_vocab_size)
    engine = StaticInferenceEngine(TextGenerationController(wrapper, tok), args.inference_max_requests)
    if args.rank != 0:                       # model-parallel peers follow rank 0's announcements
        from megatronapp_b200.inference.server import serve_worker_ranks

        return serve_worker_ranks(engine)
    if args.enable_ws_server:
        InferenceWSServer(engine, args.num_layers, tok, port=args.ws_port).run()
    else:
        MegatronServer(engine).run("0.0.0.0" if os.environ.get("B200_BIND_ALL") else "127.0.0.1", args.port)


if __name__ == "__main__":
    main()
