"""Decode-throughput test of the static inference engine (reference
``tools/run_inference_performance_test.py``): random-init GPT, synthetic prompts, reports prefill
latency and decode tokens/s (CUDA events on GPU, wall clock on CPU)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from megatronapp_b200.config import TransformerConfig
    from megatronapp_b200.data.tokenizer import NullTokenizer
    from megatronapp_b200.inference.core import (GPTInferenceWrapper, SamplingParams, StaticInferenceEngine,
                                                 TextGenerationController)
    from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed

    ap = argparse.ArgumentParser()
    ap.add_argument("--num-layers", type=int, default=4)
    ap.add_argument("--hidden-size", type=int, default=512)
    ap.add_argument("--num-attention-heads", type=int, default=8)
    ap.add_argument("--vocab-size", type=int, default=32000)
    ap.add_argument("--batch-size", type=int, default=8)
    ap.add_argument("--prompt-length", type=int, default=128)
    ap.add_argument("--num-tokens-to-generate", type=int, default=64)
    a = ap.parse_args()
    cuda = torch.cuda.is_available()
    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    dt = torch.bfloat16 if cuda else torch.float32
    cfg = TransformerConfig(num_layers=a.num_layers, hidden_size=a.hidden_size, num_attention_heads=a.num_attention_heads,
                            bf16=cuda, params_dtype=dt, use_cpu_initialization=not cuda, hidden_dropout=0.0,
                            attention_dropout=0.0)
    total = a.prompt_length + a.num_tokens_to_generate
    model = GPTModel(cfg, get_gpt_layer_local_spec(), a.vocab_size, total, position_embedding_type="rope")
    if cuda:
        model = model.cuda()
    model.eval()
    tok = NullTokenizer(a.vocab_size - 1)
    eng = StaticInferenceEngine(TextGenerationController(GPTInferenceWrapper(model, a.batch_size, total, a.vocab_size), tok),
                                max_batch_size=a.batch_size)
    g = torch.Generator().manual_seed(0)
    prompts = [" ".join(str(int(x)) for x in torch.randint(1, a.vocab_size - 2, (a.prompt_length,), generator=g))
               for _ in range(a.batch_size)]
    sp = SamplingParams(top_k=1, num_tokens_to_generate=a.num_tokens_to_generate)
    eng.generate(prompts[:1], SamplingParams(top_k=1, num_tokens_to_generate=2))      # warm-up
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.generate(prompts, sp)
    if cuda:
        torch.cuda.synchronize()
    dt_s = time.perf_counter() - t0
    print(json.dumps({"batch": a.batch_size, "prompt_length": a.prompt_length, "generated": a.num_tokens_to_generate,
                      "seconds": dt_s, "decode_tokens_per_s": a.batch_size * a.num_tokens_to_generate / dt_s}))


if __name__ == "__main__":
    main()
