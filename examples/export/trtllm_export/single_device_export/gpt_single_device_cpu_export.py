"""Export a GPT model to a TensorRT-LLM checkpoint on one device (reference
``examples/export/trtllm_export/single_device_export/gpt_single_device_cpu_export.py``): full weights -> TRT-LLM names,
re-sharded for the *inference* layout (``--export-tp``, ``--export-pp``), written as ``config.json`` + ``rank{r}.safetensors``,
plus the engine build plan (and the engine itself when ``tensorrt_llm`` is installed).

    python examples/export/trtllm_export/single_device_export/gpt_single_device_cpu_export.py --out trt_ckpt --export-tp 2
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))))

from megatronapp_b200.config import TransformerConfig  # noqa: E402
from megatronapp_b200.export import ExportConfig, TRTLLMEngineBuilder, TRTLLMHelper  # noqa: E402
from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec  # noqa: E402
from megatronapp_b200.parallel import state as ps  # noqa: E402
from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--export-tp", type=int, default=1)
    ap.add_argument("--export-pp", type=int, default=1)
    ap.add_argument("--num-layers", type=int, default=2)
    ap.add_argument("--hidden-size", type=int, default=64)
    ap.add_argument("--num-attention-heads", type=int, default=4)
    ap.add_argument("--vocab-size", type=int, default=256)
    ap.add_argument("--seq-length", type=int, default=64)
    ap.add_argument("--load-state-dict", default=None, help="torch file with a canonical state dict (default: random init)")
    ap.add_argument("--fp8", action="store_true")
    a = ap.parse_args(argv)
    if not ps.model_parallel_is_initialized():
        ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(123)
    cfg = TransformerConfig(num_layers=a.num_layers, hidden_size=a.hidden_size, num_attention_heads=a.num_attention_heads,
                            use_cpu_initialization=True, params_dtype=torch.float32, kernel_backend="torch")
    model = GPTModel(cfg, get_gpt_layer_local_spec(), a.vocab_size, a.seq_length)
    sd = torch.load(a.load_state_dict, map_location="cpu") if a.load_state_dict else model.state_dict()
    sd = {k: v for k, v in sd.items() if torch.is_tensor(v) and not k.endswith("_extra_state")}
    helper = TRTLLMHelper(cfg, model_type="gpt", position_embedding_type="learned_absolute", max_position_embeddings=a.seq_length,
                          share_embeddings_and_output_weights=True, activation="gelu")
    weights, configs = helper.get_trtllm_pretrained_config_and_model_weights(
        sd, "bfloat16", ExportConfig(a.export_tp, a.export_pp), vocab_size=a.vocab_size, fp8_quantized=a.fp8, fp8_kvcache=a.fp8)
    os.makedirs(a.out, exist_ok=True)
    with open(os.path.join(a.out, "config.json"), "w") as f:
        json.dump(configs[0], f, indent=1)
    for w, c in zip(weights, configs):
        r = c["mapping"]["rank"]
        try:
            from safetensors.torch import save_file

            save_file({k: v.contiguous() for k, v in w.items()}, os.path.join(a.out, f"rank{r}.safetensors"))
        except Exception:
            torch.save(w, os.path.join(a.out, f"rank{r}.pt"))
    plan = TRTLLMEngineBuilder.plan(configs[0], max_input_len=a.seq_length // 2, max_output_len=a.seq_length // 2, max_batch_size=4)
    with open(os.path.join(a.out, "build_config.json"), "w") as f:
        json.dump({"build": plan.build_dict(), "plugin_config": plan.plugin}, f, indent=1)
    print(f"exported {len(weights)} rank file(s) to {a.out} (tp {a.export_tp} x pp {a.export_pp})")
    try:
        import tensorrt_llm  # noqa: F401

        helper.build_and_save_engine(os.path.join(a.out, "engine"), weights[0], configs[0], max_batch_size=4)
    except ImportError:
        print("tensorrt_llm is not installed: run `trtllm-build --checkpoint_dir` on the exported directory to build the engine")
    return configs


if __name__ == "__main__":
    main()
