"""On-device distributed TensorRT-LLM export (reference ``examples/export/trtllm_export/distributed_export/
gpt_distributed_gpu_export.py``): launched with the TRAINING layout (``torchrun --nproc-per-node tp*pp``); every rank
converts only its own shard -- nothing is gathered except the vocabulary-parallel embedding -- and writes its own rank file.

    torchrun --nproc-per-node 2 examples/export/trtllm_export/distributed_export/gpt_distributed_gpu_export.py \\
        --tensor-model-parallel-size 2 --out trt_ckpt
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))))

from megatronapp_b200.config import TransformerConfig  # noqa: E402
from megatronapp_b200.export import TRTLLMHelper  # noqa: E402
from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec  # noqa: E402
from megatronapp_b200.parallel import state as ps  # noqa: E402
from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--tensor-model-parallel-size", type=int, default=1)
    ap.add_argument("--pipeline-model-parallel-size", type=int, default=1)
    ap.add_argument("--num-layers", type=int, default=2)
    ap.add_argument("--hidden-size", type=int, default=64)
    ap.add_argument("--num-attention-heads", type=int, default=4)
    ap.add_argument("--vocab-size", type=int, default=256)
    ap.add_argument("--seq-length", type=int, default=64)
    a = ap.parse_args(argv)
    cuda = torch.cuda.is_available()
    if not dist.is_initialized():
        ps.init_distributed("nccl" if cuda else "gloo")
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    ps.initialize_model_parallel(a.tensor_model_parallel_size, a.pipeline_model_parallel_size)
    model_parallel_cuda_manual_seed(123)
    cfg = TransformerConfig(num_layers=a.num_layers, hidden_size=a.hidden_size, num_attention_heads=a.num_attention_heads,
                            tensor_model_parallel_size=a.tensor_model_parallel_size,
                            pipeline_model_parallel_size=a.pipeline_model_parallel_size, use_cpu_initialization=not cuda,
                            params_dtype=torch.float32, kernel_backend="torch" if not cuda else "auto")
    model = GPTModel(cfg, get_gpt_layer_local_spec(), a.vocab_size, a.seq_length, pre_process=ps.is_pipeline_first_stage(),
                     post_process=ps.is_pipeline_last_stage())
    if cuda:
        model = model.cuda()
    sd = {k: v for k, v in model.state_dict().items() if torch.is_tensor(v) and not k.endswith("_extra_state")}
    # local layer numbering -> global, the way a training rank's checkpoint shard is keyed
    off = ps.get_pipeline_model_parallel_rank() * (a.num_layers // a.pipeline_model_parallel_size)
    sd = {(f"decoder.layers.{int(k.split('.')[2]) + off}." + k.split(".", 3)[3]) if k.startswith("decoder.layers.") else k: v for k, v in sd.items()}
    helper = TRTLLMHelper(cfg, model_type="gpt", position_embedding_type="learned_absolute", max_position_embeddings=a.seq_length,
                          share_embeddings_and_output_weights=True, activation="gelu")
    weights, configs = helper.get_trtllm_pretrained_config_and_model_weights(sd, "bfloat16", on_device_distributed_conversion=True,
                                                                              vocab_size=a.vocab_size)
    os.makedirs(a.out, exist_ok=True)
    r = configs[0]["mapping"]["rank"]
    torch.save({k: v.cpu() for k, v in weights[0].items()}, os.path.join(a.out, f"rank{r}.pt"))
    if dist.get_rank() == 0:
        with open(os.path.join(a.out, "config.json"), "w") as f:
            json.dump(configs[0], f, indent=1)
    dist.barrier()
    print(f"rank {dist.get_rank()}: wrote rank{r}.pt ({len(weights[0])} tensors)", flush=True)


if __name__ == "__main__":
    main()
