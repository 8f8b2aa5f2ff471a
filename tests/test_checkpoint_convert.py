"""tools/checkpoint/convert.py: HF Llama <-> core (logit parity with transformers) and TP/PP resharding."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ckpt_convert", os.path.join(ROOT, "tools", "checkpoint", "convert.py"))
conv = importlib.util.module_from_spec(spec)
spec.loader.exec_module(conv)


def _tiny_hf_llama(tmp):
    from transformers import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=256, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=64, rms_norm_eps=1e-5, tie_word_embeddings=False,
                      attn_implementation="eager")
    m = LlamaForCausalLM(cfg).float().eval()
    d = os.path.join(tmp, "hf")
    os.makedirs(d)
    torch.save(m.state_dict(), os.path.join(d, "pytorch_model.bin"))
    import json

    json.dump(cfg.to_dict(), open(os.path.join(d, "config.json"), "w"))
    return m, d


def _build_core_model(cfg):
    from megatronapp_b200.config import TransformerConfig
    from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    tc = TransformerConfig(num_layers=cfg["num_layers"], hidden_size=cfg["hidden_size"],
                           num_attention_heads=cfg["num_attention_heads"], num_query_groups=cfg["num_query_groups"],
                           ffn_hidden_size=cfg["ffn_hidden_size"], gated_linear_unit=True, normalization="RMSNorm",
                           add_bias_linear=False, layernorm_epsilon=cfg["layernorm_epsilon"], kernel_backend="torch", use_cpu_initialization=True,
                           hidden_dropout=0.0, attention_dropout=0.0, activation_func=torch.nn.functional.silu)
    return GPTModel(tc, get_gpt_layer_local_spec(), cfg["padded_vocab_size"], cfg["max_position_embeddings"],
                    position_embedding_type="rope", rotary_base=cfg["rotary_base"],
                    share_embeddings_and_output_weights=not cfg["untie_embeddings_and_output_weights"]).float().eval()


def test_hf_llama_to_core_logit_parity(tmp_path):
    hf, d = _tiny_hf_llama(str(tmp_path))
    ck = conv.load_hf_llama(d)
    m = _build_core_model(ck["config"])
    missing = m.load_state_dict(ck["weights"], strict=False)
    assert not [k for k in missing.missing_keys if "_extra_state" not in k], missing
    toks = torch.randint(0, 256, (2, 16))
    pos = torch.arange(16).unsqueeze(0).expand(2, -1)
    with torch.no_grad():
        ours = m(toks, pos, None)
        ref = hf(toks).logits
    assert torch.allclose(ours, ref, atol=2e-4, rtol=1e-3), float((ours - ref).abs().max())
    # and back: core -> HF reproduces the original tensors exactly
    out = os.path.join(str(tmp_path), "hf_out")
    conv.save_hf_llama(ck, out)
    back = torch.load(os.path.join(out, "pytorch_model.bin"))
    for k, v in hf.state_dict().items():
        assert torch.equal(back[k], v), k


def test_core_reshard_roundtrip(tmp_path):
    _, d = _tiny_hf_llama(str(tmp_path))
    ck = conv.load_hf_llama(d)
    a = os.path.join(str(tmp_path), "tp2pp2")
    conv.save_core(ck, a, tp=2, pp=2)
    assert os.path.isdir(os.path.join(a, "release", "mp_rank_01_001"))
    shard = torch.load(os.path.join(a, "release", "mp_rank_01_001", "model_optim_rng.pt"), weights_only=False)["model"]
    assert shard["decoder.layers.0.mlp.linear_fc1.weight"].shape == (176, 64)        # [gate/2 ; up/2]
    assert "embedding.word_embeddings.weight" not in shard and "output_layer.weight" in shard
    merged = conv.load_core(a)
    for k, v in ck["weights"].items():
        assert torch.equal(merged["weights"][k], v), k
    b = os.path.join(str(tmp_path), "tp1")
    conv.main(["--loader", "core", "--saver", "core", "--load-dir", a, "--save-dir", b])
    again = conv.load_core(b)
    assert all(torch.equal(again["weights"][k], v) for k, v in ck["weights"].items())
