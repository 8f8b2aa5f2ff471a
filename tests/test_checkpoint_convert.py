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
    # legacy (pre-mcore) naming round trip: core -> legacy tp2 -> core
    c = os.path.join(str(tmp_path), "legacy_tp2")
    conv.main(["--loader", "core", "--saver", "legacy", "--load-dir", b, "--save-dir", c, "--target-tensor-parallel-size", "2"])
    leg = torch.load(os.path.join(c, "release", "mp_rank_01", "model_optim_rng.pt"), weights_only=False)["model"]
    assert "layers.0.self_attention.query_key_value.weight" in leg["language_model"]["encoder"]
    from_legacy = conv.LOADERS["legacy"](c)
    assert all(torch.equal(from_legacy["weights"][k], v) for k, v in ck["weights"].items())


def test_trtllm_export_layout(tmp_path):
    """TRT-LLM export: [q|k|v] regrouping, gated-MLP split, TP sharding, config."""
    from types import SimpleNamespace

    from megatronapp_b200.export import TRTLLMHelper, export_trtllm_checkpoint

    _, d = _tiny_hf_llama(str(tmp_path))
    ck = conv.load_hf_llama(d)
    c = ck["config"]
    cfg = SimpleNamespace(num_layers=c["num_layers"], hidden_size=c["hidden_size"], num_attention_heads=c["num_attention_heads"],
                          num_query_groups=c["num_query_groups"], kv_channels=c["kv_channels"], ffn_hidden_size=c["ffn_hidden_size"],
                          gated_linear_unit=True, layernorm_epsilon=1e-5, add_bias_linear=False, num_moe_experts=None,
                          moe_router_topk=2)
    h = TRTLLMHelper(cfg, model_type="llama", activation="silu")
    out = export_trtllm_checkpoint(h, ck["weights"], str(tmp_path / "trt"), 256, export_tensor_parallel_size=2, dtype=torch.float32)
    import json

    conf = json.load(open(os.path.join(out, "config.json")))
    assert conf["architecture"] == "LlamaForCausalLM" and conf["mapping"]["tp_size"] == 2 and conf["num_key_value_heads"] == 2
    full = h.convert(ck["weights"], torch.float32)
    hf = torch.load(os.path.join(d, "pytorch_model.bin"))
    qkv = full["transformer.layers.0.attention.qkv.weight"]
    assert torch.equal(qkv[:64], hf["model.layers.0.self_attn.q_proj.weight"])
    assert torch.equal(qkv[64:96], hf["model.layers.0.self_attn.k_proj.weight"])
    assert torch.equal(full["transformer.layers.1.mlp.gate.weight"], hf["model.layers.1.mlp.up_proj.weight"])
    s0, s1 = h.shard(full, 2, 0), h.shard(full, 2, 1)
    assert s0["transformer.layers.0.attention.qkv.weight"].shape == (64, 64)
    assert torch.equal(torch.cat([s0["transformer.layers.0.mlp.proj.weight"], s1["transformer.layers.0.mlp.proj.weight"]], 1),
                       full["transformer.layers.0.mlp.proj.weight"])


def test_distillation_losses():
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.post_training import DistillationLossBalancer, LogitsKLLoss

    ps.initialize_model_parallel(1, 1)
    torch.manual_seed(0)
    s, t = torch.randn(4, 2, 32, requires_grad=True), torch.randn(4, 2, 32)
    kl = LogitsKLLoss(temperature=2.0)(s, t)
    ref = torch.nn.functional.kl_div(torch.log_softmax(s / 2, -1), torch.softmax(t / 2, -1), reduction="none").sum(-1) * 4
    assert torch.allclose(kl, ref, atol=1e-5) and float(LogitsKLLoss()(t, t).abs().max()) < 1e-6
    total = DistillationLossBalancer(kd_loss_scale=0.5)(torch.tensor(2.0), {"logits": kl})
    total.backward()
    assert s.grad is not None and abs(float(total) - (2.0 + 0.5 * float(kl.mean()))) < 1e-5


def test_hybrid_mamba_reshard_roundtrip(tmp_path):
    import importlib.util
    from types import SimpleNamespace

    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.models.mamba.mamba_model import MambaModel
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.training import checkpointing as ck

    spec = importlib.util.spec_from_file_location("hybrid_conversion", os.path.join(ROOT, "tools", "checkpoint", "hybrid_conversion.py"))
    hc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hc)
    if not ps.model_parallel_is_initialized():
        ps.init_distributed("gloo")
        ps.initialize_model_parallel(1, 1)
    cfg = TransformerConfig(num_layers=4, hidden_size=32, num_attention_heads=4, mamba_state_dim=8, mamba_head_dim=8,
                            mamba_num_groups=2, kernel_backend="torch", use_cpu_initialization=True,
                            params_dtype=torch.float32)
    m = MambaModel(cfg, vocab_size=64, max_sequence_length=16, hybrid_override_pattern="M*-M")
    sd = {k: v.detach().clone() for k, v in m.state_dict().items() if not k.endswith("_extra_state")}
    args = SimpleNamespace(num_layers=4, hidden_size=32, mamba_state_dim=8, mamba_head_dim=8, mamba_num_groups=2, swiglu=False,
                           tensor_model_parallel_size=1, pipeline_model_parallel_size=1)
    src = str(tmp_path / "tp1")
    path = ck.get_checkpoint_name(src, 0, release=True, pipeline_parallel=False, tensor_rank=0, pipeline_rank=0,
                                  expert_parallel=False, expert_rank=0)
    os.makedirs(os.path.dirname(path))
    torch.save({"model": sd, "args": args, "iteration": 0}, path)
    with open(ck.get_checkpoint_tracker_filename(src), "w") as f:
        f.write("release")
    mid, back = str(tmp_path / "tp2pp2"), str(tmp_path / "back")
    hc.main(["--load-dir", src, "--save-dir", mid, "--target-tp-size", "2", "--target-pp-size", "2"])
    shard = torch.load(os.path.join(mid, "release", "mp_rank_01_000", "model_optim_rng.pt"), weights_only=False)["model"]
    w_full = sd["decoder.layers.0.mixer.in_proj.weight"]
    z, x, B, C, dt = torch.split(w_full, [64, 64, 16, 16, 8], dim=0)
    want = torch.cat([t.chunk(2, 0)[1] for t in (z, x, B, C, dt)], 0)          # rank 1: its slice of EVERY segment
    assert torch.equal(shard["decoder.layers.0.mixer.in_proj.weight"], want)
    assert shard["decoder.layers.0.mixer.out_proj.weight"].shape == (32, 32)
    assert shard["decoder.layers.1.self_attention.linear_qkv.weight"].shape[0] * 2 == sd["decoder.layers.1.self_attention.linear_qkv.weight"].shape[0]
    assert "decoder.layers.2.mixer.A_log" not in shard                          # stage 0 holds layers 0-1 only
    hc.main(["--load-dir", mid, "--save-dir", back, "--target-tp-size", "1", "--target-pp-size", "1"])
    again = torch.load(os.path.join(back, "release", "mp_rank_00", "model_optim_rng.pt"), weights_only=False)["model"]
    assert again.keys() == sd.keys() and all(torch.equal(again[k], sd[k]) for k in sd)


@pytest.mark.parametrize("name,tol", [("fp8", 0.08), ("int8_sq", 0.05), ("int4_awq", 0.35), ("w4a8_awq", 0.4), ("nvfp4", 0.45)])
def test_post_training_quantization(name, tol):
    from megatronapp_b200.bench_model import build_gpt_for_bench
    from megatronapp_b200.post_training import quantization as ptq

    def build():
        torch.manual_seed(0)
        return build_gpt_for_bench(2, 128, 4, 32, 128, 1, device="cpu", dtype=torch.float32, rope=True).model[0].module

    m = build()
    ids = torch.randint(0, 128, (4, 32))
    pos = torch.arange(32).unsqueeze(0).expand(4, -1)
    m.eval()
    with torch.no_grad():
        ref = m(ids, pos, None)
    calib = [torch.randint(0, 128, (4, 32)) for _ in range(3)] + [ids]
    loop = lambda mod: [mod(c, pos, None) for c in calib]
    ptq.quantize(m, name, loop)
    with torch.no_grad():
        out = m(ids, pos, None)
    rel = float((out - ref).abs().max() / ref.abs().max())
    assert 0 < rel < tol, rel                                            # quantised for real, still the same model
    lin = m.decoder.layers[0].mlp.linear_fc1
    cfg = ptq.QUANT_CFG_CHOICES[name]
    w = lin.weight.detach()
    assert torch.allclose(ptq.fake_quant(w, lin.weight_quantizer._amax, cfg.weight), w, atol=1e-6)   # folded onto the grid
    levels = {"fp8": 256, "int8_sq": 256, "int4_awq": 16, "w4a8_awq": 16, "nvfp4": 15}[name]
    row = w[0, :cfg.weight.block] if cfg.weight.block else w[0]
    assert row.unique().numel() <= levels
    # checkpoint round trip: fresh model + same weights + quantiser state == quantised model
    state = ptq.quantizer_state_dict(m)
    assert any(k.endswith("weight_quantizer._amax") for k in state)
    m2 = build()
    m2.load_state_dict(m.state_dict(), strict=False)
    ptq.restore_quantizers(m2, name, state)
    m2.eval()
    with torch.no_grad():
        assert torch.allclose(m2(ids, pos, None), out, atol=1e-5)
    real = ptq.export_quantized_weights(m)
    k = "decoder.layers.0.mlp.linear_fc1"
    assert real[f"{k}.weight"].dtype in (torch.int8, torch.uint8)
    if cfg.weight.bits == 4:
        assert real[f"{k}.weight"].shape == (w.shape[0], w.shape[1] // 2)
        assert real[f"{k}.weights_scaling_factor"].shape == (w.shape[0], w.shape[1] // cfg.weight.block)
    if name == "int8_sq":
        s = real[f"{k}.weights_scaling_factor"]
        assert torch.allclose(real[f"{k}.weight"].float() * s.unsqueeze(1), w, atol=1e-5)
        assert f"{k}.prequant_scaling_factor" in real and f"{k}.activation_scaling_factor" in real


def test_hf_mixtral_to_core_logit_parity_and_expert_parallel_layout(tmp_path):
    """HF Mixtral -> core: logits match transformers; the per-rank files of an expert-parallel layout hold each rank's
    experts under local indices, merge back losslessly, and core -> HF restores the original tensors."""
    import json

    from transformers import MixtralConfig, MixtralForCausalLM

    from megatronapp_b200.config import TransformerConfig
    from megatronapp_b200.models.gpt import GPTModel
    from megatronapp_b200.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed

    torch.manual_seed(0)
    hcfg = MixtralConfig(vocab_size=256, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=64, rms_norm_eps=1e-5, num_local_experts=4,
                         num_experts_per_tok=2, tie_word_embeddings=False, attn_implementation="eager",
                         router_jitter_noise=0.0, sliding_window=None)
    hf = MixtralForCausalLM(hcfg).float().eval()
    d = str(tmp_path / "hf")
    os.makedirs(d)
    torch.save(hf.state_dict(), os.path.join(d, "pytorch_model.bin"))
    json.dump(hcfg.to_dict(), open(os.path.join(d, "config.json"), "w"))
    ck = conv.load_hf_mixtral(d)
    cfg = ck["config"]
    assert cfg["num_experts"] == 4 and "decoder.layers.1.mlp.experts.local_experts.3.linear_fc2.weight" in ck["weights"]

    if not ps.model_parallel_is_initialized():
        ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    tc = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, num_query_groups=2, ffn_hidden_size=96,
                           moe_ffn_hidden_size=96, num_moe_experts=4, moe_router_topk=2, gated_linear_unit=True,
                           normalization="RMSNorm", add_bias_linear=False, layernorm_epsilon=1e-5, kernel_backend="torch",
                           use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0,
                           activation_func=torch.nn.functional.silu, moe_token_dispatcher_type="allgather",
                           moe_router_pre_softmax=False)
    m = GPTModel(tc, get_gpt_layer_local_spec(4, False), 256, 64, position_embedding_type="rope", rotary_base=cfg["rotary_base"],
                 share_embeddings_and_output_weights=False).float().eval()
    missing = m.load_state_dict(ck["weights"], strict=False)
    assert not [k for k in missing.missing_keys if "_extra_state" not in k], missing
    assert not missing.unexpected_keys, missing.unexpected_keys
    toks = torch.randint(0, 256, (2, 16))
    pos = torch.arange(16).unsqueeze(0).expand(2, -1)
    with torch.no_grad():
        ours, ref = m(toks, pos, None), hf(toks).logits
    assert torch.allclose(ours, ref, atol=3e-4, rtol=1e-3), float((ours - ref).abs().max())

    ep_dir = str(tmp_path / "ep2")
    conv.save_core(ck, ep_dir, tp=1, pp=1, ep=2)
    r1 = torch.load(os.path.join(ep_dir, "release", "mp_rank_00_001", "model_optim_rng.pt"), weights_only=False)["model"]
    assert torch.equal(r1["decoder.layers.0.mlp.experts.local_experts.0.linear_fc1.weight"],
                       ck["weights"]["decoder.layers.0.mlp.experts.local_experts.2.linear_fc1.weight"])
    assert "decoder.layers.0.mlp.experts.local_experts.2.linear_fc1.weight" not in r1 and "decoder.layers.0.mlp.router.weight" in r1
    merged = conv.load_core(ep_dir)
    assert merged["weights"].keys() == ck["weights"].keys()
    assert all(torch.equal(merged["weights"][k], v) for k, v in ck["weights"].items())
    out = str(tmp_path / "hf_out")
    conv.save_hf_mixtral(merged, out)                              # written in the hub layout (block_sparse_moe ... w1/w2/w3)
    back = torch.load(os.path.join(out, "pytorch_model.bin"))
    assert "model.layers.1.block_sparse_moe.experts.3.w2.weight" in back
    again = conv.load_hf_mixtral(out)
    assert all(torch.equal(again["weights"][k], v) for k, v in ck["weights"].items())
    assert json.load(open(os.path.join(out, "config.json")))["num_local_experts"] == 4


def test_hf_llava_roundtrip_and_loads_into_the_core_model(tmp_path):
    """HF LLaVA layout (CLIP tower + MLP projector + Llama) -> core keys -> HF layout is the identity, and the core-side
    keys / shapes are exactly the ones ``LLaVAModel`` declares (reference loader_llava / saver_hf_llava)."""
    import importlib.util
    import json
    import os

    import torch

    spec = importlib.util.spec_from_file_location("conv", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                       "tools", "checkpoint", "convert.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    torch.manual_seed(0)
    vh, vl, vnh, vff, th, tl, tnh, tng, tff, V = 32, 2, 4, 64, 48, 2, 4, 2, 96, 128
    sd = {}
    vt = "model.vision_tower.vision_model."
    sd[vt + "embeddings.patch_embedding.weight"] = torch.randn(vh, 3, 14, 14)
    sd[vt + "embeddings.class_embedding"] = torch.randn(vh)
    sd[vt + "embeddings.position_embedding.weight"] = torch.randn(5, vh)
    for nm, shp in (("weight", (vh,)), ("bias", (vh,))):
        sd[vt + f"pre_layrnorm.{nm}"] = torch.randn(*shp)
    for i in range(vl):
        a = f"{vt}encoder.layers.{i}."
        for p in ("q", "k", "v", "out"):
            sd[a + f"self_attn.{p}_proj.weight"], sd[a + f"self_attn.{p}_proj.bias"] = torch.randn(vh, vh), torch.randn(vh)
        for ln in ("layer_norm1", "layer_norm2"):
            sd[a + ln + ".weight"], sd[a + ln + ".bias"] = torch.randn(vh), torch.randn(vh)
        sd[a + "mlp.fc1.weight"], sd[a + "mlp.fc1.bias"] = torch.randn(vff, vh), torch.randn(vff)
        sd[a + "mlp.fc2.weight"], sd[a + "mlp.fc2.bias"] = torch.randn(vh, vff), torch.randn(vh)
    sd["model.multi_modal_projector.linear_1.weight"], sd["model.multi_modal_projector.linear_1.bias"] = torch.randn(th, vh), torch.randn(th)
    sd["model.multi_modal_projector.linear_2.weight"], sd["model.multi_modal_projector.linear_2.bias"] = torch.randn(th, th), torch.randn(th)
    lm = "model.language_model."
    hn = th // tnh
    sd[lm + "embed_tokens.weight"], sd[lm + "norm.weight"], sd["lm_head.weight"] = torch.randn(V, th), torch.randn(th), torch.randn(V, th)
    for i in range(tl):
        a = f"{lm}layers.{i}."
        sd[a + "self_attn.q_proj.weight"], sd[a + "self_attn.o_proj.weight"] = torch.randn(tnh * hn, th), torch.randn(th, tnh * hn)
        sd[a + "self_attn.k_proj.weight"], sd[a + "self_attn.v_proj.weight"] = torch.randn(tng * hn, th), torch.randn(tng * hn, th)
        sd[a + "input_layernorm.weight"], sd[a + "post_attention_layernorm.weight"] = torch.randn(th), torch.randn(th)
        sd[a + "mlp.gate_proj.weight"], sd[a + "mlp.up_proj.weight"] = torch.randn(tff, th), torch.randn(tff, th)
        sd[a + "mlp.down_proj.weight"] = torch.randn(th, tff)
    src = tmp_path / "hf"
    src.mkdir()
    torch.save(sd, src / "pytorch_model.bin")
    json.dump({"model_type": "llava", "image_token_index": 32000,
               "vision_config": {"hidden_size": vh, "num_hidden_layers": vl, "num_attention_heads": vnh, "intermediate_size": vff,
                                 "patch_size": 14, "image_size": 28},
               "text_config": {"hidden_size": th, "num_hidden_layers": tl, "num_attention_heads": tnh, "num_key_value_heads": tng,
                               "intermediate_size": tff, "vocab_size": V, "max_position_embeddings": 64, "rope_theta": 10000.0,
                               "rms_norm_eps": 1e-5}}, open(src / "config.json", "w"))
    ck = conv.load_hf_llava(str(src))
    w = ck["weights"]
    assert w["vision_model.decoder.layers.0.self_attention.linear_qkv.weight"].shape == (3 * vh, vh)
    assert w["vision_model.decoder.layers.1.self_attention.linear_qkv.bias"].shape == (3 * vh,)
    assert w["language_model.decoder.layers.0.self_attention.linear_qkv.weight"].shape == ((tnh + 2 * tng) * hn, th)
    assert w["language_model.decoder.layers.0.mlp.linear_fc1.weight"].shape == (2 * tff, th)
    assert w["vision_projection.encoder.linear_fc2.weight"].shape == (th, th) and ck["config"]["image_token_index"] == 32000
    dst = tmp_path / "back"
    conv.save_hf_llava(ck, str(dst))
    back = torch.load(dst / "pytorch_model.bin")
    assert set(back) == set(sd)
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def _trt_cfg(c, **kw):
    from types import SimpleNamespace

    base = dict(num_layers=c["num_layers"], hidden_size=c["hidden_size"], num_attention_heads=c["num_attention_heads"],
                num_query_groups=c["num_query_groups"], kv_channels=c["kv_channels"], ffn_hidden_size=c["ffn_hidden_size"],
                gated_linear_unit=True, layernorm_epsilon=1e-5, add_bias_linear=False, num_moe_experts=None, moe_router_topk=2,
                layernorm_zero_centered_gamma=False, normalization="RMSNorm")
    base.update(kw)
    return SimpleNamespace(**base)


def _megatron_tp_shard(sd, tp, r):
    """this framework's tensor-parallel layout of a canonical state dict (what a training rank holds)"""
    out = {}
    for k, w in sd.items():
        if k.endswith("linear_qkv.weight") or k.endswith("word_embeddings.weight") or k.endswith("output_layer.weight"):
            out[k] = torch.chunk(w, tp, 0)[r].clone()
        elif k.endswith("linear_fc1.weight"):
            gate, up = torch.chunk(w, 2, 0)
            out[k] = torch.cat([torch.chunk(gate, tp, 0)[r], torch.chunk(up, tp, 0)[r]], 0)
        elif k.endswith(("linear_proj.weight", "linear_fc2.weight")):
            out[k] = torch.chunk(w, tp, 1)[r].clone()
        else:
            out[k] = w
    return out


def test_trtllm_on_device_distributed_conversion_matches_single_device_export(tmp_path):
    """Every (tp, pp) training rank converts only its own shard; the result equals the slice the single-device exporter
    produces for that rank (global layer numbering, vocabulary padding removed, head re-split)."""
    from megatronapp_b200.export import DistributedTRTLLMConverter, ExportConfig, TRTLLMHelper

    _, d = _tiny_hf_llama(str(tmp_path))
    ck = conv.load_hf_llama(d)
    sd = {k: v.float() for k, v in ck["weights"].items()}
    h = TRTLLMHelper(_trt_cfg(ck["config"]), model_type="llama", activation="silu")
    tok_vocab = 250                                                 # the table is padded to 256
    ref_w, ref_c = h.get_trtllm_pretrained_config_and_model_weights(sd, "float32", ExportConfig(2, 2), vocab_size=tok_vocab)
    assert len(ref_w) == 4 and [c["mapping"]["rank"] for c in ref_c] == [0, 1, 2, 3]
    assert "transformer.vocab_embedding.weight" in ref_w[0] and "lm_head.weight" not in ref_w[0] and "lm_head.weight" in ref_w[3]
    assert ref_w[0]["transformer.vocab_embedding.weight"].shape[0] == tok_vocab
    L = ck["config"]["num_layers"]
    for pr in range(2):
        for tr in range(2):
            local = _megatron_tp_shard(sd, 2, tr)
            lo, hi = pr * L // 2, (pr + 1) * L // 2
            stage = {}
            for k, v in local.items():
                if k.startswith("decoder.layers."):
                    if lo <= int(k.split(".")[2]) < hi:
                        stage[k] = v
                elif k.startswith("embedding.") and pr == 0:
                    stage[k] = v
                elif (k.startswith("decoder.final_layernorm") or k.startswith("output_layer")) and pr == 1:
                    stage[k] = v

            def fake_all_reduce(t, tr=tr):                         # the peer's half of the vocabulary-parallel table
                other = _megatron_tp_shard(sd, 2, 1 - tr)
                src = other["embedding.word_embeddings.weight"] if t.shape[0] == sd["embedding.word_embeddings.weight"].shape[0] and pr == 0 \
                    else other["output_layer.weight"]
                n = src.shape[0]
                t[(1 - tr) * n:(2 - tr) * n] += src

            dc = DistributedTRTLLMConverter(h, tp_rank=tr, tp_size=2, pp_rank=pr, pp_size=2, tp_all_reduce=fake_all_reduce)
            got = dc.convert(stage, torch.float32, tokenizer_vocab_size=tok_vocab)
            want = ref_w[pr * 2 + tr]
            assert set(got) == set(want), (pr, tr, set(got) ^ set(want))
            for k in want:
                assert torch.equal(got[k], want[k]), (pr, tr, k)


def test_trtllm_moe_fp8_and_engine_plan(tmp_path):
    from megatronapp_b200.export import TRTLLMEngineBuilder, TRTLLMHelper

    torch.manual_seed(0)
    H, F, E, L = 16, 24, 4, 1
    cfg = _trt_cfg(dict(num_layers=L, hidden_size=H, num_attention_heads=2, num_query_groups=2, kv_channels=8, ffn_hidden_size=F),
                   num_moe_experts=E, layernorm_zero_centered_gamma=True, normalization="LayerNorm")
    sd = {"embedding.word_embeddings.weight": torch.randn(32, H), "decoder.final_layernorm.weight": torch.zeros(H),
          "decoder.layers.0.input_layernorm.weight": torch.zeros(H), "decoder.layers.0.pre_mlp_layernorm.weight": torch.zeros(H),
          "decoder.layers.0.self_attention.linear_qkv.weight": torch.randn(3 * H, H),
          "decoder.layers.0.self_attention.linear_proj.weight": torch.randn(H, H),
          "decoder.layers.0.mlp.router.weight": torch.randn(E, H)}
    for e in range(E):
        sd[f"decoder.layers.0.mlp.experts.local_experts.{e}.linear_fc1.weight"] = torch.randn(2 * F, H)
        sd[f"decoder.layers.0.mlp.experts.local_experts.{e}.linear_fc2.weight"] = torch.randn(H, F)
    h = TRTLLMHelper(cfg, model_type="mixtral", activation="silu")
    w = h.convert(sd, torch.float32)
    assert torch.equal(w["transformer.layers.0.input_layernorm.weight"], torch.ones(H))           # zero-centred gamma + 1
    fc, proj = w["transformer.layers.0.mlp.fc.weight"], w["transformer.layers.0.mlp.proj.weight"]
    assert fc.shape == (E, 2 * F, H) and proj.shape == (E, H, F) and w["transformer.layers.0.mlp.router.weight"].dtype == torch.float32
    gate, up = torch.chunk(sd["decoder.layers.0.mlp.experts.local_experts.2.linear_fc1.weight"], 2, 0)
    assert torch.equal(fc[2, :F], up) and torch.equal(fc[2, F:], gate)                             # MoE plugin order: [up ; gate]
    s1 = h.shard(w, 2, 1)
    assert torch.equal(s1["transformer.layers.0.mlp.fc.weight"][2, :F // 2], up[F // 2:]) and s1["transformer.layers.0.mlp.proj.weight"].shape == (E, H, F // 2)
    # the grouped-GEMM parameter layout converts to the same tensors
    g = {k: v for k, v in sd.items() if "local_experts" not in k}
    g["decoder.layers.0.mlp.experts.weight1"] = torch.stack([sd[f"decoder.layers.0.mlp.experts.local_experts.{e}.linear_fc1.weight"].t() for e in range(E)])
    g["decoder.layers.0.mlp.experts.weight2"] = torch.stack([sd[f"decoder.layers.0.mlp.experts.local_experts.{e}.linear_fc2.weight"].t() for e in range(E)])
    wg = h.convert(g, torch.float32)
    assert torch.equal(wg["transformer.layers.0.mlp.fc.weight"], fc) and torch.equal(wg["transformer.layers.0.mlp.proj.weight"], proj)
    # fp8: weights quantised with per-tensor scales that reproduce them; config advertises the quant mode
    ws, cs = h.get_trtllm_pretrained_config_and_model_weights(sd, "bfloat16", fp8_quantized=True, fp8_kvcache=True)
    q = ws[0]["transformer.layers.0.attention.qkv.weight"]
    assert q.dtype == torch.float8_e4m3fn and cs[0]["quantization"] == {"quant_algo": "FP8", "kv_cache_quant_algo": "FP8"}
    deq = q.float() * ws[0]["transformer.layers.0.attention.qkv.weights_scaling_factor"]
    assert (deq - w["transformer.layers.0.attention.qkv.weight"]).abs().max() < 0.07 * w["transformer.layers.0.attention.qkv.weight"].abs().max()
    assert ws[0]["transformer.layers.0.attention.kv_cache_scaling_factor"].item() == 1.0
    # engine plan: token limits follow TRT-LLM's defaulting; the plan is written even though tensorrt_llm is absent
    conf = cs[0]
    plan = TRTLLMEngineBuilder.plan(conf, max_input_len=512, max_output_len=256, max_batch_size=8, paged_context_fmha=True, tokens_per_block=64)
    assert plan.max_seq_len == 768 and plan.max_num_tokens == 8 * 768 and plan.opt_num_tokens == 8 and plan.plugin["tokens_per_block"] == 64
    assert TRTLLMEngineBuilder.plan(conf, remove_input_padding=False).max_num_tokens is None
    try:
        import tensorrt_llm  # noqa: F401
    except ImportError:
        import json

        import pytest

        with pytest.raises(RuntimeError, match="build_config.json"):
            h.build_and_save_engine(str(tmp_path / "engine"), ws[0], conf, max_batch_size=2)
        saved = json.load(open(tmp_path / "engine" / "build_config.json"))
        assert saved["build"]["max_batch_size"] == 2 and saved["plugin_config"]["paged_kv_cache"] is True
