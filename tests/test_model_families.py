"""BERT, T5 and Mamba/hybrid models build, run forward+backward and learn on CPU."""
import pytest
import torch
import torch.nn.functional as F

from megatronapp_b200.config.transformer_config import TransformerConfig
from megatronapp_b200.parallel import state as ps
from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed


def _cfg(**kw):
    base = dict(num_layers=2, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True,
                hidden_dropout=0.0, attention_dropout=0.0)
    base.update(kw)
    return TransformerConfig(**base)


def _init():
    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    torch.manual_seed(0)


def _train(model, loss_fn, steps=8, lr=3e-3):
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        l = loss_fn()
        l.backward()
        opt.step()
        losses.append(float(l))
    return losses


def test_bert_mlm_and_binary_head():
    from megatronapp_b200.models.bert import BertModel, bert_layer_local_spec

    _init()
    m = BertModel(_cfg(), num_tokentypes=2, transformer_layer_spec=bert_layer_local_spec(), vocab_size=64,
                  max_sequence_length=16)
    ids = torch.randint(0, 64, (3, 16))
    mask = torch.ones(3, 16)
    mask[1, 10:] = 0
    types = torch.zeros(3, 16, dtype=torch.long)

    def loss():
        lm, binary = m(ids, mask, tokentype_ids=types, lm_labels=ids)
        return (lm * mask).sum() / mask.sum() + F.cross_entropy(binary.float(), torch.tensor([0, 1, 0]))

    ls = _train(m, loss)
    assert ls[-1] < ls[0]
    # padded keys must not influence real tokens
    m.eval()
    with torch.no_grad():
        a = m(ids, mask, tokentype_ids=types)[0][1, :10]
        ids2 = ids.clone(); ids2[1, 10:] = 7
        b = m(ids2, mask, tokentype_ids=types)[0][1, :10]
    assert torch.allclose(a, b, atol=1e-5)


def test_t5_encoder_decoder():
    from megatronapp_b200.models.T5 import T5Model, get_t5_decoder_block_spec, get_t5_encoder_block_spec

    _init()
    cfg = _cfg()
    m = T5Model(cfg, cfg, get_t5_encoder_block_spec(2), get_t5_decoder_block_spec(2), vocab_size=64, max_sequence_length=16)
    enc, dec = torch.randint(0, 64, (2, 12)), torch.randint(0, 64, (2, 8))
    em = torch.ones(2, 12, 12)
    dm = torch.tril(torch.ones(8, 8)).unsqueeze(0).expand(2, -1, -1)
    edm = torch.ones(2, 8, 12)
    ls = _train(m, lambda: m(enc, dec, em, dm, edm, lm_labels=dec).mean())
    assert ls[-1] < ls[0]
    assert any("cross_attention" in n for n, _ in m.named_parameters())


def test_mamba_ssd_scan_matches_sequential_recurrence():
    from megatronapp_b200.ssm.mamba import ssd_scan

    torch.manual_seed(0)
    b, l, h, p, g, n = 2, 37, 4, 8, 2, 5
    x, dt = torch.randn(b, l, h, p), torch.rand(b, l, h) * 0.5 + 0.01
    A, B, C, D = -torch.rand(h) - 0.1, torch.randn(b, l, g, n), torch.randn(b, l, g, n), torch.randn(h)
    y, _ = ssd_scan(x, dt, A, B, C, D, chunk=8)
    S = torch.zeros(b, h, p, n)
    ref = []
    Bh, Ch = B.repeat_interleave(h // g, 2), C.repeat_interleave(h // g, 2)
    for t in range(l):
        S = S * torch.exp(dt[:, t] * A).view(b, h, 1, 1) + torch.einsum("bhp,bhn->bhpn", x[:, t] * dt[:, t].unsqueeze(-1), Bh[:, t])
        ref.append(torch.einsum("bhpn,bhn->bhp", S, Ch[:, t]) + x[:, t] * D.view(1, h, 1))
    assert torch.allclose(y, torch.stack(ref, 1), atol=1e-4)


def test_hybrid_mamba_model_and_allocation():
    from megatronapp_b200.models.mamba import MambaModel
    from megatronapp_b200.ssm.mamba import allocate_layers

    assert allocate_layers(4, 0.0, 0.0, "M*-M") == ["M", "*", "-", "M"]
    assert allocate_layers(8, 0.25, 0.25).count("*") == 2 and allocate_layers(8, 0.25, 0.25).count("-") == 2
    _init()
    cfg = _cfg(num_layers=3, mamba_state_dim=8, mamba_head_dim=8, mamba_num_groups=2)
    m = MambaModel(cfg, vocab_size=64, max_sequence_length=16, hybrid_override_pattern="M*-")
    ids = torch.randint(0, 64, (2, 16))
    pos = torch.arange(16).unsqueeze(0).expand(2, -1)
    ls = _train(m, lambda: m(ids, pos, labels=ids).mean(), steps=6)
    assert ls[-1] < ls[0]


def test_gpt_multi_token_prediction():
    """MTP heads add an auxiliary loss whose gradient reaches both the MTP layers and the trunk."""
    from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec
    from megatronapp_b200.transformer.multi_token_prediction import MTPLossLoggingHelper

    _init()
    cfg = _cfg(mtp_num_layers=2, mtp_loss_scaling_factor=0.3, kernel_backend="torch")
    m = GPTModel(cfg, get_gpt_layer_local_spec(), 64, 16, position_embedding_type="rope")
    toks = torch.randint(0, 64, (2, 16))
    pos = torch.arange(16).unsqueeze(0).expand(2, -1)
    labels = torch.roll(toks, -1, 1)
    m.train()
    loss = m(toks, pos, None, labels=labels).mean()
    loss.backward()
    g = [p.grad for n, p in m.named_parameters() if n.startswith("mtp.")]
    assert g and all(x is not None and torch.isfinite(x).all() for x in g) and any(float(x.abs().sum()) > 0 for x in g)
    assert float(MTPLossLoggingHelper.tracker["values"].sum()) > 0
    # trunk gradient differs from the no-MTP gradient (aux loss flows into the decoder)
    g_with = m.decoder.layers[0].mlp.linear_fc1.weight.grad.clone()
    m.zero_grad()
    m.eval()
    m(toks, pos, None, labels=labels).mean().backward()
    assert not torch.allclose(g_with, m.decoder.layers[0].mlp.linear_fc1.weight.grad)


def test_gpt_multi_latent_attention():
    """MLA GPT trains, and incremental decoding through the MLA KV cache matches a full forward."""
    from megatronapp_b200.config.transformer_config import MLATransformerConfig
    from megatronapp_b200.inference.core import StaticInferenceContext as InferenceContext
    from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec

    _init()
    cfg = MLATransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True,
                               hidden_dropout=0.0, attention_dropout=0.0, q_lora_rank=16, kv_lora_rank=16,
                               qk_head_dim=8, qk_pos_emb_head_dim=4, v_head_dim=8, max_position_embeddings=32,
                               rotary_scaling_factor=1.0, kernel_backend="torch")
    m = GPTModel(cfg, get_gpt_layer_local_spec(multi_latent_attention=True, qk_layernorm=True), 64, 32,
                 position_embedding_type="rope")
    toks = torch.randint(0, 64, (2, 12))
    pos = torch.arange(12).unsqueeze(0).expand(2, -1)
    losses = _train(m, lambda: m(toks, pos, None, labels=torch.roll(toks, -1, 1)).mean(), steps=10, lr=1e-2)
    assert losses[-1] < losses[0]
    m.eval()
    with torch.no_grad():
        full = m(toks, pos, None)
        ctx = InferenceContext(max_batch_size=2, max_sequence_length=32)
        a = m(toks[:, :8], pos[:, :8], None, inference_context=ctx)
        ctx.sequence_len_offset = 8
        outs = [a]
        for t in range(8, 12):
            outs.append(m(toks[:, t:t + 1], pos[:, t:t + 1], None, inference_context=ctx))
            ctx.sequence_len_offset += 1
    inc = torch.cat(outs, dim=1)
    assert torch.allclose(inc, full, atol=1e-4), float((inc - full).abs().max())


def test_llava_splices_image_embeddings():
    """LLaVA: image tokens are replaced by projected ViT embeddings; labels/loss mask are rebuilt."""
    import copy

    from megatronapp_b200.models.gpt import get_gpt_layer_local_spec
    from megatronapp_b200.models.multimodal import DEFAULT_IMAGE_TOKEN_INDEX, LLaVAModel
    from megatronapp_b200.models.vision import get_num_image_embeddings, get_vit_layer_spec
    from megatronapp_b200.parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
    from megatronapp_b200.transformer.mlp import MLPSubmodules

    _init()
    lcfg = _cfg(kernel_backend="torch")
    vcfg = copy.deepcopy(lcfg)
    vcfg.num_layers = 1
    pcfg = copy.deepcopy(lcfg)
    pcfg.ffn_hidden_size = lcfg.hidden_size
    m = LLaVAModel(lcfg, get_gpt_layer_local_spec(), 64, 48, vcfg, get_vit_layer_spec(), True, pcfg,
                   MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear), img_h=14, img_w=14,
                   patch_dim=7)
    n_img = get_num_image_embeddings(14, 14, 7, "clip", True)
    assert n_img == 4
    toks = torch.randint(1, 64, (2, 10))
    toks[:, 2] = DEFAULT_IMAGE_TOKEN_INDEX
    labels = torch.randint(1, 64, (2, 10))
    pos = torch.arange(10).unsqueeze(0).expand(2, -1)
    loss, mask = m(torch.rand(2, 3, 14, 14), toks, pos, None, labels=labels, loss_mask=torch.ones(2, 10))
    assert loss.shape == (2, 10 - 1 + n_img) and mask.shape == loss.shape
    assert mask[:, 2:2 + n_img].sum() == 0 and mask[:, 1].sum() == 0 and mask[:, 0].sum() == 2
    (loss * mask).sum().backward()
    assert m.vision_model.conv1.weight.grad.abs().sum() > 0 and m.vision_projection.encoder.linear_fc1.weight.grad is not None
    # RADIO tower: 8 summary tokens kept in front of the 4 patch embeddings
    r = LLaVAModel(lcfg, get_gpt_layer_local_spec(), 64, 48, vcfg, get_vit_layer_spec(), False, pcfg,
                   MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear), img_h=14, img_w=14,
                   patch_dim=7, vision_model_type="radio", max_img_h=28, max_img_w=28)
    assert r.img_seq_len == 12
    loss, mask = r(torch.rand(2, 3, 14, 14), toks, pos, None, labels=labels, loss_mask=torch.ones(2, 10))
    assert loss.shape == (2, 10 - 1 + 12)


def test_fp8_linear_emulation_matches_bf16_within_quantisation_error():
    """fp8 quantise / GEMM / linear (CPU emulation of the device arithmetic)."""
    from megatronapp_b200.ops import fp8

    torch.manual_seed(0)
    x = torch.randn(64, 128, requires_grad=True)
    w = (torch.randn(96, 128) * 0.1).requires_grad_(True)
    b = torch.randn(96)
    q, qt, sinv, am = fp8.quantize(x.detach(), transpose=True)
    assert q.dtype == torch.uint8 and qt.shape == (128, 64) and abs(float(am) - float(x.abs().max())) < 1e-6
    assert (fp8.dequantize(q, sinv) - x.detach()).abs().max() < 0.07 * x.detach().abs().max()
    y = fp8.fp8_linear(x, w, b)
    ref = x @ w.t() + b
    assert ((y - ref).abs().max() / ref.abs().max()) < 0.06
    dy = torch.randn_like(ref)
    y.backward(dy)
    gx, gw = x.grad.clone(), w.grad.clone()
    x.grad = None
    w.grad = None
    ref.backward(dy)
    assert ((gx - x.grad).abs().max() / x.grad.abs().max()) < 0.08
    assert ((gw - w.grad).abs().max() / w.grad.abs().max()) < 0.08


def test_retro_chunked_cross_attention_is_causal_and_trains():
    """Retro: neighbours of chunk i may only influence tokens from the last position of chunk i onwards."""
    from megatronapp_b200.models.retro import RetroConfig, RetroModel

    _init()
    cfg = RetroConfig(num_layers=2, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True, hidden_dropout=0.0,
                      attention_dropout=0.0, retro_chunk_length=4, retro_num_neighbors=2, retro_retrieved_length=6,
                      retro_encoder_num_layers=1, retro_encoder_hidden_dropout=0.0, retro_encoder_attention_dropout=0.0,
                      kernel_backend="torch")
    assert cfg.retro_layer_numbers() == [1]
    m = RetroModel(cfg, vocab_size=64, max_sequence_length=16)
    b, s, mch, k, r = 2, 16, 4, 2, 6
    l = s // mch
    toks = torch.randint(0, 64, (b, s))
    pos = torch.arange(s).unsqueeze(0).expand(b, -1)
    nb = torch.randint(0, 64, (b, l, k, r))
    m.eval()
    with torch.no_grad():
        base = m(toks, pos, None, context_input_ids=nb)
        nb2 = nb.clone()
        nb2[:, 2] = torch.randint(0, 64, (b, k, r))                 # change the neighbours of chunk 2 (tokens 8..11)
        out = m(toks, pos, None, context_input_ids=nb2)
    changed = (out - base).abs().amax(dim=(0, 2)) > 1e-6            # per position
    assert not changed[: 2 * mch + mch - 1].any()                   # nothing before the last token of chunk 2
    assert changed[2 * mch + mch - 1:].any()
    m.train()
    losses = _train(m, lambda: m(toks, pos, None, context_input_ids=nb, labels=torch.roll(toks, -1, 1)).mean(), steps=10, lr=1e-2)
    assert losses[-1] < losses[0]
    assert m.retro_encoder.layers[0].cross_attention.linear_q.weight.grad is not None


def test_heterogeneous_gpt_layers():
    """Per-layer NAS configs: skipped attention, linear-replaced MLP, per-layer GQA / FFN sizes."""
    from megatronapp_b200.models.gpt import GPTModel
    from megatronapp_b200.models.gpt.heterogeneous import get_gpt_heterogeneous_layer_spec, parse_heterogeneous_config

    _init()
    cfg = _cfg(num_layers=4, kernel_backend="torch")
    blocks = parse_heterogeneous_config({"block_configs": [
        {"attention": {"n_heads_in_group": 2}, "ffn": {"intermediate_size": 64}},
        {"attention": {"no_op": True}, "ffn": {"ffn_mult": 1.5}},
        {"attention": {"replace_with_linear": True}, "ffn": {"replace_with_linear": True}},
        {"attention": {}, "ffn": {"no_op": True}}]}, cfg.num_attention_heads, cfg.hidden_size)
    assert blocks[0].num_query_groups == 2 and blocks[1].attn_no_op and blocks[1].ffn_hidden_size == 256
    m = GPTModel(cfg, get_gpt_heterogeneous_layer_spec(cfg, blocks), 64, 16)
    L = m.decoder.layers
    assert L[0].mlp.linear_fc1.weight.shape[0] == 64 and L[0].self_attention.linear_qkv.weight.shape[0] == 32 + 2 * 16
    assert not any(p.requires_grad for p in L[1].self_attention.parameters()) and L[1].mlp.linear_fc1.weight.shape[0] == 256
    assert not list(L[3].mlp.parameters())
    toks = torch.randint(0, 64, (2, 16))
    pos = torch.arange(16).unsqueeze(0).expand(2, -1)
    losses = _train(m, lambda: m(toks, pos, None, labels=torch.roll(toks, -1, 1)).mean(), steps=8, lr=1e-2)
    assert losses[-1] < losses[0]


def test_bert_embedder_mean_pools_non_padding():
    import importlib.util
    import os

    from megatronapp_b200.models.bert.bert_model import BertModel, bert_layer_local_spec

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bert_embed", os.path.join(root, "tools", "bert_embedding", "embed.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _init()
    m = BertModel(_cfg(kernel_backend="torch"), 2, bert_layer_local_spec(), 64, 16)
    emb = mod.BertEmbedder(m, batch_size=2)
    ids = torch.randint(1, 64, (3, 16))
    ids[0, 10:] = 0
    e = emb(ids)
    assert e.shape == (3, 32) and torch.isfinite(e).all()
    ids2 = ids.clone()
    ids2[0, 10:] = 0                      # identical -> identical embedding; padding content is masked out
    assert torch.allclose(emb(ids2)[0], e[0], atol=1e-6)


def test_relative_position_embedding_buckets_and_t5_relative():
    from megatronapp_b200.models.common.embeddings import RelativePositionEmbedding
    from megatronapp_b200.models.T5 import T5Model, get_t5_decoder_block_spec, get_t5_encoder_block_spec

    init = lambda w: torch.nn.init.normal_(w, std=0.5)
    bi = RelativePositionEmbedding(True, init, 4, 32, 128)
    rel = torch.tensor([0, 1, -1, 7, 8, 20, 60, 128, 500, -500])
    b = bi.bucket(rel)
    # bidirectional: 16 buckets per sign, 8 exact, log-spaced afterwards, saturating at max_distance
    assert b.tolist()[:5] == [0, 17, 1, 23, 24] and b[6] < b[7] == b[8] == 31 and b[9] == 15
    uni = RelativePositionEmbedding(False, init, 4, 32, 128)
    u = uni.bucket(torch.tensor([3, 0, -1, -15, -16, -1000]))
    assert u.tolist()[:4] == [0, 0, 1, 15] and u[4] == 16 and u[5] == 31
    bias = bi(6, 9)
    assert bias.shape == (1, 4, 6, 9)
    assert torch.equal(bias[0, :, 2, 5], bias[0, :, 0, 3])        # depends on the offset only

    _init()
    cfg = _cfg()
    m = T5Model(cfg, cfg, get_t5_encoder_block_spec(2), get_t5_decoder_block_spec(2), vocab_size=64, max_sequence_length=16,
                position_embedding_type="relative")
    assert not any("position_embeddings" in n for n, _ in m.named_parameters())
    enc, dec = torch.randint(0, 64, (2, 12)), torch.randint(0, 64, (2, 8))
    em = torch.ones(2, 12, 12)
    dm = torch.tril(torch.ones(8, 8)).unsqueeze(0).expand(2, -1, -1)
    edm = torch.ones(2, 8, 12)
    ls = _train(m, lambda: m(enc, dec, em, dm, edm, lm_labels=dec).mean())
    assert ls[-1] < ls[0]
    assert m.encoder_relative_pos_emb.relative_attention_bias.weight.grad is not None


def test_linear_with_frozen_weight_grads_input_only():
    from megatronapp_b200.parallel.tp_layers import linear_with_frozen_weight

    _init()
    x = torch.randn(5, 2, 16, requires_grad=True)
    w, b = torch.randn(24, 16), torch.randn(24)
    y = linear_with_frozen_weight(x, w, b)
    y.square().sum().backward()
    xr = x.detach().clone().requires_grad_()
    torch.nn.functional.linear(xr, w, b).square().sum().backward()
    assert torch.allclose(x.grad, xr.grad, atol=1e-4) and w.grad is None
    with pytest.raises(RuntimeError):
        linear_with_frozen_weight(x, w.requires_grad_(), b)


def test_legacy_model_shells_and_state_dict_naming():
    from types import SimpleNamespace

    from megatronapp_b200.legacy.model import BertModel as LBert, GPTModel as LGPT
    from megatronapp_b200.legacy.state_dict import core_to_legacy, legacy_to_core

    _init()
    args = SimpleNamespace(padded_vocab_size=64, max_position_embeddings=16, position_embedding_type="learned_absolute",
                           untie_embeddings_and_output_weights=False, rotary_percent=1.0, rotary_base=10000)
    torch.manual_seed(0)
    m = LGPT(_cfg(), args=args)
    sd = m.state_dict_for_save_checkpoint()
    enc = sd["language_model"]["encoder"]
    assert "weight" in sd["language_model"]["embedding"]["word_embeddings"]
    for k in ("layers.0.input_norm.weight", "layers.1.self_attention.query_key_value.weight", "layers.0.self_attention.dense.bias",
              "layers.0.post_attention_norm.weight", "layers.1.mlp.dense_h_to_4h.weight", "layers.1.mlp.dense_4h_to_h.weight",
              "final_norm.weight"):
        assert k in enc, k
    # rename is a bijection on the core keys
    core_sd = {k: v for k, v in m.language_model.state_dict().items() if not k.endswith("_extra_state")}
    back = legacy_to_core(core_to_legacy(core_sd))
    assert back.keys() == core_sd.keys() and all(back[k] is core_sd[k] for k in back)
    # a second, differently initialised shell reproduces the logits after loading the legacy dict
    torch.manual_seed(1)
    m2 = LGPT(_cfg(), args=args)
    ids = torch.randint(0, 64, (2, 8))
    pos = torch.arange(8).unsqueeze(0).expand(2, -1)
    m.eval(), m2.eval()
    with torch.no_grad():
        a = m(ids, pos, None)
        assert not torch.allclose(a, m2(ids, pos, None))
        m2.load_state_dict(sd)
        assert torch.allclose(a, m2(ids, pos, None))
    lb = LBert(_cfg(kernel_backend="torch"), args=args)
    bsd = lb.state_dict_for_save_checkpoint()
    assert "layers.0.self_attention.query_key_value.weight" in bsd["language_model"]["encoder"]
    assert "lm_head" in bsd and "binary_head" in bsd and "pooler" in bsd["language_model"]
    lb.load_state_dict(bsd)


def test_mxfp8_block_scaling_quantisation_and_linear():
    from megatronapp_b200.ops import fp8

    torch.manual_seed(0)
    x = torch.randn(64, 128)
    x[:, 32:64] *= 1.0e6                       # a hot block: block scaling keeps the neighbours alive
    x[3, 96:128] = 0.0
    q, s = fp8.quantize_mx(x)
    assert q.shape == (64, 128) and s.shape == (64, 4) and s.dtype == torch.uint8
    deq = fp8.dequantize_mx(q, s)
    blk = lambda t: t.view(64, 4, 32)
    rel = (blk(deq) - blk(x)).abs().amax(-1) / blk(x).abs().amax(-1).clamp_min(1e-30)
    assert float(rel.max()) < 2 ** -3                                  # e4m3: 3 mantissa bits, every block on its own
    assert torch.equal(deq[3, 96:], torch.zeros(32)) and int(s[3, 3]) == 0
    # power-of-two scales: the largest element of a block lands in [2^8, 2^9) (e4m3 emax = 8)
    top = (blk(x).abs().amax(-1) * torch.exp2(127.0 - s.float()))[s > 0]
    assert float(top.min()) >= 256.0 and float(top.max()) < 512.0
    tens, _, sinv, _ = fp8.quantize(x)
    err_tensorwise = (fp8.dequantize(tens, sinv) - x)[:, :32].abs().max()
    assert (deq - x)[:, :32].abs().max() < 0.1 * err_tensorwise          # what the per-tensor scale loses next to the hot block

    rec = fp8.Fp8Recipe(fmt="hybrid", recipe="mxfp8")
    xx = torch.randn(64, 128, requires_grad=True)
    w = (torch.randn(96, 128) * 0.1).requires_grad_(True)
    b = torch.randn(96)
    y = fp8.fp8_linear(xx, w, b, rec)
    ref = xx @ w.t() + b
    assert ((y - ref).abs().max() / ref.abs().max()) < 0.06
    dy = torch.randn_like(ref)
    y.backward(dy)
    gx, gw = xx.grad.clone(), w.grad.clone()
    xx.grad = w.grad = None
    ref.backward(dy)
    assert ((gx - xx.grad).abs().max() / xx.grad.abs().max()) < 0.1
    assert ((gw - w.grad).abs().max() / w.grad.abs().max()) < 0.1

    from megatronapp_b200.parallel.tp_layers import ColumnParallelLinear

    _init()
    cfg = _cfg(fp8="hybrid", fp8_recipe="mxfp8")
    lin = ColumnParallelLinear(128, 96, config=cfg, init_method=cfg.init_method, bias=True, skip_bias_add=False)
    assert lin.fp8_recipe.recipe == "mxfp8"
    inp = torch.randn(8, 4, 128)
    out, _ = lin(inp)
    want = torch.nn.functional.linear(inp, lin.weight, lin.bias)
    assert 0 < float((out - want).abs().max()) < 0.06 * float(want.abs().max())      # quantised, and close
    assert type(lin._fp8_cache).__name__ == "_MxWeightCache"


def test_radio_vit_resolution_agnostic_position_embeddings():
    from megatronapp_b200.models.vision.radio import RADIOViTModel, patchify

    _init()
    torch.manual_seed(0)
    cfg = _cfg(kernel_backend="torch")
    m = RADIOViTModel(cfg, patch_dim=4, img_h=16, img_w=16, max_img_h=32, max_img_w=32, class_token_len=3)
    img = torch.arange(2 * 3 * 8 * 8, dtype=torch.float32).view(2, 3, 8, 8)
    p = patchify(img, 4)
    assert p.shape == (2, 4, 48) and torch.equal(p[1, 3].view(3, 4, 4), img[1, :, 4:, 4:])
    m.eval()
    with torch.no_grad():
        assert m(torch.randn(2, 3, 16, 16)).shape == (2, 3 + 16, 32)
        assert m(torch.randn(2, 3, 8, 24)).shape == (2, 3 + 12, 32)          # other resolutions just work
        full = m.get_pos_enc(1, input_size=(32, 32))
        assert full.data_ptr() == m.position_embeddings.data_ptr()           # max grid: the table itself
        a, b = m.get_pos_enc(2, input_size=(16, 16)), m.get_pos_enc(2, input_size=(16, 16))
        assert torch.equal(a, b)                                             # eval: deterministic resize
    m.train()
    a, b = m.get_pos_enc(2, input_size=(16, 16)), m.get_pos_enc(2, input_size=(16, 16))
    assert a.shape == (2, 16, 32) and not torch.equal(a, b) and not torch.equal(a[0], a[1])     # CPE: random windows
    idx = torch.tensor([[0, 5], [7, 2]])
    g = m.get_pos_enc(2, idx, input_size=(32, 32))
    assert torch.equal(g[1, 0], m.position_embeddings[0, 7])
    out = m(torch.randn(2, 3, 16, 16))
    out.square().mean().backward()
    assert m.position_embeddings.grad is not None and m.class_token.grad is not None


def test_packed_thd_sequences_match_separate_forward():
    """Three sequences packed into one THD row (cu_seqlens) == the three sequences run on their own: attention stays
    inside each sequence and rotary positions restart at the boundaries."""
    from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec
    from megatronapp_b200.utils.packed_seq_params import PackedSeqParams

    _init()
    torch.manual_seed(0)
    cfg = _cfg(kernel_backend="torch")
    m = GPTModel(cfg, get_gpt_layer_local_spec(), 64, 64, position_embedding_type="rope").eval()
    lens = [5, 9, 3]
    seqs = [torch.randint(0, 64, (1, n)) for n in lens]
    with torch.no_grad():
        ref = torch.cat([m(s, torch.arange(s.shape[1]).unsqueeze(0), None) for s in seqs], dim=1)
        cu = torch.tensor([0, 5, 14, 17], dtype=torch.int32)
        psp = PackedSeqParams(qkv_format="thd", cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=9, max_seqlen_kv=9)
        packed = torch.cat(seqs, dim=1)
        pos = torch.cat([torch.arange(n) for n in lens]).unsqueeze(0)
        out = m(packed, pos, None, packed_seq_params=psp)
        assert torch.allclose(out, ref, atol=1e-5), float((out - ref).abs().max())
        naive = m(packed, torch.arange(17).unsqueeze(0), None)            # without the boundaries tokens see their neighbours
        assert not torch.allclose(naive[:, 5:], ref[:, 5:], atol=1e-3)
    # gradients flow through the per-sequence path
    m.train()
    loss = m(packed, pos, None, labels=packed, packed_seq_params=psp).mean()
    loss.backward()
    assert m.decoder.layers[0].self_attention.linear_qkv.weight.grad.abs().sum() > 0


def test_mrope_reduces_to_rope_for_text_and_separates_streams():
    """M-RoPE (reference rotary_pos_embedding.py:224): with identical temporal/height/width ids it IS ordinary RoPE; with
    distinct ids every channel section follows its own position stream."""
    import torch

    from megatronapp_b200 import ops
    from megatronapp_b200.models.common.embeddings import MultimodalRotaryEmbedding, RotaryEmbedding

    torch.manual_seed(0)
    s, b, h, d = 12, 2, 3, 16
    x = torch.randn(s, b, h, d)
    pid = torch.arange(s).view(1, 1, s).expand(3, b, s)
    m = MultimodalRotaryEmbedding(d)
    table = m(pid, [2, 3, 3])
    assert table.shape == (s, b, d)
    ref = ops.rope_ref(x, RotaryEmbedding(d)(s, device="cpu"))
    assert torch.allclose(ops.rope(x, table), ref, atol=1e-6)
    pid2 = torch.stack([pid[0], pid[1] * 2, pid[2] * 3])
    t2 = m(pid2, [2, 3, 3])
    half = d // 2
    assert torch.allclose(t2[..., :2], table[..., :2])                      # temporal section unchanged
    assert not torch.allclose(t2[..., 2:5], table[..., 2:5])                # height section follows the doubled ids
    ang = torch.atan2(t2[..., half + 2:half + 5], t2[..., 2:5])
    ang_ref = torch.atan2(table[..., half + 2:half + 5], table[..., 2:5])
    assert torch.allclose(torch.remainder(ang[1], 6.283185307), torch.remainder(2 * ang_ref[1], 6.283185307), atol=1e-4)


def test_huggingface_towers_and_llava_with_an_hf_vision_encoder(tmp_path):
    """``core/models/huggingface``: a transformers CLIP vision tower (config-only directory -> random weights, no download) is
    used as LLaVA's image encoder through ``hf://``; a Qwen2 causal LM wraps as a language model; the multimodal
    sequence-parallel / context-parallel padding helpers give the reference's values."""
    import json

    import torch
    from transformers import CLIPVisionConfig, Qwen2Config

    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatronapp_b200.models.huggingface import build_hf_model, get_hf_model_type
    from megatronapp_b200.models.multimodal.context_parallel import get_packed_seq_params, get_padding
    from megatronapp_b200.models.multimodal.llava_model import LLaVAModel
    from megatronapp_b200.models.multimodal.llava_spec import decoder_model_with_local_default_spec
    from megatronapp_b200.models.vision.vit_layer_specs import get_vit_layer_with_local_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed

    if not ps.model_parallel_is_initialized():
        ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(11)
    torch.manual_seed(0)
    vdir = tmp_path / "clip"
    vdir.mkdir()
    vc = CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, image_size=28, patch_size=14)
    d = vc.to_dict()
    d["architectures"] = ["CLIPVisionModel"]
    (vdir / "config.json").write_text(json.dumps(d))
    assert get_hf_model_type(f"hf://{vdir}") == "clip"
    kw = dict(use_cpu_initialization=True, params_dtype=torch.float32, kernel_backend="torch", hidden_dropout=0.0, attention_dropout=0.0)
    lcfg = TransformerConfig(num_layers=2, hidden_size=48, num_attention_heads=4, **kw)
    vcfg = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, **kw)
    pcfg = TransformerConfig(num_layers=1, hidden_size=48, num_attention_heads=4, ffn_hidden_size=64, **kw)
    tower = build_hf_model(vcfg, f"hf://{vdir}")
    img = torch.randn(2, 3, 28, 28)
    feats = tower(img)
    assert feats.shape == (2, 4 + 1, 32)                                       # 2x2 patches + class token
    from megatronapp_b200.models.vision.multimodal_projector import MultimodalProjector  # noqa: F401
    from megatronapp_b200.transformer.mlp import MLPSubmodules
    from megatronapp_b200.parallel.tp_layers import ColumnParallelLinear, RowParallelLinear

    m = LLaVAModel(lcfg, decoder_model_with_local_default_spec(), 128, 64, vcfg, get_vit_layer_with_local_spec(), True, pcfg,
                   MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear), vision_projection_type="mlp",
                   img_h=28, img_w=28, patch_dim=14, vision_model_type=f"hf://{vdir}", image_token_index=5)
    assert m.img_seq_len == 4                                                   # class token dropped
    ids = torch.randint(6, 128, (2, 12))
    ids[:, 3] = 5
    labels = torch.randint(6, 128, (2, 12))
    out = m(img, ids, torch.arange(12).unsqueeze(0).expand(2, -1), labels=labels, loss_mask=torch.ones(2, 12),
            num_image_tiles=torch.ones(2, dtype=torch.int))
    loss = out[0] if isinstance(out, tuple) else out
    assert loss.shape[0] == 2 and loss.shape[1] == 12 + 4 - 1 and torch.isfinite(loss).all()
    loss.float().mean().backward()
    assert any(p.grad is not None and p.grad.abs().sum() > 0 for p in m.vision_model.parameters())
    # Qwen2 as a wrapped language model
    qdir = tmp_path / "qwen"
    qdir.mkdir()
    qd = Qwen2Config(vocab_size=128, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                     num_key_value_heads=2, max_position_embeddings=64).to_dict()
    qd["architectures"] = ["Qwen2ForCausalLM"]
    (qdir / "config.json").write_text(json.dumps(qd))
    q = build_hf_model(lcfg, f"hf://{qdir}")
    emb = q.embedding(ids)
    assert emb.shape == (12, 2, 32) and q(decoder_input=emb).shape == (2, 12, 128) and q(input_ids=ids, labels=labels).shape == (2, 12)
    # padding rules
    assert get_padding(1001, 1, 4, True) == 3 and get_padding(1001, 2, 1, False) == 3 and get_padding(1001, 2, 4, True) == 7
    assert get_padding(1000, 1, 4, True, decoder_tp_comm_overlap=True, decoder_seq_len=1024) == 24 and get_padding(7, 1, 1, False) == 0
    ps_ = get_packed_seq_params(torch.zeros(2, 10, dtype=torch.long), img_seq_len=6, padding_needed=4, cp_size=2)
    assert ps_.qkv_format == "thd" and ps_.cu_seqlens_q.tolist() == [0, 12, 24] and ps_.cu_seqlens_q_padded.tolist() == [0, 16, 32]
    assert get_packed_seq_params(torch.zeros(2, 10, dtype=torch.long), 6, 0, 1).qkv_format == "sbhd"


def test_first_and_last_layers_stay_bf16_in_an_fp8_model():
    """--first-last-layers-bf16 (reference fp8_utils.is_first_last_bf16_layer): only the inner layers get an fp8 recipe."""
    from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec

    _init()
    cfg = _cfg(num_layers=5, fp8="hybrid", first_last_layers_bf16=True, num_layers_at_start_in_bf16=1, num_layers_at_end_in_bf16=2)
    m = GPTModel(config=cfg, transformer_layer_spec=get_gpt_layer_local_spec(), vocab_size=64, max_sequence_length=16)
    assert [l.mlp.linear_fc1.fp8_recipe is not None for l in m.decoder.layers] == [False, True, True, False, False]
    assert [l.self_attention.linear_proj.fp8_recipe is not None for l in m.decoder.layers] == [False, True, True, False, False]
    ids = torch.randint(0, 64, (2, 16))
    loss = m(ids, torch.arange(16)[None].expand(2, -1), None, labels=ids).mean()
    loss.backward()
    assert torch.isfinite(loss)
    plain = GPTModel(config=_cfg(num_layers=5, fp8="hybrid"), transformer_layer_spec=get_gpt_layer_local_spec(), vocab_size=64,
                     max_sequence_length=16)
    assert all(l.mlp.linear_fc1.fp8_recipe is not None for l in plain.decoder.layers)
