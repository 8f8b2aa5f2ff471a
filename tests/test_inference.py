"""Inference: KV-cache decode == full recompute, sampling rules, scheduler, REST + WS servers."""
import json
import threading
import time
import urllib.request

import pytest
import torch

from megatronapp_b200.data.tokenizer import NullTokenizer
from megatronapp_b200.inference import (GPTInferenceWrapper, SamplingParams, Scheduler, StaticInferenceEngine,
                                        TextGenerationController, sample_from_logits)


@pytest.fixture()
def engine():
    from megatronapp_b200.bench_model import build_gpt_for_bench

    st = build_gpt_for_bench(2, 64, 4, 64, 128, 1, device="cpu", dtype=torch.float32, rope=True)
    model = st.model[0].module
    tok = NullTokenizer(127)
    ctl = TextGenerationController(GPTInferenceWrapper(model, max_sequence_length=64), tok)
    return StaticInferenceEngine(ctl, max_batch_size=4), model


def test_kv_cache_decode_matches_full_forward(engine):
    eng, model = engine
    res = eng.generate(["5 9 23 7", "11 3"], SamplingParams(top_k=1, num_tokens_to_generate=6))
    model.eval()
    for r in res:
        toks = list(r.prompt_tokens)
        for _ in range(r.generated_length):
            ids = torch.tensor([toks])
            pos = torch.arange(len(toks)).unsqueeze(0)
            with torch.no_grad():
                nxt = int(model(ids, pos, None)[0, -1, :128].argmax())
            toks.append(nxt)
        assert toks[len(r.prompt_tokens):] == r.generated_tokens.tolist()


def test_sampling_rules():
    logits = torch.tensor([[0.1, 3.0, 0.2, 2.5]])
    assert int(sample_from_logits(logits, SamplingParams(top_k=1))) == 1
    g = torch.Generator().manual_seed(0)
    draws = {int(sample_from_logits(logits, SamplingParams(top_k=2), generator=g)) for _ in range(50)}
    assert draws <= {1, 3}
    draws = {int(sample_from_logits(logits, SamplingParams(top_p=0.5), generator=g)) for _ in range(50)}
    assert draws == {1}
    with pytest.raises(ValueError):
        sample_from_logits(logits, SamplingParams(top_k=2, top_p=0.5))


def test_scheduler_pools():
    s = Scheduler(max_batch_size=2)
    ids = [s.add_request(f"p{i}", [i], SamplingParams()) for i in range(3)]
    assert list(s.active_request_pool) == ids[:2] and list(s.waiting_request_pool) == ids[2:]
    from megatronapp_b200.inference import Status

    s.active_request_pool[ids[0]].status = Status.COMPLETED
    s.update_requests_pools()
    assert ids[0] in s.completed_request_pool and ids[2] in s.active_request_pool


def test_rest_and_ws_servers(engine):
    from websockets.sync.client import connect

    from megatronapp_b200.inference.server import InferenceWSServer, MegatronServer

    eng, _ = engine
    rest = MegatronServer(eng)
    rest.run(port=18811, block=False)
    ws_srv = InferenceWSServer(eng, num_layers=2, port=18812)
    ws_srv.run(block=False)
    time.sleep(0.3)
    try:
        body = json.dumps({"prompts": ["1 2 3"], "tokens_to_generate": 4, "top_k": 1}).encode()
        req = urllib.request.Request("http://127.0.0.1:18811/api", data=body, method="PUT",
                                     headers={"Content-Type": "application/json"})
        out = json.loads(urllib.request.urlopen(req, timeout=30).read())
        assert len(out["tokens"][0]) >= 4
        bad = urllib.request.Request("http://127.0.0.1:18811/api", data=json.dumps({"prompts": []}).encode(), method="PUT")
        with pytest.raises(urllib.error.HTTPError):
            urllib.request.urlopen(bad, timeout=30)
        with connect("ws://127.0.0.1:18812") as ws:
            ws.send(json.dumps({"type": "generate", "prompts": ["4 5 6"], "tokens_to_generate": 2, "top_k": 1,
                                "visualization_flags": {"MLP1_mat_mul": "true", "MLP2_mat_mul": "false"},
                                "compressor_config": {"MLP": {"pixels": 4, "method": "mean"}}}))
            kinds = []
            while True:
                m = json.loads(ws.recv(timeout=30))
                kinds.append((m["type"], m.get("update_type")))
                if m["type"] == "finish":
                    assert out["tokens"][0][:0] == []
                    break
            assert kinds[0][0] == "start" and ("update", 4) in kinds and ("update", 6) in kinds
    finally:
        rest.shutdown()
        ws_srv.shutdown()


def test_beam_search_finds_higher_likelihood_than_greedy(engine):
    from megatronapp_b200.inference.beam_search import beam_search

    _, model = engine
    w = GPTInferenceWrapper(model, 4, 32, None)
    prompt = [3, 7, 11]
    hyps, scores = beam_search(w, prompt, beam_size=4, tokens_to_generate=6, stop_token=10 ** 6, num_return_gen=2)
    assert len(hyps) == 2 and hyps[0].shape[0] == 9 and scores[0] >= scores[1]
    # greedy continuation likelihood under the same model
    toks = torch.tensor([prompt + [0] * 6])
    w.prep_model_for_inference(toks)
    lp, start = 0.0, 0
    for t in range(3, 9):
        logit = w.run_one_forward_step(toks, start, t)[:, -1].float()
        p = torch.log_softmax(logit, -1)
        nxt = int(p.argmax())
        lp += float(p[0, nxt])
        toks[0, t] = nxt
        start = t
    assert float(scores[0]) * 6 >= lp - 1e-4       # beam search never does worse than greedy (length penalty 1)


def test_openai_completions_endpoint(engine):
    from megatronapp_b200.inference.server import MegatronServer

    eng, _ = engine
    srv = MegatronServer(eng)
    srv.run(port=5093, block=False)
    time.sleep(0.3)
    try:
        req = urllib.request.Request("http://127.0.0.1:5093/v1/completions", method="POST",
                                     data=json.dumps({"prompt": "5 9 23", "max_tokens": 4, "temperature": 0}).encode(),
                                     headers={"Content-Type": "application/json"})
        out = json.loads(urllib.request.urlopen(req).read())
        assert out["object"] == "text_completion" and len(out["choices"]) == 1
        assert len(out["choices"][0]["text"].split()) == 4
    finally:
        srv.shutdown()


def test_dynamic_engine_paged_kv_matches_static(engine):
    """Continuous batching + chunked prefill + paged KV give the same greedy tokens as the static engine."""
    from megatronapp_b200.inference import DynamicInferenceEngine

    eng, model = engine
    prompts = ["5 9 23 7 8 1 2 3 4 6 10 12", "11 3", "40 41 42 43 44", "7"]
    sp = SamplingParams(top_k=1, num_tokens_to_generate=7)
    ref = [eng.generate([p], sp)[0].generated_tokens.tolist() for p in prompts]
    dyn = DynamicInferenceEngine(model, NullTokenizer(127), max_sequence_length=64, num_blocks=32, block_size=4,
                                 max_tokens_per_step=6, max_requests=3)
    out = dyn.generate(prompts, sp)
    assert [o.generated_tokens.tolist() for o in out] == ref
    assert dyn.context.allocator.free == 32 and not dyn.context.seqs        # every block returned
    assert dyn.steps < sum(len(p.split()) + 7 for p in prompts)                # requests really shared steps


def test_dynamic_engine_preempts_and_recovers_when_out_of_blocks(engine):
    from megatronapp_b200.inference import DynamicInferenceEngine

    eng, model = engine
    prompts = ["5 9 23 7", "11 3 8 2", "1 2 3 4"]
    sp = SamplingParams(top_k=1, num_tokens_to_generate=10)
    ref = [eng.generate([p], sp)[0].generated_tokens.tolist() for p in prompts]
    # 3 requests x (4 + 10) tokens need 12 blocks of 4; give 8 so that decode growth forces an eviction
    dyn = DynamicInferenceEngine(model, NullTokenizer(127), max_sequence_length=64, num_blocks=8, block_size=4,
                                 max_tokens_per_step=16, max_requests=3)
    out = dyn.generate(prompts, sp)
    assert [o.generated_tokens.tolist() for o in out] == ref
    assert dyn.num_preemptions >= 1 and dyn.context.allocator.free == 8


def test_dynamic_engine_behind_rest_server(engine):
    from types import SimpleNamespace

    from megatronapp_b200.inference import DynamicInferenceEngine
    from megatronapp_b200.inference.server import MegatronServer

    eng, model = engine
    args = SimpleNamespace(inference_dynamic_batching_chunk_size=8, bf16=False, fp16=False, tensor_model_parallel_size=1,
                           pipeline_model_parallel_size=1, inference_dynamic_batching_buffer_size_gb=1e-3,
                           inference_dynamic_batching_buffer_overflow_factor=None,
                           inference_dynamic_batching_max_requests_override=4,
                           inference_dynamic_batching_max_tokens_override=32, inference_max_seq_length=64, seed=1)
    dyn = DynamicInferenceEngine.from_args(model, NullTokenizer(127), args, model.config)
    per_tok = 2 * 2 * 4 * 16 * 4
    assert dyn.context.num_blocks == int(1e-3 * (1 << 30)) // (per_tok * 8)
    srv = MegatronServer(dyn)
    srv.run(port=18813, block=False)
    time.sleep(0.3)
    try:
        body = json.dumps({"prompts": ["1 2 3", "9 8"], "tokens_to_generate": 4, "top_k": 1}).encode()
        req = urllib.request.Request("http://127.0.0.1:18813/api", data=body, method="PUT",
                                     headers={"Content-Type": "application/json"})
        out = json.loads(urllib.request.urlopen(req, timeout=30).read())
        ref = eng.generate(["1 2 3"], SamplingParams(top_k=1, num_tokens_to_generate=4))[0]
        assert out["tokens"][0][3:] == ref.generated_tokens.tolist()
    finally:
        srv.shutdown()


def test_function_style_generation_api_matches_the_engine(engine):
    """``generate_and_post_process`` / ``beam_search_and_post_process`` (the reference's megatron.inference.text_generation
    API) over a bare model give the engine's greedy continuation and the beam search module's best hypothesis."""
    from megatronapp_b200.inference.beam_search import beam_search as bs
    from megatronapp_b200.inference.text_generation import beam_search_and_post_process, generate_and_post_process

    eng, model = engine
    tok = NullTokenizer(127)
    prompts = ["5 9 23 7", "11 3"]
    ref = eng.generate(prompts, SamplingParams(top_k=1, num_tokens_to_generate=5))
    texts, segments, logp, tokens = generate_and_post_process(model, prompts=prompts, tokens_to_generate=5, top_k_sampling=1,
                                                              return_output_log_probs=True, tokenizer=tok, max_sequence_length=32)
    for i, r in enumerate(ref):
        assert tokens[i] == r.prompt_tokens + r.generated_tokens.tolist()
        assert texts[i] == tok.detokenize(tokens[i]) and segments[i] == [str(t) for t in tokens[i]] and len(logp[i]) == 5
    w = GPTInferenceWrapper(model, 4, 32, None)
    hyps, scores = bs(w, [3, 7, 11], beam_size=3, tokens_to_generate=4, stop_token=10 ** 6, num_return_gen=2)
    t2, seg2, sc2 = beam_search_and_post_process(model, prompts=["3 7 11"], tokens_to_generate=4, beam_size=3, stop_token=10 ** 6,
                                                 num_return_gen=2, tokenizer=tok, max_sequence_length=32)
    assert t2[0] == tok.detokenize(hyps[0].tolist()) and abs(sc2[0] - float(scores[0])) < 1e-5 and len(seg2[0]) == 7


def test_token_streaming_engine_asyncio_and_sse(engine):
    """Streaming: tokens are published as they are sampled -- (a) synchronous generator on a stand-alone dynamic engine,
    (b) asyncio consumer of a TokenStream fed by the serving loop, (c) server-sent events from ``/v1/completions`` with
    ``"stream": true`` (dynamic engine) and from the static engine; every path yields the greedy continuation."""
    import asyncio

    from megatronapp_b200.inference import DynamicInferenceEngine, TokenStream
    from megatronapp_b200.inference.server import MegatronServer

    eng, model = engine
    sp = SamplingParams(top_k=1, num_tokens_to_generate=5, return_log_probs=True)
    ref = eng.generate(["5 9 23 7"], SamplingParams(top_k=1, num_tokens_to_generate=5))[0].generated_tokens.tolist()
    dyn = DynamicInferenceEngine(model, NullTokenizer(127), max_sequence_length=64, num_blocks=32, block_size=4,
                                 max_tokens_per_step=6, max_requests=3)
    other = dyn.add_request("11 3", SamplingParams(top_k=1, num_tokens_to_generate=3))          # shares the steps
    chunks = list(dyn.generate_stream("5 9 23 7", sp))
    assert [c["token"] for c in chunks] == ref and all(c["logprob"] is not None and c["text"] == str(c["token"]) for c in chunks)
    while dyn.has_unfinished_requests():
        dyn.step()
    assert other in dyn.completed
    # (b) asyncio consumer while a background thread steps the engine
    dyn.start_background_loop()
    try:
        async def consume():
            st = TokenStream().bind_loop()
            dyn._inbox.put(("5 9 23 7", sp, [5, 9, 23, 7], [], threading.Event(), st))
            return [c["token"] async for c in st.generator()], st

        toks, st = asyncio.run(consume())
        assert toks == ref and st.finished and st.result.generated_tokens.tolist() == ref
        # (c) SSE from the HTTP server on the same (serving) engine
        srv = MegatronServer(dyn)
        srv.run(port=18815, block=False)
        time.sleep(0.3)
        try:
            body = json.dumps({"prompt": "5 9 23 7", "max_tokens": 5, "temperature": 0, "stream": True}).encode()
            req = urllib.request.Request("http://127.0.0.1:18815/v1/completions", data=body, method="POST",
                                         headers={"Content-Type": "application/json"})
            with urllib.request.urlopen(req, timeout=30) as r:
                assert r.headers["Content-Type"] == "text/event-stream"
                lines = [ln.decode().strip() for ln in r if ln.strip()]
            assert lines[-1] == "data: [DONE]"
            texts = [json.loads(ln[6:])["choices"][0]["text"] for ln in lines[:-1]]
            assert [int(t) for t in texts if t] == ref
        finally:
            srv.shutdown()
    finally:
        dyn.stop_loop()
    srv = MegatronServer(eng)                                       # static engine: generate, then replay as events
    srv.run(port=18816, block=False)
    time.sleep(0.3)
    try:
        body = json.dumps({"prompt": "5 9 23 7", "max_tokens": 5, "temperature": 0, "stream": True}).encode()
        req = urllib.request.Request("http://127.0.0.1:18816/v1/completions", data=body, method="POST", headers={"Content-Type": "application/json"})
        with urllib.request.urlopen(req, timeout=30) as r:
            lines = [ln.decode().strip() for ln in r if ln.strip()]
        assert [int(json.loads(ln[6:])["choices"][0]["text"]) for ln in lines[:-2]] == ref and lines[-1] == "data: [DONE]"
    finally:
        srv.shutdown()


def test_mamba_hybrid_recurrent_decode_matches_full_forward():
    """Conv + SSM state caching (Mamba layers) next to the KV cache (attention layers) == recomputing the prefix."""
    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.models.mamba.mamba_model import MambaModel
    from megatronapp_b200.parallel import state as ps

    if not ps.model_parallel_is_initialized():
        ps.init_distributed("gloo")
        ps.initialize_model_parallel(1, 1)
    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=4, hidden_size=32, num_attention_heads=4, mamba_state_dim=8, mamba_head_dim=8,
                            mamba_num_groups=2, kernel_backend="torch", use_cpu_initialization=True, hidden_dropout=0.0,
                            attention_dropout=0.0, params_dtype=torch.float32)
    m = MambaModel(cfg, vocab_size=128, max_sequence_length=64, hybrid_override_pattern="M*-M", position_embedding_type="rope")
    tok = NullTokenizer(127)
    eng = StaticInferenceEngine(TextGenerationController(GPTInferenceWrapper(m, max_sequence_length=64), tok), max_batch_size=4)
    res = eng.generate(["5 9 23 7 11", "3 4"], SamplingParams(top_k=1, num_tokens_to_generate=6))
    m.eval()
    for r in res:
        toks = list(r.prompt_tokens)
        for _ in range(r.generated_length):
            ids = torch.tensor([toks])
            with torch.no_grad():
                nxt = int(m(ids, torch.arange(len(toks)).unsqueeze(0), None)[0, -1, :128].argmax())
            toks.append(nxt)
        assert toks[len(r.prompt_tokens):] == r.generated_tokens.tolist()


def test_vlm_generation_with_kv_cache_matches_full_forward():
    import copy

    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.inference import VLMInferenceWrapper
    from megatronapp_b200.models.gpt import get_gpt_layer_local_spec
    from megatronapp_b200.models.multimodal import DEFAULT_IMAGE_TOKEN_INDEX, LLaVAModel
    from megatronapp_b200.models.vision import get_vit_layer_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
    from megatronapp_b200.transformer.mlp import MLPSubmodules

    if not ps.model_parallel_is_initialized():
        ps.init_distributed("gloo")
        ps.initialize_model_parallel(1, 1)
    torch.manual_seed(0)
    lcfg = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, kernel_backend="torch",
                             use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, params_dtype=torch.float32)
    vcfg = copy.deepcopy(lcfg)
    vcfg.num_layers = 1
    pcfg = copy.deepcopy(lcfg)
    pcfg.ffn_hidden_size = 32
    m = LLaVAModel(lcfg, get_gpt_layer_local_spec(), 128, 64, vcfg, get_vit_layer_spec(), True, pcfg,
                   MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear), img_h=14, img_w=14,
                   patch_dim=7, language_position_embedding_type="rope")
    m.eval()
    images = torch.rand(2, 3, 14, 14)
    w = VLMInferenceWrapper(m, max_batch_size=2, max_sequence_length=64)
    w.set_images(images)
    eng = StaticInferenceEngine(TextGenerationController(w, NullTokenizer(127)), max_batch_size=2)
    img = str(DEFAULT_IMAGE_TOKEN_INDEX)
    res = eng.generate([f"5 {img} 9 23", f"7 {img} 11 3 8"], SamplingParams(top_k=1, num_tokens_to_generate=5))
    assert w._extra == 3                                           # one image -> 4 embeddings in place of 1 token
    for i, r in enumerate(res):
        toks = list(r.prompt_tokens)
        for _ in range(r.generated_length):
            ids = torch.tensor([toks])
            with torch.no_grad():
                logits = m(images[i:i + 1], ids, torch.arange(len(toks)).unsqueeze(0), None)
            toks.append(int(logits[0, -1, :128].argmax()))
        assert toks[len(r.prompt_tokens):] == r.generated_tokens.tolist()


def test_dynamic_engine_serves_concurrent_clients_in_shared_steps(engine):
    """Two HTTP clients at once: their requests are admitted into ONE running batch (fewer engine steps than the sum of
    their standalone runs) and each gets exactly its standalone greedy continuation."""
    import threading

    from megatronapp_b200.inference import DynamicInferenceEngine
    from megatronapp_b200.inference.server import MegatronServer

    eng, model = engine
    sp = SamplingParams(top_k=1, num_tokens_to_generate=30)
    prompts = ["5 9 23 7 8 1 2 3", "11 3 40 41"]
    ref = [eng.generate([p], sp)[0].generated_tokens.tolist() for p in prompts]
    dyn = DynamicInferenceEngine(model, NullTokenizer(127), max_sequence_length=64, num_blocks=32, block_size=4,
                                 max_tokens_per_step=4, max_requests=4)
    dyn.start_background_loop()
    srv = MegatronServer(dyn)
    srv.run(port=18815, block=False)
    time.sleep(0.3)
    outs = [None, None]

    gate = threading.Barrier(2)

    def client(i):
        gate.wait()                                                  # both requests leave at the same moment
        body = json.dumps({"prompts": [prompts[i]], "tokens_to_generate": 30, "top_k": 1}).encode()
        req = urllib.request.Request("http://127.0.0.1:18815/api", data=body, method="PUT",
                                     headers={"Content-Type": "application/json"})
        outs[i] = json.loads(urllib.request.urlopen(req, timeout=120).read())

    try:
        ts = [threading.Thread(target=client, args=(i,)) for i in range(2)]
        [t.start() for t in ts]
        [t.join(timeout=120) for t in ts]
        for i, p in enumerate(prompts):
            assert outs[i]["tokens"][0][len(p.split()):] == ref[i]
        standalone_steps = sum(-(-len(p.split()) // 4) + len(r) - 1 for p, r in zip(prompts, ref))
        assert dyn.steps < standalone_steps                        # the two requests really overlapped
    finally:
        srv.shutdown()
        dyn.stop_loop()


def test_t5_encoder_decoder_generation():
    """Encode once, decode step by step: equals calling the full model on the growing decoder prefix."""
    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.inference import T5InferenceWrapper
    from megatronapp_b200.models.T5 import T5Model, get_t5_decoder_block_spec, get_t5_encoder_block_spec
    from megatronapp_b200.parallel import state as ps

    if not ps.model_parallel_is_initialized():
        ps.init_distributed("gloo")
        ps.initialize_model_parallel(1, 1)
    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, kernel_backend="torch",
                            use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, params_dtype=torch.float32)
    m = T5Model(cfg, cfg, get_t5_encoder_block_spec(2), get_t5_decoder_block_spec(2), vocab_size=64, max_sequence_length=32).eval()
    enc = torch.randint(1, 64, (2, 10))
    enc[1, 7:] = 0                                                   # padding in the second prompt
    w = T5InferenceWrapper(m, decoder_start_token=1, pad_token=0, padded_vocab_size=64)
    out = w.generate(enc, SamplingParams(top_k=1, num_tokens_to_generate=5))
    assert out.shape == (2, 5)
    pad = (enc != 0).float()
    em = pad.unsqueeze(1) * pad.unsqueeze(2)
    dec = torch.full((2, 1), 1)
    with torch.no_grad():
        for t in range(5):
            n = dec.shape[1]
            dm = torch.tril(torch.ones(n, n)).unsqueeze(0).expand(2, -1, -1)
            logits = m(enc, dec, em, dm, pad.unsqueeze(1).expand(2, n, -1))
            dec = torch.cat([dec, logits[:, -1].argmax(-1, keepdim=True)], 1)
    assert torch.equal(out, dec[:, 1:])
    eos = int(out[0, 1])
    stopped = w.generate(enc, SamplingParams(top_k=1, num_tokens_to_generate=5), eos_token=eos)
    first = (stopped[0] == eos).nonzero()[0, 0]
    assert stopped.shape[1] <= 5 and torch.equal(stopped[0, :first + 1], out[0, :first + 1]) and (stopped[0, first:] == eos).all()


def test_legacy_generation_rules(engine):
    """The legacy REST API's generation rules (reference text_generation/generation.py:236-300): stop at an end of line / a double
    end of line, no newline right after a colon, decaying top-p, and the --max-tokens-to-oom guard."""
    from types import SimpleNamespace

    from megatronapp_b200.inference.text_generation import generate
    from megatronapp_b200.training import global_vars

    eng, model = engine
    torch.manual_seed(2)                   # the freshly initialised (tied) model just echoes its input: scramble the matrices
    for p in model.parameters():
        if p.dim() > 1:
            p.data.normal_(0, 0.3)
    prompt = "5 9 23 7"
    g = eng.generate([prompt], SamplingParams(top_k=1, num_tokens_to_generate=6))[0].generated_tokens.tolist()
    assert len(set(g[:4])) == 4 and not set(g) & {5, 9, 23, 7}, g

    class Tok(NullTokenizer):
        def __init__(self, special):
            super().__init__(127)
            self.special = special

        def tokenize(self, text):
            return [self.special[text]] if text in self.special else super().tokenize(text)

    def gen(special, **kw):
        toks, lens, _, _ = generate(model, prompts=[prompt], tokens_to_generate=6, top_k_sampling=1, tokenizer=Tok(special),
                                    max_sequence_length=32, **kw)
        return toks[0, 4:int(lens[0])].tolist()

    unused = next(t for t in range(100) if t not in g and t not in (5, 9, 23, 7))
    assert gen({"\n": g[1], "\n\n": unused, ":": unused}, stop_on_eol=True) == g[:2]
    assert gen({"\n": unused, "\n\n": g[2], ":": unused}, stop_on_eol=True) == g[:3]
    assert gen({"\n": g[1], "\n\n": unused, ":": unused}, stop_on_double_eol=True) == g                 # a single eol does not stop
    assert gen({"\n": unused, "\n\n": g[2], ":": unused}, stop_on_double_eol=True) == g[:3]
    out = gen({"\n": g[0], "\n\n": unused, ":": 7}, prevent_newline_after_colon=True)                  # the prompt ends with ":"
    assert out[0] != g[0] and len(out) == 6
    assert gen({"\n": g[0], "\n\n": unused, ":": unused}, prevent_newline_after_colon=True) == g
    # top-p decays to a nucleus of one: after the first (sampled) token the continuation is the greedy one
    toks, lens, _, _ = generate(model, prompts=[prompt], tokens_to_generate=5, top_p_sampling=0.9, top_p_decay=1e-9, random_seed=3,
                                tokenizer=NullTokenizer(127), max_sequence_length=32)
    row = toks[0, :int(lens[0])].tolist()
    cont = eng.generate([" ".join(map(str, row[:5]))], SamplingParams(top_k=1, num_tokens_to_generate=4))[0].generated_tokens.tolist()
    assert row[5:] == cont[:len(row) - 5]
    old = global_vars._ARGS
    try:
        global_vars.set_args(SimpleNamespace(max_tokens_to_oom=9))
        with pytest.raises(ValueError, match="Too many tokens"):
            generate(model, prompts=[prompt], tokens_to_generate=6, tokenizer=NullTokenizer(127), max_sequence_length=32)
        global_vars.set_args(SimpleNamespace(max_tokens_to_oom=10))
        generate(model, prompts=[prompt], tokens_to_generate=6, top_k_sampling=1, tokenizer=NullTokenizer(127), max_sequence_length=32)
    finally:
        global_vars.set_args(old)


@pytest.mark.parametrize("mode", [[], ["--stream"]])
def test_inference_performance_tool(mode):
    """tools/run_inference_performance_test.py on the regular argument system: static engine and streamed dynamic engine."""
    import json
    import os
    import socket
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "run_inference_performance_test.py"), "--num-layers", "2",
                        "--hidden-size", "64", "--num-attention-heads", "4", "--max-batch-size", "3", "--num-input-tokens", "12",
                        "--num-tokens-to-generate", "6", "--distributed-backend", "gloo", *mode],
                       cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    line = json.loads([x for x in p.stdout.splitlines() if x.startswith("{")][-1])
    assert line["batch"] == 3 and line["generated_tokens"] == 18 and line["tokens_per_s"] > 0
    assert line["engine"] == ("dynamic" if mode else "static") and (("time_to_first_token_s" in line) == bool(mode))


def test_token_level_generation_and_scoring(engine):
    """generate_tokens_probs_and_return_on_first_stage / score_and_return_on_first_stage (reference generation.py): continuation
    written in place into padded token rows; teacher-forced log-probabilities equal the generation log-probabilities."""
    from megatronapp_b200.inference.text_generation.api import (generate_tokens_probs_and_return_on_first_stage,
                                                                 score_and_return_on_first_stage)

    eng, model = engine
    tok = NullTokenizer(127)
    ref = eng.generate(["5 9 23 7", "11 3"], SamplingParams(top_k=1, num_tokens_to_generate=4, return_log_probs=True))
    tokens = torch.full((2, 8), tok.eod, dtype=torch.long)
    tokens[0, :4] = torch.tensor([5, 9, 23, 7])
    tokens[1, :2] = torch.tensor([11, 3])
    lengths = torch.tensor([4, 2])
    out, out_len, logp, _ = generate_tokens_probs_and_return_on_first_stage(model, tokens=tokens, lengths=lengths, top_k=1,
                                                                             return_output_log_probs=True, tokenizer=tok,
                                                                             prevent_newline_after_colon=False)
    assert out is tokens and out[0, 4:8].tolist() == ref[0].generated_tokens.tolist()[:4]
    assert out[1, 2:6].tolist() == ref[1].generated_tokens.tolist()[:4] and out_len.tolist() == [8, 8]      # rows are filled
    _, _, score, _ = score_and_return_on_first_stage(model, out[:1], torch.tensor([8]))
    assert score.shape == (1, 7) and torch.allclose(score[0, 3:7], logp[0, 3:7], atol=1e-4)
    assert torch.allclose(logp[0, 3:7], torch.tensor(ref[0].generated_log_probs[:4]), atol=1e-5)
