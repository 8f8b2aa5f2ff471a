"""Text-generation servers: REST (``PUT /api``) and the MegaScope inference WebSocket.

Parity: reference ``megatron/inference/text_generation_server.py`` -- ``MegatronServer`` (Flask
REST: ``{"prompts": [...], "tokens_to_generate": N, "temperature", "top_k", "top_p", "logprobs",
"add_BOS", "random_seed"}`` -> ``{"text": [...], "segments", "logprobs"}``) and
``InferenceWSServer`` (:273-): request ``{"type":"generate", prompts, tokens_to_generate,
temperature/top_k/top_p, visualization_flags, disturbance_configs, compressor_config}``, replies
``start`` (prompt tokens), streamed ``update``s, final ``finish``.  Flask is not available in this
image, so the REST server is a stdlib ``http.server`` with the same route and payload.
"""
from __future__ import annotations

import json
import queue
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Optional

import torch

from ..scope import hooks
from ..scope.capture import TensorTracers
from ..scope.disturbance import Disturbance
from .core import SamplingParams, StaticInferenceEngine

_LOCK = threading.Lock()


def _validate(req: dict):
    if "prompts" not in req:
        return "prompts argument required"
    if not isinstance(req["prompts"], list) or len(req["prompts"]) == 0:
        return "prompts is empty or not a list"
    if len(req["prompts"]) > 128:
        return "Maximum number of prompts is 128"
    n = req.get("tokens_to_generate", 64)
    if not isinstance(n, int) or n < 0:
        return "tokens_to_generate must be an integer >= 0"
    t = req.get("temperature", 1.0)
    if not isinstance(t, (int, float)) or not (0.0 < t <= 100.0):
        return "temperature must be a positive number less than or equal to 100.0"
    k = req.get("top_k", 0)
    if not isinstance(k, int) or not (0 <= k <= 1000):
        return "top_k must be an integer equal to or greater than 0 and less than or equal to 1000"
    p = req.get("top_p", 0.0)
    if not isinstance(p, (int, float)) or not (0.0 <= p <= 1.0):
        return "top_p must be less than or equal to 1 and greater than or equal to 0"
    if k > 0 and p > 0.0:
        return "cannot set both top-k and top-p samplings."
    d = req.get("top_p_decay", 0.0)
    if not isinstance(d, (int, float)) or not (0.0 <= d <= 1.0):
        return "top_p_decay must be less than or equal to 1.0 and greater than or equal to 0.0"
    if d > 0.0 and p == 0.0:
        return "top_p_decay cannot be set without top_p"
    b = req.get("top_p_bound", 0.0)
    if not isinstance(b, (int, float)) or not (0.0 <= b <= 1.0) or (p > 0.0 and b > p):
        return "top_p_bound must be greater than or equal to 0 and less than or equal to top_p"
    for flag in ("add_BOS", "stop_on_double_eol", "stop_on_eol", "prevent_newline_after_colon", "logprobs"):
        if flag in req and not isinstance(req[flag], bool):
            return f"{flag} must be a boolean value"
    if req.get("random_seed") is not None and (not isinstance(req["random_seed"], int) or req["random_seed"] < 0):
        return "random_seed must be a non-negative integer"
    return None


_GENERATE, _SHUTDOWN = 0, 1


def _world():
    import torch.distributed as dist

    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def _announce(kind: int, req: Optional[dict] = None):
    """Rank 0 tells the other model-parallel ranks what to run next (reference ``text_generation_server.py``:
    ``send_do_generate`` / the ``while True: choice = broadcast(...)`` loop in ``run_text_generation_server.py``)."""
    import torch.distributed as dist

    if _world() > 1:
        dist.broadcast_object_list([kind, req], src=0)


def serve_worker_ranks(engine) -> None:
    """Ranks > 0: block on rank 0's announcements and take part in every generation (all ranks must execute the
    model-parallel forward passes); returns on shutdown."""
    import torch.distributed as dist

    while True:
        box = [None, None]
        dist.broadcast_object_list(box, src=0)
        if box[0] == _SHUTDOWN:
            return
        req = box[1]
        scope = req.get("_megascope")
        if scope is None:
            run_generate(engine, req, announce=False)
            continue
        # a MegaScope request: this rank must run the same capture hooks (their TP gathers are collectives) and the
        # same perturbations as rank 0; only rank 0 reports to the client (reference broadcasts the configs per request)
        from ..scope import hooks
        from ..scope.capture import TensorTracers
        from ..scope.disturbance import Disturbance

        tracers, dis = TensorTracers(scope["num_layers"], None), Disturbance()
        tracers.flags.set_by_configs(req.get("visualization_flags", {}))
        dis.set_by_configs(req.get("disturbance_configs", {}))
        if req.get("compressor_config"):
            tracers.compressor.set_by_configs(req["compressor_config"])
        tracers.set_report(lambda *a, **k: None)
        hooks.install(tracers, dis)
        try:
            run_generate(engine, req, announce=False)
        finally:
            hooks.install(None, None)


def run_generate(engine: StaticInferenceEngine, req: dict, announce: bool = True):
    params = SamplingParams(temperature=float(req.get("temperature", 1.0)), top_k=int(req.get("top_k", 0)),
                            top_p=float(req.get("top_p", 0.0)), return_log_probs=bool(req.get("logprobs", False)),
                            num_tokens_to_generate=int(req.get("tokens_to_generate", 64)),
                            top_p_decay=float(req.get("top_p_decay", 0.0)), top_p_bound=float(req.get("top_p_bound", 0.0)),
                            stop_on_eol=bool(req.get("stop_on_eol", False)), stop_on_double_eol=bool(req.get("stop_on_double_eol", False)),
                            prevent_newline_after_colon=bool(req.get("prevent_newline_after_colon", False)))
    if req.get("random_seed") is not None:
        torch.manual_seed(int(req["random_seed"]))
    if getattr(engine, "serving", lambda: False)():
        # continuous batching: the engine's own loop (run by every rank) admits this request into the running batch;
        # concurrent HTTP clients share steps instead of queueing behind a lock
        results = engine.generate(req["prompts"], params, add_BOS=bool(req.get("add_BOS", False)))
        return _format(results, params)
    with _LOCK:                                   # one generation at a time; the announcement is part of it
        if announce:
            _announce(_GENERATE, req)
        results = engine.generate(req["prompts"], params, add_BOS=bool(req.get("add_BOS", False)))
    return _format(results, params)


def _format(results, params):
    out = {"text": [r.prompt + r.generated_text for r in results],
           "tokens": [r.prompt_tokens + r.generated_tokens.tolist() for r in results]}
    if params.return_log_probs:
        out["logprobs"] = [r.generated_log_probs for r in results]
    return out


class MegatronServer:
    def __init__(self, engine: StaticInferenceEngine):
        self.engine = engine
        self.httpd: Optional[ThreadingHTTPServer] = None

    def run(self, host: str = "127.0.0.1", port: int = 5000, block: bool = True):
        engine = self.engine

        class Handler(BaseHTTPRequestHandler):
            def log_message(self, *a):
                pass

            def _send(self, code, obj):
                body = json.dumps(obj).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def do_PUT(self):
                path = self.path.rstrip("/")
                if path in ("/completions", "/v1/completions"):
                    return self._completions()
                if path != "/api":
                    return self._send(404, {"message": "not found"})
                try:
                    req = json.loads(self.rfile.read(int(self.headers.get("Content-Length", 0)) or 0))
                except Exception:
                    return self._send(400, {"message": "invalid json"})
                err = _validate(req)
                if err:
                    return self._send(400, {"message": err})
                try:
                    self._send(200, run_generate(engine, req))
                except ValueError as e:
                    self._send(400, {"message": str(e)})

            def _completions(self):
                """OpenAI-style ``POST /v1/completions`` (reference ``inference/endpoints/completions.py``)."""
                try:
                    req = json.loads(self.rfile.read(int(self.headers.get("Content-Length", 0)) or 0))
                except Exception:
                    return self._send(400, {"error": "invalid json"})
                prompts = req.get("prompt", "")
                prompts = [prompts] if isinstance(prompts, str) else list(prompts)
                temp = float(req.get("temperature", 1.0))
                inner = {"prompts": prompts, "tokens_to_generate": int(req.get("max_tokens", 16)),
                         "temperature": temp if temp > 0 else 1.0, "top_k": 1 if temp == 0 else int(req.get("top_k", 0)),
                         "top_p": float(req.get("top_p", 0.0)) if temp > 0 and not req.get("top_k") else 0.0,
                         "logprobs": bool(req.get("logprobs")), "random_seed": req.get("seed")}
                if inner["top_p"] >= 1.0:
                    inner["top_p"] = 0.0
                err = _validate(inner)
                if err:
                    return self._send(400, {"error": err})
                if req.get("stream"):
                    return self._stream_completion(req, inner, prompts)
                out = run_generate(engine, inner)
                echo = bool(req.get("echo", False))
                choices = []
                for i, (p_, full) in enumerate(zip(prompts, out["text"])):
                    c = {"index": i, "text": full if echo else full[len(p_):], "finish_reason": "length"}
                    if "logprobs" in out:
                        c["logprobs"] = {"token_logprobs": [float(x) for x in out["logprobs"][i]]}
                    choices.append(c)
                self._send(200, {"object": "text_completion", "model": req.get("model", "megatronapp_b200"),
                                 "choices": choices})

            def _stream_completion(self, req, inner, prompts):
                """``"stream": true``: server-sent events, one ``data: {choices: [{text}]}`` line per generated token, then
                ``data: [DONE]`` (OpenAI wire format).  The dynamic engine publishes tokens as they are sampled while other
                clients keep sharing its steps; the static engine generates first and then replays the tokens."""
                if len(prompts) != 1:
                    return self._send(400, {"error": "streaming takes exactly one prompt"})
                params = SamplingParams(temperature=inner["temperature"], top_k=inner["top_k"], top_p=inner["top_p"],
                                        return_log_probs=inner["logprobs"], num_tokens_to_generate=inner["tokens_to_generate"])
                self.send_response(200)
                self.send_header("Content-Type", "text/event-stream")
                self.send_header("Cache-Control", "no-cache")
                self.end_headers()

                def emit(obj):
                    self.wfile.write(("data: " + (obj if isinstance(obj, str) else json.dumps(obj)) + "\n\n").encode())
                    self.wfile.flush()

                model = req.get("model", "megatronapp_b200")
                try:
                    if hasattr(engine, "generate_stream"):
                        chunks = engine.generate_stream(prompts[0], params)
                    else:
                        res = run_generate(engine, inner)
                        tok = engine.controller.tokenizer
                        n_prompt = len(tok.tokenize(prompts[0]))
                        chunks = ({"token": t, "text": tok.detokenize([t]), "logprob": None} for t in res["tokens"][0][n_prompt:])
                    for ch in chunks:
                        c = {"index": 0, "text": ch["text"], "finish_reason": None}
                        if ch.get("logprob") is not None:
                            c["logprobs"] = {"token_logprobs": [float(ch["logprob"])]}
                        emit({"object": "text_completion", "model": model, "choices": [c]})
                    emit({"object": "text_completion", "model": model, "choices": [{"index": 0, "text": "", "finish_reason": "length"}]})
                    emit("[DONE]")
                except (BrokenPipeError, ConnectionResetError):
                    pass                                  # client went away: the generator's close() cancels the request

            do_POST = do_PUT

        self.httpd = ThreadingHTTPServer((host, port), Handler)
        if block:
            self.httpd.serve_forever()
        else:
            threading.Thread(target=self.httpd.serve_forever, daemon=True).start()

    def shutdown(self):
        if self.httpd:
            self.httpd.shutdown()
            with _LOCK:
                _announce(_SHUTDOWN)


class InferenceWSServer:
    def __init__(self, engine: StaticInferenceEngine, num_layers: int, tokenizer=None, host="127.0.0.1", port=5000):
        self.engine, self.host, self.port = engine, host, port
        self.tracers = TensorTracers(num_layers, tokenizer)
        self.disturbance = Disturbance()
        self._srv = None

    def run(self, block: bool = True):
        from websockets.sync.server import serve

        def handler(ws):
            for raw in ws:
                try:
                    msg = json.loads(raw)
                except json.JSONDecodeError:
                    ws.send(json.dumps({"type": "error", "message": "invalid json"}))
                    continue
                if msg.get("type") == "ping":
                    ws.send(json.dumps({"type": "pong"}))
                    continue
                if msg.get("type") != "generate":
                    ws.send(json.dumps({"type": "error", "message": "unknown request type"}))
                    continue
                err = _validate(msg)
                if err:
                    ws.send(json.dumps({"type": "error", "message": err}))
                    continue
                try:
                    self.tracers.flags.set_by_configs(msg.get("visualization_flags", {}))
                    self.disturbance.set_by_configs(msg.get("disturbance_configs", {}))
                    if msg.get("compressor_config"):
                        self.tracers.compressor.set_by_configs(msg["compressor_config"])
                except ValueError as e:
                    ws.send(json.dumps({"type": "error", "message": str(e)}))
                    continue
                tok = self.engine.controller.tokenizer
                ws.send(json.dumps({"type": "start", "prompt_tokens": [tok.tokenize(p) for p in msg["prompts"]],
                                    "num_layers": self.tracers.flags.num_layers}))

                def report(name, args, data):
                    layer, flag = name
                    if torch.is_tensor(data):
                        data = data.detach().float().cpu().tolist()
                    if int(flag) == 6:
                        ws.send(json.dumps({"type": "update", "update_type": 6, "layer_id": 0, "sampled": args, "result": data}))
                    else:
                        ws.send(json.dumps({"type": "update", "update_type": int(flag), "layer_id": int(layer),
                                            "args": args, "result": data}))

                self.tracers.set_report(report)
                hooks.install(self.tracers, self.disturbance)
                msg["_megascope"] = {"num_layers": self.tracers.flags.num_layers}      # tells worker ranks to mirror the hooks
                try:
                    out = run_generate(self.engine, msg)
                finally:
                    hooks.install(None, None)
                    self.tracers.set_report(None)
                ws.send(json.dumps({"type": "finish", **out}))

        self._srv = serve(handler, self.host, self.port)
        if block:
            self._srv.serve_forever()
        else:
            threading.Thread(target=self._srv.serve_forever, daemon=True).start()

    def shutdown(self):
        if self._srv:
            self._srv.shutdown()
