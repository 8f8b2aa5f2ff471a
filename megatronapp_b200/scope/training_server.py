"""MegaScope training WebSocket server.

Protocol parity with reference ``training_wsserver.py`` / ``training.py:1975-2023``:
client ``{"type":"ping"}`` -> ``{"type":"pong"}``; client ``{"type":"run_training_step",
"visualization_flags":{...}, "disturbance_configs":{...}, "compressor_config":{...}}`` releases one
training step; server sends ``{"type":"start", micro_batch_size, seq_length, num_layers}`` then a
stream of ``{"type":"update","update_type":<FlagType int>,"layer_id","args","result":[...]}``.

The socket lives on the (pp-first, tp-0) rank; the configs are broadcast to every rank
(``broadcast_object_list``).  Binding defaults to 127.0.0.1 (the reference binds 0.0.0.0).
"""
from __future__ import annotations

import json
import queue
import threading
from typing import Optional

import torch
import torch.distributed as dist

from ..parallel import state as ps
from . import hooks
from .capture import FlagType, TensorTracers
from .disturbance import Disturbance


class TrainingWSServer:
    def __init__(self, port: int, num_layers: int, micro_batch_size: int, seq_length: int, host: str = "127.0.0.1",
                 tokenizer=None):
        self.port, self.host = port, host
        self.meta = {"micro_batch_size": micro_batch_size, "seq_length": seq_length, "num_layers": num_layers}
        self.tracers = TensorTracers(num_layers, tokenizer)
        self.disturbance = Disturbance()
        self.requests: "queue.Queue[dict]" = queue.Queue()
        self.outbox: "queue.Queue[Optional[dict]]" = queue.Queue()
        self.is_server_rank = ps.is_pipeline_first_stage(ignore_virtual=True) and ps.get_tensor_model_parallel_rank() == 0 \
            and ps.get_data_parallel_rank() == 0
        self._srv = None
        self._conns = set()
        self._lock = threading.Lock()
        hooks.install(self.tracers, self.disturbance)
        self.tracers.set_report(self._report)

    # -- capture callback (any rank that is TP rank 0 of its group) -------------------
    def _report(self, name, args, tensor_data):
        layer, flag = name
        if torch.is_tensor(tensor_data):
            tensor_data = tensor_data.detach().to("cpu", non_blocking=True).tolist()
        msg = {"type": "update", "update_type": int(flag), "layer_id": int(layer), "args": args, "result": tensor_data}
        self.outbox.put(msg)

    # -- socket side ----------------------------------------------------------------
    def start(self):
        if not self.is_server_rank:
            return
        from websockets.sync.server import serve

        def handler(ws):
            with self._lock:
                self._conns.add(ws)
            try:
                for raw in ws:
                    try:
                        msg = json.loads(raw)
                    except json.JSONDecodeError:
                        ws.send(json.dumps({"type": "error", "message": "invalid json"}))
                        continue
                    if msg.get("type") == "ping":
                        ws.send(json.dumps({"type": "pong"}))
                    elif msg.get("type") == "run_training_step":
                        self.requests.put(msg)
                    else:
                        ws.send(json.dumps({"type": "error", "message": f"unknown type {msg.get('type')!r}"}))
            finally:
                with self._lock:
                    self._conns.discard(ws)

        self._srv = serve(handler, self.host, self.port)
        threading.Thread(target=self._srv.serve_forever, name="megascope-ws", daemon=True).start()
        threading.Thread(target=self._sender, name="megascope-sender", daemon=True).start()

    def _sender(self):
        while True:
            msg = self.outbox.get()
            if msg is None:
                return
            data = json.dumps(msg)
            with self._lock:
                conns = list(self._conns)
            for ws in conns:
                try:
                    ws.send(data)
                except Exception:
                    pass

    # -- training-loop side -----------------------------------------------------------
    def wait_for_step_request(self, model=None):
        """Block until the UI releases a step; broadcast its configs to all ranks."""
        payload = [None]
        if self.is_server_rank:
            req = self.requests.get()
            payload = [{"visualization_flags": req.get("visualization_flags", {}),
                        "disturbance_configs": req.get("disturbance_configs", {}),
                        "compressor_config": req.get("compressor_config", {})}]
        if dist.is_initialized() and dist.get_world_size() > 1:
            src = 0
            dist.broadcast_object_list(payload, src=src)
        cfg = payload[0] or {}
        self.tracers.flags.set_by_configs(cfg.get("visualization_flags", {}))
        self.disturbance.set_by_configs(cfg.get("disturbance_configs", {}))
        if cfg.get("compressor_config"):
            self.tracers.compressor.set_by_configs(cfg["compressor_config"])
        if self.is_server_rank:
            self.outbox.put({"type": "start", **self.meta})

    def shutdown(self):
        hooks.install(None, None)
        self.outbox.put(None)
        if self._srv is not None:
            self._srv.shutdown()


def maybe_start_training_server(args, model):
    srv = TrainingWSServer(args.training_ws_port, args.num_layers, args.micro_batch_size, args.seq_length)
    srv.start()
    return srv
