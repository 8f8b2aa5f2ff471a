"""Serve the MegaScope front-end (static page) -- the reference's ``npm run dev`` (:5173) step.

    python tools/serve_megascope_ui.py [--port 5173]
"""
import argparse
import functools
import http.server
import os

UI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "megatronapp_b200", "scope", "ui")

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=5173)
    a = ap.parse_args()
    handler = functools.partial(http.server.SimpleHTTPRequestHandler, directory=UI)
    print(f"MegaScope UI on http://{a.host}:{a.port}/  (connect it to the inference / training WS server)")
    http.server.ThreadingHTTPServer((a.host, a.port), handler).serve_forever()
