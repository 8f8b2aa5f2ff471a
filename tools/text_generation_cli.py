"""Client of the REST text-generation server (reference ``tools/text_generation_cli.py``):

    python tools/text_generation_cli.py 127.0.0.1:5000                         # interactive
    python tools/text_generation_cli.py 127.0.0.1:5000 --prompt "Hello" --tokens-to-generate 16   # one request (scripts)
"""
import argparse
import json
import urllib.request


def ask(url: str, prompt: str, n: int) -> str:
    req = urllib.request.Request(url, data=json.dumps({"prompts": [prompt], "tokens_to_generate": n}).encode(),
                                 headers={"Content-Type": "application/json"}, method="PUT")
    with urllib.request.urlopen(req, timeout=120) as r:
        return json.loads(r.read())["text"][0]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("server", help="host:port")
    ap.add_argument("--prompt", default=None)
    ap.add_argument("--tokens-to-generate", type=int, default=32)
    a = ap.parse_args()
    url = "http://" + a.server + "/api"
    if a.prompt is not None:
        print(ask(url, a.prompt, a.tokens_to_generate))
        raise SystemExit(0)
    while True:
        try:
            sentence = input("Enter prompt: ")
            n = int(input("Enter number of tokens to generate: "))
        except EOFError:
            break
        print("Megatron Response: ")
        print(ask(url, sentence, n))
