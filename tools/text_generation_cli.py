"""Interactive client of the REST text-generation server (reference ``tools/text_generation_cli.py``):
    python tools/text_generation_cli.py 127.0.0.1:5000
"""
import json
import sys
import urllib.request

if __name__ == "__main__":
    url = "http://" + sys.argv[1] + "/api"
    while True:
        try:
            sentence = input("Enter prompt: ")
            n = int(input("Enter number of tokens to generate: "))
        except EOFError:
            break
        req = urllib.request.Request(url, data=json.dumps({"prompts": [sentence], "tokens_to_generate": n}).encode(),
                                     headers={"Content-Type": "application/json"}, method="PUT")
        with urllib.request.urlopen(req) as r:
            print("Megatron Response: ")
            print(json.loads(r.read())["text"][0])
