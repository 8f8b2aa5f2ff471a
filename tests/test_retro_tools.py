"""Retro tooling: project inspector, retrieval-augmented generation loop, QA metrics, SFT dataset + entry point."""
import json
import os
import re
import socket
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _project(tmp_path):
    from megatronapp_b200.data.indexed_dataset import IndexedDatasetBuilder
    from tools.retro import preprocess_data as rp

    rng = np.random.RandomState(0)
    b = IndexedDatasetBuilder(str(tmp_path / "c.bin"), dtype=np.int32)
    for _ in range(40):
        n = int(rng.randint(40, 90))
        b.add_document(rng.randint(1, 200, n), [n])
    b.finalize(str(tmp_path / "c.idx"))
    return rp.main(["--retro-project-dir", str(tmp_path / "proj"), "--data-path", str(tmp_path / "c"), "--seq-length", "32",
                    "--retro-gpt-chunk-length", "8", "--retro-retrieved-length", "16", "--train-samples", "10", "--vocab-size", "256",
                    "--retro-index-str", "Flat", "--retro-embed-dim", "32", "--split", "100,0,0"])


def test_cli_inspector_and_generation_loop(tmp_path, capsys):
    cfg = _project(tmp_path)
    from tools.retro.cli import retro

    retro.init(cfg.retro_project_dir)
    assert retro.get_db_num_chunks() > 100 and retro.get_db_num_indexed_datasets() == 1
    assert len(retro.get_db_chunk_gpt(3)) == 8 and len(retro.get_db_chunk_and_continuation_text(3)) == 2
    n, c = retro.get_pt_num_samples_and_chunks()
    assert c == 4 * n
    t = retro.get_neighbor_tokens(0, 1)
    assert len(t["chunk_tokens"]) == 8 and np.asarray(t["neighbor_tokens"]).shape == (2, 16)
    retro.print_neighbor_texts(0, 1)
    retro.print_usage()
    assert "NEIGHBOR_CHUNKS" in capsys.readouterr().out
    # generation loop with a stub model that "copies" the first token of the first neighbour of the current chunk: checks
    # that neighbours are (re-)retrieved as the sequence grows into new chunks and cached for complete chunks
    from tools.retro.text_generation.retro_generation import Retriever, retro_generate

    rt = Retriever(cfg.retro_project_dir)
    calls = []

    def retrieve(chunks):
        calls.append(chunks.shape[0])
        return rt(chunks)

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, ids, pos, mask, context_input_ids=None):
            b, s = ids.shape
            assert context_input_ids.shape == (1, s // 8, 2, 16)
            logits = torch.zeros(b, s, 256)
            for c in range(s // 8):
                logits[0, c * 8:(c + 1) * 8, int(context_input_ids[0, c, 0, 0])] = 1.0
            return logits

    prompt = retro.get_db_chunk_gpt(5) + retro.get_db_chunk_gpt(6)[:3]
    gen = retro_generate(Stub(), prompt, retrieve, 8, max_new_tokens=9)
    assert len(gen) == 9 and all(0 <= g < 256 for g in gen)
    assert calls[0] == 2 and calls[1] == 1 and sum(calls) <= 2 + 9 + 1      # complete chunks are looked up once


def test_generation_metrics_and_evaluate_cli(tmp_path):
    from tools.retro.text_generation import evaluate
    from tools.retro.text_generation.metrics import exact_match, f1_score

    assert exact_match("The Eiffel Tower.", ["eiffel tower", "paris"]) == 1.0 and f1_score("in Paris France", ["Paris"]) == 0.5
    (tmp_path / "gen.txt").write_text("Question: x Answer: the Eiffel tower<|endoftext|> junk\nAnswer: Berlin\n")
    (tmp_path / "ref.json").write_text(json.dumps([{"answers": ["Eiffel Tower"]}, {"answer": "Munich"}]))
    f1, em = evaluate.main(["--generations", str(tmp_path / "gen.txt"), "--references", str(tmp_path / "ref.json")])
    assert em == 50.0 and f1 == 50.0


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_retro_sft_dataset_and_entrypoint(tmp_path):
    from megatronapp_b200.data.tokenizer import GPT2BPETokenizer
    from tools.retro.sft.dataset_conv import RetroSFTDataset

    rows = [{"question": f"what did {c} report ?", "answers": [f"{c} reports revenue"],
             "ctxs": [{"title": c, "text": f"{c} reports third-quarter revenue of $12.5 million."}, {"title": "other", "text": "The board approved a higher dividend."}]}
            for c in ("Northwind Holdings", "Alder Creek Energy", "Blue Harbor Logistics", "Cobalt Ridge Mining")]
    path = tmp_path / "qa.json"
    path.write_text(json.dumps(rows))
    tok = GPT2BPETokenizer(os.path.join(ROOT, "datasets_gpt", "vocab.json"), os.path.join(ROOT, "datasets_gpt", "merges.txt"))
    ds = RetroSFTDataset(str(path), tok, 64, 16, 2, 32)
    s = ds[0]
    assert s["tokens"].shape == (64,) and s["neighbor_tokens"].shape == (4, 2, 32)
    on = s["loss_mask"].bool()
    assert 0 < int(on.sum()) < 20 and tok.detokenize(s["labels"][on].tolist()).startswith(" Northwind Holdings reports revenue")
    assert torch.equal(s["neighbor_tokens"][0], s["neighbor_tokens"][3])              # every chunk sees the sample's passages
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    args = ("--num-layers 2 --hidden-size 32 --num-attention-heads 4 --seq-length 64 --max-position-embeddings 64 --micro-batch-size 2 "
            "--global-batch-size 4 --train-iters 3 --lr 1e-3 --log-interval 1 --eval-iters 1 --eval-interval 100 --distributed-backend gloo "
            "--no-gradient-accumulation-fusion --retro-chunk-length 16 --retro-retrieved-length 32 --retro-num-neighbors 2").split()
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "retro", "sft", "sft_retro.py"), *args, "--train-data-path", str(path),
                        "--vocab-file", os.path.join(ROOT, "datasets_gpt", "vocab.json"), "--merge-file", os.path.join(ROOT, "datasets_gpt", "merges.txt")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert re.search(r"iteration\s+3/\s+3 \|.*lm loss: [0-9.E+\-]+", out)
