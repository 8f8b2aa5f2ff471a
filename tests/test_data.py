"""Data: .bin/.idx round trip (byte format), native index builders, GPT dataset, blending, samplers,
tokenizers, preprocess tool."""
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

from megatronapp_b200.data.gpt_dataset import BlendedDataset, GPTDataset, _helpers, parse_blend, parse_split, split_document_ranges
from megatronapp_b200.data.indexed_dataset import IndexedDataset, IndexedDatasetBuilder
from megatronapp_b200.data.samplers import MegatronPretrainingRandomSampler, MegatronPretrainingSampler
from megatronapp_b200.data.tokenizer import GPT2BPETokenizer, NullTokenizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_corpus(tmp_path, n_docs=40, seed=0, name="c"):
    rng = np.random.default_rng(seed)
    prefix = str(tmp_path / name)
    b = IndexedDatasetBuilder(prefix + ".bin", dtype=np.uint16)
    docs = []
    for _ in range(n_docs):
        d = rng.integers(1, 1000, size=int(rng.integers(5, 60))).astype(np.uint16)
        docs.append(d)
        b.add_document(d, [len(d)])
    b.finalize(prefix + ".idx")
    return prefix, docs


def test_idx_file_format_and_roundtrip(tmp_path):
    prefix, docs = _make_corpus(tmp_path)
    raw = open(prefix + ".idx", "rb").read()
    assert raw[:9] == b"MMIDIDX\x00\x00"
    assert struct.unpack("<Q", raw[9:17])[0] == 1 and raw[17] == 8          # version 1, dtype code 8 = uint16
    assert struct.unpack("<Q", raw[18:26])[0] == len(docs) and struct.unpack("<Q", raw[26:34])[0] == len(docs) + 1
    ds = IndexedDataset(prefix)
    assert len(ds) == len(docs)
    for i in (0, 7, len(docs) - 1):
        assert np.array_equal(ds[i], docs[i])
    assert np.array_equal(ds.get(3, 2, 3), docs[3][2:5])


def test_build_sample_idx_matches_python_reference(tmp_path):
    h = _helpers()
    sizes = np.array([5, 3, 9, 4, 7], dtype=np.int32)
    doc_idx = np.array([2, 0, 4, 1, 3, 0, 1, 2, 3, 4], dtype=np.int32)
    seq, epochs, tpe = 4, 2, int(sizes.sum())
    got = h.build_sample_idx_int32(sizes, doc_idx, seq, epochs, tpe, True, 1)
    # python reference: walk the concatenated token stream
    stream = np.concatenate([np.full(sizes[d], i) for i, d in enumerate(doc_idx)])   # doc position per token
    n = (epochs * tpe - 1) // seq
    assert got.shape == (n + 1, 2)
    starts = np.concatenate([[0], np.cumsum(sizes[doc_idx])])
    for i in range(n + 1):
        tok = i * seq
        di = int(stream[tok]) if tok < len(stream) else len(doc_idx) - 1
        assert got[i, 0] == di and got[i, 1] == tok - starts[di]


def test_gpt_dataset_samples_are_contiguous_token_windows(tmp_path):
    prefix, docs = _make_corpus(tmp_path)
    ds = IndexedDataset(prefix)
    g = GPTDataset(ds, (0, 30), num_samples=50, seq_length=16, seed=1, cache_dir=str(tmp_path / "cache"))
    assert len(g) >= 50
    item = g[3]
    assert item["tokens"].shape == (16,) and torch.equal(item["tokens"][1:], item["labels"][:-1])
    g2 = GPTDataset(ds, (0, 30), num_samples=50, seq_length=16, seed=1, cache_dir=str(tmp_path / "cache"))
    assert torch.equal(g2[3]["tokens"], item["tokens"])        # index cache reload is identical
    seen = set()
    for i in range(len(g)):
        seen.add(int(g.shuffle_idx[i]))
    assert len(seen) == len(g)


def test_blending_respects_weights(tmp_path):
    p1, _ = _make_corpus(tmp_path, 40, 0, "a")
    p2, _ = _make_corpus(tmp_path, 40, 1, "b")
    d1 = GPTDataset(IndexedDataset(p1), (0, 40), 100, 8, 1)
    d2 = GPTDataset(IndexedDataset(p2), (0, 40), 100, 8, 1)
    b = BlendedDataset([d1, d2], [0.75, 0.25], 200)
    frac = float((b.dataset_index == 0).mean())
    assert abs(frac - 0.75) < 0.02 and b[5]["tokens"].shape == (8,)
    assert parse_blend(["0.3", "x", "0.7", "y"]) == (["x", "y"], [0.3, 0.7])
    assert parse_split("969,30,1") == [0.969, 0.03, 0.001]
    assert split_document_ranges(1000, [0.969, 0.03, 0.001]) == [(0, 969), (969, 999), (999, 1000)]


def test_samplers_partition_batches_across_dp():
    a = list(MegatronPretrainingSampler(32, 0, 2, 0, 2))
    b = list(MegatronPretrainingSampler(32, 0, 2, 1, 2))
    assert a[0] == [0, 1] and b[0] == [2, 3] and len(a) == 8
    assert not set(sum(a, [])) & set(sum(b, []))
    r = MegatronPretrainingRandomSampler(list(range(32)), 32, 0, 2, 0, 2, seed=3)
    assert len(list(r)) == 8


def test_gpt2_bpe_tokenizer_roundtrip(tmp_path):
    from megatronapp_b200.data.tokenizer import bytes_to_unicode

    be = bytes_to_unicode()
    vocab = {c: i for i, c in enumerate(sorted(set(be.values())))}
    for tok in ("he", "ll", "hell", "hello", "Ġw", "Ġwo"):
        vocab[tok] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "merges.txt").write_text("#version: 0.2\nh e\nl l\nhe ll\nhell o\nĠ w\nĠw o\n")
    t = GPT2BPETokenizer(str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt"))
    ids = t.tokenize("hello world")
    assert ids[0] == vocab["hello"] and t.detokenize(ids) == "hello world"
    assert t.decoder[vocab["hello"]] == "hello" and t.eod == vocab["<|endoftext|>"]
    assert NullTokenizer(100).tokenize("3 7 9") == [3, 7, 9]


def test_preprocess_tool(tmp_path):
    src = tmp_path / "in.jsonl"
    src.write_text("\n".join(json.dumps({"text": " ".join(str(x) for x in range(i, i + 12))}) for i in range(20)))
    out = tmp_path / "corp"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "preprocess_data.py"), "--input", str(src),
                           "--tokenizer-type", "NullTokenizer", "--vocab-size", "500", "--append-eod",
                           "--output-prefix", str(out)])
    ds = IndexedDataset(str(out) + "_text_document")
    assert len(ds) == 20 and ds[2].tolist() == list(range(2, 14)) + [500]


def _sentence_corpus(tmp_path, n_docs=12):
    import numpy as np

    from megatronapp_b200.data.indexed_dataset import IndexedDatasetBuilder, get_bin_path, get_idx_path

    prefix = str(tmp_path / "sent")
    b = IndexedDatasetBuilder(get_bin_path(prefix), dtype=np.int32)
    rng = np.random.RandomState(0)
    for _ in range(n_docs):
        for _ in range(rng.randint(3, 7)):
            b.add_item(torch.from_numpy(rng.randint(10, 200, size=rng.randint(5, 20)).astype(np.int64)))
        b.end_document()
    b.finalize(get_idx_path(prefix))
    return prefix


class _Tok:
    vocab_size, cls, sep, mask, pad, bos, eos = 256, 1, 2, 3, 0, 4, 5
    inv_vocab = None


def test_bert_masked_dataset(tmp_path):
    from megatronapp_b200.data.indexed_dataset import IndexedDataset
    from megatronapp_b200.data.masked_dataset import BERTMaskedWordPieceDataset, MaskedWordPieceDatasetConfig

    ds = BERTMaskedWordPieceDataset(IndexedDataset(_sentence_corpus(tmp_path)), 40,
                                    MaskedWordPieceDatasetConfig(sequence_length=64), _Tok())
    assert len(ds) == 40
    frac, rnd = [], 0
    for i in range(len(ds)):
        s = ds[i]
        n = int(s["padding_mask"].sum())
        assert s["text"][0] == 1 and (s["text"][:n] == 2).sum() == 2 and (s["text"][n:] == 0).all()
        assert s["types"][:n].max() == 1 and int(s["loss_mask"].sum()) >= 1
        m = s["loss_mask"].bool()
        assert ((s["text"][m] == 3) | (s["labels"][m] > 0)).all()
        frac.append(float(m.sum()) / (n - 3))
        rnd += int(s["is_random"])
        a, b = ds[i], ds[i]
        assert torch.equal(a["text"], b["text"])          # deterministic per index
    assert 0.08 < sum(frac) / len(frac) < 0.25 and 5 < rnd < 35


def test_t5_masked_dataset(tmp_path):
    from megatronapp_b200.data.indexed_dataset import IndexedDataset
    from megatronapp_b200.data.masked_dataset import MaskedWordPieceDatasetConfig, T5MaskedWordPieceDataset

    sent = list(range(240, 256))
    ds = T5MaskedWordPieceDataset(IndexedDataset(_sentence_corpus(tmp_path)), 20,
                                  MaskedWordPieceDatasetConfig(sequence_length=64, sequence_length_decoder=32), _Tok(),
                                  sentinel_tokens=sent)
    for i in range(len(ds)):
        s = ds[i]
        enc, dec, lab = s["text_enc"].tolist(), s["text_dec"].tolist(), s["labels"].tolist()
        k = sum(t in sent for t in enc)
        assert k >= 1 and dec[0] == 4 and dec[1] == sent[0]
        n_out = int(s["loss_mask"].sum())
        assert lab[n_out - 1] == 5 and lab[:n_out - 1] == dec[1:n_out]           # teacher forcing shift
        assert s["dec_mask"][0, 1] == 0 and s["dec_mask"][1, 0] == 1               # causal decoder mask
        assert s["enc_dec_mask"].shape == (32, 64)


def test_retro_neighbor_database(tmp_path):
    """tools/retro/build_db.py: chunking, exact neighbour search, same-document exclusion."""
    import importlib.util
    import os

    import numpy as np

    from megatronapp_b200.data.indexed_dataset import IndexedDatasetBuilder, get_bin_path, get_idx_path

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("retro_build_db", os.path.join(root, "tools", "retro", "build_db.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    prefix = str(tmp_path / "corpus")
    b = IndexedDatasetBuilder(get_bin_path(prefix), dtype=np.int32)
    rng = np.random.RandomState(0)
    base = rng.randint(5, 50, size=32)
    base2 = rng.randint(50, 100, size=32)
    for d in range(6):
        # documents 0/1 and 2/3 share their opening chunk (near duplicates); 4, 5 are unrelated
        head = base if d < 2 else (base2 if d < 4 else rng.randint(100, 200, size=32))
        b.add_item(torch.from_numpy(np.concatenate([head, rng.randint(200, 250, size=32)]).astype(np.int64)))
        b.end_document()
    b.finalize(get_idx_path(prefix))
    nbr, retrieved = mod.main(["--data-prefix", prefix, "--out", str(tmp_path / "work"), "--chunk-length", "32",
                               "--num-neighbors", "1", "--retrieved-length", "48", "--vocab-size", "256"])
    doc = np.load(tmp_path / "work" / "chunk_doc_ids.npy")
    assert len(doc) == 12 and retrieved.shape == (12, 1, 48)
    first = {d: int(np.where(doc == d)[0][0]) for d in range(6)}
    assert doc[nbr[first[0], 0]] == 1 and doc[nbr[first[1], 0]] == 0        # near-duplicate openings find each other
    assert doc[nbr[first[2], 0]] == 3
    assert all(doc[nbr[i, 0]] != doc[i] for i in range(12))                  # never the same document


def test_tiktokenizer_rank_bpe_and_multimodal_tokenizers(tmp_path):
    import base64
    import json

    from megatronapp_b200.data.tokenizer_extra import (DEFAULT_IMAGE_TOKEN_INDEX, IGNORE_INDEX, MultimodalTokenizer,
                                                       NullMultimodalTokenizer, TikTokenizer)

    # 256 byte tokens + a few merges, in the reference's JSON layout
    merges = [b"he", b"ll", b"hell", b"hello", b" w", b"or", b" wor", b"ld", b" world"]
    vocab = [bytes([i]) for i in range(256)] + merges
    path = tmp_path / "tok.json"
    path.write_text(json.dumps([{"rank": i, "token_bytes": base64.b64encode(t).decode(), "token_str": repr(t)}
                                for i, t in enumerate(vocab)]))
    tk = TikTokenizer(str(path), "v2", None, num_special_tokens=8)
    assert tk.vocab_size == 8 + len(vocab) and (tk.unk, tk.bos, tk.eos) == (0, 1, 2)
    ids = tk.tokenize("hello world", bos=True, eos=True)
    assert ids == [1, 8 + 256 + 3, 8 + 256 + 8, 2]                    # <s> "hello" " world" </s>
    assert tk.detokenize(ids) == "<s>hello world</s>"
    s = "héllo wörld ☃ 12 x"
    assert tk.detokenize(tk.tokenize(s)) == s                          # byte fallback round-trips any utf-8
    # lowest-rank pair merges first: "hel" -> "he" + "l" (not "h" + "el")
    assert tk.tokenize("hel") == [8 + 256, 8 + ord("l")]

    mm = MultimodalTokenizer(tk, "chatml", special_tokens=["<|im_start|>", "<|im_end|>"], image_tag_type="internvl")
    conv = [{"role": "system", "content": "be brief"}, {"role": "user", "content": "<image>\nwhat?"},
            {"role": "assistant", "content": "a cat"}]
    toks, tgt = mm.tokenize_conversation(conv, True, False)
    assert len(toks) == len(tgt) and DEFAULT_IMAGE_TOKEN_INDEX in toks
    kept = tgt != IGNORE_INDEX
    assert mm.detokenize(toks[kept]) == "a cat<|im_end|>\n"             # only the assistant answer is a target
    assert mm.detokenize(toks).startswith("<|im_start|>system\nbe brief<|im_end|>\n<|im_start|>user\n<img><image></img>")
    gen = mm.tokenize(conv[:2])
    assert mm.detokenize(gen).endswith("<|im_start|>assistant\n")
    assert mm.convert_tokens_to_ids("<image>") == DEFAULT_IMAGE_TOKEN_INDEX
    assert mm.vocab_size == tk.vocab_size + 2

    nm = NullMultimodalTokenizer(100)
    assert nm.tokenize("5 <image> 7") == [5, DEFAULT_IMAGE_TOKEN_INDEX, 7] and nm.detokenize([5, -200, 7]) == "5 <image> 7"
    assert nm.vocab_size == 101 and nm.eod == 100


def test_blended_megatron_dataset_builder(tmp_path):
    from megatronapp_b200.data.blended_builder import (BlendedMegatronDatasetBuilder, GPTDatasetConfig,
                                                      get_blend_from_list)

    p1, _ = _make_corpus(tmp_path, 60, 0, "a")
    p2, _ = _make_corpus(tmp_path, 60, 1, "b")
    assert get_blend_from_list(["0.3", p1, "0.7", p2]) == ([p1, p2], [0.3, 0.7])
    assert get_blend_from_list([p1, p2]) == ([p1, p2], None)
    cfg = GPTDatasetConfig(random_seed=1, sequence_length=8, blend=([p1, p2], [0.25, 0.75]), split="8,1,1",
                           path_to_cache=str(tmp_path / "cache"))
    assert cfg.split_matrix[0] == (0.0, 0.8) and abs(cfg.split_matrix[2][1] - 1.0) < 1e-9
    tr, va, te = BlendedMegatronDatasetBuilder(GPTDataset, [200, 20, None], lambda: True, cfg).build()
    assert len(tr) == 200 and abs(float((tr.dataset_index == 1).mean()) - 0.75) < 0.02
    assert len(va) == 20 and te is not None and te[0]["tokens"].shape == (8,)      # None size = one exhaustive epoch
    assert BlendedMegatronDatasetBuilder(GPTDataset, [200, 20, 5], lambda: False, cfg).build() == [None, None, None]
    # independent blends per split
    cfg2 = GPTDatasetConfig(random_seed=1, sequence_length=8, blend_per_split=[([p1], None), ([p2], None), None])
    tr2, va2, te2 = BlendedMegatronDatasetBuilder(GPTDataset, [50, 10, 10], lambda: True, cfg2).build()
    assert len(tr2) >= 50 and len(va2) >= 10 and te2 is None
    with pytest.raises(ValueError):
        GPTDatasetConfig(random_seed=1, sequence_length=8, blend=([p1], None), blend_per_split=[([p1], None), None, None])
    # no blend at all -> mock data
    cfg3 = GPTDatasetConfig(random_seed=1, sequence_length=8, tokenizer=NullTokenizer(99))
    tr3, _, _ = BlendedMegatronDatasetBuilder(GPTDataset, [7, 0, 0], lambda: True, cfg3).build()
    assert cfg3.mock and len(tr3) == 7 and int(tr3[0]["tokens"].max()) < 100


def test_preprocess_mmdata_roundtrip(tmp_path):
    import importlib.util
    import json

    spec = importlib.util.spec_from_file_location("preprocess_mmdata", os.path.join(ROOT, "tools", "preprocess_mmdata.py"))
    mm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mm)
    (tmp_path / "img").mkdir()
    blobs = [bytes(range(7)), bytes(range(100, 112))]                 # 7 bytes (needs padding) and 12 bytes
    for i, b in enumerate(blobs):
        (tmp_path / "img" / f"{i:03d}.jpg").write_bytes(b)
    (tmp_path / "cap.jsonl").write_text("\n".join(json.dumps({"text": t}) for t in ["5 6 7", "8 9 10 11 12 13"]) + "\n")
    prefix = str(tmp_path / "mm")
    mm.main(["--input", str(tmp_path / "cap.jsonl"), "--input-image", str(tmp_path / "img"), "--pad-length", "4",
             "--tokenizer-type", "NullTokenizer", "--vocab-size", "50", "--output-prefix", prefix])
    ds = IndexedDataset(prefix, multimodal=True)
    assert len(ds) == 4 and ds.document_indices.tolist() == [0, 2, 4]
    assert ds.index.sequence_modes.tolist() == [0, 1, 0, 1]
    assert ds[0].tolist() == [5, 6, 7, 50] and ds[2].tolist() == [8, 9, 10, 11]     # padded with eod / truncated
    assert mm.words_to_image_bytes(ds[1]) == blobs[0] and mm.words_to_image_bytes(ds[3]) == blobs[1]


def test_multimodal_dataset_reads_preprocessed_caption_image_pairs(tmp_path):
    """tools/preprocess_mmdata.py output -> legacy MultiModalDataset: captions come back as token ids, images are decoded
    from the stored file bytes and transformed to ``[3, h, w]`` CLIP-normalised tensors."""
    import importlib.util
    import json

    import numpy as np
    from PIL import Image

    from megatronapp_b200.legacy.data.multimodal_dataset import MultiModalDataset

    spec = importlib.util.spec_from_file_location("preprocess_mmdata", os.path.join(ROOT, "tools", "preprocess_mmdata.py"))
    mm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mm)
    (tmp_path / "img").mkdir()
    rng = np.random.RandomState(0)
    for i in range(3):
        a = np.zeros((20, 30, 3), dtype=np.uint8)
        a[..., i] = 255                                               # pure red / green / blue images
        Image.fromarray(a).save(tmp_path / "img" / f"{i:03d}.png")
    (tmp_path / "cap.jsonl").write_text("\n".join(json.dumps({"text": f"{i + 3} {i + 4}"}) for i in range(3)) + "\n")
    mm.main(["--input", str(tmp_path / "cap.jsonl"), "--input-image", str(tmp_path / "img"), "--pad-length", "4",
             "--tokenizer-type", "NullTokenizer", "--vocab-size", "50", "--output-prefix", str(tmp_path / "mm")])
    ds = MultiModalDataset("train", str(tmp_path / "mm"), img_h=8, img_w=8)
    assert len(ds) == 3
    for i in range(3):
        s = ds[i]
        assert s["text"].tolist() == [i + 3, i + 4, 50, 50] and s["img"].shape == (3, 8, 8)
        assert int(np.argmax(s["img"].mean(axis=(1, 2)))) == i      # the dominant channel survives decode + normalise


def test_preprocess_data_partitions_and_sentence_splitter(tmp_path):
    """tools/preprocess_data.py: --partitions with --keep-sequential-samples reproduces the single-pass files byte for byte,
    round-robin partitions hold the same documents, and the rule-based sentence splitter respects abbreviations / initials."""
    import filecmp
    import importlib.util
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.path.join(root, "tools", "preprocess_data.py")
    spec = importlib.util.spec_from_file_location("preprocess_data_tool", tool)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    text = 'Dr. Smith went to Washington. He said "Hello there." Then he left! Was it 5 p.m. already? Yes.\nJ. R. Tolkien wrote it.'
    assert mod.rule_based_sentences(text) == ["Dr. Smith went to Washington.", 'He said "Hello there."', "Then he left!",
                                              "Was it 5 p.m. already?", "Yes.", "J. R. Tolkien wrote it."]
    assert mod.rule_based_sentences(text, keep_newlines=True)[4] == "Yes.\n"
    src = tmp_path / "in.jsonl"
    with open(os.path.join(root, "datasets_gpt", "dataset.json")) as f, open(src, "w") as g:
        g.writelines(next(f) for _ in range(60))
    tok = ["--tokenizer-type", "GPT2BPETokenizer", "--vocab-file", os.path.join(root, "datasets_gpt", "vocab.json"), "--merge-file",
           os.path.join(root, "datasets_gpt", "merges.txt"), "--append-eod"]

    def run(prefix, *extra):
        p = subprocess.run([sys.executable, tool, "--input", str(src), "--output-prefix", str(tmp_path / prefix), *tok, *extra],
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]

    run("one")
    run("seq", "--partitions", "4", "--keep-sequential-samples", "--workers", "2")
    run("rr", "--partitions", "3")
    for ext in (".bin", ".idx"):
        assert filecmp.cmp(tmp_path / f"one_text_document{ext}", tmp_path / f"seq_text_document{ext}", shallow=False)
    from megatronapp_b200.data.indexed_dataset import IndexedDataset

    a, b = IndexedDataset(str(tmp_path / "one_text_document")), IndexedDataset(str(tmp_path / "rr_text_document"))
    assert len(a) == len(b) == 60 and sorted(len(a[i]) for i in range(60)) == sorted(len(b[i]) for i in range(60))
    assert not [f for f in os.listdir(tmp_path) if "_part" in f]
