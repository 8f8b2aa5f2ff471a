"""Zero-shot evaluation plumbing (tasks/zeroshot_gpt)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tasks.zeroshot_gpt.evaluate import LMDataset, build_dataset, evaluate, summarize, wikitext_detokenizer  # noqa: E402


class _Tok:
    eod = 0

    def tokenize(self, text):
        return [1 + (sum(map(ord, w)) % 60) for w in text.split()]


def _model():
    from megatronapp_b200.config import TransformerConfig
    from megatronapp_b200.models.gpt import GPTModel, get_gpt_layer_local_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True,
                            hidden_dropout=0.0, attention_dropout=0.0, kernel_backend="torch")
    return GPTModel(cfg, get_gpt_layer_local_spec(), 64, 16)


def test_sliding_window_scores_every_token_once():
    ds = LMDataset(list(range(1, 60)), seq_len=16, pad_idx=0, num_original_tokens=50, num_tokenized_tokens=59,
                   overlapping_eval=4)
    scored = sum(float(ds[i]["pad_mask"].sum()) for i in range(len(ds)))
    assert scored == 58           # every target token exactly once
    assert wikitext_detokenizer("a @-@ b , c . ") == "a-b, c. "


def test_wikitext_perplexity_and_lambada_accuracy(tmp_path):
    m = _model()
    p = tmp_path / "wiki.txt"
    p.write_text(" ".join(f"w{i % 17}" for i in range(200)))
    ds = build_dataset("WIKITEXT103", str(p), _Tok(), 16, overlapping_eval=8)
    out = evaluate(m, torch.utils.data.DataLoader(ds, batch_size=4), "loss")
    res = summarize("WIKITEXT103", out, ds)
    assert 3.0 < res["avg loss"] < 5.5 and abs(res["ppl"] - math.exp(res["avg loss"])) < 1e-6   # ~ln(64) at init
    lam = tmp_path / "lambada.jsonl"
    lam.write_text("\n".join(json.dumps({"text": " ".join(f"x{j}" for j in range(i + 5))}) for i in range(6)))
    ds2 = build_dataset("LAMBADA", str(lam), _Tok(), 16)
    acc = summarize("LAMBADA", evaluate(m, torch.utils.data.DataLoader(ds2, batch_size=3), "accuracy"), ds2)
    assert acc["total examples"] == 6 and 0.0 <= acc["avg accuracy"] <= 1.0


def test_glue_classification_finetune_learns(tmp_path):
    """Classification head + finetune loop: a separable synthetic sentence-pair task reaches high accuracy."""
    from megatronapp_b200.config import TransformerConfig
    from megatronapp_b200.models.bert.bert_model import bert_layer_local_spec
    from megatronapp_b200.models.bert.classification import Classification, MultipleChoice
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed
    from tasks.finetune_utils import finetune
    from tasks.glue.data import GLUEDataset, build_tokens_types_paddings

    ids, types, pad = build_tokens_types_paddings([5, 6, 7, 8], [9, 10, 11], 8, 1, 2, 0)
    assert ids == [1, 5, 6, 7, 2, 9, 10, 2] and types == [0, 0, 0, 0, 0, 1, 1, 1] and pad == [1] * 8

    class Tok:
        cls, sep, pad, vocab_size = 1, 2, 0, 64

        def tokenize(self, t):
            return [3 + (sum(map(ord, w)) % 60) for w in t.split()]

    def write(path, n, seed):
        import random

        r = random.Random(seed)
        with open(path, "w") as f:
            f.write("id\tquestion1\tquestion2\tis_duplicate\n")
            for i in range(n):
                lab = r.randint(0, 1)
                a = " ".join(r.choice(["red", "blue", "green"]) for _ in range(4))
                b = a if lab else " ".join(r.choice(["cat", "dog", "bird"]) for _ in range(4))
                f.write(f"{i}\t{a}\t{b}\t{lab}\n")

    write(tmp_path / "train.tsv", 256, 0)
    write(tmp_path / "dev.tsv", 64, 1)
    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(1)
    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True,
                            hidden_dropout=0.0, attention_dropout=0.0, kernel_backend="torch")
    m = Classification(cfg, bert_layer_local_spec(), 64, 16, num_classes=2)
    tr = GLUEDataset("QQP", "training", [str(tmp_path / "train.tsv")], Tok(), 12)
    va = GLUEDataset("QQP", "validation", [str(tmp_path / "dev.tsv")], Tok(), 12)
    hist = finetune(m, tr, va, epochs=4, batch_size=32, lr=3e-3, log_interval=0)
    assert hist[-1]["accuracy"] > 0.9, hist
    mc = MultipleChoice(cfg, bert_layer_local_spec(), 64, 16)
    scores = mc(torch.randint(3, 64, (2, 4, 12)), torch.ones(2, 4, 12, dtype=torch.long), torch.zeros(2, 4, 12, dtype=torch.long))
    assert scores.shape == (2, 4)


def test_race_dataset_rows(tmp_path):
    from tasks.race.data import RaceDataset

    class Tok:
        cls, sep, pad = 1, 2, 0

        def tokenize(self, t):
            return [3 + (sum(map(ord, w)) % 60) for w in t.split()]

    p = tmp_path / "r.txt"
    p.write_text(json.dumps({"article": "the cat sat on the mat " * 5, "questions": ["where did the cat sit ?", "the cat sat on the _ ."],
                             "options": [["mat", "roof", "car", "tree"]] * 2, "answers": ["A", "A"]}) + "\n")
    ds = RaceDataset("validation", [str(p)], Tok(), 24)
    assert len(ds) == 2
    s = ds[1]
    assert s["text"].shape == (4, 24) and int(s["label"]) == 0 and s["text"][:, 0].eq(1).all()
    assert (s["types"][0] == 1).sum() > 0 and s["padding_mask"].shape == (4, 24)


def test_biencoder_ict_learns_in_batch_retrieval(tmp_path):
    import numpy as np

    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.data.ict_dataset import ICTDataset
    from megatronapp_b200.data.indexed_dataset import IndexedDataset, IndexedDatasetBuilder, get_bin_path, get_idx_path
    from megatronapp_b200.legacy.model.biencoder_model import BiEncoderModel, ict_loss
    from megatronapp_b200.parallel import state as ps

    if not ps.model_parallel_is_initialized():
        ps.init_distributed("gloo")
        ps.initialize_model_parallel(1, 1)
    # 12 documents; every document talks about its own small vocabulary slice -> a sentence identifies its block
    rng = np.random.default_rng(0)
    prefix = str(tmp_path / "sents")
    b = IndexedDatasetBuilder(get_bin_path(prefix), dtype=np.int32)
    for d in range(12):
        for _ in range(4):
            b.add_item(rng.integers(10 + 8 * d, 18 + 8 * d, size=6).astype(np.int32))
        b.end_document()
    b.finalize(get_idx_path(prefix))
    ds = ICTDataset(IndexedDataset(prefix), None, None, max_seq_length=32, query_in_block_prob=0.0, seed=1, cls_id=1,
                    sep_id=2, pad_id=0)
    assert len(ds) == 12
    s = ds[0]
    q_body = s["query_tokens"][1:7].tolist()
    assert s["query_tokens"][0] == 1 and s["query_tokens"][7] == 2 and s["query_pad_mask"].sum() == 8
    assert s["context_pad_mask"].sum() == 2 + 18                      # the query sentence was removed from its block
    blk = ds.get_block(0, 4, 0)[0]
    assert any(blk[i:i + 6].tolist() == q_body for i in range(1, 20))   # ... and it did come from that block

    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True,
                            hidden_dropout=0.0, attention_dropout=0.0, kernel_backend="torch", params_dtype=torch.float32)
    m = BiEncoderModel(cfg, 128, 32, biencoder_projection_dim=16)
    assert m.query_model is not m.context_model and m.query_model.projection_enc.out_features == 16
    batch = {k: torch.from_numpy(np.stack([ds[i][k] for i in range(12)])) for k in ds[0]}
    opt = torch.optim.Adam(m.parameters(), lr=3e-3)
    first = None
    for _ in range(40):
        q, c = m(batch["query_tokens"], batch["query_pad_mask"], None, batch["context_tokens"], batch["context_pad_mask"], None)
        loss, stats = ict_loss(q, c, 32, score_scaling=True, topk=(1, 5))
        first = first if first is not None else float(loss)
        opt.zero_grad()
        loss.backward()
        opt.step()
    assert float(loss) < 0.5 * first and float(stats["top1_acc"]) > 80.0
    sd = m.state_dict_for_save_checkpoint()
    assert set(sd) == {"query_model", "context_model"}
    shared = BiEncoderModel(cfg, 128, 32, biencoder_shared_query_context_model=True)
    assert set(shared.state_dict_for_save_checkpoint()) == {"model"}


def test_msdp_metrics_prompting_and_orqa_matching(tmp_path):
    import json

    from tasks.msdp import prompt
    from tasks.msdp.evaluate import evaluate_f1
    from tasks.msdp.metrics import corpus_prf1, normalize_answer, prf1
    from tasks.orqa.qa_utils import calculate_matches, has_answer

    assert normalize_answer("The  Cat, sat -- on an APPLE!") == "cat sat on apple"
    assert prf1("the cat sat", "a cat sat down") == (1.0, 2 / 3, 0.8)
    assert prf1("x", "") is None and prf1("", "y") == (0.0, 0.0, 0.0)
    assert corpus_prf1(["cat", "dog", "z"], ["cat", "cat", ""]) == (0.5, 0.5, 0.5)
    (tmp_path / "g.txt").write_text("a red fox<|endoftext|>\nnothing\n")
    (tmp_path / "a.txt").write_text("red fox\nno_passages_used\n")
    assert evaluate_f1(str(tmp_path / "g.txt"), str(tmp_path / "a.txt"))[2] == 1.0

    # stage-1 prompt: few-shot block chosen by "<topic> <last turn>", query in the "( turn ) topic =>" form
    (tmp_path / "kp.jsonl").write_text(json.dumps({"Cats do cats purr?": ["( hi ) Cats => cats are felids", "e2", "e3"]}) + "\n")
    (tmp_path / "in.txt").write_text("Cats\thello [SEP] do cats purr?\tCats purr when content.\n")
    seen = []

    def fake_generate(prompts, n):
        seen.extend(prompts)
        return [" they purr \n( next ) X => junk"] * len(prompts)

    n = prompt.generate_samples(fake_generate, str(tmp_path / "in.txt"), str(tmp_path / "out.txt"), str(tmp_path / "kp.jsonl"),
                                "knowledge", num_prompt_examples=2)
    assert n == 1 and (tmp_path / "out.txt").read_text() == "they purr\n"
    assert seen[0] == "( hi ) Cats => cats are felids \ne2 \n( do cats purr? ) Cats =>"
    (tmp_path / "rp.txt").write_text("shot one\nshot two\nshot three\n")
    seen.clear()
    prompt.generate_samples(fake_generate, str(tmp_path / "in.txt"), str(tmp_path / "out2.txt"), str(tmp_path / "rp.txt"),
                            "response", num_prompt_examples=2)
    assert seen[0] == ("shot one \nshot two \nTopic: Cats. User says: do cats purr? We know that: Cats purr when content. "
                       "System replies:")

    assert has_answer(["New York"], "He moved to new york, in 1990.") and not has_answer(["York New"], "new york")
    assert has_answer([r"19\d\d"], "born in 1987", "regex") and not has_answer(["(unclosed"], "text", "regex")
    docs = {1: ("paris is the capital of france", ""), 2: ("berlin", ""), 3: ("the seine flows through paris", "")}
    top_k, per_q = calculate_matches(docs, [["Paris"], ["Rome"]], [([2, 3, 1], [3., 2., 1.]), ([1, 2, 3], [3., 2., 1.])])
    assert top_k == [0, 1, 1] and per_q[0] == [False, True, True] and per_q[1] == [False] * 3


def test_orqa_evaluator_retrieves_with_biencoder():
    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.legacy.model.biencoder_model import BiEncoderModel
    from megatronapp_b200.parallel import state as ps
    from tasks.orqa.evaluate import ORQAEvaluator

    if not ps.model_parallel_is_initialized():
        ps.init_distributed("gloo")
        ps.initialize_model_parallel(1, 1)

    class Tok:                      # word-level toy tokenizer
        cls, sep, pad = 1, 2, 0

        def __init__(self):
            self.v = {}

        def tokenize(self, t):
            return [self.v.setdefault(w, 3 + len(self.v)) for w in t.lower().split()]

    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True,
                            hidden_dropout=0.0, attention_dropout=0.0, kernel_backend="torch", params_dtype=torch.float32)
    m = BiEncoderModel(cfg, 128, 16, biencoder_shared_query_context_model=True)     # shared tower: same text -> same vector
    docs = {i: (f"passage number {w} about {w}", "") for i, w in enumerate(["zebra", "quartz", "lemon", "violin", "harbor"])}
    ev = ORQAEvaluator(m, Tok(), seq_length=16, batch_size=2)
    ev.index_evidence(docs)
    assert ev.doc_emb.shape == (5, 32)
    closest = ev.search(["passage number lemon about lemon"], top_k=3)
    assert closest[0][0][0] == 2 and closest[0][1][0] >= closest[0][1][1]            # identical text ranks first
    acc = ev.evaluate(["passage number lemon about lemon", "passage number harbor about harbor"], [["lemon"], ["nowhere"]], docs,
                      report_topk=(1, 3))
    assert acc[1] == 50.0 and acc[3] == 50.0


def test_supervised_retriever_finetune_on_dpr_json_learns_and_ranks(tmp_path):
    """RET-FINETUNE-NQ: DPR-format json -> bi-encoder training with hard negatives -> average-rank / top-k validation."""
    import random

    from functools import partial

    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.legacy.model.biencoder_model import BiEncoderModel
    from megatronapp_b200.parallel import state as ps
    from tasks.finetune_utils import finetune
    from tasks.orqa.supervised import NQSupervisedDataset, retrieval_forward_step, retrieval_metrics, task_collate_fn

    if not ps.model_parallel_is_initialized():
        ps.init_distributed("gloo")
        ps.initialize_model_parallel(1, 1)

    class Tok:
        cls, sep, pad = 1, 2, 0

        def __init__(self):
            self.v = {}

        def tokenize(self, t):
            return [self.v.setdefault(w, 3 + len(self.v)) for w in t.lower().split()]

    random.seed(0)
    topics = ["zebra", "quartz", "lemon", "violin", "harbor", "comet", "saddle", "walnut", "ember", "glacier", "orchid", "anvil"]
    rows = []
    for i, w in enumerate(topics):
        others = [t for t in topics if t != w]
        rows.append({"question": f"what is known about {w} ?", "answers": [w],
                     "positive_ctxs": [{"title": w, "text": f"the {w} entry describes {w} in detail"}],
                     "hard_negative_ctxs": [{"title": o, "text": f"the {o} entry describes {o} in detail"} for o in others[:3]],
                     "negative_ctxs": [{"title": o, "text": f"unrelated note on {o}"} for o in others[3:6]]})
    path = tmp_path / "nq.json"
    path.write_text(json.dumps(rows))
    tok = Tok()
    train = NQSupervisedDataset("training", [str(path)], tok, 16, train_with_neg=True, train_hard_neg=2)
    valid = NQSupervisedDataset("validation", [str(path)], tok, 16, evaluate=True, val_av_rank_hard_neg=3, val_av_rank_other_neg=3)
    s0 = train[0]
    assert s0["query"][0] == 1 and len(s0["neg"]) == 2 and len(valid[0]["neg"]) == 6
    assert "?" not in tok.v                                              # the trailing question mark is stripped
    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True,
                            hidden_dropout=0.0, attention_dropout=0.0, kernel_backend="torch", params_dtype=torch.float32)
    m = BiEncoderModel(cfg, 128, 16, biencoder_shared_query_context_model=True)
    loader = torch.utils.data.DataLoader(valid, batch_size=4, collate_fn=task_collate_fn)
    before = retrieval_metrics(m, loader, topk=(1, 3))
    hist = finetune(m, train, valid, epochs=40, batch_size=4, lr=3e-3, log_interval=0, clip_grad=0.0, collate_fn=task_collate_fn,
                    forward_step=partial(retrieval_forward_step, train_with_neg=True), evaluate=partial(retrieval_metrics, topk=(1, 3)))
    after = hist[-1]
    assert after["total"] == 12 and after["rank"] < before["rank"] and after["top3_acc"] >= 90.0 > before["top3_acc"], (before, after)


def test_vision_classification_finetune_learns_from_an_image_folder(tmp_path):
    import numpy as np

    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.models.vision.vit_classification import VitClassificationModel
    from megatronapp_b200.parallel import state as ps
    from tasks.finetune_utils import finetune
    from tasks.vision.classification import ImageFolderTensors, top1, vision_forward_step

    if not ps.model_parallel_is_initialized():
        ps.init_distributed("gloo")
        ps.initialize_model_parallel(1, 1)
    rng = np.random.default_rng(0)
    for split in ("train", "val"):
        for c, mean in (("bright", 1.5), ("dark", -1.5)):
            os.makedirs(tmp_path / split / c)
            for i in range(8 if split == "train" else 4):
                np.save(tmp_path / split / c / f"{i}.npy", (rng.standard_normal((3, 20, 20)) * 0.3 + mean).astype(np.float32))
    train = ImageFolderTensors(str(tmp_path / "train"), 16, 16)
    valid = ImageFolderTensors(str(tmp_path / "val"), 16, 16, train.classes)
    assert train.classes == ["bright", "dark"] and len(train) == 16 and train[0]["image"].shape == (3, 16, 16)   # resized
    torch.manual_seed(0)
    cfg = TransformerConfig(num_layers=1, hidden_size=32, num_attention_heads=4, use_cpu_initialization=True,
                            hidden_dropout=0.0, attention_dropout=0.0, kernel_backend="torch", params_dtype=torch.float32)
    model = VitClassificationModel(cfg, 2, img_h=16, img_w=16, patch_dim=4, finetune=True)
    vloader = torch.utils.data.DataLoader(valid, batch_size=4)
    hist = finetune(model, train, valid, epochs=6, batch_size=4, lr=5e-3, forward_step=vision_forward_step,
                    evaluate=lambda m, _l: top1(m, vloader), warmup_fraction=0.0)
    assert hist[-1]["total"] == 8 and hist[-1]["accuracy"] == 1.0
