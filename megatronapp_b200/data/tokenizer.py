"""Tokenizers behind ``--tokenizer-type`` (reference ``training/tokenizer/tokenizer.py:26-96``).

Implemented natively: ``NullTokenizer``, ``GPT2BPETokenizer`` (byte-level BPE from
vocab.json/merges.txt, with the ``decoder``/``encoder`` maps MegaScope needs),
``BertWordPieceLowerCase/Case``; ``SentencePieceTokenizer`` / ``GPTSentencePieceTokenizer`` /
``Llama2Tokenizer`` wrap the ``sentencepiece`` package; ``HuggingFaceTokenizer`` wraps
``transformers``; ``TikTokenizer`` is gated on ``tiktoken`` being installed.
"""
from __future__ import annotations

import json
import os
import unicodedata
from functools import lru_cache
from typing import Dict, List

import regex as re


def _vocab_size_with_padding(orig_vocab_size: int, args) -> int:
    mult = args.make_vocab_size_divisible_by * args.tensor_model_parallel_size
    after = (orig_vocab_size + mult - 1) // mult * mult
    if getattr(args, "rank", 0) == 0:
        print(f" > padded vocab (size: {orig_vocab_size}) with {after - orig_vocab_size} dummy tokens "
              f"(new size: {after})", flush=True)
    return after


class MegatronTokenizer:
    def __init__(self, name):
        self.name = name

    @property
    def vocab_size(self):
        raise NotImplementedError

    def tokenize(self, text):
        raise NotImplementedError

    def detokenize(self, ids):
        raise NotImplementedError

    @property
    def eod(self):
        raise NotImplementedError

    @property
    def eos(self):
        return self.eod

    @property
    def pad(self):
        return self.eod


class NullTokenizer(MegatronTokenizer):
    def __init__(self, vocab_size):
        super().__init__("NullTokenizer")
        self._vocab_size_without_eod = int(vocab_size)
        self._eod_id = int(vocab_size)

    def tokenize(self, text):
        return [int(x) for x in text.split(" ")]

    def detokenize(self, ids):
        return " ".join(str(i) for i in ids)

    def offsets(self, ids, text):
        out, start = [], 0
        for i in ids:
            out.append(start)
            start += 1 + len(str(i))
        return out

    @property
    def vocab_size(self):
        return self._vocab_size_without_eod + 1

    @property
    def vocab(self):
        raise NotImplementedError

    @property
    def eod(self):
        return self._eod_id


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


class GPT2BPETokenizer(MegatronTokenizer):
    def __init__(self, vocab_file, merge_file):
        super().__init__("GPT2 BPE")
        with open(vocab_file, encoding="utf-8") as f:
            self.encoder: Dict[str, int] = json.load(f)
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        with open(merge_file, encoding="utf-8") as f:
            merges = [tuple(line.split()) for line in f.read().split("\n")[1:] if line and not line.startswith("#")]
        self.bpe_ranks = {m: i for i, m in enumerate(merges) if len(m) == 2}
        self.cache: Dict[str, str] = {}
        self.pat = re.compile(r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")
        self.eod_id = self.encoder.get("<|endoftext|>", len(self.encoder) - 1)

    def _bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word = tuple(token)
        if len(word) < 2:
            return token
        while True:
            pairs = {(word[i], word[i + 1]) for i in range(len(word) - 1)}
            if not pairs:
                break
            bigram = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if bigram not in self.bpe_ranks:
                break
            first, second = bigram
            new, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    new.append(first + second)
                    i += 2
                else:
                    new.append(word[i])
                    i += 1
            word = tuple(new)
            if len(word) == 1:
                break
        out = " ".join(word)
        self.cache[token] = out
        return out

    def tokenize(self, text):
        ids = []
        for tok in re.findall(self.pat, text):
            tok = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self._bpe(tok).split(" "))
        return ids

    def detokenize(self, ids):
        text = "".join(self.decoder[i] for i in ids)
        return bytearray([self.byte_decoder[c] for c in text]).decode("utf-8", errors="replace")

    def offsets(self, ids, text):
        out, pos = [], 0
        for i in ids:
            out.append(pos)
            pos += len(bytearray([self.byte_decoder[c] for c in self.decoder[i]]).decode("utf-8", errors="replace"))
        return out

    @property
    def vocab_size(self):
        return len(self.encoder)

    @property
    def vocab(self):
        return self.encoder

    @property
    def inv_vocab(self):
        return self.decoder

    @property
    def eod(self):
        return self.eod_id


class BertWordPieceTokenizer(MegatronTokenizer):
    def __init__(self, vocab_file, lower_case=True, vocab_extra_ids=0):
        super().__init__("BERT Lower Case" if lower_case else "BERT Upper Case")
        self.vocab_: Dict[str, int] = {}
        with open(vocab_file, encoding="utf-8") as f:
            for i, line in enumerate(f):
                self.vocab_[line.rstrip("\n")] = i
        self.lower = lower_case
        for k in range(vocab_extra_ids):
            self.vocab_.setdefault(f"<extra_id_{k}>", len(self.vocab_))
        self.inv = {v: k for k, v in self.vocab_.items()}
        self.cls_id, self.sep_id = self.vocab_["[CLS]"], self.vocab_["[SEP]"]
        self.pad_id, self.mask_id = self.vocab_["[PAD]"], self.vocab_["[MASK]"]
        self.unk = self.vocab_.get("[UNK]", 0)

    def _basic(self, text):
        if self.lower:
            text = "".join(c for c in unicodedata.normalize("NFD", text.lower()) if unicodedata.category(c) != "Mn")
        return re.findall(r"\w+|[^\w\s]", text)

    def tokenize(self, text):
        ids = []
        for w in self._basic(text):
            start, sub = 0, []
            bad = False
            while start < len(w):
                end = len(w)
                cur = None
                while start < end:
                    piece = w[start:end] if start == 0 else "##" + w[start:end]
                    if piece in self.vocab_:
                        cur = piece
                        break
                    end -= 1
                if cur is None:
                    bad = True
                    break
                sub.append(self.vocab_[cur])
                start = end
            ids.extend([self.unk] if bad else sub)
        return ids

    def detokenize(self, ids):
        toks = [self.inv[i] for i in ids if i not in (self.pad_id, self.cls_id, self.sep_id)]
        out = ""
        for t in toks:
            out += t[2:] if t.startswith("##") else (" " + t if out else t)
        return out

    def decode_token_ids(self, ids):
        return self.detokenize(ids)

    @property
    def vocab_size(self):
        return len(self.vocab_)

    @property
    def vocab(self):
        return self.vocab_

    @property
    def inv_vocab(self):
        return self.inv

    @property
    def cls(self):
        return self.cls_id

    @property
    def sep(self):
        return self.sep_id

    @property
    def pad(self):
        return self.pad_id

    @property
    def mask(self):
        return self.mask_id

    @property
    def eod(self):
        return self.sep_id


class SentencePieceTokenizer(MegatronTokenizer):
    def __init__(self, model_file, vocab_extra_ids=0, legacy=True):
        super().__init__("SentencePieceTokenizer")
        import sentencepiece

        self.sp = sentencepiece.SentencePieceProcessor(model_file=model_file)
        self._extra = vocab_extra_ids
        self._eod = self.sp.eos_id() if self.sp.eos_id() >= 0 else self.sp.pad_id()

    def tokenize(self, text):
        return self.sp.encode(text)

    def detokenize(self, ids):
        return self.sp.decode([i for i in ids if i < self.sp.get_piece_size()])

    @property
    def vocab_size(self):
        return self.sp.get_piece_size() + self._extra

    @property
    def bos(self):
        return self.sp.bos_id()

    @property
    def eod(self):
        return self._eod


class HuggingFaceTokenizer(MegatronTokenizer):
    def __init__(self, name_or_path, **kw):
        super().__init__("HuggingFaceTokenizer")
        import transformers

        self.tk = transformers.AutoTokenizer.from_pretrained(name_or_path, **kw)

    def tokenize(self, text, **kw):
        return self.tk(text, **kw).input_ids

    def detokenize(self, ids, **kw):
        return self.tk.decode(ids, **kw)

    @property
    def vocab_size(self):
        return len(self.tk)

    @property
    def vocab(self):
        return self.tk.get_vocab()

    @property
    def eod(self):
        return self.tk.eos_token_id


def build_tokenizer(args):
    ttype = args.tokenizer_type
    if ttype is None:
        if args.vocab_size is not None or args.padded_vocab_size is not None or args.mock_data:
            ttype = "NullTokenizer"
        else:
            return None
    if ttype == "NullTokenizer":
        vs = args.vocab_size if args.vocab_size is not None else 50256
        tok = NullTokenizer(vs)
    elif ttype == "GPT2BPETokenizer":
        tok = GPT2BPETokenizer(args.vocab_file, args.merge_file)
    elif ttype in ("BertWordPieceLowerCase", "BertWordPieceCase"):
        tok = BertWordPieceTokenizer(args.vocab_file, ttype.endswith("LowerCase"), args.vocab_extra_ids)
    elif ttype in ("SentencePieceTokenizer", "GPTSentencePieceTokenizer", "Llama2Tokenizer"):
        tok = SentencePieceTokenizer(args.tokenizer_model, args.vocab_extra_ids)
    elif ttype == "HuggingFaceTokenizer":
        tok = HuggingFaceTokenizer(args.tokenizer_model)
    elif ttype == "TikTokenizer":
        from .tokenizer_extra import TikTokenizer

        tok = TikTokenizer(args.tokenizer_model, getattr(args, "tiktoken_pattern", None) or "v2", args.vocab_size,
                           getattr(args, "tiktoken_num_special_tokens", 1000), getattr(args, "tiktoken_special_tokens", None))
    elif ttype == "MultimodalTokenizer":
        from .tokenizer_extra import MultimodalTokenizer

        inner = HuggingFaceTokenizer(args.tokenizer_model) if args.tokenizer_model else NullTokenizer(args.vocab_size or 50256)
        tok = MultimodalTokenizer(inner, getattr(args, "tokenizer_prompt_format", None) or "chatml",
                                  getattr(args, "special_tokens", None), getattr(args, "image_tag_type", None) or "")
    elif ttype == "NullMultimodalTokenizer":
        from .tokenizer_extra import NullMultimodalTokenizer

        tok = NullMultimodalTokenizer(args.vocab_size if args.vocab_size is not None else 50256)
    else:
        raise NotImplementedError(f"{ttype} tokenizer is not implemented")
    if getattr(args, "padded_vocab_size", None) is None:
        args.padded_vocab_size = _vocab_size_with_padding(tok.vocab_size, args)
    return tok


# -- class / function names of the reference's ``tokenizer/gpt2_tokenization.py`` and ``bert_tokenization.py`` ------------------------
def get_pairs(word):
    """Adjacent symbol pairs of a word (tuple of symbols)."""
    return {(word[i], word[i + 1]) for i in range(len(word) - 1)}


class GPT2Tokenizer(GPT2BPETokenizer):
    """``GPT2Tokenizer(vocab_file, merges_file)`` with the reference's method names."""

    def __init__(self, vocab_file, merges_file, errors="replace", special_tokens=None, max_len=None):
        super().__init__(vocab_file, merges_file)
        self.max_len = max_len if max_len is not None else int(1e12)

    def encode(self, text):
        return self.tokenize(text)

    def decode(self, tokens):
        return self.detokenize(tokens)

    def convert_tokens_to_ids(self, tokens):
        return [self.encoder.get(t, 0) for t in ([tokens] if isinstance(tokens, str) else tokens)]

    def convert_ids_to_tokens(self, ids, skip_special_tokens=False):
        return [self.decoder[i] for i in ids]

    def __len__(self):
        return len(self.encoder)


def load_vocab(vocab_file):
    """Ordered ``token -> id`` of a WordPiece vocabulary file."""
    vocab = {}
    with open(vocab_file, encoding="utf-8") as f:
        for i, line in enumerate(f):
            vocab[line.rstrip("\n")] = i
    return vocab


def convert_to_unicode(text):
    return text.decode("utf-8", "ignore") if isinstance(text, bytes) else str(text)


def printable_text(text):
    return convert_to_unicode(text)


def whitespace_tokenize(text):
    return text.strip().split()


def convert_by_vocab(vocab, items):
    return [vocab[i] for i in items]


def convert_tokens_to_ids(vocab, tokens):
    return convert_by_vocab(vocab, tokens)


def convert_ids_to_tokens(inv_vocab, ids):
    return convert_by_vocab(inv_vocab, ids)


class FullTokenizer:
    """Basic (punctuation / lower-casing) + WordPiece tokenisation to *tokens* (strings), as the reference's BERT tokenizer."""

    def __init__(self, vocab_file, do_lower_case=True):
        self._impl = BertWordPieceTokenizer(vocab_file, lower_case=do_lower_case)
        self.vocab, self.inv_vocab = self._impl.vocab_, self._impl.inv

    def tokenize(self, text):
        return [self.inv_vocab[i] for i in self._impl.tokenize(text)]

    def convert_tokens_to_ids(self, tokens):
        return convert_by_vocab(self.vocab, tokens)

    def convert_ids_to_tokens(self, ids):
        return convert_by_vocab(self.inv_vocab, ids)

    @staticmethod
    def convert_tokens_to_string(tokens, clean_up_tokenization_spaces=True):
        return " ".join(tokens).replace(" ##", "").strip()

    def vocab_size(self):
        return len(self.vocab)


class BasicTokenizer:
    def __init__(self, do_lower_case=True):
        self.do_lower_case = do_lower_case

    def tokenize(self, text):
        if self.do_lower_case:
            text = "".join(c for c in unicodedata.normalize("NFD", text.lower()) if unicodedata.category(c) != "Mn")
        return re.findall(r"\w+|[^\w\s]", text)


class WordpieceTokenizer:
    def __init__(self, vocab, unk_token="[UNK]", max_input_chars_per_word=200):
        self.vocab, self.unk_token, self.max_input_chars_per_word = vocab, unk_token, max_input_chars_per_word

    def tokenize(self, text):
        out = []
        for w in whitespace_tokenize(text):
            if len(w) > self.max_input_chars_per_word:
                out.append(self.unk_token)
                continue
            start, sub, bad = 0, [], False
            while start < len(w):
                end, cur = len(w), None
                while start < end:
                    piece = w[start:end] if start == 0 else "##" + w[start:end]
                    if piece in self.vocab:
                        cur = piece
                        break
                    end -= 1
                if cur is None:
                    bad = True
                    break
                sub.append(cur)
                start = end
            out.extend([self.unk_token] if bad else sub)
        return out
