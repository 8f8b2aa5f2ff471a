"""Rank-BPE ("tiktoken"-format) and multimodal conversation tokenizers.

Parity: reference ``megatron/training/tokenizer/tokenizer.py`` (``CustomTikTokenizer`` :692, ``reload_mergeable_ranks``
:653, ``_NullMultimodalTokenizer`` :884) and ``multimodal_tokenizer.py`` (``MultimodalTokenizer`` :71).

* ``TikTokenizer`` implements the byte-level rank-BPE itself (regex pre-split, then repeatedly merge the adjacent
  pair whose concatenation has the lowest rank) -- no dependency on the ``tiktoken`` package; ids are offset by the
  special-token block exactly as the reference lays them out (specials first, then ranks).
* ``MultimodalTokenizer`` renders conversations turn by turn with small Python formatters (no jinja templates), so the
  loss-target mask falls out of the same pass that produces the tokens.
"""
from __future__ import annotations

import base64
import json
from dataclasses import dataclass
from functools import lru_cache
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .tokenizer import MegatronTokenizer

IMAGE_TOKEN = "<image>"
IGNORE_INDEX = -100
DEFAULT_IMAGE_TOKEN_INDEX = -200

PATTERN_TIKTOKEN = r"[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+"
PATTERN_TIKTOKEN_V2 = (r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+|"
                       r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*|\p{N}|"
                       r" ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+")


def load_mergeable_ranks(path: str, max_vocab: Optional[int] = None) -> Dict[bytes, int]:
    """Either the reference's JSON list ``[{rank, token_bytes(b64), token_str}]`` or tiktoken's text format
    (``<b64 token> <rank>`` per line)."""
    ranks: Dict[bytes, int] = {}
    if path.endswith(".json"):
        with open(path) as f:
            vocab = json.load(f)
        if max_vocab is not None:
            vocab = vocab[:max_vocab]
        for i, x in enumerate(vocab):
            if x["rank"] != i:
                raise ValueError(f"{path}: entry {i} has rank {x['rank']}")
            ranks[base64.b64decode(x["token_bytes"])] = i
    else:
        with open(path, "rb") as f:
            for line in f:
                if line.strip():
                    tok, rank = line.split()
                    if max_vocab is None or int(rank) < max_vocab:
                        ranks[base64.b64decode(tok)] = int(rank)
    if sorted(ranks.values()) != list(range(len(ranks))):
        raise ValueError(f"{path}: ranks are not a permutation of 0..{len(ranks) - 1}")
    return ranks


class TikTokenizer(MegatronTokenizer):
    def __init__(self, path: str, pattern: str = "v2", vocab_size: Optional[int] = None, num_special_tokens: int = 1000,
                 special_tokens: Optional[List[str]] = None):
        super().__init__("TikTokenizer")
        import regex

        base = ["<unk>", "<s>", "</s>"]
        special_tokens = list(special_tokens) if special_tokens is not None else list(base)
        if not set(base) <= set(special_tokens) or len(set(special_tokens)) != len(special_tokens):
            raise ValueError(f"special tokens must be unique and include {base}")
        if len(special_tokens) > num_special_tokens:
            raise ValueError("more special tokens than num_special_tokens")
        special_tokens += [f"<SPECIAL_{i}>" for i in range(len(special_tokens), num_special_tokens)]
        ranks = load_mergeable_ranks(path, None if vocab_size is None else vocab_size - num_special_tokens)
        self._vocab_size = vocab_size if vocab_size is not None else len(ranks) + num_special_tokens
        self._ranks = ranks
        self._off = num_special_tokens
        self._special = {t: i for i, t in enumerate(special_tokens)}
        self._id_to_bytes = {i + self._off: b for b, i in ranks.items()}
        self._id_to_special = {i: t for t, i in self._special.items()}
        self._pat = regex.compile({"v1": PATTERN_TIKTOKEN, "v2": PATTERN_TIKTOKEN_V2}.get(pattern, pattern))
        self._bpe = lru_cache(maxsize=1 << 16)(self._bpe_uncached)

    def _bpe_uncached(self, piece: bytes) -> Tuple[int, ...]:
        r = self._ranks.get(piece)
        if r is not None:
            return (r,)
        parts = [piece[i:i + 1] for i in range(len(piece))]
        while len(parts) > 1:
            best, at = None, -1
            for i in range(len(parts) - 1):
                rk = self._ranks.get(parts[i] + parts[i + 1])
                if rk is not None and (best is None or rk < best):
                    best, at = rk, i
            if best is None:
                break
            parts[at:at + 2] = [parts[at] + parts[at + 1]]
        return tuple(self._ranks[p] for p in parts)

    def tokenize(self, s: str, bos: bool = False, eos: bool = False) -> List[int]:
        out = [self.bos] if bos else []
        for m in self._pat.findall(s):
            out.extend(r + self._off for r in self._bpe(m.encode("utf-8")))
        if eos:
            out.append(self.eos)
        return out

    def detokenize(self, ids: Sequence[int]) -> str:
        buf = bytearray()
        for i in ids:
            i = int(i)
            if i in self._id_to_bytes:
                buf += self._id_to_bytes[i]
            elif i in self._id_to_special:
                buf += self._id_to_special[i].encode()
        return buf.decode("utf-8", errors="replace")

    def offsets(self, ids: Sequence[int], text: str) -> List[int]:
        out, pos = [], 0
        for i in ids:
            out.append(pos)
            pos += len(self.detokenize([i]))
        return out

    @property
    def vocab_size(self):
        return self._vocab_size

    @property
    def vocab(self):
        v = {b: r + self._off for b, r in self._ranks.items()}
        v.update(self._special)
        return v

    @property
    def inv_vocab(self):
        return {**self._id_to_bytes, **self._id_to_special}

    @property
    def bos(self):
        return self._special["<s>"]

    @property
    def eos(self):
        return self._special["</s>"]

    @property
    def unk(self):
        return self._special["<unk>"]

    @property
    def eod(self):
        return self.eos


# ---------------------------------------------------------------------------------------------------------------
# multimodal
# ---------------------------------------------------------------------------------------------------------------
IMAGE_TAGS = {"nvlm": ("<Image>", "</Image>"), "internvl": ("<img>", "</img>"), "": None}


@dataclass
class PromptFormat:
    """``open(role)`` / ``close(role)`` wrap every turn; ``assistant_prefix`` is masked out of the loss target."""
    bos: str
    open: Callable[[str], str]
    close: Callable[[str], str]
    has_system_role: bool = True


def _chatml(start: str, end: str, bos: str = "") -> PromptFormat:
    return PromptFormat(bos, lambda r: f"{start}{r}\n", lambda r: f"{end}\n")


PROMPT_FORMATS: Dict[str, PromptFormat] = {
    "mistral": PromptFormat("<s>", lambda r: "[INST] " if r == "user" else " ",
                            lambda r: "[/INST]" if r == "user" else "</s>", has_system_role=False),
    "llama3": PromptFormat("<|begin_of_text|>", lambda r: f"<|start_header_id|>{r}<|end_header_id|>\n\n",
                           lambda r: "<|eot_id|>"),
    "chatml": _chatml("<|im_start|>", "<|im_end|>"),
    "nemotron5": _chatml("<SPECIAL_14>", "<SPECIAL_15>", "<s>"),
}
for _alias, _to in {"llama3p1": "llama3", "qwen2p0": "chatml", "qwen2p5": "chatml", "nvlm-yi-34b": "chatml"}.items():
    PROMPT_FORMATS[_alias] = PROMPT_FORMATS[_to]


class MultimodalTokenizer(MegatronTokenizer):
    """Wraps a text tokenizer (anything with ``tokenize(str) -> ids`` / ``detokenize``), adds image placeholders,
    image tags and chat rendering with loss-target masking."""

    def __init__(self, tokenizer, prompt_format: str = "chatml", special_tokens: Optional[List[str]] = None,
                 image_tag_type: str = "", image_token_id: int = DEFAULT_IMAGE_TOKEN_INDEX):
        super().__init__("MultimodalTokenizer")
        if prompt_format not in PROMPT_FORMATS:
            raise NotImplementedError(f"unknown multimodal prompt format {prompt_format!r}")
        self._tok, self._fmt = tokenizer, PROMPT_FORMATS[prompt_format]
        self._image_tag = IMAGE_TAGS[image_tag_type]
        self._extra: Dict[str, int] = {IMAGE_TOKEN: image_token_id}
        base = tokenizer.vocab_size
        for t in special_tokens or []:
            if t not in self._extra:
                self._extra[t] = base + len(self._extra) - 1
        self._extra_inv = {i: t for t, i in self._extra.items()}

    # text with <image> (and other registered specials) -> ids; specials never go through the text tokenizer
    def _encode(self, text: str) -> List[int]:
        if self._image_tag is not None:
            text = text.replace(IMAGE_TOKEN, f"{self._image_tag[0]}{IMAGE_TOKEN}{self._image_tag[1]}")
        pieces: List[Union[str, int]] = [text]
        for tok, tid in self._extra.items():
            nxt: List[Union[str, int]] = []
            for p in pieces:
                if isinstance(p, int):
                    nxt.append(p)
                    continue
                parts = p.split(tok)
                for j, q in enumerate(parts):
                    if q:
                        nxt.append(q)
                    if j + 1 < len(parts):
                        nxt.append(tid)
            pieces = nxt
        out: List[int] = []
        for p in pieces:
            out.extend([p] if isinstance(p, int) else self._tok.tokenize(p))
        return out

    def tokenize(self, text: Union[str, List[Dict]]):
        if isinstance(text, list):
            return self.tokenize_conversation(text, False, True).tolist()
        return self._encode(text)

    def tokenize_conversation(self, conversation: List[Dict], return_target: bool, add_generation_prompt: bool):
        f = self._fmt
        conv = [dict(t) for t in conversation]
        if not f.has_system_role and conv and conv[0]["role"] == "system":
            conv = conv[1:]
        tokens: List[int] = self._encode(f.bos) if f.bos else []
        target: List[int] = [IGNORE_INDEX] * len(tokens)
        for turn in conv:
            role, content = turn["role"].lower(), turn["content"]
            if not content:
                raise ValueError(f"empty turn in conversation: {conversation}")
            head, body = self._encode(f.open(role)), self._encode(content + f.close(role))
            tokens += head + body
            if role == "assistant":
                if IMAGE_TOKEN in content:
                    raise RuntimeError(f"{IMAGE_TOKEN} not allowed in assistant content")
                target += [IGNORE_INDEX] * len(head) + body
            else:
                target += [IGNORE_INDEX] * (len(head) + len(body))
        if add_generation_prompt:
            head = self._encode(f.open("assistant"))
            tokens += head
            target += [IGNORE_INDEX] * len(head)
        tokens_np = np.asarray(tokens, dtype=np.int64)
        return (tokens_np, np.asarray(target, dtype=np.int64)) if return_target else tokens_np

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            tokens = [tokens]
        ids = [self._extra[t] if t in self._extra else self._tok.tokenize(t)[0] for t in tokens]
        return ids if len(ids) > 1 else ids[0]

    def detokenize(self, tokens: Sequence[int]) -> str:
        out, run = [], []
        for t in tokens:
            t = int(t)
            if t in self._extra_inv:
                if run:
                    out.append(self._tok.detokenize(run))
                    run = []
                out.append(self._extra_inv[t])
            else:
                run.append(t)
        if run:
            out.append(self._tok.detokenize(run))
        return "".join(out)

    def get_special_tokens(self):
        return list(self._extra)

    @property
    def pad(self):
        return getattr(self._tok, "pad", 0)

    @property
    def eod(self):
        return self._tok.eod

    @property
    def vocab(self):
        return getattr(self._tok, "vocab", {})

    @property
    def inv_vocab(self):
        return getattr(self._tok, "inv_vocab", {})

    @property
    def vocab_size(self):
        return self._tok.vocab_size + len(self._extra) - 1


class NullMultimodalTokenizer(MegatronTokenizer):
    """Space-separated integer "text" with an image placeholder (reference ``_NullMultimodalTokenizer``)."""

    def __init__(self, vocab_size: int, image_token: Optional[str] = None, image_token_id: Optional[int] = None):
        super().__init__("NullMultimodalTokenizer")
        self._n = int(vocab_size)
        self._image_token = image_token or IMAGE_TOKEN
        self._image_token_id = image_token_id if image_token_id is not None else DEFAULT_IMAGE_TOKEN_INDEX

    def tokenize(self, text):
        return [self._image_token_id if t == self._image_token else int(t) for t in text.split()]

    def detokenize(self, ids):
        return " ".join(self._image_token if int(i) == self._image_token_id else str(int(i)) for i in ids)

    def offsets(self, ids, text):
        out, pos = [], 0
        for i in ids:
            out.append(pos)
            pos += len(self._image_token if int(i) == self._image_token_id else str(int(i))) + 1
        return out

    def convert_tokens_to_ids(self, tokens):
        ids = self.tokenize(tokens if isinstance(tokens, str) else " ".join(tokens))
        return ids if len(ids) > 1 else ids[0]

    @property
    def vocab_size(self):
        return self._n + 1

    @property
    def eod(self):
        return self._n

    @property
    def pad(self):
        return 0
