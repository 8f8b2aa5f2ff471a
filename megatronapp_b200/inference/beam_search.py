"""Beam search over the KV-cached decode loop (reference
``megatron/inference/text_generation/generation.py beam_search_and_return_on_first_stage`` +
``beam_utils.BeamHypotheses`` with length penalty; single prompt, ``beam_size`` hypotheses share the
batch dimension and the KV cache rows are re-gathered by ``swap_key_value_dict`` each step)."""
from __future__ import annotations

from typing import List, Tuple

import torch


class BeamHypotheses:
    def __init__(self, num_beams: int, length_penalty: float = 1.0):
        self.num_beams, self.length_penalty = num_beams, length_penalty
        self.beams: List[Tuple[float, torch.Tensor]] = []
        self.worst = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp: torch.Tensor, sum_logprobs: float, length: int):
        score = sum_logprobs / (max(length, 1) ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                self.beams.sort(key=lambda x: x[0])
                self.beams.pop(0)
            self.worst = min(s for s, _ in self.beams)

    def is_done(self, best_sum_logprobs: float, cur_len: int) -> bool:
        if len(self) < self.num_beams:
            return False
        return self.worst >= best_sum_logprobs / (max(cur_len, 1) ** self.length_penalty)


@torch.no_grad()
def beam_search(wrapper, prompt_tokens: List[int], beam_size: int, tokens_to_generate: int, stop_token: int,
                length_penalty: float = 1.0, num_return_gen: int = 1):
    """Returns ``(tokens [num_return_gen, len], scores [num_return_gen])``."""
    dev = next(wrapper.model.parameters()).device
    plen = len(prompt_tokens)
    total = plen + tokens_to_generate
    tokens = torch.zeros(beam_size, total, dtype=torch.int64, device=dev)
    tokens[:, :plen] = torch.tensor(prompt_tokens, device=dev)
    wrapper.prep_model_for_inference(tokens)
    hyps = BeamHypotheses(beam_size, length_penalty)
    scores = torch.zeros(beam_size, device=dev)
    start = 0
    done = False
    for ctx_len in range(plen, total):
        logits = wrapper.run_one_forward_step(tokens, start, ctx_len)[:, -1].float()
        logp = torch.log_softmax(logits, dim=-1)
        V = logp.shape[-1]
        cand = logp + scores.unsqueeze(1)
        flat = cand.view(-1) if start > 0 else cand[0]            # first step: all beams are identical
        best_scores, best_idx = torch.topk(flat, 2 * beam_size)
        beam_ids, words = best_idx // V, best_idx % V
        nxt = []
        for rank, (w, b, s) in enumerate(zip(words.tolist(), beam_ids.tolist(), best_scores.tolist())):
            if w == stop_token:
                if rank < beam_size:
                    hyps.add(tokens[b, :ctx_len].clone(), s, ctx_len - plen)
            else:
                nxt.append((w, b, s))
            if len(nxt) == beam_size:
                break
        if hyps.is_done(float(best_scores.max()), ctx_len + 1 - plen):
            done = True
            break
        bidx = torch.tensor([b for _, b, _ in nxt], device=dev)
        tokens = tokens[bidx]
        tokens[:, ctx_len] = torch.tensor([w for w, _, _ in nxt], device=dev)
        scores = torch.tensor([s for _, _, s in nxt], device=dev)
        wrapper.context.swap_key_value_dict(bidx)
        start = ctx_len
    if not done:
        for b in range(beam_size):
            hyps.add(tokens[b].clone(), float(scores[b]), total - plen)
    best = sorted(hyps.beams, key=lambda x: x[0], reverse=True)[:num_return_gen]
    return [h for _, h in best], torch.tensor([s for s, _ in best])
