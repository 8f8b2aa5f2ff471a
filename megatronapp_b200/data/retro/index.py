"""Approximate nearest-neighbour indexes for the Retro chunk embeddings (reference ``core/datasets/retro/index/`` wraps
``faiss.index_factory("OPQ..,IVF..,PQ..")``; faiss is not available here, so the IVF-PQ index is implemented directly).

``IVFPQIndex`` (inner-product metric on L2-normalised embeddings):

* train: spherical k-means coarse quantiser (``nlist`` centroids), then a product quantiser -- ``m`` sub-spaces x 256
  centroids -- on the residuals ``x - c[assign]``;
* add: assign, encode residuals to ``m`` bytes, append to the inverted list of the centroid;
* search: the ``nprobe`` best lists per query from one GEMM, then asymmetric distance computation -- per query a
  ``[m, 256]`` table of sub-vector inner products, per candidate ``q . c + sum_j table[j, code_j]`` -- grouped by list so
  every list is visited once per batch with a gather; running top-k merged per query.

All heavy steps are matrix products / gathers on ``device`` (the GPU when present)."""
from __future__ import annotations

import re
from typing import Optional, Tuple

import numpy as np
import torch


def _device(device=None):
    return torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")


def kmeans(x: torch.Tensor, k: int, iters: int = 12, seed: int = 0, spherical: bool = False) -> torch.Tensor:
    """Lloyd iterations with k-means++-free random init; empty clusters are re-seeded from the largest one."""
    n = x.shape[0]
    g = torch.Generator(device="cpu").manual_seed(seed)
    k = min(k, n)
    c = x[torch.randperm(n, generator=g)[:k].to(x.device)].clone()
    for _ in range(iters):
        a = _assign(x, c)
        sums = torch.zeros_like(c).index_add_(0, a, x)
        cnt = torch.zeros(k, device=x.device, dtype=x.dtype).index_add_(0, a, torch.ones(n, device=x.device, dtype=x.dtype))
        empty = cnt == 0
        c = torch.where(empty[:, None], c, sums / cnt.clamp_min(1)[:, None])
        if empty.any():
            big = int(cnt.argmax())
            pool = (a == big).nonzero().flatten()
            pick = pool[torch.randint(0, len(pool), (int(empty.sum()),), generator=g).to(x.device)]
            c[empty] = x[pick] + 1e-4 * torch.randn(int(empty.sum()), x.shape[1], generator=g).to(x.device, x.dtype)
        if spherical:
            c = torch.nn.functional.normalize(c, dim=-1)
    return c


def _assign(x: torch.Tensor, c: torch.Tensor, block: int = 65536) -> torch.Tensor:
    c2 = (c * c).sum(1)
    out = []
    for i in range(0, x.shape[0], block):
        out.append((c2[None, :] - 2.0 * (x[i:i + block] @ c.t())).argmin(1))       # argmin ||x - c||^2
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.long, device=x.device)


class FlatIndex:
    """Exact inner-product search (the reference's ``Flat`` / validation index)."""

    def __init__(self, dim: int, device=None):
        self.dim, self.device = dim, _device(device)
        self.x = torch.zeros(0, dim, device=self.device)
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return self.x.shape[0]

    def train(self, x) -> None:
        pass

    def add(self, x) -> None:
        self.x = torch.cat([self.x, torch.as_tensor(x, dtype=torch.float32, device=self.device)])

    def search(self, q, k: int, nprobe: Optional[int] = None, block: int = 4096) -> Tuple[np.ndarray, np.ndarray]:
        q = torch.as_tensor(q, dtype=torch.float32, device=self.device)
        k = min(k, self.ntotal)
        S, I = [], []
        for i in range(0, q.shape[0], block):
            s, idx = (q[i:i + block] @ self.x.t()).topk(k, dim=1)
            S.append(s)
            I.append(idx)
        return torch.cat(S).cpu().numpy(), torch.cat(I).cpu().numpy()

    def state(self):
        return {"kind": "Flat", "dim": self.dim, "x": self.x.cpu().numpy()}

    @classmethod
    def from_state(cls, st, device=None):
        ix = cls(int(st["dim"]), device)
        ix.add(st["x"])
        return ix


class IVFPQIndex:
    def __init__(self, dim: int, nlist: int, m: int = 0, device=None, seed: int = 0):
        if m and dim % m:
            raise ValueError(f"PQ needs dim % m == 0 (dim {dim}, m {m})")
        self.dim, self.nlist, self.m, self.seed, self.device = dim, nlist, m, seed, _device(device)
        self.centroids: Optional[torch.Tensor] = None          # [nlist, dim]
        self.codebooks: Optional[torch.Tensor] = None          # [m, 256, dim / m]
        self.is_trained = False
        self._pending = []                                     # (assign, codes | residual vectors, ids)
        self._n = 0
        self.list_offsets = self.list_ids = self.list_codes = None

    @property
    def ntotal(self) -> int:
        return self._n

    # ---- train -------------------------------------------------------------------------------------------------------
    def train(self, x) -> None:
        x = torch.as_tensor(x, dtype=torch.float32, device=self.device)
        self.nlist = min(self.nlist, max(1, x.shape[0] // 4))
        self.centroids = kmeans(x, self.nlist, seed=self.seed)
        self.nlist = self.centroids.shape[0]
        if self.m:
            r = x - self.centroids[_assign(x, self.centroids)]
            ds = self.dim // self.m
            self.codebooks = torch.stack([kmeans(r[:, j * ds:(j + 1) * ds], 256, seed=self.seed + 1 + j) for j in range(self.m)])
            if self.codebooks.shape[1] < 256:                  # tiny training sets: pad the codebook, codes stay valid
                pad = 256 - self.codebooks.shape[1]
                self.codebooks = torch.cat([self.codebooks, self.codebooks[:, :1].expand(-1, pad, -1)], 1)
        self.is_trained = True

    def _encode(self, r: torch.Tensor) -> torch.Tensor:
        ds = self.dim // self.m
        return torch.stack([_assign(r[:, j * ds:(j + 1) * ds].contiguous(), self.codebooks[j]) for j in range(self.m)], 1).to(torch.uint8)

    # ---- add ---------------------------------------------------------------------------------------------------------
    def add(self, x) -> None:
        if not self.is_trained:
            raise RuntimeError("train the index before adding")
        x = torch.as_tensor(x, dtype=torch.float32, device=self.device)
        a = _assign(x, self.centroids)
        r = x - self.centroids[a]
        payload = self._encode(r) if self.m else r
        ids = torch.arange(self._n, self._n + x.shape[0], device=self.device)
        self._pending.append((a, payload, ids))
        self._n += x.shape[0]
        self.list_offsets = None

    def _finalize(self) -> None:
        if self.list_offsets is not None:
            return
        if not self._pending:                                   # trained, nothing added yet
            self._pending = [(torch.zeros(0, dtype=torch.long, device=self.device),
                              torch.zeros(0, self.m if self.m else self.dim, dtype=torch.uint8 if self.m else torch.float32, device=self.device),
                              torch.zeros(0, dtype=torch.long, device=self.device))]
        a = torch.cat([p[0] for p in self._pending])
        payload = torch.cat([p[1] for p in self._pending])
        ids = torch.cat([p[2] for p in self._pending])
        order = torch.argsort(a, stable=True)
        self._pending = [(a[order], payload[order], ids[order])]            # one sorted part: later adds append to it
        self.list_ids, self.list_codes = ids[order], payload[order]
        counts = torch.bincount(a[order], minlength=self.nlist)
        self.list_offsets = torch.cat([torch.zeros(1, dtype=torch.long, device=self.device), counts.cumsum(0)])

    # ---- search ------------------------------------------------------------------------------------------------------
    def search(self, q, k: int, nprobe: int = 16, block: int = 1024) -> Tuple[np.ndarray, np.ndarray]:
        self._finalize()
        q = torch.as_tensor(q, dtype=torch.float32, device=self.device)
        nprobe = min(nprobe, self.nlist)
        S = torch.full((q.shape[0], k), -float("inf"), device=self.device)
        I = torch.full((q.shape[0], k), -1, dtype=torch.long, device=self.device)
        off = self.list_offsets.tolist()
        ds = self.dim // self.m if self.m else 0
        for b0 in range(0, q.shape[0], block):
            qb = q[b0:b0 + block]
            coarse = qb @ self.centroids.t()                                       # [nq, nlist]
            probe = coarse.topk(nprobe, dim=1).indices
            lut = torch.einsum("qjd,jcd->qjc", qb.view(-1, self.m, ds), self.codebooks) if self.m else None
            best_s, best_i = S[b0:b0 + block], I[b0:b0 + block]
            hit = torch.zeros(qb.shape[0], self.nlist, dtype=torch.bool, device=self.device).scatter_(1, probe, True)
            for l in hit.any(0).nonzero().flatten().tolist():
                lo, hi = off[l], off[l + 1]
                if hi == lo:
                    continue
                qs = hit[:, l].nonzero().flatten()
                if self.m:
                    codes = self.list_codes[lo:hi].long()                          # [n_l, m]
                    adc = lut[qs][:, torch.arange(self.m, device=self.device)[None, :], codes].sum(-1)   # [nq_l, n_l]
                else:
                    adc = qb[qs] @ self.list_codes[lo:hi].t()
                sc = adc + coarse[qs, l][:, None]
                kk = min(k, hi - lo)
                s_l, i_l = sc.topk(kk, dim=1)
                cat_s = torch.cat([best_s[qs], s_l], 1)
                cat_i = torch.cat([best_i[qs], self.list_ids[lo:hi][i_l]], 1)
                top = cat_s.topk(k, dim=1)
                best_s[qs], best_i[qs] = top.values, cat_i.gather(1, top.indices)
        return S.cpu().numpy(), I.cpu().numpy()

    # ---- persistence -------------------------------------------------------------------------------------------------
    def state(self):
        self._finalize()
        return {"kind": "IVFPQ", "dim": self.dim, "nlist": self.nlist, "m": self.m, "n": self._n,
                "centroids": self.centroids.cpu().numpy(), "codebooks": self.codebooks.cpu().numpy() if self.m else np.zeros(0),
                "assign": self._pending[0][0].cpu().numpy(), "payload": self._pending[0][1].cpu().numpy(),
                "ids": self._pending[0][2].cpu().numpy()}

    @classmethod
    def from_state(cls, st, device=None):
        ix = cls(int(st["dim"]), int(st["nlist"]), int(st["m"]), device)
        ix.centroids = torch.as_tensor(st["centroids"], device=ix.device)
        ix.codebooks = torch.as_tensor(st["codebooks"], device=ix.device) if ix.m else None
        ix.is_trained, ix._n = True, int(st["n"])
        if ix._n:
            ix._pending = [(torch.as_tensor(st["assign"], device=ix.device), torch.as_tensor(st["payload"], device=ix.device),
                            torch.as_tensor(st["ids"], device=ix.device))]
        return ix


def index_factory(spec: str, dim: int, device=None, seed: int = 0):
    """``"Flat"`` | ``"IVF<nlist>,Flat"`` | ``"IVF<nlist>,PQ<m>"`` (an ``OPQ..`` / ``_HNSW..`` decoration of the reference's
    default string is accepted and ignored: the coarse search here is already an exact GEMM)."""
    parts = [p for p in spec.split(",") if not p.startswith("OPQ")]
    if parts == ["Flat"]:
        return FlatIndex(dim, device)
    m_ivf = re.match(r"IVF(\d+)", parts[0])
    if not m_ivf:
        raise ValueError(f"unsupported index string {spec!r}")
    m_pq = re.match(r"PQ(\d+)", parts[1]) if len(parts) > 1 else None
    return IVFPQIndex(dim, int(m_ivf.group(1)), int(m_pq.group(1)) if m_pq else 0, device, seed)


def save_index(ix, path: str) -> None:
    np.savez(path, **ix.state())


def load_index(path: str, device=None):
    st = dict(np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=False))
    return (FlatIndex if str(st["kind"]) == "Flat" else IVFPQIndex).from_state(st, device)
