"""Node-local non-persistent checkpoints (``--non-persistent-ckpt-type local``).

The reference delegates this to ``nvidia_resiliency_ext``'s ``LocalCheckpointManager`` (checkpointing.py:363,494-532,867,895)
with optional ``CliqueReplicationStrategy`` (``--replication --replication-jump J --replication-factor F``); that package is
not part of this stack, so the manager is implemented here with the same behaviour:

* every rank writes its *own* complete training state (model chunks, optimizer, scheduler, RNG) as one file under
  ``--non-persistent-local-ckpt-dir`` -- local SSD / ramdisk, no shared file system, no cross-rank traffic -- atomically
  (temp name + ``os.replace``) and keeps only the newest iteration;
* with replication, rank r also hands its blob to ranks ``r + J, r + 2J, ...`` (``F - 1`` holders), which store it next to
  their own file: after losing a node the state of its ranks is retrieved from the surviving holders;
* ``find_latest`` is a collective: the newest iteration for which EVERY rank's state is retrievable (own file or a
  replica somewhere) -- otherwise -1 and the caller falls back to the persistent checkpoint;
* the parallel layout is stored and must match on load (local checkpoints are a same-layout fast path by design).
"""
from __future__ import annotations

import io
import os
import re
import shutil
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

_ITER_RE = re.compile(r"iter_(\d{7})$")


def _rank_world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


class LocalCheckpointManager:
    def __init__(self, local_dir: str, replication: bool = False, replication_jump: Optional[int] = None,
                 replication_factor: int = 2):
        self.dir = local_dir
        self.rank, self.world = _rank_world()
        self.holders: List[int] = []          # ranks that keep a replica of MY state
        self.sources: List[int] = []          # ranks whose replica I keep
        if replication and self.world > 1:
            if not replication_jump:
                raise ValueError("--replication requires --replication-jump")
            for k in range(1, max(1, replication_factor)):
                h, s = (self.rank + k * replication_jump) % self.world, (self.rank - k * replication_jump) % self.world
                if h != self.rank and h not in self.holders:
                    self.holders.append(h)
                if s != self.rank and s not in self.sources:
                    self.sources.append(s)
        os.makedirs(self.dir, exist_ok=True)

    # ---- paths -------------------------------------------------------------------------------------------------------
    def _iter_dir(self, it: int) -> str:
        return os.path.join(self.dir, f"iter_{it:07d}")

    def _file(self, it: int, owner: int) -> str:
        # files of different processes never collide even when several ranks share one local directory
        suffix = "" if owner == self.rank else f".held_by_{self.rank:05d}"
        return os.path.join(self._iter_dir(it), f"rank_{owner:05d}{suffix}.pt")

    def _mine(self) -> Dict[int, List[int]]:
        """iteration -> owners whose state this process can serve"""
        out: Dict[int, List[int]] = {}
        if not os.path.isdir(self.dir):
            return out
        for name in os.listdir(self.dir):
            m = _ITER_RE.match(name)
            if not m:
                continue
            it = int(m.group(1))
            for f in os.listdir(os.path.join(self.dir, name)):
                fm = re.match(r"rank_(\d{5})(?:\.held_by_(\d{5}))?\.pt$", f)
                if fm and (fm.group(2) is None and int(fm.group(1)) == self.rank or fm.group(2) is not None and int(fm.group(2)) == self.rank):
                    out.setdefault(it, []).append(int(fm.group(1)))
        return out

    # ---- save --------------------------------------------------------------------------------------------------------
    @staticmethod
    def _atomic_write(blob: bytes, path: str) -> None:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = f"{path}.tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(blob)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, path)

    def save(self, state: dict, iteration: int) -> None:
        buf = io.BytesIO()
        torch.save(state, buf)
        blob = buf.getvalue()
        self._atomic_write(blob, self._file(iteration, self.rank))
        if self.holders or self.sources:
            # ring exchange: send mine to my holders, receive my sources' (object collectives carry the byte blobs)
            got: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(got, blob if self.holders else None)
            for s in self.sources:
                if got[s] is not None:
                    self._atomic_write(got[s], self._file(iteration, s))
        for it in [i for i in self._mine() if i < iteration]:
            for owner in self._mine().get(it, []):
                try:
                    os.remove(self._file(it, owner))
                except OSError:
                    pass
            try:
                os.rmdir(self._iter_dir(it))                      # only succeeds once every co-located rank has cleaned up
            except OSError:
                pass

    # ---- find / load ---------------------------------------------------------------------------------------------------
    def find_latest(self) -> int:
        mine = self._mine()
        if self.world == 1:
            return max([it for it, owners in mine.items() if self.rank in owners], default=-1)
        everyone: List[Optional[dict]] = [None] * self.world
        dist.all_gather_object(everyone, {it: sorted(o) for it, o in mine.items()})
        best = -1
        for it in sorted({i for d in everyone for i in d}, reverse=True):
            served = {o for d in everyone for o in d.get(it, [])}
            if len(served) == self.world:
                best = it
                break
        return best

    def load(self, iteration: int) -> dict:
        """This rank's state of ``iteration``: from its own file, or from a surviving holder (collective when replicas are
        needed anywhere)."""
        own = self._file(iteration, self.rank)
        have = os.path.exists(own)
        if self.world > 1:
            flags: List[Optional[bool]] = [None] * self.world
            dist.all_gather_object(flags, have)
            if not all(flags):
                # holders publish the blobs of the ranks that lost theirs
                missing = [r for r, ok in enumerate(flags) if not ok]
                offer = {}
                for r in missing:
                    p = self._file(iteration, r)
                    if r != self.rank and os.path.exists(p):
                        offer[r] = open(p, "rb").read()
                offers: List[Optional[dict]] = [None] * self.world
                dist.all_gather_object(offers, offer)
                if not have:
                    for o in offers:
                        if o and self.rank in o:
                            self._atomic_write(o[self.rank], own)      # restore the local copy as well
                            have = True
                            break
        if not have:
            raise FileNotFoundError(f"rank {self.rank}: no local checkpoint for iteration {iteration} (and no replica)")
        return torch.load(own, map_location="cpu", weights_only=False)

    def cleanup(self) -> None:
        shutil.rmtree(self.dir, ignore_errors=True)
