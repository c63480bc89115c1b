"""Multi-GPU sharding of a batch by signature index: contiguous split, one context and one
host thread per GPU, statuses written straight into the caller's slice, no collective on the data path."""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .api import BatchContext, Ciphersuite


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """GPU g gets [g*n/world, (g+1)*n/world) -- the split the north star names."""
    return [(g * n // world, (g + 1) * n // world) for g in range(world)]


class ShardedVerifier:
    """`verify_batch` over several GPUs of one box (one BatchContext per device)."""

    def __init__(self, suite: Ciphersuite, pk: bytes, header: bytes, n_messages: int, devices: Sequence[int],
                 lib_path: Optional[str] = None):
        self.suite = suite
        self.ctxs = [BatchContext(suite, pk, header, n_messages, device=d, lib_path=lib_path) for d in devices]

    def verify_batch(self, signatures, messages: Sequence[Sequence[bytes]]) -> np.ndarray:
        sigs = np.frombuffer(signatures, dtype=np.uint8) if not isinstance(signatures, np.ndarray) else signatures.reshape(-1)
        sb = self.suite.signature_bytes
        n = sigs.size // sb
        out = np.full(n, 255, dtype=np.uint8)
        errs: list = []

        def work(ctx, lo, hi):
            try:
                if hi > lo:
                    out[lo:hi] = ctx.verify_batch(sigs[lo * sb: hi * sb], messages[lo:hi])
            except Exception as e:  # pragma: no cover
                errs.append(e)

        ts = [threading.Thread(target=work, args=(c, lo, hi)) for c, (lo, hi) in zip(self.ctxs, shard_bounds(n, len(self.ctxs)))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]
        return out

    def close(self):
        for c in self.ctxs:
            c.close()
