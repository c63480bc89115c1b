#!/usr/bin/env python3
"""Latency of small verify batches through the host-buffer entry point (bbs_verify_batch, pinned buffers): what a caller
with a handful of signatures sees.  BLS12-381, L = 10; median of 20 calls per size."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bbs_sign_b200 import _native, api  # noqa: E402

lib = _native.load()
L = 10
ctx = api.BatchContext(api.BLS12_381, bench.IRTF_PK, header=b"", n_messages=L)
for n in (1, 8, 32, 33, 256, 1024, 4096, 16384, 65536):
    msgs, offs, sigs, expect = bench.make_workload(ctx, lib, n, L, seed=3)
    p = [torch.from_numpy(a).pin_memory() for a in (sigs, msgs, offs.view(np.int64))]
    st = torch.zeros(n, dtype=torch.uint8).pin_memory()
    ts = []
    for _ in range(23):
        t0 = time.perf_counter()
        assert lib.bbs_verify_batch(ctx.handle, n, bench.ptr(p[0]), bench.ptr(p[1]), bench.ptr(p[2]), L, bench.ptr(st)) == 0
        ts.append(time.perf_counter() - t0)
    assert np.array_equal(st.numpy(), expect)
    ts = sorted(ts[3:])
    print(f"n = {n:6d}: median {ts[len(ts) // 2] * 1e3:8.3f} ms   ({n / ts[len(ts) // 2]:10.0f} verifies/s)", flush=True)
ctx.close()
