#!/usr/bin/env python3
"""Times bbs_sign_batch (host buffers) on cuda:0: BLS12-381, n x L=10 messages of 32 bytes (BASELINE configs[4] sign leg)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bbs_sign_b200 import api, _native
lib = _native.load()
n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 10
ctx = api.BatchContext(api.BLS12_381, bench.IRTF_PK, header=b"", n_messages=L)
rng = np.random.default_rng(1)
msgs = rng.integers(0, 256, size=n * L * 32, dtype=np.uint8)
offs = np.arange(n * L + 1, dtype=np.uint64) * 32
sigs = np.zeros(n * 80, dtype=np.uint8); st = np.zeros(n, dtype=np.uint8)
sk = np.frombuffer(bench.IRTF_SK.to_bytes(32, "little"), dtype=np.uint8).copy()
import ctypes as C
import torch
pin = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).pin_memory().numpy()
msgs, sigs, st = pin(msgs), pin(sigs), pin(st)
offs = pin(offs).view(np.uint64)
lib.bbs_ctx_set_profiling(ctx.handle, 1)
for rep in range(3):
    t0 = time.perf_counter()
    rc = lib.bbs_sign_batch(ctx.handle, bench.ptr(sk), n, bench.ptr(msgs), bench.ptr(offs), L, bench.ptr(sigs), None, bench.ptr(st))
    dt = time.perf_counter() - t0
    assert rc == 0 and (st == 1).all()
kt = (C.c_float * 3)()
lib.bbs_ctx_kernel_times(ctx.handle, kt, 3)
print(f"sign_batch (pinned host buffers): {n / dt:,.0f} signatures/s ({dt * 1e3:.1f} ms for {n}); kernels: msg_to_scalars {kt[0]:.2f} ms, sign {kt[1]:.2f} ms")
