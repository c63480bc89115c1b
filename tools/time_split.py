#!/usr/bin/env python3
"""verify G1 half: one thread per item against the two-task split (bbs_ctx_set_g1_split), by batch size, on one context."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bbs_sign_b200 import _native, api  # noqa: E402

lib = _native.load()
dev = torch.device("cuda", 0)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L = 10
ctx = api.BatchContext(api.BLS12_381, bench.IRTF_PK, header=b"", n_messages=L)
lib.bbs_ctx_set_profiling(ctx.handle, 1)
for n in [int(x) for x in (sys.argv[1:] or "16384 32768 65536 131072 262144".split())]:
    msgs, offs, sigs, expect = bench.make_workload(ctx, lib, n, L, seed=3)
    d_m, d_o, d_s = torch.from_numpy(msgs).to(dev), torch.from_numpy(offs.view(np.int64)).to(dev), torch.from_numpy(sigs).to(dev)
    d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    row = {}
    for name, lim in (("one-thread", 0), ("split", 1 << 40)):
        ctx.set_g1_split(lim)
        ts = []
        for _ in range(4):
            d_st.zero_()
            assert lib.bbs_verify_batch_dev(ctx.handle, n, bench.ptr(d_s), bench.ptr(d_m), bench.ptr(d_o), L, bench.ptr(d_st), sp) == 0
            torch.cuda.synchronize()
            kt = (C.c_float * 3)()
            lib.bbs_ctx_kernel_times(ctx.handle, kt, 3)
            ts.append(round(kt[1], 2))
        assert np.array_equal(d_st.cpu().numpy(), expect), name
        row[name] = ts[1:]
    print(n, row, flush=True)
ctx.close()
