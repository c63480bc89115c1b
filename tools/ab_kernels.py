#!/usr/bin/env python3
"""A/B timing of the per-thread kernels between two CUDA builds of the library (tuning experiments): for each library path
given on the command line, time sign (n = 524,288, L = 10) and verify (n = 65,536) through the *_dev entry points with the
context's CUDA-event profiling, three repetitions each, and print the SM clock seen during the run."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bbs_sign_b200 import _native, api  # noqa: E402


def clocks():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip()
    return out


def run(path):
    lib = _native.load(path)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    L = 10
    ctx = api.BatchContext(api.BLS12_381, bench.IRTF_PK, header=b"", n_messages=L, lib_path=path)
    lib.bbs_ctx_set_profiling(ctx.handle, 1)
    res = {}
    # sign
    n = 524288
    rng = np.random.default_rng(1)
    sc = rng.integers(0, 256, size=(n * L, 32), dtype=np.uint8)
    sc[:, 31] &= 0x3f
    d_sc = torch.from_numpy(sc.reshape(-1)).to(dev)
    d_sig = torch.zeros(n * 80, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    sk = np.frombuffer(bench.IRTF_SK.to_bytes(32, "little"), dtype=np.uint8).copy()
    ts = []
    for _ in range(5):
        assert lib.bbs_core_sign_batch_dev(ctx.handle, bench.ptr(sk), n, bench.ptr(d_sc), L, bench.ptr(d_sig), None, bench.ptr(d_st), sp) == 0
        torch.cuda.synchronize()
        kt = (C.c_float * 3)()
        lib.bbs_ctx_kernel_times(ctx.handle, kt, 3)
        ts.append(round(kt[1], 2))
    res["sign_524288_ms"] = ts[2:]
    res["clk_after_sign"] = clocks()
    # verify
    n = 65536
    msgs, offs, sigs, expect = bench.make_workload(ctx, lib, n, L, seed=3)
    d_m = torch.from_numpy(msgs).to(dev)
    d_o = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_s = torch.from_numpy(sigs).to(dev)
    d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(5):
        assert lib.bbs_verify_batch_dev(ctx.handle, n, bench.ptr(d_s), bench.ptr(d_m), bench.ptr(d_o), L, bench.ptr(d_st), sp) == 0
        torch.cuda.synchronize()
        kt = (C.c_float * 3)()
        lib.bbs_ctx_kernel_times(ctx.handle, kt, 3)
        ts.append([round(x, 2) for x in kt])
    assert np.array_equal(d_st.cpu().numpy(), expect)
    res["verify_65536_ms[h2s,g1,pairing]"] = ts[2:]
    res["clk_after_verify"] = clocks()
    ctx.close()
    # BN254 core_verify, L = 31
    key = np.load(os.path.join(ROOT, "tests", "golden", "bn254_bench_key.npz"))
    n = 131072
    ctx = api.BatchContext(api.BN254, bytes(key["pk"]), header=b"", n_messages=31, lib_path=path)
    lib.bbs_ctx_set_profiling(ctx.handle, 1)
    sc, sigs, expect, _ = bench.make_bn254_workload(ctx, lib, n, seed=9, L=31)
    d_s, d_c = torch.from_numpy(sigs).to(dev), torch.from_numpy(sc).to(dev)
    d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(4):
        assert lib.bbs_core_verify_batch_dev(ctx.handle, n, bench.ptr(d_s), bench.ptr(d_c), 31, bench.ptr(d_st), sp) == 0
        torch.cuda.synchronize()
        kt = (C.c_float * 3)()
        lib.bbs_ctx_kernel_times(ctx.handle, kt, 3)
        ts.append([round(x, 2) for x in kt][1:])
    assert np.array_equal(d_st.cpu().numpy(), expect)
    res["bn254_131072_ms[g1,pairing]"] = ts[2:]
    ctx.close()
    return res


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(p, run(os.path.abspath(p)), flush=True)
