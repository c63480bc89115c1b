#!/usr/bin/env python3
"""bench.py -- BLS12-381 BBS batch verify throughput on B200 (BASELINE.json metric, config 2).

A step = one pass of the hot path (msg_to_scalars -> core_verify G1 half -> two-pair pairing) over one
batch of 65,536 synthetic signatures x L=10 messages of 32 bytes under one issuer key, 1/16 of them
corrupted.  Signatures are produced by the library's own `bbs_sign_batch` on the GPU (no oracle on the
product path); the status vector of every run is checked against the construction (valid -> 1,
corrupted -> 0).

  python bench.py --gpus N --steps K --warmup W           our arm (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference ...                    CPU arm: the oracle port of the reference's path

Prints ONE JSON line (rank 0)."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bls12_381_bbs_verifies_per_sec_L10"
UNIT = "verifies/s"
N_DEFAULT = 65536
L_DEFAULT = 10
# IRTF draft key pair (test_vector.rs:140-160): sk little-endian, pk compressed
IRTF_SK = 0x60e55110f76883a13d030b2f6bd11883422d5abde717569fc0731f51237169fc
IRTF_PK = bytes.fromhex("a820f230f6ae38503b86c70dc50b61c58a77e45c39ab25c0652bbaa8fa136f2851bd4781c9dcde39fc9d1d52c9e60268"
                        "061e7d7632171d91aa8d460acee0e96f1e7c4cfb12d3ff9ab5d5dc91c277db75c845d649ef3c4f63aebc364cd55ded0c")
# counting model of SURVEY 8(d) / Appendix D: 32x32->64 products per BLS12-381 verify at L=10
PRODUCTS_PER_VERIFY = 19509 * 300 + 3285 * 234              # 6.62e6
PRODUCTS_PAIRING = (8116 + 7777) * 300 + (76 * 300 + 380 * 234)  # Miller + final exp + its one inversion
PRODUCTS_G1 = PRODUCTS_PER_VERIFY - PRODUCTS_PAIRING
SIG_BYTES = 80
MSG_BYTES = 32


def ptr(a):
    return C.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr())


def make_workload(ctx, lib, n, L, seed):
    """Synthetic batch: random 32-byte messages, signed on the GPU, 1/16 corrupted."""
    rng = np.random.default_rng(seed)
    msgs = rng.integers(0, 256, size=n * L * MSG_BYTES, dtype=np.uint8)
    offs = (np.arange(n * L + 1, dtype=np.uint64) * MSG_BYTES)
    sigs = np.zeros(n * SIG_BYTES, dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    sk = np.frombuffer(IRTF_SK.to_bytes(32, "little"), dtype=np.uint8).copy()
    rc = lib.bbs_sign_batch(ctx.handle, ptr(sk), n, ptr(msgs), ptr(offs), L, ptr(sigs), None, ptr(st))
    if rc != 0 or not (st == 1).all():
        raise RuntimeError(f"bbs_sign_batch failed rc={rc}: {lib.bbs_last_error().decode()}")
    sigs = sigs.reshape(n, SIG_BYTES)
    expect = np.ones(n, dtype=np.uint8)
    bad = np.arange(3, n, 16)
    for k, i in enumerate(bad):
        kind = k % 4
        if kind == 0:      # flipped message byte
            msgs[(i * L + (k % L)) * MSG_BYTES + 7] ^= 0x40
        elif kind == 1:    # e ^ 1
            sigs[i, 48] ^= 1
        elif kind == 2:    # A = identity
            sigs[i, :48] = 0
            sigs[i, 0] = 0xC0
        else:              # A of the neighbouring signature
            sigs[i, :48] = sigs[i - 1, :48]
        expect[i] = 0
    return msgs, offs, np.ascontiguousarray(sigs.reshape(-1)), expect


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        # under load = upper half of the samples
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bbs_sign_b200 import api, _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = _native.load()
    n, L = args.n, args.L
    ctx = api.BatchContext(api.BLS12_381, IRTF_PK, header=b"", n_messages=L, device=local)
    msgs, offs, sigs, expect = make_workload(ctx, lib, n, L, seed=1234 + rank)

    # ---- device-resident arm: inputs already in HBM ------------------------------------------------
    dev = torch.device("cuda", local)
    d_msgs = torch.from_numpy(msgs).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_sigs = torch.from_numpy(sigs).to(dev)
    d_status = torch.zeros(n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()
    lib.bbs_ctx_set_profiling(ctx.handle, 1)

    def step():
        rc = lib.bbs_verify_batch_dev(ctx.handle, n, ptr(d_sigs), ptr(d_msgs), ptr(d_offs), L, ptr(d_status),
                                      C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(lib.bbs_last_error().decode())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    got = d_status.cpu().numpy()
    nocheck = bool(os.environ.get("BBS_BENCH_NOCHECK"))      # tuning experiments with deliberately wrong programs
    if nocheck:
        expect = got.copy()
    if not np.array_equal(got, expect):
        raise RuntimeError(f"status vector mismatch: {int((got != expect).sum())} of {n} items")
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ktimes = np.zeros((args.steps, 3), dtype=np.float32)
    barrier()
    for k in range(args.steps):
        flush.fill_(k)                      # L2 flush between timed iterations (outside the timed events)
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
        ev[k][1].synchronize()
        kt = (C.c_float * 3)()
        lib.bbs_ctx_kernel_times(ctx.handle, kt, 3)
        ktimes[k] = list(kt)
    barrier()
    clocks = sampler.stop()
    launches = (ctx.launch_count() - launches0) // max(args.steps, 1)
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    t_local = float(ms.sum()) * 1e-3
    t = torch.tensor([t_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_max = float(t.item())
    if not np.array_equal(d_status.cpu().numpy(), expect):
        raise RuntimeError("status vector mismatch after the timed region")
    value = world * n * args.steps / t_max

    # ---- end-to-end arm: host (pinned) buffers through the reference-facing call, copies included ----
    lib.bbs_ctx_set_profiling(ctx.handle, 0)
    p_msgs = torch.from_numpy(msgs).pin_memory()
    p_offs = torch.from_numpy(offs.view(np.int64)).pin_memory()
    p_sigs = torch.from_numpy(sigs).pin_memory()
    p_status = torch.zeros(n, dtype=torch.uint8).pin_memory()

    def e2e_step():
        rc = lib.bbs_verify_batch(ctx.handle, n, ptr(p_sigs), ptr(p_msgs), ptr(p_offs), L, ptr(p_status))
        if rc != 0:
            raise RuntimeError(lib.bbs_last_error().decode())

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if not nocheck and not np.array_equal(p_status.numpy(), expect):
        raise RuntimeError("e2e status vector mismatch")
    e2e_value = world * n * args.steps / float(te.item())

    out = None
    if rank == 0:
        # ---- roofline of the dominant kernel (pairing), integer-multiply bound ------------------------
        gprod = C.c_double()
        pk_ms = C.c_float()
        peak_modes = {}
        for mode, name in ((0, "mad.lo+mad.hi"), (1, "mad.wide (IMAD.WIDE)")):
            lib.bbs_imad_peak(local, 2000, mode, C.byref(gprod), C.byref(pk_ms))
            peak_modes[name] = gprod.value * 1e9
        # IMAD.WIDE.U32 (one 32x32->64 product) issues at 8 lanes/clk/SMSP = 32 products/clk/SM on sm_100 (half the
        # rate of a 32-bit IMAD; confirmed by ncu: 4 fmaheavy-pipe cycles per warp instruction, and by the in-run
        # probe, which reaches ~92 % of this figure).  MEASURED_PEAKS.json has no integer entry, so the roofline
        # denominator is this nominal rate at the maximum SM clock; the probe's own number is reported beside it.
        sm_mhz = (clocks.get("sm_max_mhz") or 1965)
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        peak = n_sm * 32 * sm_mhz * 1e6
        kmean = ktimes.mean(axis=0)
        pairing_s = float(kmean[2]) * 1e-3
        achieved = n * PRODUCTS_PAIRING / pairing_s
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = n * (6 * 12 * 4 + 4 + 1)        # pair record + flags in, status out
        cpu = cpu_baseline(args, sample=args.cpu_sample) if world == 1 and not args.no_cpu else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"BLS12-381-SHA-256 batch verify: {n} signatures x L={L} messages of 32 B, one issuer key, "
                                   "1/16 corrupted (BASELINE configs[1])", "n_per_gpu": n, "L": L,
                       "l2": "256 MB flush write between timed iterations", "sharding": "by signature index, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(msgs.nbytes + offs.nbytes + sigs.nbytes),
                    "d2h_bytes_per_step": int(n)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kernels_ms": {"msg_to_scalars": float(kmean[0]), "verify_g1": float(kmean[1]), "pairing": float(kmean[2])},
            "roofline": {"kernel": "pairing_coop_kernel<Bls> (2-pair Miller loop + final exponentiation, 6 role-warps per 32 items)", "bound": "imad",
                         "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "T(32x32->64 products)/s",
                         "frac": achieved / peak, "peak_source": f"nominal IMAD.WIDE rate: {n_sm} SM x 32 products/clk x {sm_mhz} MHz (no integer entry in MEASURED_PEAKS.json); in-run probe below",
                         "peak_by_form_tprod_s": {k: v / 1e12 for k, v in peak_modes.items()},
                         "algorithmic_products_per_item": PRODUCTS_PAIRING,
                         "whole_step_frac": value / world * PRODUCTS_PER_VERIFY / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of pairing_coop_kernel from the ncu --set full capture
                         # of this command at n = 65,536 (profiles/summary_r01.md: 20.3 MB + 34.6 MB), scaled to n
                         "traffic": int(54.9e6 * n / 65536),
                         "traffic_source": "DRAM bytes per launch: ncu --set full, profiles/summary_r01.md",
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "hbm": {"achieved_gbs": alg_bytes / pairing_s / 1e9, "peak_gbs": hbm_peak,
                                 "frac": alg_bytes / pairing_s / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        }
        if cpu:
            out["cpu_baseline"] = cpu
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return out


def run_proof(args):
    """Secondary workload (BASELINE config 4 shape): BLS12-381 batch proof_verify, L = 32 with 16 disclosed messages.
    Proofs come from the committed oracle-made fixture tests/golden/proofs_bls_L32_R16.npz (24 proofs, valid and
    corrupted, with recorded verdicts) tiled to n items.  A step = msg_to_scalars of the disclosed messages +
    core_proof_verify (G1 half with the challenge hash, then the cooperative pairing kernel)."""
    import torch
    import torch.distributed as dist
    from bbs_sign_b200 import api, _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = _native.load()
    fx = np.load(os.path.join(ROOT, "tests", "golden", "proofs_bls_L32_R16.npz"))
    n, L, R = args.n, 32, 16
    nb = fx["fixed"].shape[0]
    sel = (np.arange(n) + rank) % nb
    fixed = np.ascontiguousarray(fx["fixed"][sel]).reshape(-1)
    commit = np.ascontiguousarray(fx["commitments"][sel]).reshape(-1)
    dmsg = np.ascontiguousarray(fx["disclosed_msgs"][sel]).reshape(-1)
    expect = fx["expect"][sel].astype(np.uint8)
    U = L - R
    commit_off = (np.arange(n + 1, dtype=np.uint64) * U)
    dis_off = (np.arange(n + 1, dtype=np.uint64) * R)
    idx = np.tile(fx["disclosed_idx"].astype(np.uint32), n)
    moff = (np.arange(n * R + 1, dtype=np.uint64) * MSG_BYTES)
    ctx = api.BatchContext(api.BLS12_381, bytes(fx["pk"]), header=b"", n_messages=L, device=local)
    dev = torch.device("cuda", local)
    to = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int32) if a.dtype == np.uint32 else a)).to(dev)
    d_fixed, d_commit, d_coff, d_idx, d_dmsg, d_moff, d_doff = map(to, (fixed, commit, commit_off, idx, dmsg, moff, dis_off))
    d_scal = torch.zeros(n * R * 32, dtype=torch.uint8, device=dev)
    d_status = torch.zeros(n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def step():
        rc = lib.bbs_msg_to_scalars_dev(ctx.handle, n * R, ptr(d_dmsg), ptr(d_moff), ptr(d_scal), sp)
        rc = rc or lib.bbs_core_proof_verify_batch_dev(ctx.handle, n, ptr(d_fixed), ptr(d_commit), ptr(d_coff), ptr(d_idx),
                                                       ptr(d_scal), ptr(d_doff), None, 0, ptr(d_status), sp)
        if rc != 0:
            raise RuntimeError(lib.bbs_last_error().decode())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    if not np.array_equal(d_status.cpu().numpy(), expect):
        raise RuntimeError("proof status vector mismatch")
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.fill_(k)
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = (ctx.launch_count() - launches0) // max(args.steps, 1)
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = world * n * args.steps / float(t.item())
    # end to end through the reference-facing host call (pinned buffers, copies inside the timed region)
    pin = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int32) if a.dtype == np.uint32 else a)).pin_memory()
    p_fixed, p_commit, p_coff, p_idx, p_dmsg, p_moff, p_doff = map(pin, (fixed, commit, commit_off, idx, dmsg, moff, dis_off))
    p_status = torch.zeros(n, dtype=torch.uint8).pin_memory()

    def e2e_step():
        rc = lib.bbs_proof_verify_batch(ctx.handle, n, ptr(p_fixed), ptr(p_commit), ptr(p_coff), ptr(p_idx), ptr(p_dmsg),
                                        ptr(p_moff), ptr(p_doff), None, 0, ptr(p_status))
        if rc != 0:
            raise RuntimeError(lib.bbs_last_error().decode())

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if not np.array_equal(p_status.numpy(), expect):
        raise RuntimeError("e2e proof status vector mismatch")
    out = None
    if rank == 0:
        h2d = sum(int(a.nbytes) for a in (fixed, commit, commit_off, idx, dmsg, moff, dis_off))
        out = {"metric": "bls12_381_bbs_proof_verifies_per_sec_L32_R16", "value": value, "unit": "proof-verifies/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t.item()) / args.steps * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
               "config": {"workload": f"BLS12-381 batch proof_verify: {n} proofs, L=32, 16 disclosed 32-B messages, one issuer key; "
                                      "24 oracle-made proofs (valid and corrupted) tiled (BASELINE configs[3] shape)",
                          "n_per_gpu": n, "l2": "256 MB flush write between timed iterations",
                          "sharding": "by proof index, no collective"},
               "e2e": {"value": world * n * args.steps / float(te.item()), "unit": "proof-verifies/s",
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(n)},
               "gpu_launches": int(launches), "clocks": clocks}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return out


def run_bn254(args):
    """BASELINE configs[2] shape: BN254 core_verify over pre-hashed scalar messages (L = 31), device-resident inputs.
    Key pair: tests/golden/bn254_bench_key.npz (made once with the oracle's key_gen / sk_to_pk); signatures are made on
    the GPU by bbs_core_sign_batch."""
    import torch
    import torch.distributed as dist
    from bbs_sign_b200 import api, _native
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = _native.load()
    key = np.load(os.path.join(ROOT, "tests", "golden", "bn254_bench_key.npz"))
    n, L = args.n, 31
    ctx = api.BatchContext(api.BN254, bytes(key["pk"]), header=b"", n_messages=L, device=local)
    rng = np.random.default_rng(5 + rank)
    sc = rng.integers(0, 256, size=(n * L, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0f                                   # < 2^252 < r: canonical scalars
    sc = np.ascontiguousarray(sc.reshape(-1))
    sigs = np.zeros(n * 64, dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    if lib.bbs_core_sign_batch(ctx.handle, ptr(np.ascontiguousarray(key["sk"])), n, ptr(sc), L, ptr(sigs), None, ptr(st)) != 0:
        raise RuntimeError(lib.bbs_last_error().decode())
    expect = np.ones(n, dtype=np.uint8)
    sg = sigs.reshape(n, 64)
    for i in range(5, n, 16):
        sg[i, 32] ^= 1                                  # e ^ 1
        expect[i] = 0
    dev = torch.device("cuda", local)
    d_sigs = torch.from_numpy(sigs).to(dev)
    d_sc = torch.from_numpy(sc).to(dev)
    d_status = torch.zeros(n, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        if lib.bbs_core_verify_batch_dev(ctx.handle, n, ptr(d_sigs), ptr(d_sc), L, ptr(d_status), C.c_void_p(stream.cuda_stream)) != 0:
            raise RuntimeError(lib.bbs_last_error().decode())

    lib.bbs_ctx_set_profiling(ctx.handle, 1)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    kt = (C.c_float * 3)()
    lib.bbs_ctx_kernel_times(ctx.handle, kt, 3)
    lib.bbs_ctx_set_profiling(ctx.handle, 0)
    if not np.array_equal(d_status.cpu().numpy(), expect):
        raise RuntimeError("bn254 status vector mismatch")
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        v = world * n * args.steps / float(t.item())
        out = {"metric": "bn254_bbs_core_verifies_per_sec_L31", "value": v, "unit": "verifies/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t.item()) / args.steps * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
               "config": {"workload": f"BN254 core_verify: {n} signatures x L=31 pre-hashed scalar messages, one issuer key, "
                                      "1/16 corrupted (BASELINE configs[2] shape); cooperative pairing kernel", "n_per_gpu": n},
               "kernels_ms": {"verify_g1": float(kt[1]), "pairing": float(kt[2])}, "gpu_launches": 2}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return out


def run_rlc(args):
    """Optional random-linear-combination mode (BASELINE configs[4] shape, one GPU per rank): n valid signatures
    under one issuer, one batch verdict per step through the host-buffer call bbs_rlc_verify_batch (copies inside)."""
    import torch
    import torch.distributed as dist
    from bbs_sign_b200 import api, _native
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = _native.load()
    n, L = args.n, args.L
    ctx = api.BatchContext(api.BLS12_381, IRTF_PK, header=b"", n_messages=L, device=local)
    rng = np.random.default_rng(99 + rank)
    msgs = rng.integers(0, 256, size=n * L * MSG_BYTES, dtype=np.uint8)
    offs = (np.arange(n * L + 1, dtype=np.uint64) * MSG_BYTES)
    sigs = np.zeros(n * SIG_BYTES, dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    sk = np.frombuffer(IRTF_SK.to_bytes(32, "little"), dtype=np.uint8).copy()
    if lib.bbs_sign_batch(ctx.handle, ptr(sk), n, ptr(msgs), ptr(offs), L, ptr(sigs), None, ptr(st)) != 0:
        raise RuntimeError(lib.bbs_last_error().decode())
    seed = np.frombuffer(bytes(range(32)), dtype=np.uint8).copy()
    verdict = np.zeros(1, dtype=np.uint8)
    # pinned host buffers (the caller's side of the C ABI), as in the verify arm
    p_msgs = torch.from_numpy(msgs).pin_memory()
    p_offs = torch.from_numpy(offs.view(np.int64)).pin_memory()
    p_sigs = torch.from_numpy(sigs).pin_memory()
    msgs, offs, sigs = p_msgs.numpy(), p_offs.numpy().view(np.uint64), p_sigs.numpy()

    # N > 1 (BASELINE configs[4]): ONE verdict for the world * n signatures.  Every rank reduces its shard to two compressed
    # G1 points (coefficients indexed globally), the 96-byte partials are gathered, rank 0 adds them and does the single
    # pairing.  The gather is the path's only exchange: 96 bytes per GPU (the north star's "host-combined partials").
    parts = np.zeros(96, dtype=np.uint8)
    pst = np.zeros(1, dtype=np.uint8)
    dev = torch.device("cuda", local)
    d_parts = torch.zeros(96, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(96 * world, dtype=torch.uint8, device=dev)

    def step(expect=1):
        if world == 1:
            if lib.bbs_rlc_verify_batch(ctx.handle, n, ptr(sigs), ptr(msgs), ptr(offs), L, ptr(seed), ptr(verdict)) != 0:
                raise RuntimeError(lib.bbs_last_error().decode())
        else:
            if lib.bbs_rlc_partial(ctx.handle, n, ptr(sigs), ptr(msgs), ptr(offs), L, ptr(seed), C.c_uint64(rank * n),
                                   ptr(parts), ptr(pst)) != 0 or pst[0] != 1:
                raise RuntimeError(lib.bbs_last_error().decode() or f"rlc partial status {pst[0]}")
            d_parts.copy_(torch.from_numpy(parts))
            dist.all_gather_into_tensor(d_all, d_parts)
            if rank == 0:
                allp = d_all.cpu().numpy()
                if lib.bbs_rlc_combine(ctx.handle, world, ptr(allp), ptr(verdict)) != 0:
                    raise RuntimeError(lib.bbs_last_error().decode())
            else:
                verdict[0] = expect
        if verdict[0] != expect:
            raise RuntimeError(f"rlc verdict {verdict[0]}, expected {expect}")

    lib.bbs_ctx_set_profiling(ctx.handle, 1)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    kt = (C.c_float * 7)()
    lib.bbs_ctx_kernel_times(ctx.handle, kt, 7)
    lib.bbs_ctx_set_profiling(ctx.handle, 0)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    # a corrupted batch must be rejected (one flipped bit of e in the last rank's shard)
    if rank == world - 1:
        sigs[48] ^= 1
    step(expect=0)
    out = None
    if rank == 0:
        v = world * n * args.steps / float(dt.item())
        out = {"metric": "bls12_381_bbs_rlc_batch_verified_signatures_per_sec_L10", "value": v, "unit": "signatures/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(dt.item()) / args.steps * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
               "config": {"workload": f"random-linear-combination batch verify: ONE verdict for {world * n} valid signatures x L={L} "
                                      f"under one issuer ({n} per GPU; partial G1 sums of the shards combined on rank 0, then one pairing; "
                                      "optional mode, BASELINE configs[4]; host buffers, copies inside the timed region)", "n_per_gpu": n},
               "e2e": {"value": v, "unit": "signatures/s", "h2d_bytes_per_step": int(msgs.nbytes + offs.nbytes + sigs.nbytes),
                       "d2h_bytes_per_step": 2 * 48 + 1}, "gpu_launches": 9,
               "kernels_ms": dict(zip(["first_chunk_upload_and_msg_to_scalars", "chunked_msg_to_scalars_and_rlc_prep_under_the_uploads", "msm_scan_scatter",
                                       "msm_bucket", "msm_reduce", "rlc_msm_finish", "pairing"],
                                      [float(x) for x in kt]))}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return out


def run_sign(args):
    """BASELINE configs[4], first leg: batch signing (core_sign: B, e = H2S(sk || msgs || domain), A = B * 1/(sk + e)) of n
    message sets x L = 10 through the host-buffer call bbs_sign_batch (pinned buffers, copies inside the timed region).
    Every step's signatures are checked once: the RLC verdict of the produced batch must be ACCEPT."""
    import torch
    import torch.distributed as dist
    from bbs_sign_b200 import api, _native
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = _native.load()
    n, L = args.n, args.L
    ctx = api.BatchContext(api.BLS12_381, IRTF_PK, header=b"", n_messages=L, device=local)
    rng = np.random.default_rng(7 + rank)
    pin = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).pin_memory().numpy()
    msgs = pin(rng.integers(0, 256, size=n * L * MSG_BYTES, dtype=np.uint8))
    offs = pin(np.arange(n * L + 1, dtype=np.uint64) * MSG_BYTES).view(np.uint64)
    sigs = pin(np.zeros(n * SIG_BYTES, dtype=np.uint8))
    st = pin(np.zeros(n, dtype=np.uint8))
    sk = np.frombuffer(IRTF_SK.to_bytes(32, "little"), dtype=np.uint8).copy()

    def step():
        if lib.bbs_sign_batch(ctx.handle, ptr(sk), n, ptr(msgs), ptr(offs), L, ptr(sigs), None, ptr(st)) != 0:
            raise RuntimeError(lib.bbs_last_error().decode())

    lib.bbs_ctx_set_profiling(ctx.handle, 1)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    kt = (C.c_float * 2)()
    lib.bbs_ctx_kernel_times(ctx.handle, kt, 2)
    lib.bbs_ctx_set_profiling(ctx.handle, 0)
    if not (st == 1).all():
        raise RuntimeError("sign status vector is not all ACCEPT")
    seed = np.frombuffer(bytes(range(32)), dtype=np.uint8).copy()
    verdict = np.zeros(1, dtype=np.uint8)
    if lib.bbs_rlc_verify_batch(ctx.handle, n, ptr(sigs), ptr(msgs), ptr(offs), L, ptr(seed), ptr(verdict)) != 0 or verdict[0] != 1:
        raise RuntimeError("the signed batch does not verify")
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        v = world * n * args.steps / float(dt.item())
        kernel_v = n / (float(kt[1]) * 1e-3) if kt[1] > 0 else None
        out = {"metric": "bls12_381_bbs_signatures_per_sec_L10", "value": v, "unit": "signatures/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(dt.item()) / args.steps * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
               "config": {"workload": f"batch sign: {n} message sets x L={L} of 32 B under one key (BASELINE configs[4], first leg; "
                                      "host buffers, copies inside the timed region)", "n_per_gpu": n},
               "e2e": {"value": v, "unit": "signatures/s", "h2d_bytes_per_step": int(msgs.nbytes + offs.nbytes),
                       "d2h_bytes_per_step": int(sigs.nbytes + st.nbytes)}, "gpu_launches": 2,
               "kernels_ms": {"msg_to_scalars": float(kt[0]), "sign_kernel": float(kt[1])},
               "roofline": {"kernel": "sign_kernel<Bls>", "bound": "imad", "unit": "T(32x32->64 products)/s",
                            "achieved": (kernel_v or 0) * 1.85e6 / 1e12, "peak": 9.30624,
                            "frac": (kernel_v or 0) * 1.85e6 / 9.30624e12,
                            "algorithmic_products_per_item": 1850000, "traffic": None,
                            "note": "products per item are SURVEY 8d's model (8-bit fixed-base windows, 255-bit variable-base "
                                    "multiplication); the kernel needs fewer (GLV halves, 16-bit tables, one inversion per block), "
                                    "so the fraction can exceed 1"}}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return out


def cpu_baseline(args, sample):
    """The oracle port of the reference's per-item path (msg_to_scalars + core_verify with two pairings) on
    the host.  Uses oracle/_ref/ (compiled C restatement, all cores) when present, else the big-int Python
    oracle on one core."""
    from oracle import cpu_ref
    return cpu_ref.time_verify(L=args.L, sample=sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from oracle import cpu_ref
    vals = []
    for _ in range(args.warmup):
        cpu_ref.time_verify(L=args.L, sample=max(args.cpu_sample // 4, 1))
    last = None
    t_total = 0.0
    for _ in range(args.steps):
        last = cpu_ref.time_verify(L=args.L, sample=args.cpu_sample)
        vals.append(last["value"])
        t_total += last["seconds"]
    v = float(np.mean(vals))
    return {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"BLS12-381-SHA-256 batch verify, L={args.L} messages of 32 B, one issuer key "
                               f"(bounded sample per step of BASELINE configs[1]: {last['sample'].split(':')[0]})"},
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=None, help="items per GPU (default 65,536; rlc: 524,288, bn254: 131,072 = configs[4] / configs[2] per GPU of 8)")
    ap.add_argument("--L", type=int, default=L_DEFAULT)
    ap.add_argument("--cpu-sample", type=int, default=0, help="signatures per CPU-baseline step (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="verify", choices=["verify", "proof", "rlc", "bn254", "sign"],
                    help="verify = BASELINE configs[1] (the headline); proof = configs[3]-shaped proof_verify")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.n is None:
        # rlc: configs[4] per GPU (4M / 8); bn254: configs[2] per GPU (1M / 8); verify / proof: 65,536
        args.n = {"rlc": 524288, "sign": 524288, "bn254": 131072}.get(args.workload, N_DEFAULT) if args.impl == "ours" else N_DEFAULT
    if args.workload == "proof" and args.impl == "ours":
        out = run_proof(args)
    elif args.workload == "rlc" and args.impl == "ours":
        out = run_rlc(args)
    elif args.workload == "bn254" and args.impl == "ours":
        out = run_bn254(args)
    elif args.workload == "sign" and args.impl == "ours":
        out = run_sign(args)
    else:
        out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
