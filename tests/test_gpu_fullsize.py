"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle needs ~1 s per pairing, so it
cannot label these batches): items are produced by the GPU signer / the oracle-made proof fixture (both pinned to the
oracle at small sizes by test_gpu_parity.py), a known subset is corrupted, and the status vector must accept exactly the
untouched items.  A random sample of the big batch is then re-derived by the oracle without pairings:
verify accepts  <=>  A (sk + e) == B  (pairing-free ground truth)."""
import argparse
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from bbs_sign_b200 import _native
    return _native.load()


def test_config2_bls_verify_65536(lib):
    """configs[1]: 65,536 signatures x L = 10, one issuer, 1/16 corrupted; oracle spot check of 6 items"""
    import bench
    from bbs_sign_b200 import api as A
    from oracle import bbs_oracle as O
    n, L = 65536, 10
    ctx = A.BatchContext(A.BLS12_381, bench.IRTF_PK, header=b"", n_messages=L)
    msgs, offs, sigs, expect = bench.make_workload(ctx, lib, n, L, seed=11)
    st = np.zeros(n, dtype=np.uint8)
    assert lib.bbs_verify_batch(ctx.handle, n, bench.ptr(sigs), bench.ptr(msgs), bench.ptr(offs), L, bench.ptr(st)) == 0
    assert np.array_equal(st, expect)
    assert int(expect.sum()) == n - len(range(3, n, 16))
    # pairing-free oracle truth on a sample (valid and corrupted items)
    cs = O.BLS12_381
    gens = O.create_generators_cached(cs, L + 1, cs.api_id)
    pk = O.sk_to_pk(cs, bench.IRTF_SK)
    dom = O.calculate_domain(cs, pk, gens[0], gens[1:], b"", cs.api_id)
    rnd = random.Random(3)
    sample = [3, 19, 35, 51] + [rnd.randrange(n) for _ in range(4)]      # the four corruption kinds + random items
    sg = sigs.reshape(n, bench.SIG_BYTES)
    for i in sample:
        m = [bytes(msgs[(i * L + j) * bench.MSG_BYTES:(i * L + j + 1) * bench.MSG_BYTES]) for j in range(L)]
        B = O.compute_B(cs, gens, dom, O.msg_to_scalars(cs, m, cs.api_id))
        a_bytes = bytes(sg[i, :48])
        e = int.from_bytes(bytes(sg[i, 48:80]), "little")
        Apt = cs.g1_decompress(a_bytes)
        truth = Apt is not None and O.ec_mul(cs.F1, Apt, (bench.IRTF_SK + e) % cs.r) == B
        assert bool(st[i]) == truth, i
    ctx.close()


def test_config3_bn254_core_verify_1m(lib):
    """configs[2]: 1,048,576 signatures x L = 31 pre-hashed scalars on BN254, 1/16 corrupted over all four signature
    rejection classes; the status vector must equal the construction, and 12 items (every class + random ones) are
    re-derived by the oracle without pairings: accept <=> A (sk + e) == B."""
    import bench
    from bbs_sign_b200 import api as A
    from oracle import bbs_oracle as O
    n, L = 1 << 20, 31
    key = np.load(os.path.join(ROOT, "tests", "golden", "bn254_bench_key.npz"))
    ctx = A.BatchContext(A.BN254, bytes(key["pk"]), header=b"", n_messages=L)
    sc, sigs, expect, _ = bench.make_bn254_workload(ctx, lib, n, seed=21, L=L)
    st = np.zeros(n, dtype=np.uint8)
    assert lib.bbs_core_verify_batch(ctx.handle, n, bench.ptr(sigs), bench.ptr(sc), L, bench.ptr(st)) == 0
    assert np.array_equal(st, expect)
    assert int(expect.sum()) == n - len(range(5, n, 16))
    cs = O.BN254
    sk = int.from_bytes(bytes(key["sk"]), "little")
    pk = cs.g2_decompress(bytes(key["pk"]))
    assert O.sk_to_pk(cs, sk) == pk
    gens = O.create_generators_cached(cs, L + 1, cs.api_id)
    dom = O.calculate_domain(cs, pk, gens[0], gens[1:], b"", cs.api_id)
    assert ctx.domain() == cs.scalar_le(dom)
    rnd = random.Random(4)
    sample = [5, 21, 37, 53, 69, 85, 101, 117] + [rnd.randrange(n) for _ in range(4)]      # 2 x the four kinds + random items
    sg = sigs.reshape(n, 64)
    scm = sc.reshape(n, L, 32)
    for i in sample:
        m = [int.from_bytes(bytes(scm[i, j]), "little") for j in range(L)]
        B = O.compute_B(cs, gens, dom, m)
        Apt = cs.g1_decompress(bytes(sg[i, :32]))
        e = int.from_bytes(bytes(sg[i, 32:64]), "little")
        truth = Apt is not None and O.ec_mul(cs.F1, Apt, (sk + e) % cs.r) == B
        assert bool(st[i]) == truth, i
    ctx.close()


def test_config4_bls_proof_verify_262144(lib):
    """configs[3]: 262,144 DISTINCT proofs (L = 32, 16 disclosed) made by bbs_sign_batch + bbs_proof_gen_batch, 1/16
    corrupted over the seven rejection classes; statuses must equal the construction.  A sample is checked against the
    oracle: the proof bytes equal the oracle's proof_gen for the same signature and random scalars, the untouched ones
    satisfy the pairing-free trapdoor Bbar == sk * Abar, and the oracle's proof_verify (trapdoor mode) agrees."""
    import bench
    from bbs_sign_b200 import api as A
    from oracle import bbs_oracle as O
    n, L, R = 262144, 32, 16
    U = L - R
    ctx = A.BatchContext(A.BLS12_381, bench.IRTF_PK, header=b"", n_messages=L)
    w = bench.make_proof_workload(ctx, lib, n, seed=31, L=L, R=R)
    st = np.zeros(n, dtype=np.uint8)
    rc = lib.bbs_proof_verify_batch(ctx.handle, n, bench.ptr(w["fixed"]), bench.ptr(w["commit"]), bench.ptr(w["commit_off"]),
                                    bench.ptr(w["idx"]), bench.ptr(w["dmsg"]), bench.ptr(w["moff"]), bench.ptr(w["dis_off"]),
                                    None, 0, bench.ptr(st))
    assert rc == 0
    assert np.array_equal(st, w["expect"])
    assert int(w["expect"].sum()) == n - len(range(5, n, 16))
    cs = O.BLS12_381
    sk = bench.IRTF_SK
    pk = O.sk_to_pk(cs, sk)
    dis = list(range(0, L, 2))
    fx = w["fixed"].reshape(n, bench.PROOF_FIXED)
    cm = w["commit"].reshape(n, U * 32)
    msgs = w["msgs"].reshape(n, L, 32)
    dm = w["dmsg"].reshape(n, R, 32)
    rand = w["rand"].reshape(n, 5 + U, 32)
    sg = w["sigs"].reshape(n, 80)
    rnd = random.Random(9)
    corrupted = [5 + 16 * k for k in range(7)]
    for i in corrupted + [rnd.randrange(n) for _ in range(3)]:
        pr = O.Proof(cs.g1_decompress(bytes(fx[i, 0:48])), cs.g1_decompress(bytes(fx[i, 48:96])), cs.g1_decompress(bytes(fx[i, 96:144])),
                     int.from_bytes(bytes(fx[i, 144:176]), "little"), int.from_bytes(bytes(fx[i, 176:208]), "little"),
                     int.from_bytes(bytes(fx[i, 208:240]), "little"),
                     [int.from_bytes(bytes(cm[i, 32 * k:32 * k + 32]), "little") for k in range(U)],
                     int.from_bytes(bytes(fx[i, 240:272]), "little"))
        truth = O.proof_verify(cs, pk, pr, b"", b"", [bytes(dm[i, k]) for k in range(R)], dis, trapdoor_sk=sk)
        assert bool(st[i] == 1) == bool(truth), i
        if w["expect"][i]:
            assert pr.b_bar == O.ec_mul(cs.F1, pr.a_bar, sk), i                           # trapdoor: Bbar == sk * Abar
            sig = (cs.g1_decompress(bytes(sg[i, :48])), int.from_bytes(bytes(sg[i, 48:80]), "little"))
            rs = [int.from_bytes(bytes(rand[i, k]), "little") for k in range(5 + U)]
            want = O.proof_gen(cs, pk, sig, b"", b"", [bytes(msgs[i, j]) for j in range(L)], dis, random_scalars=rs)
            wb = A.ProofBytes.from_canonical(A.BLS12_381, O.proof_to_bytes(cs, want))
            assert wb.fixed == bytes(fx[i]) and wb.commitments == bytes(cm[i]), i
    ctx.close()


def test_config5_bls_sign_and_rlc_524288(lib):
    """configs[4] per-GPU share (4M / 8): sign 524,288 message sets (device-resident and host-buffer paths must agree byte
    for byte and the batch must pass the random-linear-combination check), then one RLC verdict per step; the valid batch
    must be accepted and the same batch with one flipped bit rejected (bench_sign / bench_rlc raise otherwise)"""
    import bench
    env = bench.Env()
    try:
        out = bench.bench_sign(env, 524288, 1, 3)
        assert out["value"] > 0 and out["checked"]
        out = bench.bench_rlc(env, 524288, 1, 3)
        assert out["value"] > 0 and out["checked"]
    finally:
        env.close()


def test_chunked_host_signing_equals_one_shot(lib):
    """bbs_sign_batch / bbs_core_sign_batch with >= 524,288 items go through the chunked, overlapped host-buffer path
    (capi.cu sign_chunked): signatures, B points and statuses must equal those of one-shot calls on the two halves,
    with a ragged tail, ragged message lengths and a few non-canonical scalars in the scalar-level call"""
    import bench
    from bbs_sign_b200 import api as A
    L = 2
    n = 2 * 262144 + 777
    ctx = A.BatchContext(A.BLS12_381, bench.IRTF_PK, header=b"chunk", n_messages=L)
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 9, size=n * L)
    offs = np.zeros(n * L + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens, dtype=np.uint64)
    msgs = rng.integers(0, 256, size=int(offs[-1]) + 1, dtype=np.uint8)
    sk = np.frombuffer(bench.IRTF_SK.to_bytes(32, "little"), dtype=np.uint8).copy()

    def run(i0, i1):
        cnt = i1 - i0
        o = (offs[i0 * L: i1 * L + 1] - offs[i0 * L]).astype(np.uint64)
        m = msgs[int(offs[i0 * L]): int(offs[i1 * L]) + 1].copy()
        sig = np.zeros(cnt * bench.SIG_BYTES, dtype=np.uint8)
        b = np.zeros(cnt * 48, dtype=np.uint8)
        st = np.full(cnt, 255, dtype=np.uint8)
        assert lib.bbs_sign_batch(ctx.handle, bench.ptr(sk), cnt, bench.ptr(m), bench.ptr(o), L, bench.ptr(sig), bench.ptr(b), bench.ptr(st)) == 0
        return sig, b, st

    sig, b, st = run(0, n)                      # chunked
    h = n // 2 - 5
    s0, b0, t0 = run(0, h)                      # one shot (below the threshold)
    s1, b1, t1 = run(h, n)
    assert np.array_equal(sig, np.concatenate([s0, s1])) and np.array_equal(b, np.concatenate([b0, b1]))
    assert np.array_equal(st, np.concatenate([t0, t1])) and st.min() == 1
    # scalar-level entry point, some scalars >= r
    sc = rng.integers(0, 256, size=(n * L, 32), dtype=np.uint8)
    sc[:, 31] &= 0x3f
    bad = rng.choice(n, size=50, replace=False)
    sc[bad * L + 1, :] = 0xff
    sc = sc.reshape(-1)
    sig2 = np.zeros(n * bench.SIG_BYTES, dtype=np.uint8)
    st2 = np.full(n, 255, dtype=np.uint8)
    assert lib.bbs_core_sign_batch(ctx.handle, bench.ptr(sk), n, bench.ptr(sc), L, bench.ptr(sig2), None, bench.ptr(st2)) == 0
    sig3 = np.zeros(h * bench.SIG_BYTES, dtype=np.uint8)
    st3 = np.full(h, 255, dtype=np.uint8)
    assert lib.bbs_core_sign_batch(ctx.handle, bench.ptr(sk), h, bench.ptr(sc), L, bench.ptr(sig3), None, bench.ptr(st3)) == 0
    assert np.array_equal(sig2[: h * bench.SIG_BYTES], sig3) and np.array_equal(st2[:h], st3)
    assert sorted(np.nonzero(st2 != 1)[0].tolist()) == sorted(bad.tolist()) and set(st2[bad].tolist()) == {A.ST_ERR_MALFORMED}
    ctx.close()


def test_chunked_host_verify_equals_device_path(lib):
    """bbs_verify_batch / bbs_core_verify_batch with >= 262,144 items upload in chunks under the G1 half (capi.cu
    verify_chunked) and run ONE pairing launch: the status vector must equal the construction and the device-buffer entry
    point's (never chunked), with a ragged number of items"""
    import ctypes as C
    import torch
    import bench
    from bbs_sign_b200 import api as A
    L = 3
    n = 262144 + 131072 + 555
    ctx = A.BatchContext(A.BLS12_381, bench.IRTF_PK, header=b"", n_messages=L)
    msgs, offs, sigs, expect = bench.make_workload(ctx, lib, n, L, seed=17)
    st = np.full(n, 255, dtype=np.uint8)
    assert lib.bbs_verify_batch(ctx.handle, n, bench.ptr(sigs), bench.ptr(msgs), bench.ptr(offs), L, bench.ptr(st)) == 0
    assert np.array_equal(st, expect) and 0 < int((expect == 0).sum()) < n
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(a).to(dev) for a in (sigs, msgs, offs.view(np.int64))]
    d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.bbs_verify_batch_dev(ctx.handle, n, bench.ptr(d[0]), bench.ptr(d[1]), bench.ptr(d[2]), L, bench.ptr(d_st), sp) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_st.cpu().numpy(), st)
    # scalar-level entry point on the same batch
    sc = np.zeros(n * L * 32, dtype=np.uint8)
    assert lib.bbs_msg_to_scalars(ctx.handle, n * L, bench.ptr(msgs), bench.ptr(offs), bench.ptr(sc)) == 0
    st2 = np.full(n, 255, dtype=np.uint8)
    assert lib.bbs_core_verify_batch(ctx.handle, n, bench.ptr(sigs), bench.ptr(sc), L, bench.ptr(st2)) == 0
    assert np.array_equal(st2, expect)
    ctx.close()
