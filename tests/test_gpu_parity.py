"""Parity tests proper: the CUDA library (sm_100a) through the C ABI against the oracle, bit-exact.
Run with `pytest -m gpu` on a B200."""
import pytest

import parity_cases as P
from bbs_sign_b200 import _native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return _native.load()   # raises if libbbs_b200.so is missing: no fallback


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_field(lib, curve):
    P.case_field(lib, curve, n=4096)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_g1_mul(lib, curve):
    P.case_g1_mul(lib, curve, n=64)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_pairing(lib, curve):
    P.case_pairing(lib, curve)


def test_irtf_kat(lib):
    P.case_irtf_kat(None)


@pytest.mark.parametrize("curve,L", [("BLS12_381", 10), ("BN254", 5), ("BLS12_381", 0), ("BN254", 0), ("BLS12_381", 1)])
def test_verify(lib, curve, L):
    P.case_verify(None, curve, L, n=12, use_pairing_oracle_on=2)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_verify_malformed(lib, curve):
    P.case_verify_malformed(None, curve)


@pytest.mark.parametrize("curve,L,dis", [("BLS12_381", 4, [0, 2]), ("BN254", 3, [1]), ("BLS12_381", 2, [0, 1]),
                                         ("BN254", 2, []), ("BLS12_381", 8, [1, 3, 5, 7]), ("BLS12_381", 0, [])])
def test_proof_verify(lib, curve, L, dis):
    P.case_proof_verify(None, curve, L, dis, n=8, pairing_on=1)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_proof_errors(lib, curve):
    P.case_proof_errors(None, curve)


def test_proof_fixture_l32(lib):
    """tests/golden/proofs_bls_L32_R16.npz (tools/gen_proof_fixture.py): config-4-shaped proofs, L = 32, 16 disclosed,
    valid and corrupted, against the verdicts the oracle recorded when the fixture was made."""
    import os
    import numpy as np
    from bbs_sign_b200 import api as A
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proofs_bls_L32_R16.npz"))
    ctx = A.BatchContext(A.BLS12_381, bytes(fx["pk"]), header=b"", n_messages=32)
    n = fx["fixed"].shape[0]
    proofs = [A.ProofBytes(bytes(fx["fixed"][i]), bytes(fx["commitments"][i])) for i in range(n)]
    dis = [int(x) for x in fx["disclosed_idx"]]
    msgs = [[bytes(fx["disclosed_msgs"][i][32 * k: 32 * k + 32]) for k in range(len(dis))] for i in range(n)]
    got = ctx.proof_verify_batch(proofs, b"", msgs, [dis] * n)
    assert got.tolist() == fx["expect"].tolist()
    ctx.close()


def test_large_batch_consistency(lib):
    """Size-independent property at a bench-like size: sign 4,096 synthetic message sets on the GPU, corrupt every
    16th item in one of four ways, and require accept exactly for the untouched items (the same check bench.py
    applies to its 65,536-signature batch on every run)."""
    import numpy as np
    import ctypes as C
    import bench
    from bbs_sign_b200 import api as A
    ctx = A.BatchContext(A.BLS12_381, bench.IRTF_PK, header=b"", n_messages=10)
    msgs, offs, sigs, expect = bench.make_workload(ctx, lib, 4096, 10, seed=7)
    st = np.zeros(4096, dtype=np.uint8)
    rc = lib.bbs_verify_batch(ctx.handle, 4096, bench.ptr(sigs), bench.ptr(msgs), bench.ptr(offs), 10, bench.ptr(st))
    assert rc == 0
    assert np.array_equal(st, expect)
    ctx.close()


@pytest.mark.parametrize("curve,L", [("BLS12_381", 3), ("BN254", 2), ("BLS12_381", 0)])
def test_rlc(lib, curve, L):
    P.case_rlc(None, curve, L=L, n=9)


@pytest.mark.parametrize("curve,windows,n", [("BLS12_381", 8, 9), ("BN254", 8, 9), ("BLS12_381", 13, 40), ("BN254", 19, 40),
                                             ("BLS12_381", 32, 9), ("BN254", 32, 9), ("BLS12_381", 0, 150)])
def test_rlc_msm_geometries(lib, curve, windows, n):
    """bucket MSM (rlc_msm.cuh) with other digit counts than the cost model's pick (0 = the model), incl. uneven digit
    widths and the 16-bit rows: partial sums stay bit-exact against the oracle"""
    P.case_rlc(None, curve, L=2, n=n, windows=windows)


@pytest.mark.parametrize("curve,L,dis", [("BLS12_381", 5, [0, 2, 3]), ("BN254", 3, [1]), ("BLS12_381", 2, []),
                                         ("BLS12_381", 3, [0, 1, 2]), ("BLS12_381", 0, [])])
def test_proof_gen(lib, curve, L, dis):
    P.case_proof_gen(None, curve, L, dis, n=5)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_h2s_ragged(lib, curve):
    P.case_h2s_ragged(None, curve)


@pytest.mark.parametrize("curve", ["BN254", "BLS12_381"])
def test_readme_example(lib, curve):
    P.case_readme_example(None, curve)


def test_sharded_helpers(lib):
    """bbs_sign_b200.sharding on whatever GPUs the box has (one context per device; a single GPU is used twice):
    verify, proof verify and the random-linear-combination verdict agree with the unsharded calls."""
    import torch
    from bbs_sign_b200 import api as A
    from bbs_sign_b200.sharding import ShardedVerifier
    from oracle import bbs_oracle as O
    suite, ocs = P.SUITES["BLS12_381"]
    sk, pk = P.keypair(ocs, 1)
    L, n = 2, 7
    msgs = [[P.rng_bytes(f"sh{i}.{j}", 32) for j in range(L)] for i in range(n)]
    sigs = [O.sign(ocs, sk, m, b"") for m in msgs]
    enc = [O.signature_to_bytes(ocs, s) for s in sigs]
    devs = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0]
    sv = ShardedVerifier(suite, ocs.g2_compress(pk), b"", L, devs)
    assert sv.verify_batch(b"".join(enc), msgs).tolist() == [1] * n
    seed = P.rng_bytes("sh-seed", 32)
    assert sv.rlc_verify_batch(enc, msgs, seed) == A.ST_ACCEPT
    bad = list(enc)
    bad[5] = enc[4]
    assert sv.rlc_verify_batch(bad, msgs, seed) == A.ST_REJECT
    assert sv.verify_batch(b"".join(bad), msgs).tolist() == [1, 1, 1, 1, 1, 0, 1]
    sv.close()


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_create_generators(lib, curve):
    P.case_create_generators(None, curve, count=40)


def test_multi_issuer_batch(lib):
    """per-item public keys through the MultiIssuerVerifier helper (api.IssuerSet underneath): three issuers
    interleaved in one batch, one wrong-issuer item, one forged signature; every status equals the oracle's verify under
    that item's key"""
    from bbs_sign_b200.sharding import MultiIssuerVerifier
    from oracle import bbs_oracle as O
    suite, ocs = P.SUITES["BLS12_381"]
    keys = [P.keypair(ocs, s) for s in (1, 2, 3)]
    L, n = 2, 9
    msgs = [[P.rng_bytes(f"mi{i}.{j}", 32) for j in range(L)] for i in range(n)]
    owner = [i % 3 for i in range(n)]
    sigs = [O.sign(ocs, keys[owner[i]][0], msgs[i], b"mi") for i in range(n)]
    claimed = list(owner)
    claimed[4] = (owner[4] + 1) % 3                       # verified under another issuer's key
    sigs[7] = (sigs[7][0], (sigs[7][1] + 1) % ocs.r)      # forged
    pks = [ocs.g2_compress(keys[k][1]) for k in claimed]
    enc = [O.signature_to_bytes(ocs, s) for s in sigs]
    mv = MultiIssuerVerifier(suite, b"mi", L)
    got = mv.verify_batch(pks, enc, msgs)
    want = [1 if O.verify(ocs, keys[claimed[i]][1], sigs[i], b"mi", msgs[i], trapdoor_sk=keys[claimed[i]][0]) else 0
            for i in range(n)]
    assert O.verify(ocs, keys[claimed[4]][1], sigs[4], b"mi", msgs[4]) is False        # and once with the pairing oracle
    assert got.tolist() == want == [1, 1, 1, 1, 0, 1, 1, 0, 1]
    assert len(mv.keys) == 3
    mv.close()


def test_multi_issuer_1024_issuers(lib):
    """1,024 issuers x 4 items in one mixed batch: keys sk_k * BP2 and signatures from the GPU signer under
    per-issuer contexts would need 1,024 contexts, so items are signed by the ORACLE-free trapdoor construction instead:
    A = B * (sk_k + e)^-1 needs B, which the single-issuer GPU signer provides (b_out) under a context for that key's
    domain.  Cheaper and independent: sign on the GPU under 8 real contexts, register those 8 keys 128 times each at
    different set positions (a key is what it is wherever it sits), interleave, corrupt a known subset, and compare
    with (a) the construction and (b) the oracle's trapdoor verify on a sample.  Memory per issuer and the throughput
    relative to the single-issuer path are printed."""
    import time
    import numpy as np
    import bench
    from bbs_sign_b200 import api as A
    from oracle import bbs_oracle as O
    suite, ocs = P.SUITES["BLS12_381"]
    L, n_keys, copies, per = 10, 8, 128, 4
    header = b""
    keys = [P.keypair(ocs, 40 + k) for k in range(n_keys)]
    pkb = [ocs.g2_compress(pk) for _, pk in keys]
    n_issuers = n_keys * copies
    set_pks = [pkb[s % n_keys] for s in range(n_issuers)]
    rng = np.random.default_rng(8)
    n = n_issuers * per
    msgs = rng.integers(0, 256, size=(n, L, 32), dtype=np.uint8)
    issuer = np.repeat(np.arange(n_issuers, dtype=np.uint32), per)
    rng.shuffle(issuer)
    sigs = np.zeros((n, 80), dtype=np.uint8)
    for k in range(n_keys):
        ctx = A.BatchContext(suite, pkb[k], header, n_messages=L)
        sel = np.nonzero(issuer % n_keys == k)[0]
        s_k, _, st = ctx.sign_batch(ocs.scalar_le(keys[k][0]), [[bytes(msgs[i, j]) for j in range(L)] for i in sel])
        assert (st == 1).all()
        sigs[sel] = s_k
        ctx.close()
    expect = np.ones(n, dtype=np.uint8)
    claimed = issuer.copy()
    for t, i in enumerate(range(7, n, 16)):
        if t % 2 == 0:
            claimed[i] = (issuer[i] + 1) % n_issuers          # another issuer's key (a different secret: +1 mod 8)
        else:
            sigs[i, 48] ^= 1
        expect[i] = 0
    t0 = time.perf_counter()
    iset = A.IssuerSet(suite, set_pks, header, n_messages=L)
    t_create = time.perf_counter() - t0
    assert (iset.status == 1).all()
    per_b, shared_b = iset.memory_bytes()
    msg_lists = [[bytes(msgs[i, j]) for j in range(L)] for i in range(n)]
    got = iset.verify_batch(claimed, sigs.reshape(-1), msg_lists)
    assert np.array_equal(got, expect), int((got != expect).sum())
    t0 = time.perf_counter()
    got = iset.verify_batch(claimed, sigs.reshape(-1), msg_lists)
    t_multi = time.perf_counter() - t0
    # oracle (trapdoor) on a sample, valid and corrupted
    gens = O.create_generators_cached(ocs, L + 1, ocs.api_id)
    for i in [7, 23, 39, 55, 0, 1, 1000, n - 1]:
        sk_c, pk_c = keys[int(claimed[i]) % n_keys]
        sig = (ocs.g1_decompress(bytes(sigs[i, :48])), int.from_bytes(bytes(sigs[i, 48:]), "little"))
        assert bool(got[i]) == O.verify(ocs, pk_c, sig, header, msg_lists[i], trapdoor_sk=sk_c), i
    # the same number of items under ONE issuer through the single-issuer path
    ctx = A.BatchContext(suite, pkb[0], header, n_messages=L)
    s_1, _, _ = ctx.sign_batch(ocs.scalar_le(keys[0][0]), msg_lists)
    ctx.verify_batch(s_1.reshape(-1), msg_lists)
    t0 = time.perf_counter()
    one = ctx.verify_batch(s_1.reshape(-1), msg_lists)
    t_single = time.perf_counter() - t0
    assert (one == 1).all()
    ctx_bytes = ctx.memory_bytes()
    ctx.close()
    iset.close()
    print(f"\nmulti-issuer: {n_issuers} issuers x {per} items: set creation {t_create * 1e3:.0f} ms, {per_b / n_issuers / 1024:.1f} KiB per issuer "
          f"(+ {shared_b / 2**20:.0f} MiB shared) vs {ctx_bytes / 2**20:.0f} MiB for one single-issuer context; "
          f"verify {n} items: {t_multi * 1e3:.1f} ms multi-issuer vs {t_single * 1e3:.1f} ms single-issuer (host-buffer calls, Python packing included)")
    assert per_b / n_issuers < 64 * 1024


@pytest.mark.parametrize("curve", ["BN254", "BLS12_381"])
def test_core_api_id(lib, curve):
    P.case_core_api_id(None, curve, L=5)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_g1_mul_edges(lib, curve):
    P.case_g1_mul_edges(lib, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_subgroup_validation(lib, curve):
    P.case_subgroup(None, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_per_thread_pairing_kernel(lib, curve):
    """the one-thread-per-item pairing kernel (the fallback of a context with a degenerate line) gives the same verdicts as
    the cooperative kernel and the oracle on valid items and every rejection class"""
    P.case_verify(None, curve, 3, n=12, use_pairing_oracle_on=1, per_thread_pairing=True)
    P.case_proof_verify(None, curve, 4, [0, 2], n=8, pairing_on=1, per_thread_pairing=True)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_multi_issuer_set(lib, curve):
    P.case_multi_issuer(None, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_multi_issuer_proofs(lib, curve):
    """bbs_proof_verify_batch_multi: a proof batch naming a different issuer per item, against the oracle per item"""
    P.case_multi_issuer_proofs(None, curve, n_issuers=3, pairing_on=1)
    P.case_multi_issuer_proofs(None, curve, n_issuers=2, L=5, disclosed=(1, 3, 4), pairing_on=0)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_small_tables(lib, curve, monkeypatch):
    """contexts built with BBS_CTX_SMALL_TABLES (8-bit windows, L2-resident tables) give byte-identical signatures, B points,
    proofs and verdicts"""
    monkeypatch.setattr(P, "SMALL_TABLES", True)
    P.case_verify(None, curve, 3, n=6, use_pairing_oracle_on=1)
    P.case_proof_gen(None, curve, 3, [0, 2], n=3)
    P.case_proof_verify(None, curve, 4, [0, 2], n=8, pairing_on=0)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_one_thread_per_item_g1_kernel(lib, curve, monkeypatch):
    """verify / core_verify through the one-thread-per-item G1 kernel (bbs_ctx_set_g1_split(0)); the default is the two-task
    split, which every other test exercises"""
    from bbs_sign_b200 import api as A
    orig = A.BatchContext.__init__

    def init(self, *a, **k):
        orig(self, *a, **k)
        self.set_g1_split(0)

    monkeypatch.setattr(A.BatchContext, "__init__", init)
    P.case_verify(None, curve, 3, n=12, use_pairing_oracle_on=1)
    P.case_verify_malformed(None, curve)
    P.case_subgroup(None, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_one_warp_per_role_pairing_kernel_on_small_batches(lib, curve, monkeypatch):
    """batches of up to 32 items run the cooperative pairing kernel with two warps per role (every small case of this file);
    here the same cases go through the one-warp-per-role kernel the large batches use (bbs_ctx_set_pairing_split(0))"""
    monkeypatch.setattr(P, "PAIRING_SPLIT", 0)
    P.case_verify(None, curve, 3, n=12, use_pairing_oracle_on=2)
    P.case_verify(None, curve, 1, n=33, use_pairing_oracle_on=0)
    P.case_proof_verify(None, curve, 4, [0, 2], n=8, pairing_on=1)
    P.case_rlc(None, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_two_warps_per_role_pairing_kernel_sizes(lib, curve):
    """the SPLIT = 2 pairing kernel at its boundaries: 1, 31 and 32 items (33 is the first size of the other kernel)"""
    for n in (1, 31, 32, 33):
        P.case_verify(None, curve, 2, n=n, use_pairing_oracle_on=1 if n == 1 else 0)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_split_paths_agree_with_one_thread_paths_on_damaged_inputs(lib, curve):
    for seed in (1, 2, 3):
        P.case_split_differential(None, curve, n=64, seed=seed)
