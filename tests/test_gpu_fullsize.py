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
    """configs[2]: 1,048,576 signatures x L = 31 pre-hashed scalars on BN254; bench.run_bn254 raises on any status that
    differs from the construction (every 16th item has e ^ 1)"""
    import bench
    out = bench.run_bn254(argparse.Namespace(n=1 << 20, steps=1, warmup=1))
    assert out["config"]["n_per_gpu"] == 1 << 20 and out["value"] > 0


def test_config4_bls_proof_verify_262144(lib):
    """configs[3]: 262,144 proofs, L = 32, 16 disclosed: the oracle-made fixture (valid + every rejection class) tiled to
    full size; bench.run_proof raises unless every verdict equals the one the oracle recorded"""
    import bench
    out = bench.run_proof(argparse.Namespace(n=262144, steps=1, warmup=1))
    assert out["value"] > 0


def test_config5_bls_sign_and_rlc_524288(lib):
    """configs[4] per-GPU share (4M / 8): sign 524,288 message sets, one random-linear-combination verdict; the valid batch
    must be accepted and the same batch with one flipped bit rejected (bench.run_rlc raises otherwise)"""
    import bench
    out = bench.run_rlc(argparse.Namespace(n=524288, L=10, steps=1, warmup=1))
    assert out["value"] > 0
