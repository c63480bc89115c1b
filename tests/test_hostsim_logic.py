"""GPU-less logic tests: the CUDA sources compiled for x86 (tests/hostsim.py) driven through the same C
ABI and the same parity cases the GPU tests use.  They check the pipeline logic, byte plumbing and
the oracle agreement of everything except the PTX arithmetic itself (which only a GPU can run)."""
import numpy as np
import pytest

import hostsim
import parity_cases as P
from bbs_sign_b200 import _native


@pytest.fixture(scope="module")
def lib_path():
    path = hostsim.build()
    _native.load(path, allow_host_simulation=True)      # tests only: the product loader refuses non-CUDA builds
    return path


@pytest.fixture(scope="module")
def lib(lib_path):
    return _native.load(lib_path, allow_host_simulation=True)


def test_product_loader_refuses_the_host_simulation(lib_path):
    _native._CACHE.pop(lib_path, None)
    with pytest.raises(_native.NativeLibraryMissing):
        _native.load(lib_path)
    _native.load(lib_path, allow_host_simulation=True)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_field(lib, curve):
    P.case_field(lib, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_g1_mul(lib, curve):
    P.case_g1_mul(lib, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_pairing(lib, curve):
    P.case_pairing(lib, curve)


def test_irtf_kat(lib_path):
    P.case_irtf_kat(lib_path)


@pytest.mark.parametrize("curve,L", [("BLS12_381", 3), ("BN254", 2), ("BLS12_381", 0)])
def test_verify(lib_path, curve, L):
    P.case_verify(lib_path, curve, L, n=6, use_pairing_oracle_on=1)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_verify_malformed(lib_path, curve):
    P.case_verify_malformed(lib_path, curve)


@pytest.mark.parametrize("curve,L,dis", [("BLS12_381", 4, [0, 2]), ("BN254", 3, [1]), ("BLS12_381", 2, [0, 1]),
                                         ("BN254", 2, [])])
def test_proof_verify(lib_path, curve, L, dis):
    P.case_proof_verify(lib_path, curve, L, dis, n=8, pairing_on=1)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_proof_errors(lib_path, curve):
    P.case_proof_errors(lib_path, curve)


@pytest.mark.parametrize("curve,L,dis", [("BLS12_381", 3, [0, 2]), ("BN254", 2, [1])])
def test_proof_gen(lib_path, curve, L, dis):
    P.case_proof_gen(lib_path, curve, L, dis, n=3)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_h2s_ragged(lib_path, curve):
    P.case_h2s_ragged(lib_path, curve)


@pytest.mark.parametrize("curve", ["BN254", "BLS12_381"])
def test_readme_example(lib_path, curve):
    P.case_readme_example(lib_path, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_create_generators(lib_path, curve):
    P.case_create_generators(lib_path, curve, count=3)


@pytest.mark.parametrize("curve", ["BN254", "BLS12_381"])
def test_core_api_id(lib_path, curve):
    P.case_core_api_id(lib_path, curve, L=3)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_g1_mul_edges(lib, curve):
    P.case_g1_mul_edges(lib, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_subgroup_validation(lib_path, curve):
    P.case_subgroup(lib_path, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_multi_issuer_set(lib_path, curve):
    P.case_multi_issuer(lib_path, curve)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_multi_issuer_proofs(lib_path, curve):
    P.case_multi_issuer_proofs(lib_path, curve, n_issuers=2, pairing_on=0)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_small_tables(lib_path, curve, monkeypatch):
    """contexts built with BBS_CTX_SMALL_TABLES (8-bit windows, L2-resident tables) give byte-identical signatures, B points,
    proofs and verdicts"""
    monkeypatch.setattr(P, "SMALL_TABLES", True)
    P.case_verify(lib_path, curve, 3, n=6, use_pairing_oracle_on=1)
    P.case_proof_gen(lib_path, curve, 3, [0, 2], n=3)
    P.case_proof_verify(lib_path, curve, 4, [0, 2], n=8, pairing_on=0)


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_g1_task_split(lib_path, curve, monkeypatch):
    """the G1 halves of verify / proof-verify as independent tasks + a join (kernels.cuh verify_task_* / proof_task_*): the
    CUDA build's default; here the same task functions run in sequence"""
    monkeypatch.setattr(P, "G1_SPLIT", 1 << 62)
    P.case_verify(lib_path, curve, 3, n=8, use_pairing_oracle_on=1)
    P.case_verify_malformed(lib_path, curve)
    P.case_subgroup(lib_path, curve)
    P.case_proof_verify(lib_path, curve, 4, [0, 2], n=10, pairing_on=0)
    P.case_proof_verify(lib_path, curve, 3, [], n=6, pairing_on=0)
    P.case_proof_verify(lib_path, curve, 3, [0, 1, 2], n=6, pairing_on=0)
    P.case_proof_errors(lib_path, curve)


def test_public_key_keeps_a_bounded_number_of_contexts(lib_path):
    """PublicKey.context: header and L are per call in the reference (verify.rs:18-30), a context per distinct pair is device
    memory; the least recently used one is destroyed beyond MAX_CONTEXTS"""
    from bbs_sign_b200 import api as A
    suite, ocs = P.SUITES["BLS12_381"]
    sk = P.O.key_gen(ocs, bytes([7] * 32), b"", b"BBS-SIG-KEYGEN-SALT-")
    pk = A.PublicKey(suite, ocs.g2_compress(P.O.sk_to_pk(ocs, sk)), lib_path=lib_path)
    first = pk.context(b"h0", 1)
    ctxs = [pk.context(b"h%d" % i, 1) for i in range(1, A.PublicKey.MAX_CONTEXTS)]
    assert pk.context(b"h0", 1) is first and first._h                      # a hit refreshes the entry
    pk.context(b"one more", 1)
    assert len(pk._ctx) == A.PublicKey.MAX_CONTEXTS
    assert ctxs[0]._h is None and first._h                                 # h1 was the least recently used
    msgs = [[b"m"]]
    sig = P.O.signature_to_bytes(ocs, P.O.sign(ocs, sk, msgs[0], b"h1"))
    assert pk.verify_batch(np.frombuffer(sig, dtype=np.uint8).reshape(1, -1), b"h1", msgs).tolist() == [1]              # rebuilt on demand


@pytest.mark.parametrize("curve", ["BLS12_381", "BN254"])
def test_split_paths_agree_with_one_thread_paths_on_damaged_inputs(lib_path, curve):
    P.case_split_differential(lib_path, curve, n=24, seed=2)
