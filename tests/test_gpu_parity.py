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
