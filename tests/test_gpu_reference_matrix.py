"""The reference's own `#[test_case]` matrices, case for case, on both curves, through the C ABI on the GPU:
  proof_verify_tests.rs:50-57 (BN254) / bbs_over_bls_tests.rs:41-48 (BLS12-381)  (count, disclosed_indexes, header)
  sign_verify_tests.rs:39-43                                                     (count, header)
  core_sign_tests.rs:38-45                                                       (count, api_id, header)
Each case runs the reference's flow (sign -> verify -> proof_gen -> proof_verify with ph = &[], 5-byte messages as
`generate_random_msg` makes them; core_sign / core_verify over `Fr::from(u64)` messages with generators made for the
case's api_id) and additionally pins every produced byte to the oracle."""
import numpy as np
import pytest

import parity_cases as P
from bbs_sign_b200 import _native
from bbs_sign_b200 import api as A
from oracle import bbs_oracle as O

pytestmark = pytest.mark.gpu
CURVES = ["BN254", "BLS12_381"]


@pytest.fixture(scope="module")
def lib():
    return _native.load()


PROOF_MATRIX = [(0, [], b""), (0, [], b"abc"), (1, [0], b"abc"), (1, [], b"abc"), (10, [0, 1, 2], b""),
                (10, [0, 4, 7, 9], b"def"), (5, [0, 4], b"defghjsdjdbcjbejd"), (5, [0, 1, 2, 3, 4], b"def")]


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("count,disclosed,header", PROOF_MATRIX)
def test_proof_verify_matrix(lib, curve, count, disclosed, header):
    suite, ocs = P.SUITES[curve]
    sk, pk = P.keypair(ocs, 1)
    ctx, gens = P.make_ctx(None, suite, ocs, pk, header, count)
    msgs = [P.rng_bytes(f"pm{count}.{j}.{header!r}", 5) for j in range(count)]
    sigs, _, st = ctx.sign_batch(ocs.scalar_le(sk), [msgs])
    assert st.tolist() == [1]
    osig = O.sign(ocs, sk, msgs, header)
    assert sigs[0].tobytes() == O.signature_to_bytes(ocs, osig)
    assert ctx.verify_batch(sigs, [msgs]).tolist() == [1]
    U = count - len(disclosed)
    rs = O.seeded_random_scalars(ocs, f"pm{count}{disclosed}".encode(), b"rs-dst", 5 + U)
    got, gst = ctx.proof_gen_batch([sigs[0].tobytes()], [msgs], [list(disclosed)], [[ocs.scalar_le(x) for x in rs]], b"")
    assert gst.tolist() == [1]
    want = O.proof_gen(ocs, pk, osig, header, b"", msgs, list(disclosed), random_scalars=rs)
    wb = A.ProofBytes.from_canonical(suite, O.proof_to_bytes(ocs, want))
    assert got[0].fixed == wb.fixed and got[0].commitments == wb.commitments
    dm = [msgs[i] for i in disclosed]
    assert ctx.proof_verify_batch(got, b"", [dm], [list(disclosed)]).tolist() == [1]
    assert O.proof_verify(ocs, pk, want, header, b"", dm, list(disclosed)) is True          # the oracle's own pairing check
    ctx.close()


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("count,header", [(5, b""), (0, b""), (0, b"abc"), (10, b""), (10, b"abc")])
def test_sign_and_verify_matrix(lib, curve, count, header):
    suite, ocs = P.SUITES[curve]
    sk, pk = P.keypair(ocs, 1)
    ctx, gens = P.make_ctx(None, suite, ocs, pk, header, count)
    msgs = [P.rng_bytes(f"sv{count}.{j}", 5) for j in range(count)]
    sigs, _, st = ctx.sign_batch(ocs.scalar_le(sk), [msgs])
    assert st.tolist() == [1]
    osig = O.sign(ocs, sk, msgs, header)
    assert sigs[0].tobytes() == O.signature_to_bytes(ocs, osig)
    assert ctx.verify_batch(sigs, [msgs]).tolist() == [1]
    assert O.verify(ocs, pk, osig, header, msgs) is True
    ctx.close()


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("count,api_id,header", [(5, b"", b""), (0, b"", b""), (0, b"abc", b""), (0, b"", b"abc"), (10, b"", b""),
                                                 (10, b"abc", b"def"), (10, b"", b"def"), (10, b"abc", b"")])
def test_core_sign_and_verify_matrix(lib, curve, count, api_id, header):
    suite, ocs = P.SUITES[curve]
    sk, pk = P.keypair(ocs, 1)
    gens = O.create_generators_cached(ocs, count + 1, api_id)
    gb = P.gens_bytes(ocs, gens)
    assert suite.derive_generators(count + 1, api_id=api_id) == gb            # create_generators on the device
    ctx = A.BatchContext(suite, ocs.g2_compress(pk), header, generators=gb, api_id=api_id)
    scal = [(0x9E3779B97F4A7C15 * (j + 1)) % (1 << 64) for j in range(count)]     # Fr::from(u64)
    flat = np.frombuffer(b"".join(ocs.scalar_le(x) for x in scal), dtype=np.uint8) if count else np.zeros(0, np.uint8)
    sigs, _, st = ctx.core_sign_batch(ocs.scalar_le(sk), flat, 1, count)
    assert st.tolist() == [1]
    osig = O.core_sign(ocs, sk, gens, header, scal, api_id)
    assert sigs[0].tobytes() == O.signature_to_bytes(ocs, osig)
    assert ctx.core_verify_batch(sigs.reshape(-1), flat, count).tolist() == [1]
    assert O.core_verify(ocs, pk, osig, gens, header, scal, api_id) is True
    ctx.close()
