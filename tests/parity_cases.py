"""Parity cases shared by the GPU tests (`-m gpu`, CUDA library through the C ABI) and the GPU-less
logic tests (host simulation of the same sources).  Every case compares the library with the oracle
(oracle/bbs_oracle.py) on the same seeded inputs; the bar is bit-exact bytes / identical status vectors.

The cases mirror the reference's own tests: test_vector.rs (KATs), core_sign_tests.rs, sign_verify_tests.rs,
proof_verify_tests.rs, bbs_over_bls_tests.rs (round trips and the rejection classes)."""
import ctypes as C
import hashlib
import random

import numpy as np

from bbs_sign_b200 import api as A
from oracle import bbs_oracle as O

SUITES = {"BLS12_381": (A.BLS12_381, O.BLS12_381), "BN254": (A.BN254, O.BN254)}

MSG = bytes.fromhex("9872ad089e452c7b6e283dfac2a80d58e8d0ff71cc4d5e310a1debdda4a45f02")
HEADER = bytes.fromhex("11223344556677889900aabbccddeeff")
PH = bytes.fromhex("bed231d880675ed101ead304512e043ade9958dd0241ea70b4b3957fba941501")
IRTF_SK = 0x60e55110f76883a13d030b2f6bd11883422d5abde717569fc0731f51237169fc
IRTF_SIG = bytes.fromhex("84773160b824e194073a57493dac1a20b667af70cd2352d8af241c77658da5253aa8458317cca0eae615690d55b1f271"
                         "64657dcafee1d5c1973947aa70e2cfbb4c892340be5969920d0916067b4565a0")
IRTF_B = bytes.fromhex("92d264aed02bf23de022ebe778c4f929fddf829f504e451d011ed89a313b8167ac947332e1648157ceffc6e6e41ab255")
IRTF_DOMAIN = 0x25d57fab92a8274c68fde5c3f16d4b275e4a156f211ae34b3ab32fbaf506ed5c


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def rng_bytes(seed: str, n: int) -> bytes:
    out = b""
    i = 0
    while len(out) < n:
        out += hashlib.sha256(f"{seed}/{i}".encode()).digest()
        i += 1
    return out[:n]


def keypair(ocs, seed=1):
    km = hashlib.sha256(b"bbs-b200-key" + seed.to_bytes(4, "big")).digest()
    sk = O.key_gen(ocs, km, b"", b"BBS-SIG-KEYGEN-SALT-")
    return sk, O.sk_to_pk(ocs, sk)


def gens_bytes(ocs, gens):
    return b"".join(ocs.g1_compress(g) for g in gens)


G1_SPLIT = None             # None: the library default; else bbs_ctx_set_g1_split on every context make_ctx builds
PAIRING_SPLIT = None        # None: the library default (batches of <= 32 items: two warps per role); else bbs_ctx_set_pairing_split
SMALL_TABLES = False      # test_small_tables flips this: every context of a case is then built with BBS_CTX_SMALL_TABLES


def make_ctx(lib_path, suite, ocs, pk, header, L, api_id=None, gens=None):
    api_id = ocs.api_id if api_id is None else api_id
    if gens is None:
        gens = O.create_generators_cached(ocs, L + 1, api_id)
    ctx = A.BatchContext(suite, ocs.g2_compress(pk), header, generators=gens_bytes(ocs, gens), api_id=api_id,
                         lib_path=lib_path, small_tables=SMALL_TABLES)
    if G1_SPLIT is not None:
        ctx.set_g1_split(G1_SPLIT)
    if PAIRING_SPLIT is not None:
        ctx.set_pairing_split(PAIRING_SPLIT)
    return ctx, gens


# ---------------------------------------------------------------------------------------------------
def case_field(lib, curve_name, n=64):
    suite, ocs = SUITES[curve_name]
    rnd = random.Random(7)
    for op, mod, width in [(0, ocs.p, ocs.fp_bytes), (1, ocs.p, ocs.fp_bytes), (2, ocs.p, ocs.fp_bytes),
                           (3, ocs.p, ocs.fp_bytes), (4, ocs.p, ocs.fp_bytes), (5, ocs.r, 32), (6, ocs.r, 32),
                           (7, ocs.p, ocs.fp_bytes), (8, ocs.r, 32)]:     # 7, 8: the variable-time inversion
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << (8 * width)) % mod, (mod - 1) // 2, (mod + 1) // 2]
        xs = edge + [rnd.randrange(mod) for _ in range(n - len(edge))]
        ys = list(reversed(edge)) + [rnd.randrange(mod) for _ in range(n - len(edge))]
        if op == 4:
            xs = [x * x % mod for x in xs[: n // 2]] + xs[n // 2:]
        a = np.frombuffer(b"".join(x.to_bytes(width, "little") for x in xs), dtype=np.uint8)
        b = np.frombuffer(b"".join(y.to_bytes(width, "little") for y in ys), dtype=np.uint8)
        out = np.zeros(n * width, dtype=np.uint8)
        rc = lib.bbs_selftest_field(suite.curve_id, 0, op, n, ptr(a), ptr(b), ptr(out))
        assert rc == 0, lib.bbs_last_error()
        got = [int.from_bytes(out[i * width:(i + 1) * width].tobytes(), "little") for i in range(n)]
        for i, (x, y, g) in enumerate(zip(xs, ys, got)):
            if op in (0, 5):
                want = x * y % mod
            elif op == 1:
                want = (x + y) % mod
            elif op == 2:
                want = (x - y) % mod
            elif op in (3, 6, 7, 8):
                want = pow(x, mod - 2, mod)
            else:
                s = pow(x, (mod + 1) // 4, mod)
                want = s if s * s % mod == x else 0
                if want and g != want:
                    want = mod - want  # either root is a square root; the kernel returns a^((p+1)/4)
            assert g == want, (curve_name, op, i, hex(x), hex(y), hex(g), hex(want))


def case_g1_mul(lib, curve_name, n=12):
    suite, ocs = SUITES[curve_name]
    rnd = random.Random(11)
    pts, ks = [], []
    for i in range(n):
        P = O.ec_mul(ocs.F1, ocs.BP1, rnd.randrange(1, ocs.r))
        k = [0, 1, 2, ocs.r - 1, ocs.r, (1 << 256) - 1][i] if i < 6 else rnd.randrange(ocs.r)
        if i == 7:
            P = None
        pts.append(P)
        ks.append(k)
    a = np.frombuffer(b"".join(ocs.g1_compress(P) for P in pts), dtype=np.uint8)
    b = np.frombuffer(b"".join(k.to_bytes(32, "little") for k in ks), dtype=np.uint8)
    out = np.zeros(n * suite.g1_bytes, dtype=np.uint8)
    rc = lib.bbs_selftest_g1_mul(suite.curve_id, 0, n, ptr(a), ptr(b), ptr(out))
    assert rc == 0, lib.bbs_last_error()
    for i in range(n):
        want = ocs.g1_compress(O.ec_mul(ocs.F1, pts[i], ks[i]))
        got = out[i * suite.g1_bytes:(i + 1) * suite.g1_bytes].tobytes()
        assert got == want, (curve_name, i, got.hex(), want.hex())


def case_pairing(lib, curve_name):
    suite, ocs = SUITES[curve_name]
    rnd = random.Random(5)
    a_, b_ = rnd.randrange(1, ocs.r), rnd.randrange(1, ocs.r)
    Q = O.ec_mul(ocs.F2, ocs.BP2, b_)
    P = O.ec_mul(ocs.F1, ocs.BP1, a_)
    R_good = O.ec_neg(ocs.F1, O.ec_mul(ocs.F1, ocs.BP1, a_ * b_ % ocs.r))      # e(P,Q) e(R,BP2) = 1
    R_bad = O.ec_mul(ocs.F1, ocs.BP1, (a_ * b_ + 1) % ocs.r)
    cases = [(P, R_good, 1), (P, R_bad, 0), (None, R_good, 0), (P, None, 0), (None, None, 1)]
    p = np.frombuffer(b"".join(ocs.g1_compress(c[0]) for c in cases), dtype=np.uint8)
    r = np.frombuffer(b"".join(ocs.g1_compress(c[1]) for c in cases), dtype=np.uint8)
    q = np.frombuffer(ocs.g2_compress(Q), dtype=np.uint8)
    st = np.full(len(cases), 255, dtype=np.uint8)
    rc = lib.bbs_selftest_pairing(suite.curve_id, 0, len(cases), ptr(p), ptr(r), ptr(q), ptr(st))
    assert rc == 0, lib.bbs_last_error()
    assert st.tolist() == [c[2] for c in cases], st.tolist()
    # identity Q: first pair contributes 1
    q = np.frombuffer(ocs.g2_compress(None), dtype=np.uint8)
    rc = lib.bbs_selftest_pairing(suite.curve_id, 0, len(cases), ptr(p), ptr(r), ptr(q), ptr(st))
    assert rc == 0
    assert st.tolist() == [0, 0, 0, 1, 1], st.tolist()


def case_irtf_kat(lib_path):
    """test_vector.rs: msg->scalar, domain, B, signature and proof fixtures through the ABI (BLS12-381)."""
    suite, ocs = SUITES["BLS12_381"]
    pk = O.sk_to_pk(ocs, IRTF_SK)
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, HEADER, 1)
    sc = ctx.msg_to_scalars([MSG, b""])
    assert sc[0].tobytes()[::-1].hex() == "1cb5bb86114b34dc438a911617655a1db595abafac92f47c5001799cf624b430"
    assert sc[1].tobytes()[::-1].hex() == "08e3afeb2b4f2b5f907924ef42856616e6f2d5f1fb373736db1cca32707a7d16"
    assert int.from_bytes(ctx.domain(), "little") == IRTF_DOMAIN
    sigs, b, st = ctx.sign_batch(IRTF_SK.to_bytes(32, "little"), [[MSG]], want_b=True)
    assert st.tolist() == [1]
    got = sigs[0].tobytes()
    assert got[:48] + got[48:][::-1] == IRTF_SIG            # test_vector.rs:187-191 (e is big-endian there)
    assert b[0].tobytes() == IRTF_B
    assert ctx.verify_batch(sigs, [[MSG]]).tolist() == [1]
    # proof fixture (test_vector.rs:199-260): produced by the oracle (pinned to the KAT), verified by the library
    sig = (ocs.g1_decompress(IRTF_SIG[:48]), int.from_bytes(IRTF_SIG[48:], "big"))
    proof = O.proof_gen(ocs, pk, sig, HEADER, PH, [MSG], [0])
    pb = A.ProofBytes.from_canonical(suite, O.proof_to_bytes(ocs, proof))
    assert ctx.proof_verify_batch([pb], PH, [[MSG]], [[0]]).tolist() == [1]
    assert ctx.proof_verify_batch([pb], b"", [[MSG]], [[0]]).tolist() == [0]
    # ... and produced by the library itself from the mocked random scalars (core_utilities.rs:84-113, the stream the
    # reference uses under feature testvector_bls12_381): byte-identical to the IRTF proof fixture
    rs = [ocs.scalar_le(x) for x in O.mocked_calculate_random_scalars(ocs, 5)]
    gen, gst = ctx.proof_gen_batch([sigs[0].tobytes()], [[MSG]], [[0]], [rs], PH)
    assert gst.tolist() == [1]
    assert gen[0].fixed == pb.fixed and gen[0].commitments == pb.commitments
    ctx.close()


def _corrupt_sig(ocs, sig, kind, other_sig):
    a, e = sig
    if kind == "e+1":
        return (a, (e + 1) % ocs.r)
    if kind == "A=identity":
        return (None, e)
    if kind == "wrong-issuer":
        return other_sig
    return sig


def case_verify(lib_path, curve_name, L, n=10, header=b"hdr", use_pairing_oracle_on=2, per_thread_pairing=False):
    """sign -> verify round trips and the rejection classes of sign_verify_tests.rs / core_sign_tests.rs."""
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    sk2, pk2 = keypair(ocs, 2)
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, header, L)
    if per_thread_pairing:
        ctx.use_per_thread_pairing(True)
    msgs = [[rng_bytes(f"m{i}.{j}", 32 if j % 3 else 5) for j in range(L)] for i in range(n)]
    if L:
        msgs[0][0] = b""   # empty message (test_vector.rs:115-119)
    # library signs; oracle signs; bytes must agree
    sigs, b_pts, st = ctx.sign_batch(sk.to_bytes(32, "little"), msgs, want_b=True)
    assert st.tolist() == [1] * n
    osigs = [O.sign(ocs, sk, m, header) for m in msgs]
    for i in range(n):
        assert sigs[i].tobytes() == O.signature_to_bytes(ocs, osigs[i]), (curve_name, L, i)
        ms = O.msg_to_scalars(ocs, msgs[i], ocs.api_id)
        dom = O.calculate_domain(ocs, pk, gens[0], gens[1:], header, ocs.api_id)
        assert b_pts[i].tobytes() == ocs.g1_compress(O.compute_B(ocs, gens, dom, ms))
    # verify: valid + corrupted, expectations from the oracle
    kinds = ["ok", "e+1", "A=identity", "wrong-issuer", "msg-flip", "ok"]
    items, exp = [], []
    for i in range(n):
        kind = kinds[i % len(kinds)]
        m = list(msgs[i])
        s = osigs[i]
        if kind == "msg-flip" and L:
            m[L - 1] = m[L - 1] + b"x"
        else:
            s = _corrupt_sig(ocs, s, kind, O.sign(ocs, sk2, msgs[i], header))
        items.append((s, m))
        if i < use_pairing_oracle_on:
            exp.append(int(O.verify(ocs, pk, s, header, m)))
        else:
            exp.append(int(O.verify(ocs, pk, s, header, m, trapdoor_sk=sk)))
    blob = b"".join(O.signature_to_bytes(ocs, s) for s, _ in items)
    got = ctx.verify_batch(np.frombuffer(blob, dtype=np.uint8), [m for _, m in items])
    assert got.tolist() == exp, (curve_name, L, got.tolist(), exp)
    # core_verify with scalar messages gives the same verdicts
    sc = b"".join(ocs.scalar_le(x) for _, m in items for x in O.msg_to_scalars(ocs, m, ocs.api_id))
    got2 = ctx.core_verify_batch(np.frombuffer(blob, dtype=np.uint8), np.frombuffer(sc, dtype=np.uint8) if sc else np.zeros(0, np.uint8), L)
    assert got2.tolist() == exp
    # generator / message count mismatch -> Err(InvalidMessageAndGeneratorsLength) (verify.rs:68-71)
    sc1 = b"".join(ocs.scalar_le(7) for _ in range(n * (L + 1)))
    got3 = ctx.core_verify_batch(np.frombuffer(blob, dtype=np.uint8), np.frombuffer(sc1, dtype=np.uint8), L + 1)
    assert got3.tolist() == [A.ST_ERR_MSG_GEN_LEN] * n
    ctx.close()
    # default (identity) public key -> Ok(false) (sign_verify_tests.rs:84-93)
    ctx0, _ = make_ctx(lib_path, suite, ocs, None, header, L)
    got4 = ctx0.verify_batch(np.frombuffer(blob, dtype=np.uint8), [m for _, m in items])
    exp4 = [int(O.verify(ocs, None, s, header, m)) if i < 1 else 0 for i, (s, m) in enumerate(items)]
    assert got4.tolist() == exp4
    ctx0.close()


def case_verify_malformed(lib_path, curve_name):
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    L = 2
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, b"", L)
    msgs = [[b"a", b"b"]] * 4
    sig = O.sign(ocs, sk, msgs[0], b"")
    good = O.signature_to_bytes(ocs, sig)
    g = suite.g1_bytes
    # x not on the curve: search a small x with no square root
    bad_pt = None
    x = 1
    while bad_pt is None:
        rhs = (x ** 3 + ocs.b) % ocs.p
        if pow(rhs, (ocs.p - 1) // 2, ocs.p) != 1:
            enc = bytearray(x.to_bytes(g, "big" if curve_name == "BLS12_381" else "little"))
            if curve_name == "BLS12_381":
                enc[0] |= 0x80
            bad_pt = bytes(enc)
        x += 1
    non_canon_e = good[:g] + (ocs.r + 5).to_bytes(32, "little")
    x_ge_p = bytearray(good)
    if curve_name == "BLS12_381":
        x_ge_p[:g] = bytes([0x9f]) + bytes([0xff]) * (g - 1)
    else:
        x_ge_p[:g] = bytes([0xff]) * (g - 1) + bytes([0x3f])
    blob = good + bad_pt + good[g:] + non_canon_e + bytes(x_ge_p)
    st = ctx.verify_batch(np.frombuffer(blob, dtype=np.uint8), msgs)
    assert st.tolist() == [1, A.ST_ERR_MALFORMED, A.ST_ERR_MALFORMED, A.ST_ERR_MALFORMED], st.tolist()
    ctx.close()


def make_proofs(ocs, sk, pk, header, ph, L, disclosed, n, seed="p"):
    out = []
    for i in range(n):
        msgs = [rng_bytes(f"{seed}{i}.{j}", 32) for j in range(L)]
        sig = O.sign(ocs, sk, msgs, header)
        rs = O.seeded_random_scalars(ocs, f"{seed}{i}".encode(), b"rs-dst", 5 + L - len(set(disclosed)))
        pr = O.proof_gen(ocs, pk, sig, header, ph, msgs, disclosed, random_scalars=rs)
        out.append((pr, msgs))
    return out


def case_proof_verify(lib_path, curve_name, L, disclosed, n=8, header=b"h", ph=b"ph", pairing_on=1, per_thread_pairing=False):
    """proof_verify_tests.rs / bbs_over_bls_tests.rs: happy paths and forged inputs."""
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, header, L)
    if per_thread_pairing:
        ctx.use_per_thread_pairing(True)
    base = make_proofs(ocs, sk, pk, header, ph, L, disclosed, n)
    dis = sorted(set(disclosed))
    kinds = ["ok", "Abar=identity", "challenge+1", "swap-commit", "wrong-msg", "ok", "Bbar-mutated", "e_cap+1"]
    proofs, dmsgs, didx, exp = [], [], [], []
    for i, (pr, msgs) in enumerate(base):
        kind = kinds[i % len(kinds)]
        p = O.Proof(pr.a_bar, pr.b_bar, pr.d, pr.e_cap, pr.r1_cap, pr.r3_cap, list(pr.commitments), pr.challenge)
        dm = [msgs[j] for j in dis]
        if kind == "Abar=identity":
            p.a_bar = None
        elif kind == "challenge+1":
            p.challenge = (p.challenge + 1) % ocs.r
        elif kind == "swap-commit" and len(p.commitments) >= 2:
            p.commitments[0], p.commitments[1] = p.commitments[1], p.commitments[0]
        elif kind == "wrong-msg" and dm:
            dm[0] = dm[0] + b"!"
        elif kind == "Bbar-mutated":
            p.b_bar = O.ec_add(ocs.F1, p.b_bar, ocs.BP1)
        elif kind == "e_cap+1":
            p.e_cap = (p.e_cap + 1) % ocs.r
        proofs.append(A.ProofBytes.from_canonical(suite, O.proof_to_bytes(ocs, p)))
        dmsgs.append(dm)
        didx.append(dis)
        td = None if i < pairing_on else sk
        exp.append(int(O.proof_verify(ocs, pk, p, header, ph, dm, dis, trapdoor_sk=td)))
    got = ctx.proof_verify_batch(proofs, ph, dmsgs, didx)
    assert got.tolist() == exp, (curve_name, L, disclosed, got.tolist(), exp)
    # scalar-level entry point agrees
    dsc = [[ocs.scalar_le(x) for x in O.msg_to_scalars(ocs, dm, ocs.api_id)] for dm in dmsgs]
    got2 = ctx.core_proof_verify_batch(proofs, ph, dsc, didx)
    assert got2.tolist() == exp
    ctx.close()


def case_proof_errors(lib_path, curve_name):
    """Err(...) paths of proof_verify_init (proof_verify.rs:139-150) and the documented panic mapping."""
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    L, dis = 4, [0, 2]
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, b"", L)
    (pr, msgs), = make_proofs(ocs, sk, pk, b"", b"", L, dis, 1)
    pb = A.ProofBytes.from_canonical(suite, O.proof_to_bytes(ocs, pr))
    dm = [msgs[0], msgs[2]]
    zero = A.ProofBytes(bytes(suite.g1_bytes * 3 + 128), b"")          # all-zero bytes: undecodable points
    if curve_name == "BN254":
        # ark's default encoding of the identity is all-zero x with the infinity flag; all-zero bytes decode to x=0
        pass
    cases = [
        ([pb], [dm], [[0, 2]], [1]),
        ([pb], [dm], [[0, 9]], [A.ST_ERR_DISCLOSED_INDEX]),            # index >= L
        ([pb], [[msgs[0]]], [[0, 2]], [A.ST_ERR_IDX_MSG_LEN]),         # fewer messages than indexes
        ([pb], [dm + [b"x"]], [[0, 2, 3]], [A.ST_ERR_MSG_GEN_LEN]),    # R + U != L
        ([pb], [[msgs[0], msgs[0]]], [[0, 0]], [A.ST_ERR_MALFORMED]),  # duplicate index: the reference panics
    ]
    for proofs, m, idx, want in cases:
        got = ctx.proof_verify_batch(proofs, b"", m, idx)
        assert got.tolist() == want, (idx, got.tolist(), want)
    # Proof::default()-style forged input: identity points, zero scalars, 2 commitments (proof_verify_tests.rs:160-184)
    ident = ocs.g1_compress(None)
    forged = A.ProofBytes(ident * 3 + bytes(128), bytes(64))
    got = ctx.proof_verify_batch([forged], b"", [dm], [[0, 2]])
    p0 = O.Proof(None, None, None, 0, 0, 0, [0, 0], 0)
    want = int(O.proof_verify(ocs, pk, p0, b"", b"", dm, [0, 2]))
    assert got.tolist() == [want]
    ctx.close()


def rlc_coeff(seed: bytes, i: int) -> int:
    """coefficient definition of include/bbs_b200.h (random-linear-combination mode)"""
    import hashlib
    r = int.from_bytes(hashlib.sha256(seed + i.to_bytes(8, "big")).digest()[:16], "big")
    return r or 1


def case_rlc(lib_path, curve_name, L=3, n=9, windows=0):
    """Random-linear-combination batch mode against the oracle: the two partial sums of a shard are compared bit-exactly
    with sum r_i A_i and sum r_i (e_i A_i - B_i) computed by the oracle, and the verdicts with the per-item truth."""
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    header = b"rlc"
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, header, L)
    if windows:
        ctx.set_rlc_windows(windows)
    msgs = [[rng_bytes(f"rlc{i}.{j}", 32) for j in range(L)] for i in range(n)]
    sigs = [O.sign(ocs, sk, m, header) for m in msgs]
    seed = rng_bytes("rlc-seed", 32)
    F1 = ocs.F1
    dom = O.calculate_domain(ocs, pk, gens[0], gens[1:], header, ocs.api_id)

    def expected_parts(sig_list, base):
        S1, S2 = None, None
        for i, (sg, m) in enumerate(zip(sig_list, msgs)):
            A, e = sg
            r = rlc_coeff(seed, base + i)
            B = O.compute_B(ocs, gens, dom, O.msg_to_scalars(ocs, m, ocs.api_id))
            if A is not None:
                S1 = O.ec_add(F1, S1, O.ec_mul(F1, A, r))
                S2 = O.ec_add(F1, S2, O.ec_mul(F1, A, r * e % ocs.r))
            S2 = O.ec_add(F1, S2, O.ec_neg(F1, O.ec_mul(F1, B, r)))
        return ocs.g1_compress(S1) + ocs.g1_compress(S2)

    enc = [O.signature_to_bytes(ocs, s) for s in sigs]
    parts, st = ctx.rlc_partial(enc, msgs, seed, 5)
    assert st == A.ST_ACCEPT
    assert parts == expected_parts(sigs, 5), "partial sums differ from the oracle"
    assert ctx.rlc_verify_batch(enc, msgs, seed) == A.ST_ACCEPT
    assert ctx.rlc_verify_batch(enc, msgs) == A.ST_ACCEPT           # seed drawn inside the library (the safe default)
    # two shards with consistent global indexes combine to the same verdict
    p0, _ = ctx.rlc_partial(enc[:4], msgs[:4], seed, 0)
    p1, _ = ctx.rlc_partial(enc[4:], msgs[4:], seed, 4)
    assert ctx.rlc_combine([p0, p1]) == A.ST_ACCEPT
    # one bad signature anywhere -> REJECT; identity A -> REJECT; malformed -> ERR
    bad = list(sigs)
    bad[3] = (bad[3][0], (bad[3][1] + 1) % ocs.r)
    assert ctx.rlc_verify_batch([O.signature_to_bytes(ocs, s) for s in bad], msgs, seed) == A.ST_REJECT
    bad = list(sigs)
    bad[7] = (None, bad[7][1])
    encb = [O.signature_to_bytes(ocs, s) for s in bad]
    assert ctx.rlc_verify_batch(encb, msgs, seed) == A.ST_REJECT
    assert ctx.rlc_verify_batch(encb, msgs) == A.ST_REJECT
    assert ctx.rlc_partial(encb, msgs, seed, 0)[0] == expected_parts(bad, 0)
    if L > 0:
        m2 = [list(m) for m in msgs]
        m2[0][0] = m2[0][0] + b"x"
        assert ctx.rlc_verify_batch(enc, m2, seed) == A.ST_REJECT
    mal = list(enc)
    mal[2] = b"\xff" * len(mal[2])
    assert ctx.rlc_verify_batch(mal, msgs, seed) == A.ST_ERR_MALFORMED
    ctx.close()


def case_proof_gen(lib_path, curve_name, L, disclosed, n=5, header=b"hg", ph=b"ph-gen"):
    """core_proof_gen / proof_gen (proof_gen.rs:78-365) with caller-supplied random scalars: byte-identical proofs
    to the oracle's, which the library then verifies; plus the Err(..) classes of proof_gen.rs:139-147, :228-234."""
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, header, L)
    U = L - len(set(disclosed))
    msgs = [[rng_bytes(f"g{i}.{j}", 32) for j in range(L)] for i in range(n)]
    sigs = [O.sign(ocs, sk, m, header) for m in msgs]
    if n > 2:
        sigs[2] = (None, sigs[2][1])              # identity A: the reference still produces a (useless) proof
    rss = [O.seeded_random_scalars(ocs, f"gen{i}".encode(), b"rs-dst", 5 + U) for i in range(n)]
    enc = [O.signature_to_bytes(ocs, s) for s in sigs]
    got, st = ctx.proof_gen_batch(enc, msgs, [list(disclosed)] * n, [[ocs.scalar_le(x) for x in rs] for rs in rss], ph)
    assert st.tolist() == [1] * n, st.tolist()
    dis = sorted(set(disclosed))
    for i in range(n):
        want = O.proof_gen(ocs, pk, sigs[i], header, ph, msgs[i], list(disclosed), random_scalars=rss[i])
        wb = A.ProofBytes.from_canonical(suite, O.proof_to_bytes(ocs, want))
        assert got[i].fixed == wb.fixed, (curve_name, i, "fixed part differs")
        assert got[i].commitments == wb.commitments, (curve_name, i, "commitments differ")
    ver = ctx.proof_verify_batch(got, ph, [[m[j] for j in dis] for m in msgs], [dis] * n)
    assert ver.tolist() == [0 if (i == 2 and n > 2) else 1 for i in range(n)], ver.tolist()
    # error classes
    one = [enc[0]]
    rs_ok = [[ocs.scalar_le(x) for x in rss[0]]]
    if L >= 1:
        _, st = ctx.proof_gen_batch(one, [msgs[0]], [[L]], [[ocs.scalar_le(1)] * (5 + L - 1)], ph)
        assert st.tolist() == [A.ST_ERR_DISCLOSED_INDEX]
        _, st = ctx.proof_gen_batch(one, [msgs[0]], [list(range(L)) + [0]], [[ocs.scalar_le(1)] * 5], ph)
        assert st.tolist() == [A.ST_ERR_DISCLOSED_LEN]
    _, st = ctx.proof_gen_batch(one, [msgs[0]], [list(disclosed)], [rs_ok[0] + [ocs.scalar_le(7)]], ph)
    assert st.tolist() == [A.ST_ERR_RANDOM_LEN]
    if L >= 2:
        # a duplicated disclosed index: the reference de-duplicates the set but sizes the random scalars by the raw
        # list, so proof_init refuses (proof_gen.rs:232-234)
        _, st = ctx.proof_gen_batch(one, [msgs[0]], [[0, 0]], [[ocs.scalar_le(3)] * (5 + L - 2)], ph)
        assert st.tolist() == [A.ST_ERR_RANDOM_LEN]
    ctx.close()


def case_h2s_ragged(lib_path, curve_name):
    """msg_to_scalars (interface_utilities.rs:76-88) over ragged message lengths around every SHA-256 padding boundary
    of expand_message_xmd's first block chain, and an empty batch."""
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, b"", 1)
    lens = [0, 1, 2, 31, 32, 33, 54, 55, 56, 57, 63, 64, 65, 118, 119, 120, 121, 127, 128, 129, 255, 256, 1000, 4099]
    msgs = [rng_bytes(f"ragged{n}", n) for n in lens]
    got = ctx.msg_to_scalars(msgs)
    want = O.msg_to_scalars(ocs, msgs, ocs.api_id)
    for i, n in enumerate(lens):
        assert got[i].tobytes() == ocs.scalar_le(want[i]), (curve_name, n)
    assert len(ctx.msg_to_scalars([])) == 0
    assert ctx.verify_batch(b"", []).tolist() == []
    ctx.close()


def case_readme_example(lib_path, curve_name):
    """The reference's README example (README.md:64-130, BASELINE configs[0]): key_material = [5; 32], four messages,
    sign -> verify -> proof_gen with {0, 2} disclosed -> proof_verify, as batches of one through the library, every
    byte compared with the oracle."""
    suite, ocs = SUITES[curve_name]
    sk = O.key_gen(ocs, bytes([5] * 32), b"", b"BBS-SIG-KEYGEN-SALT-")
    pk = O.sk_to_pk(ocs, sk)
    msgs = [b"message1", b"message2", b"msg3", b"msg4"]
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, b"", len(msgs))
    sigs, _, st = ctx.sign_batch(ocs.scalar_le(sk), [msgs])
    assert st.tolist() == [1]
    want_sig = O.sign(ocs, sk, msgs, b"")
    assert sigs[0].tobytes() == O.signature_to_bytes(ocs, want_sig)
    assert ctx.verify_batch(sigs, [msgs]).tolist() == [1]
    rs = O.seeded_random_scalars(ocs, b"readme", b"rs-dst", 5 + 2)
    got, st = ctx.proof_gen_batch([sigs[0].tobytes()], [msgs], [[0, 2]], [[ocs.scalar_le(x) for x in rs]], b"")
    assert st.tolist() == [1]
    want = A.ProofBytes.from_canonical(suite, O.proof_to_bytes(ocs, O.proof_gen(ocs, pk, want_sig, b"", b"", msgs, [0, 2], random_scalars=rs)))
    assert got[0].fixed == want.fixed and got[0].commitments == want.commitments
    assert ctx.proof_verify_batch(got, b"", [[msgs[0], msgs[2]]], [[0, 2]]).tolist() == [1]
    assert ctx.proof_verify_batch(got, b"", [[msgs[0], msgs[3]]], [[0, 2]]).tolist() == [0]
    ctx.close()


# ---------------------------------------------------------------------------------------------------
def case_create_generators(lib_path, curve_name, count=5):
    """create_generators on the device (interface_utilities.rs:47-73, hash-to-G1 :24-44) against the oracle: the
    suite's own api_id (for BLS12-381 these are the IRTF generators of test_vector.rs:124-136), a foreign api_id, and
    the fixture tests/golden/generators_<suite>.bin (129 generators made by the oracle, tools/gen_generators.py)."""
    import os
    suite, ocs = SUITES[curve_name]
    want = gens_bytes(ocs, O.create_generators_cached(ocs, count, ocs.api_id))
    got = suite.create_generators(count, lib_path=lib_path)
    assert got == want
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"generators_{suite.name.lower()}.bin"), "rb") as f:
        fixture = f.read()
    assert fixture[: len(want)] == want
    if lib_path is None:                    # GPU: all 129 fixture generators
        assert suite.create_generators(129) == fixture
    if curve_name == "BLS12_381":
        assert got[:48].hex() == ("a9ec65b70a7fbe40c874c9eb041c2cb0a7af36ccec1bea48fa2ba4c2eb67ef7f"
                                  "9ecb17ed27d38d27cdeddff44c8137be")       # Q1, test_vector.rs:132
    other = b"OTHER_API_ID_" + ocs.api_id[:7]
    want2 = gens_bytes(ocs, O.create_generators(ocs, 3, other))
    assert suite.create_generators(3, api_id=other, lib_path=lib_path) == want2
    assert suite.derive_generators(0, lib_path=lib_path) == b""


# ---------------------------------------------------------------------------------------------------
def case_core_api_id(lib_path, curve_name, L=5):
    """core_sign_tests.rs: core_sign / core_verify over SCALAR messages with an arbitrary api_id ("" and "abc",
    :38-65), generators made for that api_id (on the device: bbs_create_generators), and its rejections: wrong api_id,
    wrong header, message[0] = 0 (:67-155).  Signatures and B points are compared byte for byte with the oracle."""
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    n = 4
    for api_id in (b"", b"abc"):
        gens = O.create_generators(ocs, L + 1, api_id)
        gb = gens_bytes(ocs, gens)
        if lib_path is None:          # GPU run: the device derivation must give the same generators
            assert suite.derive_generators(L + 1, api_id=api_id) == gb
        ctx = A.BatchContext(suite, ocs.g2_compress(pk), b"h", generators=gb, api_id=api_id, lib_path=lib_path)
        scal = [[(7 * i + j + 1) % ocs.r for j in range(L)] for i in range(n)]        # Fr::from(u64) style messages
        flat = np.frombuffer(b"".join(ocs.scalar_le(x) for row in scal for x in row), dtype=np.uint8)
        sigs, _, st = ctx.core_sign_batch(sk.to_bytes(32, "little"), flat, n, L)
        assert st.tolist() == [1] * n
        osigs = [O.core_sign(ocs, sk, gens, b"h", row, api_id) for row in scal]
        for i in range(n):
            assert sigs[i].tobytes() == O.signature_to_bytes(ocs, osigs[i])
        blob = np.frombuffer(b"".join(O.signature_to_bytes(ocs, s) for s in osigs), dtype=np.uint8)
        assert ctx.core_verify_batch(blob, flat, L).tolist() == [1] * n
        # message[0] = 0 on item 2
        bad = [list(r) for r in scal]
        bad[2][0] = 0
        flat_bad = np.frombuffer(b"".join(ocs.scalar_le(x) for row in bad for x in row), dtype=np.uint8)
        assert ctx.core_verify_batch(blob, flat_bad, L).tolist() == [1, 1, 0, 1]
        ctx.close()
        # wrong api_id (same generators) and wrong header: the domain changes -> Ok(false)
        for other_api, other_hdr in ((api_id + b"x", b"h"), (api_id, b"other")):
            ctx2 = A.BatchContext(suite, ocs.g2_compress(pk), other_hdr, generators=gb, api_id=other_api, lib_path=lib_path)
            want = [int(O.core_verify(ocs, pk, s, gens, other_hdr, row, other_api, trapdoor_sk=sk)) for s, row in zip(osigs, scal)]
            assert ctx2.core_verify_batch(blob, flat, L).tolist() == want == [0] * n
            ctx2.close()


# ---------------------------------------------------------------------------------------------------
def case_g1_mul_edges(lib, curve_name):
    """GLV split edge scalars through the production scalar multiplication (g1.cuh bls_glv_split / bn_glv_split): values
    around the endomorphism eigenvalue, powers of two around 2^128, the BN254 lattice constants, r - small."""
    suite, ocs = SUITES[curve_name]
    r = ocs.r
    if curve_name == "BLS12_381":
        x = 0xD201000000010000
        lam = x * x - 1
        special = [lam, lam - 1, lam + 1, lam * lam % r, r - lam, 2 * lam, lam * (lam - 1) % r]
    else:
        t = 4965661367192848881
        lam = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
        a, b = 6 * t * t + 2 * t, 2 * t + 1
        special = [lam, lam - 1, (lam + 1) % r, r - lam, a, b, a + b, a * b % r, 2 * b * b, r - a, r - b, (a * lam) % r]
    special += [(1 << 128) - 1, 1 << 128, (1 << 128) + 1, 1 << 127, (1 << 64) - 1, r - 1, r - 2, r // 2, r // 3]
    P = O.ec_mul(ocs.F1, ocs.BP1, 0x1234567)
    n = len(special)
    a_ = np.frombuffer(ocs.g1_compress(P) * n, dtype=np.uint8)
    b_ = np.frombuffer(b"".join(k.to_bytes(32, "little") for k in special), dtype=np.uint8)
    out = np.zeros(n * suite.g1_bytes, dtype=np.uint8)
    rc = lib.bbs_selftest_g1_mul(suite.curve_id, 0, n, ptr(a_), ptr(b_), ptr(out))
    assert rc == 0, lib.bbs_last_error()
    for i, k in enumerate(special):
        want = ocs.g1_compress(O.ec_mul(ocs.F1, P, k))
        assert out[i * suite.g1_bytes:(i + 1) * suite.g1_bytes].tobytes() == want, (curve_name, hex(k))


# ---------------------------------------------------------------------------------------------------
def off_subgroup_g1(ocs, seed=0):
    """an on-curve point of E(Fp) outside the prime-order subgroup (no cofactor clearing): small x until on curve"""
    x = 5 + seed
    while True:
        rhs = (x ** 3 + ocs.b) % ocs.p
        y = pow(rhs, (ocs.p + 1) // 4, ocs.p)
        if y * y % ocs.p == rhs and O.ec_mul(ocs.F1, (x, y), ocs.r) is not None:
            return (x, y)
        x += 1


def off_subgroup_g2(ocs, seed=0):
    x = 3 + seed
    F = ocs.F2
    while True:
        X = (x, 1)
        y = F.sqrt(F.add(F.mul(F.mul(X, X), X), ocs.b2))
        if y is not None and O.ec_mul(F, (X, y), ocs.r) is not None:
            return (X, y)
        x += 1


def case_subgroup(lib_path, curve_name):
    """ark-serialize `deserialize_compressed` (Validate::Yes; derived for Signature sign.rs:18, Proof proof_gen.rs:29,
    PublicKey key_gen.rs:12) refuses points outside the prime-order subgroup.  On-curve off-subgroup points (and valid
    points shifted by cofactor torsion, which the pairing cannot see) must give ERR_MALFORMED at every entry point that
    takes G1 bytes, and BBS_E_ARG from bbs_ctx_create for the public key and the generators."""
    suite, ocs = SUITES[curve_name]
    sk, pk = keypair(ocs, 1)
    L = 2
    header, ph = b"sg", b"sg-ph"
    ctx, gens = make_ctx(lib_path, suite, ocs, pk, header, L)
    gb = gens_bytes(ocs, gens)
    # --- public key / generators at context creation
    Qbad = off_subgroup_g2(ocs)
    assert ocs.g2_on_curve(Qbad)
    try:
        ocs.g2_decompress(ocs.g2_compress(Qbad))
        assert False, "oracle accepted an off-subgroup G2 point"
    except ValueError:
        pass
    assert ocs.g2_decompress(ocs.g2_compress(Qbad), validate=False) == Qbad
    for bad_pk in (Qbad, O.ec_add(ocs.F2, pk, Qbad)):
        try:
            A.BatchContext(suite, ocs.g2_compress(bad_pk), header, generators=gb, lib_path=lib_path)
            assert False, "off-subgroup public key accepted"
        except A.BbsError as e:
            assert "(-1)" in str(e), str(e)               # BBS_E_ARG
    if curve_name != "BLS12_381":
        ctx.close()       # BN254: E(Fp) has prime order r, every on-curve G1 point is in the subgroup
        return
    T = off_subgroup_g1(ocs)
    T3 = (0, 2)                                           # a point of order 3 on y^2 = x^3 + 4
    assert ocs.g1_on_curve(T) and ocs.g1_on_curve(T3) and O.ec_mul(ocs.F1, T3, 3) is None
    h = (ocs.p + 1 - (-0xD201000000010000 + 1)) // ocs.r  # cofactor of G1
    Ttor = O.ec_mul(ocs.F1, T, ocs.r)                     # pure cofactor torsion
    assert Ttor is not None and O.ec_mul(ocs.F1, Ttor, h) is None
    for Pb in (T, T3, Ttor):
        try:
            ocs.g1_decompress(ocs.g1_compress(Pb))
            assert False, "oracle accepted an off-subgroup G1 point"
        except ValueError:
            pass
    try:
        A.BatchContext(suite, ocs.g2_compress(pk), header, generators=gb[:48] + ocs.g1_compress(T) + gb[96:], lib_path=lib_path)
        assert False, "off-subgroup generator accepted"
    except A.BbsError as e:
        assert "(-1)" in str(e), str(e)
    # --- signatures: A replaced by torsion-shifted copies (the pairing equation still holds for A + Ttor)
    msgs = [[rng_bytes(f"sg{i}.{j}", 32) for j in range(L)] for i in range(6)]
    sigs = [O.sign(ocs, sk, m, header) for m in msgs]
    mall = [sigs[0],
            (O.ec_add(ocs.F1, sigs[1][0], Ttor), sigs[1][1]),
            (O.ec_add(ocs.F1, sigs[2][0], T3), sigs[2][1]),
            (T, sigs[3][1]),
            (T3, sigs[4][1]),
            sigs[5]]
    # without validation the malleated signature satisfies the verification equation: that is the hole being closed
    assert O.verify(ocs, pk, mall[2], header, msgs[2]) is True
    enc = [O.signature_to_bytes(ocs, s) for s in mall]
    want = [1, A.ST_ERR_MALFORMED, A.ST_ERR_MALFORMED, A.ST_ERR_MALFORMED, A.ST_ERR_MALFORMED, 1]
    blob = np.frombuffer(b"".join(enc), dtype=np.uint8)
    assert ctx.verify_batch(blob, msgs).tolist() == want
    sc = np.frombuffer(b"".join(ocs.scalar_le(x) for m in msgs for x in O.msg_to_scalars(ocs, m, ocs.api_id)), dtype=np.uint8)
    assert ctx.core_verify_batch(blob, sc, L).tolist() == want
    # --- proof_gen from a malleated signature
    rs = [[ocs.scalar_le(x) for x in O.seeded_random_scalars(ocs, f"sg{i}".encode(), b"rs-dst", 5 + L - 1)] for i in range(6)]
    proofs, st = ctx.proof_gen_batch(enc, msgs, [[0]] * 6, rs, ph)
    assert st.tolist() == want
    # --- proof_verify: each of Abar, Bbar, D off the subgroup
    good = proofs[0]
    assert ctx.proof_verify_batch([good], ph, [[msgs[0][0]]], [[0]]).tolist() == [1]
    g = suite.g1_bytes
    bad_proofs = []
    for slot in range(3):
        for Pb in (T, T3):
            f = bytearray(good.fixed)
            f[slot * g:(slot + 1) * g] = ocs.g1_compress(Pb)
            bad_proofs.append(A.ProofBytes(bytes(f), good.commitments))
        P0 = ocs.g1_decompress(good.fixed[slot * g:(slot + 1) * g])
        f = bytearray(good.fixed)
        f[slot * g:(slot + 1) * g] = ocs.g1_compress(O.ec_add(ocs.F1, P0, Ttor))
        bad_proofs.append(A.ProofBytes(bytes(f), good.commitments))
    nb = len(bad_proofs)
    got = ctx.proof_verify_batch(bad_proofs + [good], ph, [[msgs[0][0]]] * (nb + 1), [[0]] * (nb + 1))
    assert got.tolist() == [A.ST_ERR_MALFORMED] * nb + [1], got.tolist()
    # --- random-linear-combination mode (CUDA build only)
    if lib_path is None:
        seed = rng_bytes("sg-seed", 32)
        assert ctx.rlc_verify_batch([enc[0], enc[5]], [msgs[0], msgs[5]], seed) == A.ST_ACCEPT
        assert ctx.rlc_verify_batch(enc, msgs, seed) == A.ST_ERR_MALFORMED
        assert ctx.rlc_verify_batch([enc[0], enc[2]], [msgs[0], msgs[2]], seed) == A.ST_ERR_MALFORMED
    ctx.close()


# ---------------------------------------------------------------------------------------------------
def case_multi_issuer(lib_path, curve_name, n_issuers=5, per_issuer=3, L=2):
    """bbs_issuer_set_create / bbs_verify_batch_multi: items of one batch name different issuer keys (the key is `&self` of
    every call in the reference, verify.rs:18-30).  Each status must equal the oracle's verify under the item's claimed key:
    own-issuer signatures, a signature checked under another issuer's key, a forged e, the identity key, an off-subgroup
    key (refused at creation, its items come back malformed), an out-of-range issuer index."""
    suite, ocs = SUITES[curve_name]
    header = b"multi"
    keys = [keypair(ocs, 10 + k) for k in range(n_issuers)]
    pks = [ocs.g2_compress(pk) for _, pk in keys]
    bad_key = ocs.g2_compress(off_subgroup_g2(ocs))
    ident = ocs.g2_compress(None)
    all_pks = pks + [bad_key, ident]
    gens = O.create_generators_cached(ocs, L + 1, ocs.api_id)
    iset = A.IssuerSet(suite, all_pks, header, generators=gens_bytes(ocs, gens), lib_path=lib_path)
    assert iset.status.tolist() == [1] * n_issuers + [A.ST_ERR_MALFORMED, 1]
    per, shared = iset.memory_bytes()
    assert 0 < per < 64 * 1024 * len(all_pks) and shared > 0
    items = []            # (issuer index, signature, messages, expected)
    for k in range(n_issuers):
        sk, pk = keys[k]
        for t in range(per_issuer):
            m = [rng_bytes(f"mi{k}.{t}.{j}", 32) for j in range(L)]
            sig = O.sign(ocs, sk, m, header)
            claimed = k
            if t == 1 and k % 2 == 0:
                claimed = (k + 1) % n_issuers                         # verified under another issuer's key
            if t == 2 and k % 3 == 0:
                sig = (sig[0], (sig[1] + 1) % ocs.r)                  # forged
            csk, cpk = keys[claimed]
            use_pairing = (k == 0 and t < 2)                          # a few items with the pairing oracle, the rest by trapdoor
            want = int(O.verify(ocs, cpk, sig, header, m, trapdoor_sk=None if use_pairing else csk))
            items.append((claimed, sig, m, want))
    sk0, _ = keys[0]
    m = [rng_bytes(f"mi-x.{j}", 32) for j in range(L)]
    sig = O.sign(ocs, sk0, m, header)
    items.append((n_issuers, sig, m, A.ST_ERR_MALFORMED))                   # the off-subgroup key
    items.append((n_issuers + 1, sig, m, int(O.verify(ocs, None, sig, header, m))))   # identity key: Ok(false)
    items.append((n_issuers + 2, sig, m, A.ST_ERR_MALFORMED))               # no such issuer
    random.Random(3).shuffle(items)
    enc = b"".join(O.signature_to_bytes(ocs, it[1]) for it in items)
    got = iset.verify_batch([it[0] for it in items], enc, [it[2] for it in items])
    assert got.tolist() == [it[3] for it in items], (curve_name, got.tolist(), [it[3] for it in items])
    sc = b"".join(ocs.scalar_le(x) for it in items for x in O.msg_to_scalars(ocs, it[2], ocs.api_id))
    got2 = iset.core_verify_batch([it[0] for it in items], np.frombuffer(enc, dtype=np.uint8), np.frombuffer(sc, dtype=np.uint8), L)
    assert got2.tolist() == got.tolist()
    # per key, the set's domain-dependent state equals a dedicated context's: same verdicts through bbs_ctx_create
    ctx, _ = make_ctx(lib_path, suite, ocs, keys[1][1], header, L, gens=gens)
    mine = [it for it in items if it[0] == 1]
    one = ctx.verify_batch(np.frombuffer(b"".join(O.signature_to_bytes(ocs, it[1]) for it in mine), dtype=np.uint8), [it[2] for it in mine])
    assert one.tolist() == [it[3] for it in mine]
    ctx.close()
    iset.close()


def case_multi_issuer_proofs(lib_path, curve_name, n_issuers=3, L=3, disclosed=(0, 2), pairing_on=1):
    """bbs_proof_verify_batch_multi: proof i is checked under pks[item_issuer[i]] (the key is `&self` of proof_verify too,
    proof_verify.rs:19-34).  Every status equals the oracle's proof_verify under the claimed key: own-issuer proofs, a proof
    checked under another issuer's key (its challenge no longer matches: the domain differs), forged fields, ragged
    disclosure, the identity key, an off-subgroup key, an out-of-range issuer index; the scalar-level entry point agrees."""
    suite, ocs = SUITES[curve_name]
    header, ph = b"multi-proof", b"ph"
    keys = [keypair(ocs, 20 + k) for k in range(n_issuers)]
    all_pks = [ocs.g2_compress(pk) for _, pk in keys] + [ocs.g2_compress(off_subgroup_g2(ocs)), ocs.g2_compress(None)]
    gens = O.create_generators_cached(ocs, L + 1, ocs.api_id)
    iset = A.IssuerSet(suite, all_pks, header, generators=gens_bytes(ocs, gens), lib_path=lib_path)
    assert iset.status.tolist() == [1] * n_issuers + [A.ST_ERR_MALFORMED, 1]
    items = []                              # (claimed issuer, proof, disclosed messages, disclosed indexes, expected)
    kinds = ["ok", "other-issuer", "challenge+1", "ok-none-disclosed", "wrong-msg", "Abar=identity", "ok-all-disclosed"]
    n_pair = 0
    for k in range(n_issuers):
        sk, pk = keys[k]
        for t, kind in enumerate(kinds):
            dis = sorted(set(disclosed))
            if kind == "ok-none-disclosed":
                dis = []
            elif kind == "ok-all-disclosed":
                dis = list(range(L))
            (pr, msgs), = make_proofs(ocs, sk, pk, header, ph, L, dis, 1, seed=f"mp{k}.{t}.")
            p = O.Proof(pr.a_bar, pr.b_bar, pr.d, pr.e_cap, pr.r1_cap, pr.r3_cap, list(pr.commitments), pr.challenge)
            dm = [msgs[j] for j in dis]
            claimed = k
            if kind == "other-issuer":
                claimed = (k + 1) % n_issuers
            elif kind == "challenge+1":
                p.challenge = (p.challenge + 1) % ocs.r
            elif kind == "wrong-msg":
                dm[0] = dm[0] + b"!"
            elif kind == "Abar=identity":
                p.a_bar = None
            csk, cpk = keys[claimed]
            td = csk
            if kind == "ok" and n_pair < pairing_on:
                td, n_pair = None, n_pair + 1                          # a few through the oracle's pairing, the rest by trapdoor
            want = int(O.proof_verify(ocs, cpk, p, header, ph, dm, dis, trapdoor_sk=td))
            items.append((claimed, p, dm, dis, want))
    (pr, msgs), = make_proofs(ocs, keys[0][0], keys[0][1], header, ph, L, sorted(set(disclosed)), 1, seed="mpx")
    dis = sorted(set(disclosed))
    dm = [msgs[j] for j in dis]
    items.append((n_issuers, pr, dm, dis, A.ST_ERR_MALFORMED))                                  # the off-subgroup key
    items.append((n_issuers + 1, pr, dm, dis, int(O.proof_verify(ocs, None, pr, header, ph, dm, dis))))   # identity key
    items.append((n_issuers + 2, pr, dm, dis, A.ST_ERR_MALFORMED))                              # no such issuer
    items.append((0, pr, dm[:2], [0, L], A.ST_ERR_DISCLOSED_INDEX))                             # the checks come first
    random.Random(5).shuffle(items)
    proofs = [A.ProofBytes.from_canonical(suite, O.proof_to_bytes(ocs, it[1])) for it in items]
    got = iset.proof_verify_batch([it[0] for it in items], proofs, ph, [it[2] for it in items], [it[3] for it in items])
    exp = [it[4] for it in items]
    assert got.tolist() == exp, (curve_name, got.tolist(), exp)
    assert exp.count(1) >= 3 * n_issuers - 1
    dsc = [[ocs.scalar_le(x) for x in O.msg_to_scalars(ocs, it[2], ocs.api_id)] for it in items]
    got2 = iset.core_proof_verify_batch([it[0] for it in items], proofs, ph, dsc, [it[3] for it in items])
    assert got2.tolist() == exp
    # the same proofs of issuer 1 through a dedicated context give the same verdicts
    ctx, _ = make_ctx(lib_path, suite, ocs, keys[1][1], header, L, gens=gens)
    mine = [k for k, it in enumerate(items) if it[0] == 1]
    one = ctx.proof_verify_batch([proofs[k] for k in mine], ph, [items[k][2] for k in mine], [items[k][3] for k in mine])
    assert one.tolist() == [exp[k] for k in mine]
    ctx.close()
    iset.close()


def case_split_differential(lib_path, curve_name, n=48, seed=1):
    """The G1 halves exist twice in the library: one thread per item (verify_g1_item / proof_g1_item) and as independent
    tasks plus a join (verify_task_* / proof_task_*), where the statuses of an item are owned by different tasks.  Random
    byte damage anywhere in the inputs (point encodings: flags, not on the curve, off the subgroup, identity; scalars:
    non-canonical; commitments; disclosed scalars; disclosed indexes out of range or repeated) must give the SAME status
    vector through both, and the undamaged items the oracle's verdict."""
    suite, ocs = SUITES[curve_name]
    rnd = random.Random(seed)
    sk, pk = keypair(ocs, 3)
    L, header, ph = 3, b"diff", b"p"
    gens = O.create_generators_cached(ocs, L + 1, ocs.api_id)

    def ctx_with(split):
        c = A.BatchContext(suite, ocs.g2_compress(pk), header, generators=gens_bytes(ocs, gens), lib_path=lib_path)
        c.set_g1_split(split)
        return c

    def damage(b: bytes, how: int) -> bytes:
        b = bytearray(b)
        if how == 1:                                  # one random bit anywhere
            i = rnd.randrange(len(b)); b[i] ^= 1 << rnd.randrange(8)
        elif how == 2:                                # a whole field set to 0xff (non-canonical scalar / x >= p)
            f = rnd.randrange(len(b) // 16); b[16 * f: 16 * f + 16] = b"\xff" * 16
        elif how == 3:                                # flag bits of the first byte
            b[0] ^= rnd.choice([0x80, 0x40, 0x20])
        elif how == 4:                                # a run of zero bytes
            i = rnd.randrange(len(b) - 8); b[i: i + 8] = bytes(8)
        return bytes(b)

    # ---- signatures ----
    msgs = [[rng_bytes(f"d{seed}.{i}.{j}", 16) for j in range(L)] for i in range(n)]
    sigs = [O.signature_to_bytes(ocs, O.sign(ocs, sk, m, header)) for m in msgs]
    scal = [b"".join(ocs.scalar_le(x) for x in O.msg_to_scalars(ocs, m, ocs.api_id)) for m in msgs]
    clean = []
    for i in range(n):
        how = rnd.randrange(6)                        # 0 and 5: untouched
        if how in (1, 2, 3, 4):
            if rnd.random() < 0.6:
                sigs[i] = damage(sigs[i], how)
            else:
                scal[i] = damage(scal[i], how)
        else:
            clean.append(i)
    sg = np.frombuffer(b"".join(sigs), dtype=np.uint8)
    sc = np.frombuffer(b"".join(scal), dtype=np.uint8)
    res = []
    for split in (0, 1 << 62):
        c = ctx_with(split)
        res.append(c.core_verify_batch(sg, sc, L).tolist())
        c.close()
    assert res[0] == res[1], (curve_name, "verify", res)
    assert all(res[0][i] == 1 for i in clean) and len(set(res[0])) >= 2
    # ---- proofs ----
    base = make_proofs(ocs, sk, pk, header, ph, L, [0, 2], n, seed=f"dp{seed}.")
    proofs, dsc, didx, clean = [], [], [], []
    for i, (pr, m) in enumerate(base):
        pb = A.ProofBytes.from_canonical(suite, O.proof_to_bytes(ocs, pr))
        fixed, commit = pb.fixed, pb.commitments
        d = [ocs.scalar_le(x) for x in O.msg_to_scalars(ocs, [m[0], m[2]], ocs.api_id)]
        idx = [0, 2]
        how = rnd.randrange(8)
        if how in (1, 2, 3, 4):
            r = rnd.random()
            if r < 0.5:
                fixed = damage(fixed, how)
            elif r < 0.75:
                commit = damage(commit, how)
            else:
                d[rnd.randrange(2)] = damage(d[0], how)
        elif how == 5:
            idx = [0, L + rnd.randrange(3)]            # out of range
        elif how == 6:
            idx = [2, 2]                              # repeated
        else:
            clean.append(i)
        proofs.append(A.ProofBytes(fixed, commit)); dsc.append(d); didx.append(idx)
    res = []
    for split in (0, 1 << 62):
        c = ctx_with(split)
        res.append(c.core_proof_verify_batch(proofs, ph, dsc, didx).tolist())
        c.close()
    assert res[0] == res[1], (curve_name, "proof", res)
    assert all(res[0][i] == 1 for i in clean) and len(set(res[0])) >= 3
