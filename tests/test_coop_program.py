"""CPU check of the cooperative pairing PROGRAM (tools/coop_prog.py): the per-role instruction streams the CUDA
interpreter executes are emulated on Python integers (same Montgomery / accumulator / REDC / canonicalisation
arithmetic, with range assertions and a barrier-hazard checker) and their accept / reject verdict is compared
with the oracle's independent pairing (flat Fp12, plain ate, square-and-multiply final exponentiation) on the
pairing equation of core_verify (verify.rs:88-92) in the rewritten form e(A, W) e(eA - B, BP2) == 1."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import bbs_oracle as O  # noqa: E402
import coop_prog as CP  # noqa: E402
import coop_consts as CC  # noqa: E402


def build_bls():
    ids, vals = CC.const_table(CC.BLS)
    prog = CP.PairingProgram(CC.BLS, abs(CC.X_BLS), ids, CC.P_BLS - 2)
    return prog, prog.build(), ids, vals


@pytest.fixture(scope="module")
def bls_program():
    return build_bls()


def run_pairs(built, P0, Q0, P1, Q1):
    prog, streams, ids, vals = built
    cs = O.BLS12_381
    skip = (P0 is None or Q0 is None, P1 is None or Q1 is None)
    l0 = CC.line_table(CC.BLS, Q0 if Q0 is not None else cs.BP2, abs(CC.X_BLS))
    l1 = CC.line_table(CC.BLS, Q1 if Q1 is not None else cs.BP2, abs(CC.X_BLS))
    lines = [[a[0], a[1], b[0], b[1]] for a, b in zip(l0, l1)]
    em = CP.Emulator(CC.BLS, streams, vals, lines)
    for k in range(6):
        em.set_cell(prog.SLOTS[0] + k, (1, 0) if k == 0 else (0, 0))
    em.set_cell(prog.P_CELLS[0], P0 if P0 is not None else (0, 0))
    em.set_cell(prog.P_CELLS[1], P1 if P1 is not None else (0, 0))
    return em.run(skip=skip), em


def test_program_matches_oracle_pairing(bls_program):
    cs = O.BLS12_381
    sk = O.key_gen(cs, b"coop-program-test-key-material-32", b"", b"BBS-SIG-KEYGEN-SALT-")
    pk = O.sk_to_pk(cs, sk)
    msgs = [b"m1", b"m2", b"m3"]
    sig = O.sign(cs, sk, msgs, b"hdr")
    A, e = sig
    gens = O.create_generators(cs, len(msgs) + 1, cs.api_id)
    scal = O.msg_to_scalars(cs, msgs, cs.api_id)
    dom = O.calculate_domain(cs, pk, gens[0], gens[1:], b"hdr", cs.api_id)
    B = O.compute_B(cs, gens, dom, scal)
    F1 = cs.F1
    C = O.ec_add(F1, O.ec_mul(F1, A, e), O.ec_neg(F1, B))
    ok, em = run_pairs(bls_program, A, pk, C, cs.BP2)
    assert ok is True
    assert O.pairing_product_is_one(cs, [(A, pk), (C, cs.BP2)])
    # wrong e
    C2 = O.ec_add(F1, O.ec_mul(F1, A, e + 1), O.ec_neg(F1, B))
    ok2, _ = run_pairs(bls_program, A, pk, C2, cs.BP2)
    assert ok2 is False
    # identity A: pair 0 skipped, e(C, BP2) != 1
    ok3, _ = run_pairs(bls_program, None, pk, O.ec_neg(F1, B), cs.BP2)
    assert ok3 is False
    # both skipped: empty product == 1
    ok4, _ = run_pairs(bls_program, None, pk, None, cs.BP2)
    assert ok4 is True
    assert em.max_out < 16


def test_program_work_statistics(bls_program):
    prog = bls_program[0]
    w = prog.b.work
    # every role executes the same number of barriers by construction; the EP counts are the kernel's work model
    assert max(w["ep"]) < 9000 and w["bar"] < 1200


# ---- BN254: xi = 9 + u applied to the accumulators (XI), wide reduction, D-type twist, Fuentes-Castaneda hard part ----
def build_bn():
    ids, vals = CC.const_table(CC.BN)
    prog = CP.PairingProgram(CC.BN, CC.T_BN, ids, CC.P_BN - 2, CC.BN_MILLER_DIGITS, CC.BN_MILLER_TAIL)
    return prog, prog.build(), ids, vals


@pytest.fixture(scope="module")
def bn_program():
    return build_bn()


def run_pairs_bn(built, P0, Q0, P1, Q1):
    prog, streams, ids, vals = built
    cs = O.BN254
    p = CC.P_BN
    skip = (P0 is None or Q0 is None, P1 is None or Q1 is None)
    l0 = CC.line_table_bn(CC.BN, Q0 if Q0 is not None else cs.BP2)
    l1 = CC.line_table_bn(CC.BN, Q1 if Q1 is not None else cs.BP2)
    lines = [[a[0], a[1], b[0], b[1]] for a, b in zip(l0, l1)]
    em = CP.Emulator(CC.BN, streams, vals, lines)
    for k in range(6):
        em.set_cell(prog.SLOTS[0] + k, (1, 0) if k == 0 else (0, 0))

    def ratio(P):          # the kernel prologue's (x / y, 1 / y)
        if P is None:
            return (0, 0)
        yi = pow(P[1], -1, p)
        return (P[0] * yi % p, yi)

    em.set_cell(prog.P_CELLS[0], ratio(P0))
    em.set_cell(prog.P_CELLS[1], ratio(P1))
    return em.run(skip=skip), em


def test_bn_program_matches_oracle_pairing(bn_program):
    cs = O.BN254
    sk = O.key_gen(cs, b"coop-program-test-key-material-32", b"", b"BBS-SIG-KEYGEN-SALT-")
    pk = O.sk_to_pk(cs, sk)
    msgs = [b"m1", b"m2", b"m3"]
    A, e = O.sign(cs, sk, msgs, b"hdr")
    gens = O.create_generators(cs, len(msgs) + 1, cs.api_id)
    scal = O.msg_to_scalars(cs, msgs, cs.api_id)
    dom = O.calculate_domain(cs, pk, gens[0], gens[1:], b"hdr", cs.api_id)
    B = O.compute_B(cs, gens, dom, scal)
    F1 = cs.F1
    C = O.ec_add(F1, O.ec_mul(F1, A, e), O.ec_neg(F1, B))
    ok, em = run_pairs_bn(bn_program, A, pk, C, cs.BP2)
    assert ok is True
    assert O.pairing_product_is_one(cs, [(A, pk), (C, cs.BP2)])
    C2 = O.ec_add(F1, O.ec_mul(F1, A, e + 1), O.ec_neg(F1, B))
    assert run_pairs_bn(bn_program, A, pk, C2, cs.BP2)[0] is False
    assert run_pairs_bn(bn_program, None, pk, O.ec_neg(F1, B), cs.BP2)[0] is False
    assert run_pairs_bn(bn_program, None, pk, None, cs.BP2)[0] is True
    assert em.max_out < 128
    # the number of lines the program consumes equals the table the context builds (pairing.cuh ate_line_count<Bn>)
    assert len(CC.line_table_bn(CC.BN, cs.BP2)) == 65 + 2 + sum(1 for d in CC.BN_NAF[:-1] if d)


def test_bls_program_is_the_committed_one(bls_program):
    """the generalisation for BN254 must not change a single word of the BLS12-381 program the kernel ships"""
    import re
    txt = open(os.path.join(ROOT, "bbs_sign_b200", "csrc", "gen_pairing_prog.cuh")).read()
    for tag, built in (("BLS", bls_program),):
        body = txt[txt.index(f"COOP_PROG_{tag}["):txt.index(f"COOP_PROG_OFF_{tag}")]
        words = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", body)]
        assert words == [w for st in built[1] for w in st]
