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
