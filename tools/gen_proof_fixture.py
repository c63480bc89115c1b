#!/usr/bin/env python3
"""Generates tests/golden/proofs_bls_L32_R16.npz: BLS12-381 selective-disclosure proofs in the shape of BASELINE
config 4 (L = 32 messages of 32 bytes, the 16 even indexes disclosed, empty header and presentation header, IRTF
key pair), made by the oracle's proof_gen with seeded random scalars, plus corrupted variants with the oracle's own
verdict for each.  bench.py --workload proof tiles this fixture (it may not import the oracle) and
tests/test_gpu_parity.py checks the CUDA path against the recorded verdicts."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import bbs_oracle as O  # noqa: E402

IRTF_SK = 0x60e55110f76883a13d030b2f6bd11883422d5abde717569fc0731f51237169fc
L, DIS, N_BASE = 32, list(range(0, 32, 2)), 12


def main():
    cs = O.BLS12_381
    sk = IRTF_SK
    pk = O.sk_to_pk(cs, sk)
    g = 48
    fixed, commits, msgs_out, expect, kinds = [], [], [], [], []
    for i in range(N_BASE):
        import hashlib
        msgs = [hashlib.sha256(f"proof-fixture-{i}-{j}".encode()).digest() for j in range(L)]
        sig = O.sign(cs, sk, msgs, b"")
        rs = O.seeded_random_scalars(cs, f"fixture{i}".encode(), b"rs-dst", 5 + L - len(DIS))
        pr = O.proof_gen(cs, pk, sig, b"", b"", msgs, DIS, random_scalars=rs)
        for kind in ("ok", "challenge+1", "e_cap+1", "swap-commit") if i < 4 else ("ok",):
            p = O.Proof(pr.a_bar, pr.b_bar, pr.d, pr.e_cap, pr.r1_cap, pr.r3_cap, list(pr.commitments), pr.challenge)
            if kind == "challenge+1":
                p.challenge = (p.challenge + 1) % cs.r
            elif kind == "e_cap+1":
                p.e_cap = (p.e_cap + 1) % cs.r
            elif kind == "swap-commit":
                p.commitments[0], p.commitments[1] = p.commitments[1], p.commitments[0]
            dm = [msgs[j] for j in DIS]
            ok = O.proof_verify(cs, pk, p, b"", b"", dm, DIS, trapdoor_sk=None if kind == "ok" and i == 0 else sk)
            blob = O.proof_to_bytes(cs, p)
            head = blob[: 3 * g + 96]
            u = int.from_bytes(blob[3 * g + 96: 3 * g + 104], "little")
            assert u == L - len(DIS)
            fixed.append(head + blob[3 * g + 104 + 32 * u: 3 * g + 136 + 32 * u])
            commits.append(blob[3 * g + 104: 3 * g + 104 + 32 * u])
            msgs_out.append(b"".join(dm))
            expect.append(int(ok))
            kinds.append(kind)
        print(i, "done", flush=True)
    out = os.path.join(ROOT, "tests", "golden", "proofs_bls_L32_R16.npz")
    np.savez_compressed(out, pk=np.frombuffer(cs.g2_compress(pk), dtype=np.uint8),
                        fixed=np.frombuffer(b"".join(fixed), dtype=np.uint8).reshape(len(fixed), -1),
                        commitments=np.frombuffer(b"".join(commits), dtype=np.uint8).reshape(len(fixed), -1),
                        disclosed_msgs=np.frombuffer(b"".join(msgs_out), dtype=np.uint8).reshape(len(fixed), -1),
                        disclosed_idx=np.array(DIS, dtype=np.uint32), expect=np.array(expect, dtype=np.uint8),
                        kinds=np.array(kinds))
    print("wrote", out, len(fixed), "proofs", expect)


if __name__ == "__main__":
    main()
