#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): one verify batch of 3 items (pairing kernel
with two warps per role), one of 40 (one warp per role, two groups per block), a proof batch and an RLC verdict, all checked
against the construction.   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as P  # noqa: E402

which = sys.argv[1:] or ["verify3", "verify40", "proof", "rlc"]
if "verify3" in which:
    P.case_verify(None, "BLS12_381", 2, n=3, use_pairing_oracle_on=0)
if "verify40" in which:
    P.case_verify(None, "BLS12_381", 2, n=40, use_pairing_oracle_on=0)
if "proof" in which:
    P.case_proof_verify(None, "BLS12_381", 3, [0, 2], n=5, pairing_on=0)
if "rlc" in which:
    P.case_rlc(None, "BLS12_381")
if "bn" in which:
    P.case_verify(None, "BN254", 2, n=3, use_pairing_oracle_on=0)
    P.case_verify(None, "BN254", 2, n=40, use_pairing_oracle_on=0)
print("SANITIZE_RUN_OK", which)
