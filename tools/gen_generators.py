#!/usr/bin/env python3
"""Writes the test fixture tests/golden/generators_<suite>.bin: the first 129 message generators (Q1, H_1..H_128)
of each ciphersuite, compressed, in create_generators order (interface_utilities.rs:47-73).
They are constants of the ciphersuite (the IRTF draft lists them as fixtures); computed here once with
the big-int oracle (itself checked against the KATs of test_vector.rs:124-136 by tests/test_oracle_kat.py).  The package
derives its generators on the GPU (bbs_create_generators); the parity tests compare that derivation with this fixture."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import bbs_oracle as O

for cs in (O.BLS12_381, O.BN254):
    gens = O.create_generators(cs, 129, cs.api_id)
    blob = b"".join(cs.g1_compress(g) for g in gens)
    path = os.path.join(ROOT, "tests", "golden", f"generators_{cs.name.lower()}.bin")
    with open(path, "wb") as f:
        f.write(blob)
    print(path, len(blob))
