#!/usr/bin/env python3
"""Generates tests/golden/bn254_bench_key.npz: the BN254 issuer key pair used by `bench.py --workload bn254`
(sk = key_gen("bbs-b200-bn254-bench-key-material", "", "BBS-SIG-KEYGEN-SALT-"), pk = sk * BP2, compressed), made with
the oracle so that bench.py itself does not import it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bbs_oracle as O  # noqa: E402

cs = O.BN254
sk = O.key_gen(cs, b"bbs-b200-bn254-bench-key-material", b"", b"BBS-SIG-KEYGEN-SALT-")
pk = O.sk_to_pk(cs, sk)
out = os.path.join(ROOT, "tests", "golden", "bn254_bench_key.npz")
np.savez(out, sk=np.frombuffer(sk.to_bytes(32, "little"), dtype=np.uint8), pk=np.frombuffer(cs.g2_compress(pk), dtype=np.uint8))
print("wrote", out)
