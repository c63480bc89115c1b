import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import parity_cases as P
from oracle import bbs_oracle as O
suite, ocs = P.SUITES["BLS12_381"]
pk = O.sk_to_pk(ocs, P.IRTF_SK)
ctx, gens = P.make_ctx(None, suite, ocs, pk, P.HEADER, 1)
sigs, b, st = ctx.sign_batch(P.IRTF_SK.to_bytes(32, "little"), [[P.MSG]], want_b=True)
print("sign st", st.tolist())
print("verify", ctx.verify_batch(sigs, [[P.MSG]]).tolist())
import numpy as np
print("verify x3", ctx.verify_batch(np.concatenate([sigs[0]] * 3), [[P.MSG]] * 3).tolist())
