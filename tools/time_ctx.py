import time, sys, os
sys.path.insert(0, os.getcwd())
import bench
from bbs_sign_b200 import api
import numpy as np
for suite, pk, L in ((api.BLS12_381, bench.IRTF_PK, 10), (api.BLS12_381, bench.IRTF_PK, 32)):
    for rep in range(2):
        t0 = time.perf_counter(); ctx = api.BatchContext(suite, pk, header=b"", n_messages=L); dt = time.perf_counter() - t0; ctx.close()
    print(suite.name, "L =", L, "context creation %.1f ms" % (dt * 1e3))
key = np.load("tests/golden/bn254_bench_key.npz")
for rep in range(2):
    t0 = time.perf_counter(); ctx = api.BatchContext(api.BN254, bytes(key["pk"]), header=b"", n_messages=31); dt = time.perf_counter() - t0; ctx.close()
print("BN254 L = 31 context creation %.1f ms" % (dt * 1e3))
