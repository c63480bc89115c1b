"""CPU timing leg (TEST / BENCH INFRASTRUCTURE): times the oracle's restatement of the reference's per-item
path -- msg_to_scalars (interface_utilities.rs:76-88) + core_verify (verify.rs:53-93: domain per call,
L+1 double-and-add scalar multiplications, one G2 scalar multiplication, two full pairings) -- on the
host cores.  Only bench.py's cpu_baseline / --impl reference legs call this.

kind = "port": the reference itself (Rust + arkworks) cannot be built in this image (no cargo / rustc), so
the baseline is the oracle port: the compiled C restatement under oracle/cref (all host cores) when its
shared library has been built into oracle/_ref/, otherwise the big-int Python oracle on one core."""
import ctypes as C
import hashlib
import os
import time

from . import bbs_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CREF = os.path.join(HERE, "_ref", "libbbs_cref.so")


def _workload(cs, L, n, seed=7):
    sk = O.key_gen(cs, hashlib.sha256(b"bbs-b200-key" + seed.to_bytes(4, "big")).digest(), b"", b"BBS-SIG-KEYGEN-SALT-")
    pk = O.sk_to_pk(cs, sk)
    msgs = [[hashlib.sha256(f"{seed}/{i}/{j}".encode()).digest() for j in range(L)] for i in range(n)]
    return sk, pk, msgs


def time_verify(L=10, sample=0, curve="BLS12_381"):
    cs = O.SUITES[curve]
    if os.path.exists(CREF):
        from . import cref_binding
        return cref_binding.time_verify(cs, L, sample)
    n = sample or 4
    sk, pk, msgs = _workload(cs, L, n)
    sigs = [O.sign(cs, sk, m, b"") for m in msgs]
    gens = O.create_generators_cached(cs, L + 1, cs.api_id)
    t0 = time.perf_counter()
    ok = 0
    for s, m in zip(sigs, msgs):
        ms = O.msg_to_scalars(cs, m, cs.api_id)
        ok += O.core_verify(cs, pk, s, gens, b"", ms, cs.api_id)
    dt = time.perf_counter() - t0
    assert ok == n
    return {"value": n / dt, "unit": "verifies/s", "cores": 1, "kind": "port",
            "sample": f"{n} signatures, L={L}, big-int Python oracle (generators cached; domain, B, G2 mul and two "
                      "pairings per item as core_verify does)", "seconds": dt}
