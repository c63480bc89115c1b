"""ctypes binding of oracle/_ref/libbbs_cref.so (the C restatement, oracle/cref/bbs_cref.c).
TEST / BENCH INFRASTRUCTURE ONLY."""
import ctypes as C
import hashlib
import os
import time

import numpy as np

from . import bbs_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libbbs_cref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(PATH)
        l.cref_ctx_create.restype = C.c_void_p
        l.cref_ctx_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        l.cref_ctx_destroy.argtypes = [C.c_void_p]
        l.cref_verify_one.restype = C.c_int
        l.cref_verify_one.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        l.cref_verify_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_int]
        l.cref_verify_batch_as_reference.argtypes = l.cref_verify_batch.argtypes
        l.cref_create_generators.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        l.cref_sign_one.restype = C.c_int
        l.cref_sign_one.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CrefContext:
    """pk + generators decoded once (they are typed values in the reference's signature)."""

    def __init__(self, cs, pk, gens, api_id=None):
        assert cs.name == "BLS12_381", "the C restatement covers BLS12-381"
        self.cs, self.L = cs, len(gens) - 1
        api_id = cs.api_id if api_id is None else api_id
        pkb = np.frombuffer(cs.g2_compress(pk), dtype=np.uint8)
        gb = np.frombuffer(b"".join(cs.g1_compress(g) for g in gens), dtype=np.uint8)
        ab = np.frombuffer(api_id, dtype=np.uint8)
        self.h = lib().cref_ctx_create(_p(pkb), _p(gb), len(gens), _p(ab), len(api_id))
        if not self.h:
            raise ValueError("cref_ctx_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().cref_ctx_destroy(self.h)
            self.h = None

    @staticmethod
    def _pack(messages):
        flat = np.frombuffer(b"".join(m for item in messages for m in item) or b"\0", dtype=np.uint8)
        lens = [len(m) for item in messages for m in item]
        offs = np.zeros(len(lens) + 1, dtype=np.uint64)
        if lens:
            offs[1:] = np.cumsum(lens, dtype=np.uint64)
        return flat, offs

    def verify_batch(self, sig_bytes, messages, header=b"", threads=0, as_reference=False):
        """as_reference: mode (A) of BASELINE.md -- `create_generators(L + 1)` recomputed for every item, exactly as
        `PublicKey::verify` does (verify.rs:35)."""
        n = len(messages)
        flat, offs = self._pack(messages)
        sigs = np.frombuffer(sig_bytes, dtype=np.uint8)
        st = np.full(n, 255, dtype=np.uint8)
        hb = np.frombuffer(header or b"\0", dtype=np.uint8)
        fn = lib().cref_verify_batch_as_reference if as_reference else lib().cref_verify_batch
        fn(self.h, n, _p(sigs), _p(flat), _p(offs), _p(hb), len(header), _p(st), threads or (os.cpu_count() or 1))
        return st

    def sign(self, sk, msgs, header=b""):
        flat, offs = self._pack([msgs])
        skb = np.frombuffer(sk.to_bytes(32, "little"), dtype=np.uint8)
        sig = np.zeros(80, dtype=np.uint8)
        b = np.zeros(48, dtype=np.uint8)
        hb = np.frombuffer(header or b"\0", dtype=np.uint8)
        lib().cref_sign_one(self.h, _p(skb), _p(flat), _p(offs), _p(hb), len(header), _p(sig), _p(b))
        return sig.tobytes(), b.tobytes()


def create_generators(cs, count, api_id=None):
    api_id = cs.api_id if api_id is None else api_id
    ab = np.frombuffer(api_id or b"\0", dtype=np.uint8)
    out = np.zeros(48 * max(count, 1), dtype=np.uint8)
    lib().cref_create_generators(_p(ab), len(api_id), count, _p(out))
    return out[: 48 * count].tobytes()


def time_verify(cs, L, sample):
    """cpu_baseline leg: all host cores over a bounded sample of the cfg-2 workload."""
    cores = os.cpu_count() or 1
    sk = O.key_gen(cs, hashlib.sha256(b"bbs-b200-key" + (7).to_bytes(4, "big")).digest(), b"", b"BBS-SIG-KEYGEN-SALT-")
    pk = O.sk_to_pk(cs, sk)
    gens = O.create_generators_cached(cs, L + 1, cs.api_id)
    ctx = CrefContext(cs, pk, gens)
    base = 16
    msgs = [[hashlib.sha256(f"7/{i}/{j}".encode()).digest() for j in range(L)] for i in range(base)]
    sigs = [ctx.sign(sk, m)[0] for m in msgs]
    # calibrate, then size the sample for ~10-20 s of CPU work unless the caller fixed it
    t0 = time.perf_counter()
    st = ctx.verify_batch(b"".join(sigs), msgs)
    dt = time.perf_counter() - t0
    assert (st == 1).all()
    n = sample or int(max(cores * 4, min(4096, base / dt * 20)))
    reps = (n + base - 1) // base
    big_msgs = (msgs * reps)[:n]
    big_sigs = b"".join((sigs * reps)[:n])
    t0 = time.perf_counter()
    st = ctx.verify_batch(big_sigs, big_msgs)
    dt = time.perf_counter() - t0
    assert (st == 1).all()
    # mode (A) of BASELINE.md beside it: PublicKey::verify as written, i.e. create_generators(L + 1) (L + 1 hash-to-curve
    # operations) recomputed for every signature (verify.rs:35), on a quarter of the sample
    na = max(cores * 2, n // 4)
    t0 = time.perf_counter()
    sta = ctx.verify_batch(b"".join((sigs * reps)[:na]), (msgs * reps)[:na], as_reference=True)
    dta = time.perf_counter() - t0
    assert (sta == 1).all()
    return {"value": n / dt, "unit": "verifies/s", "cores": cores, "kind": "port",
            "sample": f"{n} signatures, L={L}: C restatement of msg_to_scalars + core_verify as the reference runs them per item "
                      "(domain recomputed, double-and-add, G2 mul, two Miller loops + two final exponentiations, G1 decompression "
                      "with ark's subgroup test), OpenMP over items",
            "seconds": dt,
            "as_reference": {"value": na / dta, "unit": "verifies/s", "seconds": dta,
                             "sample": f"{na} signatures: the same plus create_generators(L + 1) per signature, as PublicKey::verify "
                                       "does (verify.rs:35): mode (A) of BASELINE.md"}}
