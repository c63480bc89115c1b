"""CPU-side checks: the CUDA library loads and exports every symbol include/bbs_b200.h declares (no compute
calls: there is no GPU here), the package refuses to run without it, and the N>1 sharding path (two gloo
ranks) reproduces the single-rank status vector."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bbs_sign_b200", "libbbs_b200.so")


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bbs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bbs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True)
    lib = ctypes.CDLL(LIB)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bbs_b200.h but not exported"
    from bbs_sign_b200 import _native
    assert sorted(_native.SYMBOLS) == syms          # the Python binding covers the header exactly
    # pure size queries work without a GPU
    assert lib.bbs_g1_bytes(1) == 48 and lib.bbs_g1_bytes(2) == 32


def test_no_gpu_means_error_not_fallback():
    """Without a CUDA device the product path must fail loudly (BBS_E_CUDA), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bbs_sign_b200 import api
    with pytest.raises(api.BbsError):
        api.BatchContext(api.BLS12_381, bytes(96), b"", n_messages=1)


def test_missing_library_is_loud(tmp_path):
    from bbs_sign_b200 import _native
    with pytest.raises(_native.NativeLibraryMissing):
        _native.load(str(tmp_path / "nope.so"))


def test_shard_bounds():
    from bbs_sign_b200.sharding import shard_bounds
    for n in (0, 1, 7, 65536, 1000003):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


WORKER = r'''
import os, sys, hashlib
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import hostsim, parity_cases as P
from oracle import bbs_oracle as O
from bbs_sign_b200 import api
from bbs_sign_b200.sharding import shard_bounds
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
suite, ocs = P.SUITES["BLS12_381"]
sk, pk = P.keypair(ocs, 1)
L, n = 2, 6
msgs = [[P.rng_bytes(f"s{i}.{j}", 32) for j in range(L)] for i in range(n)]
sigs = [O.sign(ocs, sk, m, b"") for m in msgs]
sigs[1] = (sigs[1][0], (sigs[1][1] + 1) % ocs.r)
sigs[4] = (None, sigs[4][1])
blob = b"".join(O.signature_to_bytes(ocs, s) for s in sigs)
lo, hi = shard_bounds(n, world)[rank]
from bbs_sign_b200 import _native
_native.load(hostsim.build(), allow_host_simulation=True)     # tests only
ctx, _ = P.make_ctx(hostsim.build(), suite, ocs, pk, b"", L)
st = ctx.verify_batch(np.frombuffer(blob[lo * 80: hi * 80], dtype=np.uint8), msgs[lo:hi])
parts = [None] * world
dist.all_gather_object(parts, st.tolist())     # test-only gather; the data path itself has no collective
if rank == 0:
    got = [x for p in parts for x in p]
    assert got == [1, 0, 1, 1, 0, 1], got
    print("SHARD_OK")
dist.destroy_process_group()
'''


def test_two_rank_sharding_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                       capture_output=True, text=True, env=env, timeout=600)
    assert "SHARD_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's CPU path: oracle/_ref C port on the host cores) prints one JSON line
    with the keys the driver reads; runs on the CPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle", "cref")], check=True, capture_output=True)   # oracle/_ref C port
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "64"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "bls12_381_bbs_verifies_per_sec_L10" and d["unit"] == "verifies/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    sys.path.insert(0, root)
    import bench
    assert d["config"] == bench.verify_config(65536, 10)        # the same config dict as our arm (bounded sample per step)
