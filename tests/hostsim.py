"""Builds and binds the HOST SIMULATION of the CUDA sources (g++ -DBBS_HOSTSIM): the same kernels.cuh /
capi.cu compiled for x86 with a host loop instead of <<<>>>.  It exists only so the pipeline logic and
the ABI plumbing can be exercised by `-m "not gpu"` tests in a container without a GPU.  It is never
shipped, never loaded by the package, and is not a fallback: `bbs_sign_b200` always binds
libbbs_b200.so (CUDA)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "bbs_sign_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build", "libbbs_hostsim.so")


def build() -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(ROOT, "include", "bbs_b200.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-w", "-DBBS_HOSTSIM", "-x", "c++", "-fPIC", "-shared", "-o", OUT,
           os.path.join(SRC, "capi.cu")] + sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.startswith("tu_"))
    subprocess.run(cmd, check=True)
    return OUT
