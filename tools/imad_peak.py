import ctypes as C, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bbs_sign_b200 import _native
lib = _native.load()
g = C.c_double(); ms = C.c_float()
for mode in (0, 1):
    lib.bbs_imad_peak(0, 2000, mode, C.byref(g), C.byref(ms))
    print("mode", mode, "Tprod/s", g.value / 1e3, "ms", ms.value, "per clk per SM @1965:", g.value * 1e9 / 148 / 1.965e9)
