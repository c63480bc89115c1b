#!/usr/bin/env python3
"""Curve constants and host-side reference computations shared by the cooperative-pairing program generator
(tools/gen_pairing_prog.py) and its CPU test (tests/test_coop_program.py): Frobenius constants in the w-basis,
the Fp2-normalised line table of a G2 point, and the static constant-table layout."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from coop_prog import Curve  # noqa: E402

X_BLS = -0xD201000000010000
R_BLS = X_BLS ** 4 - X_BLS ** 2 + 1
P_BLS = (X_BLS - 1) ** 2 * R_BLS // 3 + X_BLS
BLS = Curve("bls", P_BLS, 12, 1)


class F2:
    """Fp2 = Fp[u]/(u^2+1) on tuples"""

    def __init__(self, p):
        self.p = p

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def neg(self, a):
        return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % self.p, (a[0] * b[1] + a[1] * b[0]) % self.p)

    def inv(self, a):
        n = pow(a[0] * a[0] + a[1] * a[1], -1, self.p)
        return (a[0] * n % self.p, (-a[1]) * n % self.p)

    def pow(self, a, e):
        r = (1, 0)
        while e:
            if e & 1:
                r = self.mul(r, a)
            a = self.mul(a, a)
            e >>= 1
        return r


def const_table(cv):
    """name -> (id, Fp2 value).  gamma_{j,k} = xi^(k (p^j - 1) / 6): (c w^k)^(p^j) = conj^j(c) gamma_{j,k} w^k."""
    f2 = F2(cv.p)
    xi = (cv.xi_c, 1)
    names = {}
    names["one"] = (1, 0)
    for j in (1, 2, 3):
        for k in range(6):
            names[f"frob{j}_{k}"] = f2.pow(xi, k * (cv.p ** j - 1) // 6)
    ids = {n: i for i, n in enumerate(names)}
    vals = {ids[n]: v for n, v in names.items()}
    for k in range(6):
        assert names[f"frob2_{k}"][1] == 0
    return ids, vals


def line_table(cv, Q, ate_abs):
    """Per ate step and G2 point: the line through T (tangent) or T, Q evaluated at P = (x, y) is
       l(P) = A + (Bc x) w^2 + y w^3  with  A = lambda x_T - y_T,  Bc = -lambda   (M-type twist; the w^3 scaling and
    any Fp2 factor die in the final exponentiation), normalised here to  1 + (Bc/A) x w^2 + (1/A) y w^3.
    Returns a list of (Bc/A, 1/A) per line; raises ZeroDivisionError if some A == 0 (degenerate; callers fall back)."""
    f2 = F2(cv.p)
    T = Q
    out = []

    def emit(lam):
        A = f2.sub(f2.mul(lam, T[0]), T[1])
        if A == (0, 0):
            raise ZeroDivisionError
        Ai = f2.inv(A)
        out.append((f2.mul(f2.neg(lam), Ai), Ai))

    def step(lam, X2):
        x3 = f2.sub(f2.sub(f2.mul(lam, lam), T[0]), X2)
        y3 = f2.sub(f2.mul(lam, f2.sub(T[0], x3)), T[1])
        return (x3, y3)

    for bit in bin(ate_abs)[3:]:
        x2 = f2.mul(T[0], T[0])
        lam = f2.mul(f2.add(f2.add(x2, x2), x2), f2.inv(f2.add(T[1], T[1])))
        emit(lam)
        T = step(lam, T[0])
        if bit == "1":
            lam = f2.mul(f2.sub(Q[1], T[1]), f2.inv(f2.sub(Q[0], T[0])))
            emit(lam)
            T = step(lam, Q[0])
    return out
