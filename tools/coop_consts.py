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
T_BN = 4965661367192848881
P_BN = 36 * T_BN ** 4 + 36 * T_BN ** 3 + 24 * T_BN ** 2 + 6 * T_BN + 1
BN = Curve("bn", P_BN, 8, 9, twist="D", wide=True)


def naf(x):
    """little-endian non-adjacent form (digits in {-1, 0, 1}); same digits as tools/gen_constants.py BN_ATE_NAF"""
    out = []
    while x:
        if x & 1:
            d = 2 - (x & 3)
            x -= d
        else:
            d = 0
        out.append(d)
        x >>= 1
    return out


BN_NAF = naf(6 * T_BN + 2)
# Miller loop of the cooperative program: digits below the leading one, MSB first, '1' = an addition step follows
BN_MILLER_DIGITS = "".join("1" if d else "0" for d in reversed(BN_NAF[:-1]))
BN_MILLER_TAIL = 2


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


def line_table_bn(cv, Q):
    """BN254 (D-type twist): line through T (tangent) or T, +-Q / pi(Q) / -pi^2(Q) evaluated at P = (x, y) is
       l(P) = y + (Bc x) w + A w^3,  A = lambda x_T - y_T,  Bc = -lambda  (pairing.cuh g2_precompute_lines, f12_mul_line),
    normalised by the Fp factor 1/y (dies in the final exponentiation) to  1 + Bc (x/y) w + A (1/y) w^3.
    Returns (Bc, A) per line, in the order of pairing.cuh; the kernel keeps (x/y, 1/y) as the item's point."""
    f2 = F2(cv.p)
    T = Q
    out = []
    nQ = (Q[0], f2.neg(Q[1]))

    def step(lam, X2):
        x3 = f2.sub(f2.sub(f2.mul(lam, lam), T[0]), X2)
        y3 = f2.sub(f2.mul(lam, f2.sub(T[0], x3)), T[1])
        return (x3, y3)

    def dbl():
        nonlocal T
        x2 = f2.mul(T[0], T[0])
        lam = f2.mul(f2.add(f2.add(x2, x2), x2), f2.inv(f2.add(T[1], T[1])))
        out.append((f2.neg(lam), f2.sub(f2.mul(lam, T[0]), T[1])))
        T = step(lam, T[0])

    def add(X):
        nonlocal T
        lam = f2.mul(f2.sub(X[1], T[1]), f2.inv(f2.sub(X[0], T[0])))
        out.append((f2.neg(lam), f2.sub(f2.mul(lam, T[0]), T[1])))
        T = step(lam, X[0])

    for d in reversed(BN_NAF[:-1]):
        dbl()
        if d:
            add(Q if d > 0 else nQ)
    xi = (cv.xi_c, 1)
    conj = lambda a: (a[0], (-a[1]) % cv.p)
    g12 = f2.pow(xi, (cv.p - 1) // 3)          # xi^(2 (p-1) / 6)
    g13 = f2.pow(xi, (cv.p - 1) // 2)          # xi^(3 (p-1) / 6)
    g22 = f2.pow(xi, (cv.p ** 2 - 1) // 3)
    g23 = f2.pow(xi, (cv.p ** 2 - 1) // 2)
    Q1 = (f2.mul(conj(Q[0]), g12), f2.mul(conj(Q[1]), g13))
    Q2 = (f2.mul(Q[0], g22), f2.neg(f2.mul(Q[1], g23)))
    add(Q1)
    add(Q2)
    return out
