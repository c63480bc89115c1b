#!/usr/bin/env python3
"""Program builder + big-int emulator for the cooperative pairing kernel (bbs_sign_b200/csrc/pairing_coop.cuh).

The kernel is an interpreter: a thread block owns 32 items (lane = item) and has one warp per ROLE; role k
owns the Fp2 coefficient c_k of every Fp12 value  f = sum c_k w^k  in  Fp2[w]/(w^6 - xi).  All values live in
shared-memory CELLS (one Fp2 per cell and item, Montgomery form, canonical).  A role executes its own stream:

  EP   acc_R += sR * X*Y, acc_I += sI * X*Y      X, Y = Fp operands derived from a cell:
                                                 form 0: c0, 1: c1, 2: c0+c1, 3: c0-c1+p, shifted left by 0..2;
                                                 Y may come from the global constant table instead of a cell
  FIN  cell[dst] = canon( REDC( 3^t * acc + zs * 2^zd * Z * R + KP[k] ) )   per component, then acc = 0
  CTL  REP n / ENDREP / NEXTLINE / BAR / GSAVE / GLOAD / CHECK / INV (Fp inversion of a cell, in place) / END

i.e. every output coefficient is ONE lazily-reduced sum of double-width products (Karatsuba at the Fp2 level:
three EPs per Fp2 product, each feeding the real and/or imaginary accumulator with a sign).  This file builds the
per-role streams for  "e(P0,Q0) e(P1,Q1) == 1"  (two-pair Miller loop over precomputed, Fp2-normalised lines +
final exponentiation), emulates them on Python integers exactly as the kernel computes (Montgomery domain, same
accumulator / REDC / canonicalisation arithmetic, with range assertions and a barrier-hazard checker), and is
imported by tools/gen_pairing_prog.py (emission) and tests/test_coop_program.py (parity with the oracle).

Replaces, together with the kernel, the two `E::pairing` calls of verify.rs:88-92 / proof_verify.rs:112-115.
"""
from dataclasses import dataclass
from fractions import Fraction

ROLES = 6
NCELLS = 24            # shared-memory cells per block (4 Fp12 values)

K_EP, K_FIN, K_CTL, K_XI = 0, 1, 2, 3
C_END, C_REP, C_ENDREP, C_NEXTLINE, C_BAR, C_GSAVE, C_GLOAD, C_CHECK, C_INV = range(9)
F_C0, F_C1, F_SUM, F_DIFF = range(4)
FORM_BOUND = {F_C0: 1, F_C1: 1, F_SUM: 2, F_DIFF: 2}     # in units of p (cells are canonical)
EP_HALF = 1 << 31      # EP only: executed by the second warp of a role pair in the SPLIT = 2 kernel (ignored otherwise)
LINE_BASE = 128        # global-constant ids >= LINE_BASE address the line table relative to the line counter
KP_HI = (0, 1, 2, 3, 4, 5, 6, 7)  # BLS12-381: KP[k] = KP_HI[k] * p * R, added before REDC so the accumulator is >= 0.  A multiple
                                 # of R touches only the HIGH half of the accumulator (13-word addition of k p instead of 25 words)
QP_ROWS = 10                     # BLS12-381: a reduction result is < R < 10 p; the kernel subtracts q^ p from a table (canon_q)
CANON_STEPS = (1, 2, 3, 4)       # canon level l: output < 2^(l+1) p; level 0: one conditional subtraction of p, above: canon_q
                                 # (BLS12-381: quotient estimate + table of q p; the step ladder 2^l p ... p remains as fallback)
# BN254: xi = 9 + u is applied to the ACCUMULATORS (instruction XI: (R, I) <- (9R - I, 9I + R)), so sums reach a few
# hundred p^2; the reduction there returns N+1 limbs (< 128 p) and canonicalises in up to 7 steps
KP_HI_WIDE = (0, 2, 4, 7, 13, 25, 49, 97)       # BN254 (R / p = 5.29): KP[k] = KP_HI_WIDE[k] * p * R >= (0, 8, ..., 512) p^2
CANON_STEPS_WIDE = (1, 2, 3, 4, 5, 6, 7)


@dataclass
class Operand:
    cell: int
    form: int = F_C0
    shift: int = 0
    glob: bool = False     # Y only: global constant table


def enc_ep(sR, sI, x, y):
    assert not x.glob
    sg = {0: 0, 1: 1, -1: 2}
    assert 0 <= x.cell < 256 and 0 <= y.cell < 256 and x.shift <= 2 and y.shift <= 2
    return (K_EP | (sg[sR] << 2) | (sg[sI] << 4) | (x.cell << 6) | (x.form << 14) | (x.shift << 16) |
            (y.cell << 18) | (y.form << 26) | (y.shift << 28) | ((1 if y.glob else 0) << 30))


def enc_fin(dst, triple=0, zsign=0, zdouble=0, zcell=0, kp=0, canon=3, bar=0, skip_pair=None, fp_only=0):
    zs = {0: 0, 1: 1, -1: 2}[zsign]
    sk = 0 if skip_pair is None else (1 | (skip_pair << 1))
    assert kp < 8 and canon < 8
    return (K_FIN | (dst << 2) | (triple << 10) | (zs << 11) | (zdouble << 13) | (zcell << 14) | ((kp & 3) << 22) |
            ((canon & 3) << 24) | (bar << 26) | (sk << 27) | (fp_only << 29) | ((canon >> 2) << 30) | ((kp >> 2) << 31))


def enc_ctl(sub, arg=0):
    return K_CTL | (sub << 2) | (arg << 6)


def dec(word):
    kind = word & 3
    sg = {0: 0, 1: 1, 2: -1}
    if kind == K_EP:
        return ("EP", sg[(word >> 2) & 3], sg[(word >> 4) & 3],
                Operand((word >> 6) & 255, (word >> 14) & 3, (word >> 16) & 3),
                Operand((word >> 18) & 255, (word >> 26) & 3, (word >> 28) & 3, bool((word >> 30) & 1)))
    if kind == K_FIN:
        return ("FIN", dict(dst=(word >> 2) & 255, triple=(word >> 10) & 1, zsign=sg[(word >> 11) & 3],
                            zdouble=(word >> 13) & 1, zcell=(word >> 14) & 255,
                            kp=((word >> 22) & 3) | (((word >> 31) & 1) << 2),
                            canon=((word >> 24) & 3) | (((word >> 30) & 1) << 2), bar=(word >> 26) & 1,
                            skip=(word >> 27) & 3, fp_only=(word >> 29) & 1))
    if kind == K_XI:
        return ("XI",)
    return ("CTL", (word >> 2) & 15, word >> 6)


@dataclass
class Curve:
    name: str
    p: int
    n_limbs: int
    xi_c: int        # xi = xi_c + u
    twist: str = "M"     # M: line = 1 + l2 w^2 + l3 w^3 ; D: line = 1 + l1 w + l3 w^3  (normalised, see coop_consts.py)
    wide: bool = False   # reduction returns N+1 limbs (R - p is too small a head-room for lazy sums: BN254)

    @property
    def R(self):
        return 1 << (32 * self.n_limbs)

    @property
    def kp_mult(self):
        """KP table in units of p^2"""
        return tuple(Fraction(k * self.R, self.p) for k in (KP_HI_WIDE if self.wide else KP_HI))

    @property
    def canon_steps(self):
        return CANON_STEPS_WIDE if self.wide else CANON_STEPS

    @property
    def acc_limit(self):
        """accumulator bound at REDC time, in units of p^2"""
        if self.wide:         # REDC output < T/R + p must stay below 2^max_step p
            return Fraction(((1 << self.canon_steps[-1]) - 1) * self.R, self.p)
        return Fraction((self.R - self.p) * self.R, self.p * self.p)


class ProgramError(Exception):
    pass


class Builder:
    """Emits the six per-role instruction streams op by op and does the static (worst-case) bound analysis."""

    def __init__(self, curve):
        self.cv = curve
        self.streams = [[] for _ in range(ROLES)]
        self.max_t = curve.acc_limit                                                # accumulator limit in p^2
        self.work = {"ep": [0] * ROLES, "fin": [0] * ROLES, "bar": 0}
        self._mult = [1]

    def ctl_all(self, sub, arg=0):
        for r in range(ROLES):
            self.streams[r].append(enc_ctl(sub, arg))

    def ctl_roles(self, sub, args):
        for r in range(ROLES):
            self.streams[r].append(enc_ctl(sub, args[r]))

    def rep(self, n):
        assert 1 <= n < (1 << 20)
        self.ctl_all(C_REP, n)
        self._mult.append(self._mult[-1] * n)

    def endrep(self):
        self.ctl_all(C_ENDREP)
        self._mult.pop()

    def op(self, per_role, bar=True):
        """per_role: role -> dict(dst=cell, eps=[(sR, sI, X, Y)], z=(sign, double, cell) | None, triple=bool,
        skip_pair=None|0|1, fp_only=bool).  Roles without an entry only take part in the barrier."""
        pu = Fraction(self.cv.p, self.cv.R)
        if bar:
            self.work["bar"] += self._mult[-1]
        for r in range(ROLES):
            d = per_role.get(r)
            if d is None:
                if bar:
                    self.streams[r].append(enc_ctl(C_BAR))
                continue
            pos = {"R": Fraction(0), "I": Fraction(0)}
            neg = {"R": Fraction(0), "I": Fraction(0)}
            groups = [(d.get("eps_xi") or [], True), (d["eps"], False)]
            for eps, is_xi in groups:
                for (sR, sI, x, y) in eps:
                    bx, by = FORM_BOUND[x.form] << x.shift, FORM_BOUND[y.form] << y.shift
                    if max(bx, by) * self.cv.p >= self.cv.R:
                        raise ProgramError("operand overflows the limb array")
                    for acc, s in (("R", sR), ("I", sI)):
                        if s > 0:
                            pos[acc] += bx * by
                        elif s < 0:
                            neg[acc] += bx * by
                    self.streams[r].append(enc_ep(sR, sI, x, y))
                    self.work["ep"][r] += self._mult[-1]
                if is_xi and eps:
                    # (R, I) <- (c R - I, c I + R)
                    c = self.cv.xi_c
                    pos, neg = ({"R": c * pos["R"] + neg["I"], "I": c * pos["I"] + pos["R"]},
                                {"R": c * neg["R"] + pos["I"], "I": c * neg["I"] + neg["R"]})
                    self.streams[r].append(K_XI)
                    self.work["xi"] = self.work.get("xi", 0) + self._mult[-1]
            tr = 3 if d.get("triple") else 1
            z = d.get("z")
            zpos = zneg = Fraction(0)
            if z:
                zb = Fraction(1 << z[1]) / pu          # (2^zd p) R in units of p^2
                if z[0] > 0:
                    zpos = zb
                else:
                    zneg = zb
            worst_neg = max(neg["R"], neg["I"]) * tr + zneg
            worst_pos = max(pos["R"], pos["I"]) * tr + zpos
            KPM, CST = self.cv.kp_mult, self.cv.canon_steps
            kp = next((i for i, m in enumerate(KPM) if m >= worst_neg), None)
            if kp is None:
                raise ProgramError(f"negative part {float(worst_neg):.1f} p^2 exceeds the KP table")
            total = worst_pos + KPM[kp]
            if total >= self.max_t:
                raise ProgramError(f"accumulator bound {float(total):.1f} p^2 >= limit {float(self.max_t):.1f}")
            out_bound = total * pu + 1          # REDC output < T/R + p, in units of p
            canon = next((i for i, st in enumerate(CST) if (1 << st) >= out_bound), None)
            if canon is None:
                raise ProgramError("output bound too large to canonicalise")
            self.streams[r].append(enc_fin(d["dst"], triple=1 if d.get("triple") else 0, zsign=z[0] if z else 0,
                                           zdouble=z[1] if z else 0, zcell=z[2] if z else 0, kp=kp, canon=canon,
                                           bar=1 if bar else 0, skip_pair=d.get("skip_pair"),
                                           fp_only=1 if d.get("fp_only") else 0))
            self.work["fin"][r] += self._mult[-1]

    def finish(self):
        self.ctl_all(C_END)
        return [self._tag_halves(s) for s in self.streams]

    @staticmethod
    def _tag_halves(stream):
        """EP_HALF bit for the kernel's two-warps-per-role variant (small batches, pairing_coop.cuh SPLIT = 2): the EPs
        between two FINs alternate between the halves of the pair; each half accumulates its share, the FIN adds the
        shares (exact: integer accumulators).  The one-warp kernel and the emulator ignore the bit."""
        out, k = [], 0
        for w in stream:
            kind = w & 3
            if kind == K_EP:
                w |= (k & 1) << 31
                k += 1
            elif kind == K_FIN:
                k = 0
            out.append(w)
        return out


# ---- Fp2-level decompositions into EPs (Karatsuba, 3 EPs per Fp2 product) --------------------------------
def _emit(coefs_forms, xc, yc, sign, yglob):
    eps = []
    for c, xf, yf in coefs_forms:
        cr, ci = sign * c["R"], sign * c["I"]
        mags = {abs(v) for v in (cr, ci) if v}
        if not mags:
            continue
        assert len(mags) == 1, (cr, ci)
        sh = {1: 0, 2: 1, 4: 2}[mags.pop()]
        eps.append(((cr > 0) - (cr < 0), (ci > 0) - (ci < 0), Operand(xc, xf, sh), Operand(yc, yf, 0, yglob)))
    return eps


def _xi(cs):
    for c in cs:                 # (r, i) -> (r - i, r + i)   [xi = 1 + u]
        c["R"], c["I"] = c["R"] - c["I"], c["R"] + c["I"]


def fp2_mul_eps(xc, yc, sign=1, xi=False, xconj=False, yconj=False, yglob=False):
    """EPs adding  sign * [xi] * conj?(x) * conj?(y)  to (R, I); x, y are cells.  sign in {+-1, +-2}."""
    sx = -1 if xconj else 1
    sy = -1 if yconj else 1
    # real = A - sx sy B ; imag = s (S - A - B) if sx == sy == s, else sy (S' - A + B) with S' = (x0 - x1)(y0 + y1)
    cA = {"R": 1, "I": 0}
    cB = {"R": -sx * sy, "I": 0}
    cS = {"R": 0, "I": 0}
    if sx == sy:
        cA["I"], cB["I"], cS["I"] = -sx, -sx, sx
        sform = F_SUM
    else:
        cA["I"], cB["I"], cS["I"] = -sy, sy, sy
        sform = F_DIFF
    if xi:
        _xi((cA, cB, cS))
    return _emit(((cA, F_C0, F_C0), (cB, F_C1, F_C1), (cS, sform, F_SUM)), xc, yc, sign, yglob)


def fp2_sqr_eps(xc, sign=1, xi=False):
    """EPs adding sign * [xi] * x^2:  real = (x0+x1)(x0-x1+p), imag = (2 x0) x1."""
    cP = {"R": 1, "I": 0}
    cQ = {"R": 0, "I": 2}
    if xi:
        _xi((cP, cQ))
    return _emit(((cP, F_SUM, F_DIFF), (cQ, F_C0, F_C1)), xc, xc, sign, False)


def fp2_mul_fp_eps(xc, yc, yform=F_C0, sign=1, xconj=False, yglob=False):
    """EPs adding sign * conj?(x) * y  with y in Fp (component `yform` of cell yc)."""
    return [(sign, 0, Operand(xc, F_C0), Operand(yc, yform, 0, yglob)),
            (0, -sign if xconj else sign, Operand(xc, F_C1), Operand(yc, yform, 0, yglob))]


# ---- symbolic Fp12 values ---------------------------------------------------------------------------------
@dataclass
class V12:
    base: int            # first of six consecutive cells
    conj: bool = False   # the value is the p^6-conjugate (odd coefficients negated) of what the cells hold

    def cell(self, k):
        return self.base + k

    def sgn(self, k):
        return -1 if (self.conj and (k & 1)) else 1


def conj(v):
    return V12(v.base, not v.conj)


class PairingProgram:
    """Streams for one curve.  Cell map: four Fp12 slots (cells 0..23); during the Miller loop slot 2 holds the
    evaluated line coefficients (cells 12..15) and the two G1 points (cells 16, 17: x in c0, y in c1).  The kernel
    prologue stores f = 1 in slot 0 and the points in cells 16, 17."""

    SLOTS = (0, 6, 12, 18)
    LINE_CELLS = (12, 13, 14, 15)      # pair 0: l2, l3 ; pair 1: l2, l3
    P_CELLS = (16, 17)

    def __init__(self, curve, ate_abs, const_index, inv_exp, miller_digits=None, miller_tail=0):
        """ate_abs: |x| (BLS12: also the Miller loop count; BN: the exponent t of the hard part).  miller_digits: string
        over '0'/'1' below the leading digit of the loop count, '1' = an addition step follows the doubling step
        (default: the bits of ate_abs); miller_tail: extra addition steps after the loop (BN: pi(Q), -pi^2(Q))."""
        self.cv = curve
        self.b = Builder(curve)
        self.ate_bits = bin(ate_abs)[3:]
        self.miller_digits = miller_digits if miller_digits is not None else self.ate_bits
        self.miller_tail = miller_tail
        self.ate_abs = ate_abs
        self.ci = const_index      # name -> global constant id (< LINE_BASE)
        self.inv_exp = inv_exp     # p - 2
        self.free = []
        self.acc_xi = curve.xi_c != 1      # xi applied to the accumulators (XI) instead of by sign routing
        self.line_exps = (2, 3) if curve.twist == "M" else (1, 3)

    # -- EP emission with the two ways of multiplying by xi -----------------------------------------------------
    def _put(self, d, eps_fn, xi, **kw):
        if xi and self.acc_xi:
            d.setdefault("eps_xi", []).extend(eps_fn(xi=False, **kw))
        else:
            d.setdefault("eps", []).extend(eps_fn(xi=xi, **kw))

    def fmul(self, d, xc, yc, sign=1, xi=False, **kw):
        self._put(d, lambda xi, **k: fp2_mul_eps(xc, yc, sign=sign, xi=xi, **k), xi, **kw)

    def fsqr(self, d, xc, sign=1, xi=False):
        self._put(d, lambda xi: fp2_sqr_eps(xc, sign=sign, xi=xi), xi)

    # -- Fp12 ops --------------------------------------------------------------------------------------------
    def mul(self, dst, A, B, roles=range(ROLES), bar=True, b_coeffs=range(ROLES)):
        """cells[dst + k] = (A * B)_k for k in roles; only coefficients j in b_coeffs of B are non-zero."""
        if A.conj == B.conj:
            out_conj, sa, sb = A.conj, (lambda k: 1), (lambda k: 1)
        else:
            out_conj, sa, sb = False, A.sgn, B.sgn
        per = {}
        for k in roles:
            d = dict(dst=dst + k, eps=[])
            for j in b_coeffs:
                i = (k - j) % 6
                self.fmul(d, A.cell(i), B.cell(j), sign=sa(i) * sb(j), xi=(i + j >= 6))
            per[k] = d
        self.b.op(per, bar=bar)
        return V12(dst, out_conj)

    def sqr(self, dst, A, extra=None):
        per = {}
        for k in range(ROLES):
            d = dict(dst=dst + k, eps=[])
            for i in range(6):
                j = (k - i) % 6
                if i > j:
                    continue
                if i == j:
                    self.fsqr(d, A.cell(i), xi=(i + j >= 6))
                else:
                    self.fmul(d, A.cell(i), A.cell(j), sign=2, xi=(i + j >= 6))
            per[k] = d
        if extra:
            self.b.op(per, bar=False)
            self.b.op(extra, bar=True)
        else:
            self.b.op(per)
        return V12(dst, A.conj)

    def sparse(self, dst, F, l2c, l3c):
        """F * (1 + la w^a + lb w^b), la / lb = line cells, (a, b) = (2, 3) on an M-type twist, (1, 3) on a D-type
        twist (line normalised to constant term 1)."""
        assert not F.conj
        per = {}
        for k in range(ROLES):
            d = dict(dst=dst + k, eps=[], z=(1, 0, F.cell(k)))
            for (j, lc) in ((self.line_exps[0], l2c), (self.line_exps[1], l3c)):
                i = (k - j) % 6
                self.fmul(d, F.cell(i), lc, xi=(i + j >= 6))
            per[k] = d
        self.b.op(per)
        return V12(dst)

    def line_eval(self):
        """LINE_CELLS[r] = lineconst[r] (Fp2, global, at the line counter) * P coordinate (Fp); 0 for skipped pairs."""
        per = {}
        for r in range(4):
            pair, which = r // 2, r % 2
            pc = Operand(self.P_CELLS[pair], F_C0 if which == 0 else F_C1)
            eps = [(1, 0, pc, Operand(LINE_BASE + r, F_C0, 0, True)), (0, 1, pc, Operand(LINE_BASE + r, F_C1, 0, True))]
            per[r] = dict(dst=self.LINE_CELLS[r], eps=eps, skip_pair=pair)
        return per

    def cyc_sqr(self, dst, A):
        """Granger-Scott squaring in the w-basis; the Fp4 pairs are (g_k, g_{k+3})."""
        g = A.cell
        plan = {0: ("A", g(0), g(3), False), 3: ("B", g(0), g(3), False),
                2: ("A", g(1), g(4), False), 5: ("B", g(1), g(4), False),
                4: ("A", g(2), g(5), False), 1: ("B", g(2), g(5), True)}
        per = {}
        for k, (kind, x, y, xi) in plan.items():
            if kind == "A":      # 3 (x^2 + xi y^2) - 2 g_k
                d = dict(dst=dst + k, eps=[], triple=True, z=(-1, 1, g(k)))
                self.fsqr(d, x)
                self.fsqr(d, y, xi=True)
            else:                # 3 [xi] (2 x y) + 2 g_k
                d = dict(dst=dst + k, eps=[], triple=True, z=(1, 1, g(k)))
                self.fmul(d, x, y, sign=2, xi=xi)
            per[k] = d
        self.b.op(per)
        return V12(dst, A.conj)

    def frob(self, dst, A, j):
        """A^(p^j): coefficient-wise Fp2-conjugation^j times gamma_{j,k} = xi^(k (p^j - 1) / 6)."""
        per = {}
        for k in range(ROLES):
            cid = self.ci[f"frob{j}_{k}"]
            if j % 2 == 0:
                eps = fp2_mul_fp_eps(A.cell(k), cid, sign=A.sgn(k), yglob=True)       # gamma in Fp
            else:
                eps = fp2_mul_eps(A.cell(k), cid, sign=A.sgn(k), xconj=True, yglob=True)
            per[k] = dict(dst=dst + k, eps=eps)
        self.b.op(per)
        return V12(dst)

    # -- Miller loop -------------------------------------------------------------------------------------------
    def miller(self):
        S0, S1 = self.SLOTS[0], self.SLOTS[1]
        b = self.b
        lc = self.LINE_CELLS
        st = {"f": V12(S0)}

        def other():
            return S1 if st["f"].base == S0 else S0

        def mul_lines():
            for pair in range(2):
                st["f"] = self.sparse(other(), st["f"], lc[2 * pair], lc[2 * pair + 1])

        def dbl_step():
            st["f"] = self.sqr(other(), st["f"], extra=self.line_eval())
            b.ctl_all(C_NEXTLINE)
            mul_lines()

        def add_step():
            b.op(self.line_eval())
            b.ctl_all(C_NEXTLINE)
            mul_lines()

        bits = self.miller_digits
        add_step()                       # f = 1 before the first doubling: its squaring is skipped
        if bits[0] == "1":
            add_step()
        i = 1
        while i < len(bits):
            run = 0
            while i + run < len(bits) and bits[i + run] == "0":
                run += 1
            if run >= 4:
                # a doubling step flips the slot, so REP over pairs of steps
                b.rep(run // 2)
                dbl_step()
                dbl_step()
                b.endrep()
                if run % 2:
                    dbl_step()
                i += run
                continue
            dbl_step()
            if bits[i] == "1":
                add_step()
            i += 1
        for _ in range(self.miller_tail):
            add_step()
        return st["f"]

    # -- final exponentiation ------------------------------------------------------------------------------------
    def take(self, *keep):
        """a slot not used by any of the values in `keep`"""
        used = {v.base for v in keep if v is not None}
        for s in self.SLOTS:
            if s not in used:
                return s
        raise ProgramError("out of slots")

    def pow_abs_x(self, base, keep=()):
        """base^|x| by cyclotomic squarings (base must be in the cyclotomic subgroup)."""
        b = self.b
        sa = self.take(base, *keep)
        sb = self.take(base, V12(sa), *keep)
        bits = self.ate_bits
        acc = base
        i = 0

        def nxt():
            return sb if acc.base == sa else sa

        while i < len(bits):
            run = 0
            while i + run < len(bits) and bits[i + run] == "0":
                run += 1
            if run >= 4 and acc.base in (sa, sb):
                b.rep(run // 2)
                acc = self.cyc_sqr(nxt(), acc)
                acc = self.cyc_sqr(nxt(), acc)
                b.endrep()
                if run % 2:
                    acc = self.cyc_sqr(nxt(), acc)
                i += run
                continue
            acc = self.cyc_sqr(nxt(), acc)
            if bits[i] == "1":
                acc = self.mul(nxt(), acc, base)
            i += 1
        return acc

    def fp_inverse(self, xcell):
        """role 0: cell[xcell].c0 = cell[xcell].c0^-1 in place (one INV instruction: Fermat inversion on registers,
        field.cuh fe_inv, instead of ~500 interpreted multiply + reduce steps); no barrier."""
        self.b.streams[0].append(enc_ctl(C_INV, xcell))

    def final_exp(self, f):
        b = self.b
        assert not f.conj
        # ---- easy part: g = f^(p^6 - 1) = conj(f) * f^-1 ; e = g^(p^2 + 1) --------------------------------------
        sn = self.take(f)
        n = self.mul(sn, f, conj(f), roles=(0, 2, 4))              # N = f conj(f) in Fp6: n0, n1, n2 at w^0, w^2, w^4
        n0, n1, n2 = sn, sn + 2, sn + 4
        t0, t1, t2 = sn + 1, sn + 3, sn + 5
        d0, d1, d2 = dict(dst=t0, eps=[]), dict(dst=t1, eps=[]), dict(dst=t2, eps=[])
        self.fsqr(d0, n0); self.fmul(d0, n1, n2, sign=-1, xi=True)
        self.fsqr(d1, n2, xi=True); self.fmul(d1, n0, n1, sign=-1)
        self.fsqr(d2, n1); self.fmul(d2, n0, n2, sign=-1)
        b.op({0: d0, 1: d1, 2: d2})
        sx = self.take(f, V12(sn))
        sy = self.take(f, V12(sn), V12(sx))
        D, DN, DI = sx, sx + 1, sx + 3
        dd = dict(dst=D, eps=[])
        self.fmul(dd, n0, t0); self.fmul(dd, n2, t1, xi=True); self.fmul(dd, n1, t2, xi=True)
        b.op({0: dd}, bar=False)
        b.op({0: dict(dst=DN, eps=[(1, 0, Operand(D, F_C0), Operand(D, F_C0)), (1, 0, Operand(D, F_C1), Operand(D, F_C1))],
                      fp_only=True)}, bar=False)
        self.fp_inverse(DN)
        b.op({0: dict(dst=DI, eps=[(1, 0, Operand(D, F_C0), Operand(DN, F_C0)), (0, -1, Operand(D, F_C1), Operand(DN, F_C0))])})
        # N^-1 = (t0, t1, t2) / d   ->  even cells of slot sy
        b.op({k: dict(dst=sy + 2 * k, eps=fp2_mul_eps((t0, t1, t2)[k], DI)) for k in range(3)})
        ninv = V12(sy)
        # h = conj(f) * N^-1 goes to slot sn (N, t are dead after the barrier above)
        h = self.mul(sn, conj(f), ninv, b_coeffs=(0, 2, 4))        # = f^-1
        g = self.mul(sx, conj(f), h)
        g2 = self.frob(self.take(g), g, 2)
        e = self.mul(self.take(g, g2), g2, g)
        r = self.hard_bn(e) if self.cv.twist == "D" else self.hard_bls12(e)
        b.ctl_roles(C_CHECK, [r.cell(k) for k in range(ROLES)])
        return r

    def gsave(self, v, slot=0):
        self.b.ctl_roles(C_GSAVE, [v.cell(k) | (slot << 8) for k in range(ROLES)])

    def gload(self, base, slot=0, conj_flag=False):
        self.b.ctl_roles(C_GLOAD, [(base + k) | (slot << 8) for k in range(ROLES)])
        self.b.ctl_all(C_BAR)
        return V12(base, conj_flag)

    def hard_bn(self, f):
        """BN, x = t > 0: the Fuentes-Castaneda et al. chain (pairing.cuh final_exp_hard<Bn>, the one ark-ec's bn model
        uses): f^(t-exponents) with three parked values (f, y1, y3) so that four resident slots suffice."""
        assert not f.conj
        self.gsave(f, 0)
        t = self.pow_abs_x(f)
        y0 = conj(t)
        y1 = self.cyc_sqr(self.take(f, y0), y0)
        y2 = self.cyc_sqr(self.take(f, y1), y1)
        y3 = self.mul(self.take(f, y1, y2), y2, y1)
        self.gsave(y1, 1)                        # (the conj flag of a parked value is carried by the program, see gload)
        y1_conj = y1.conj
        t = self.pow_abs_x(y3)
        y4 = conj(t)
        y5 = self.cyc_sqr(self.take(y3, y4), y4)
        self.gsave(y3, 2)
        y3_conj = y3.conj
        t = self.pow_abs_x(y5, keep=(y4,))
        y6 = conj(t)
        y7 = self.mul(self.take(y4, y6), conj(y6), y4)
        y3 = self.gload(self.take(y4, y7), 2, y3_conj)
        y8 = self.mul(self.take(y4, y7, y3), y7, conj(y3))
        y1 = self.gload(self.take(y4, y8), 1, y1_conj)
        y9 = self.mul(self.take(y4, y8, y1), y8, y1)
        y10 = self.mul(self.take(y4, y8, y9), y8, y4)
        y8f = self.frob(self.take(y8, y9, y10), y8, 2)
        a = self.mul(self.take(y8f, y9, y10), y8f, y10)
        f = self.gload(self.take(a, y9), 0)
        bb = self.mul(self.take(a, y9, f), a, f)
        c = self.mul(self.take(bb, y9, f), conj(f), y9)
        y12 = self.frob(self.take(bb, c, y9), y9, 1)
        dd = self.mul(self.take(bb, c, y12), y12, bb)
        y15 = self.frob(self.take(c, dd), c, 3)
        return self.mul(self.take(y15, dd), y15, dd)

    def hard_bls12(self, e):
        b = self.b
        # ---- hard part (BLS12, x < 0):  3 (p^4 - p^2 + 1) / r = (x-1)^2 (x+p) (x^2+p^2-1) + 3 ---------------------
        b.ctl_roles(C_GSAVE, [e.cell(k) for k in range(ROLES)])
        t = self.pow_abs_x(e)
        a = conj(self.mul(self.take(t, e), t, e))                   # e^x * e^-1 = conj(e^|x| * e)
        t = self.pow_abs_x(a)
        a = conj(self.mul(self.take(t, a), t, a))                   # a^(x-1)
        t = self.pow_abs_x(a)
        c = self.frob(self.take(t, a), a, 1)
        a = self.mul(self.take(t, c), conj(t), c)                   # a^(x+p)
        t = self.pow_abs_x(a)
        t = self.pow_abs_x(conj(t), keep=(a,))                      # a^(x^2)   (conj twice = identity up to the flag)
        t = conj(t)
        c = self.frob(self.take(t, a), a, 2)
        bb = self.mul(self.take(t, a, c), t, c)
        a = self.mul(self.take(bb, a), bb, conj(a))                 # a^(x^2 + p^2 - 1)
        s = self.take(a)
        b.ctl_roles(C_GLOAD, [s + k for k in range(ROLES)])
        b.ctl_all(C_BAR)
        e = V12(s)
        e2 = self.cyc_sqr(self.take(a, e), e)
        e3 = self.mul(self.take(a, e, e2), e2, e)
        return self.mul(self.take(a, e3), a, e3)

    def build(self):
        f = self.miller()
        self.final_exp(f)
        return self.b.finish()


# ---- emulator ------------------------------------------------------------------------------------------------
class Emulator:
    """Executes the streams on integers exactly as the kernel does (Montgomery domain, canonical cells)."""

    def __init__(self, curve, streams, consts, lines):
        """consts: id -> (c0, c1) normal-domain ints; lines: list over line index of 4 Fp2 tuples."""
        self.cv = curve
        self.streams = streams
        p, R = curve.p, curve.R
        self.mont = lambda v: (v * R) % p
        self.consts = {k: (self.mont(v[0]), self.mont(v[1])) for k, v in consts.items()}
        self.lines = [[(self.mont(c[0]), self.mont(c[1])) for c in ln] for ln in lines]
        self.cells = {}
        self.gscratch = {}
        self.kp = [int(m * p * p) for m in curve.kp_mult]
        assert all(k == m * p * p for k, m in zip(self.kp, curve.kp_mult))
        self.max_out = 0
        self.n_intervals = 0

    def set_cell(self, c, v):            # v normal-domain (c0, c1)
        self.cells[c] = (self.mont(v[0]), self.mont(v[1]))

    def get_cell(self, c):
        p, R = self.cv.p, self.cv.R
        ri = pow(R, -1, p)
        v = self.cells[c]
        return (v[0] * ri % p, v[1] * ri % p)

    def _operand(self, o, line_ctr, reads):
        p = self.cv.p
        if o.glob:
            v = self.lines[line_ctr][o.cell - LINE_BASE] if o.cell >= LINE_BASE else self.consts[o.cell]
        else:
            v = self.cells[o.cell]
            reads.add(o.cell)
        x = {F_C0: v[0], F_C1: v[1], F_SUM: v[0] + v[1], F_DIFF: v[0] - v[1] + p}[o.form] << o.shift
        assert 0 <= x < self.cv.R
        return x

    def _redc(self, T):
        p, R = self.cv.p, self.cv.R
        assert 0 <= T < self.cv.acc_limit * p * p, "accumulator out of range"
        m = (-T * pow(p, -1, R)) % R
        return (T + m * p) // R

    def run(self, skip=(False, False)):
        p, R = self.cv.p, self.cv.R
        st = [dict(pc=0, stack=[], line=0, R=0, I=0, done=False, check=None) for _ in range(ROLES)]
        while not all(s["done"] for s in st):
            reads = [set() for _ in range(ROLES)]
            writes = [set() for _ in range(ROLES)]
            pending = [[] for _ in range(ROLES)]
            for r in range(ROLES):
                s = st[r]
                prog = self.streams[r]
                while not s["done"]:
                    ins = dec(prog[s["pc"]])
                    s["pc"] += 1
                    if ins[0] == "EP":
                        _, sR, sI, x, y = ins
                        pr = self._operand(x, s["line"], reads[r]) * self._operand(y, s["line"], reads[r])
                        s["R"] += sR * pr
                        s["I"] += sI * pr
                    elif ins[0] == "XI":
                        c = self.cv.xi_c
                        s["R"], s["I"] = c * s["R"] - s["I"], c * s["I"] + s["R"]
                        lim = 1 << (32 * (2 * self.cv.n_limbs + 1) - 1)
                        assert -lim <= s["R"] < lim and -lim <= s["I"] < lim
                    elif ins[0] == "FIN":
                        d = ins[1]
                        out = []
                        z = None
                        if d["zsign"]:
                            z = self.cells[d["zcell"]]
                            reads[r].add(d["zcell"])
                        for comp, acc in ((0, s["R"]), (1, s["I"])):
                            if comp == 1 and d["fp_only"]:
                                out.append(0)
                                continue
                            T = acc * (3 if d["triple"] else 1)
                            if z is not None:
                                T += d["zsign"] * (z[comp] << d["zdouble"]) * R
                            T += self.kp[d["kp"]]
                            o = self._redc(T)
                            lim = (1 << self.cv.canon_steps[d["canon"]]) * p
                            assert o < lim, "canonicalisation depth too small"
                            self.max_out = max(self.max_out, o / p)
                            out.append(o % p)
                        if d["skip"] & 1 and skip[(d["skip"] >> 1) & 1]:
                            out = [0, 0]
                        self.cells[d["dst"]] = tuple(out)
                        writes[r].add(d["dst"])
                        s["R"] = s["I"] = 0
                        if d["bar"]:
                            break
                    else:
                        _, sub, arg = ins
                        if sub == C_END:
                            s["done"] = True
                        elif sub == C_REP:
                            s["stack"].append([s["pc"], arg])
                        elif sub == C_ENDREP:
                            s["stack"][-1][1] -= 1
                            if s["stack"][-1][1] > 0:
                                s["pc"] = s["stack"][-1][0]
                            else:
                                s["stack"].pop()
                        elif sub == C_NEXTLINE:
                            s["line"] += 1
                        elif sub == C_BAR:
                            break
                        elif sub == C_GSAVE:
                            self.gscratch[(r, arg >> 8)] = self.cells[arg & 255]
                            reads[r].add(arg & 255)
                        elif sub == C_GLOAD:
                            self.cells[arg & 255] = self.gscratch[(r, arg >> 8)]
                            writes[r].add(arg & 255)
                        elif sub == C_INV:
                            v = self.cells[arg]
                            reads[r].add(arg)
                            writes[r].add(arg)
                            ri = pow(R, -1, p)
                            xn = v[0] * ri % p
                            self.cells[arg] = (pow(xn, p - 2, p) * R % p, v[1])
                        elif sub == C_CHECK:
                            v = self.cells[arg]
                            reads[r].add(arg)
                            one = self.mont(1)
                            s["check"] = (v == ((one, 0) if r == 0 else (0, 0)))
                            break        # CHECK contains a barrier
                        else:
                            raise ValueError(sub)
            # hazard check of this barrier interval.  Roles run sequentially here, so a cross-role conflict would
            # silently read a value the GPU may or may not see: forbid every cross-role read/write overlap.
            for a in range(ROLES):
                for c in range(ROLES):
                    if a != c and (writes[a] & (reads[c] | writes[c])):
                        raise ProgramError(f"barrier hazard between roles {a} and {c}: cells {writes[a] & (reads[c] | writes[c])}")
            self.n_intervals += 1
        assert len({s["line"] for s in st}) == 1
        return all(s["check"] for s in st)
