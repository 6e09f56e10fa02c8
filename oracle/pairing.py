"""Pairings for the oracle's verifier: the reference's `PC::check_combinations` ends in
`KZG10::check`, a two-pairing product [U ark-poly-commit kzg10::check], and this module lets the oracle run that
check without the SRS trapdoor -- for BLS12-381 (the reference's test curve) and BN254 (BASELINE config 4).

Deliberately simple rather than fast, and the same code for both curves.  With Fq2 = Fq[u]/(u^2 + 1) and the
sextic non-residue xi = alpha + u (alpha = 1 for BLS12-381, 9 for BN254), Fq12 is the single extension
    Fq[w] / (w^12 - 2 alpha w^6 + alpha^2 + 1)            (w^6 = xi, u = w^6 - alpha),
G2 points are carried straight in E(Fq12) coordinates through the untwisting map
    M-type twist y^2 = x^3 + b xi  (BLS12-381):  (x', y') -> (x' / w^2, y' / w^3)
    D-type twist y^2 = x^3 + b / xi (BN254):      (x', y') -> (x' w^2,  y' w^3)
and the pairing is the reduced Tate pairing
    e(P, Q) = f_{r,P}(psi(Q)) ^ ((p^12 - 1) / r)
with lines through multiples of P (slopes in Fq) and denominator elimination.  No external G2 constant is
trusted: the generator is found deterministically and its order is asserted.  Bilinearity and non-degeneracy
are asserted by tests/test_oracle.py.  TEST INFRASTRUCTURE only.
"""
from .params import BLS12_381, BN254

DEG = 12


class PairingEngine:
    """Fq12 arithmetic, E(Fq12) group law and the reduced Tate pairing of one curve."""

    def __init__(self, curve, alpha, twist, g2_cofactor):
        self.curve = curve
        self.P = P = curve.fq.p
        self.R = curve.fr.p
        self.alpha = alpha
        self.twist = twist  # "M" or "D"
        self.g2_cofactor = g2_cofactor
        self.c6 = 2 * alpha % P          # w^12 = c6 w^6 - c0
        self.c0 = (alpha * alpha + 1) % P
        eng = self

        class Fq12:
            __slots__ = ("c",)

            def __init__(self, coeffs):
                self.c = [x % P for x in coeffs] + [0] * (DEG - len(coeffs))

            @staticmethod
            def from_fq(a):
                return Fq12([a])

            @staticmethod
            def from_fq2(a, b):
                """a + b u with u = w^6 - alpha"""
                return Fq12([a - alpha * b, 0, 0, 0, 0, 0, b])

            @staticmethod
            def one():
                return Fq12([1])

            @staticmethod
            def zero():
                return Fq12([0])

            def __eq__(self, o):
                return self.c == o.c

            def is_zero(self):
                return not any(self.c)

            def __add__(self, o):
                return Fq12([x + y for x, y in zip(self.c, o.c)])

            def __sub__(self, o):
                return Fq12([x - y for x, y in zip(self.c, o.c)])

            def __neg__(self):
                return Fq12([-x for x in self.c])

            def scale(self, k):
                return Fq12([x * k for x in self.c])

            def __mul__(self, o):
                a, b = self.c, o.c
                t = [0] * (2 * DEG - 1)
                for i, x in enumerate(a):
                    if x:
                        for j, y in enumerate(b):
                            t[i + j] += x * y
                # reduce with w^12 = c6 w^6 - c0, highest degree first
                for k in range(2 * DEG - 2, DEG - 1, -1):
                    v = t[k]
                    if v:
                        t[k - 6] += eng.c6 * v
                        t[k - 12] -= eng.c0 * v
                return Fq12(t[:DEG])

            def square(self):
                return self * self

            def pow(self, e):
                r = Fq12.one()
                base = self
                while e:
                    if e & 1:
                        r = r * base
                    base = base * base
                    e >>= 1
                return r

            def inv(self):
                """polynomial extended Euclid over Fq against the modulus"""
                mod = [eng.c0, 0, 0, 0, 0, 0, -eng.c6 % P, 0, 0, 0, 0, 0, 1]

                def deg(p):
                    d = len(p) - 1
                    while d >= 0 and p[d] == 0:
                        d -= 1
                    return d

                lm, hm = [1] + [0] * DEG, [0] * (DEG + 1)
                low, high = list(self.c) + [0], list(mod)
                while deg(low) > 0:
                    # high = high - q * low with polynomial long division
                    r = list(high)
                    q = [0] * (DEG + 1)
                    dl = deg(low)
                    inv_lead = pow(low[dl], -1, P)
                    for i in range(deg(r) - dl, -1, -1):
                        coef = r[dl + i] * inv_lead % P
                        q[i] = coef
                        if coef:
                            for j in range(dl + 1):
                                r[i + j] = (r[i + j] - coef * low[j]) % P
                    nm = list(hm)
                    for i, qi in enumerate(q):
                        if qi:
                            for j, lj in enumerate(lm):
                                if lj and i + j <= DEG:
                                    nm[i + j] = (nm[i + j] - qi * lj) % P
                    lm, low, hm, high = nm, r, lm, low
                assert deg(low) == 0, "not invertible"
                k = pow(low[0], -1, P)
                return Fq12([x * k for x in lm[:DEG]])

        self.Fq12 = Fq12
        w = Fq12([0, 1])
        w2, w3 = w * w, w * w * w
        # untwisting factors: divide (M) or multiply (D) by w^2 / w^3
        self.fx, self.fy = (w2.inv(), w3.inv()) if twist == "M" else (w2, w3)
        self.B12 = Fq12.from_fq(curve.b)
        self.final_exp = (P ** 12 - 1) // self.R
        self._g2 = None

    # ---- E(Fq12): y^2 = x^3 + b, affine, None = infinity ----------------------------------------------------------
    def e12_add(self, A, B):
        if A is None:
            return B
        if B is None:
            return A
        x1, y1 = A
        x2, y2 = B
        if x1 == x2:
            if (y1 + y2).is_zero():
                return None
            lam = x1.square().scale(3) * (y1.scale(2)).inv()
        else:
            lam = (y2 - y1) * (x2 - x1).inv()
        x3 = lam.square() - x1 - x2
        return (x3, lam * (x1 - x3) - y1)

    @staticmethod
    def e12_neg(A):
        return None if A is None else (A[0], -A[1])

    def e12_mul(self, k, A):
        acc = None
        for bit in bin(k)[2:]:
            acc = self.e12_add(acc, acc)
            if bit == "1":
                acc = self.e12_add(acc, A)
        return acc

    def e12_on_curve(self, A):
        return A is None or A[1].square() == A[0].square() * A[0] + self.B12

    def untwist(self, x2, y2):
        """psi: E'(Fq2) -> E(Fq12);  x2, y2 are (a, b) pairs meaning a + b u."""
        return (self.Fq12.from_fq2(*x2) * self.fx, self.Fq12.from_fq2(*y2) * self.fy)

    # ---- Fq2 helpers for the generator search -----------------------------------------------------------------------
    def _fq2_mul(self, a, b):
        P = self.P
        return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)

    def _fq2_inv(self, a):
        P = self.P
        n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
        return (a[0] * n % P, -a[1] * n % P)

    def _fq2_sqrt(self, a):
        """square root in Fq2 = Fq[u]/(u^2+1), p = 3 mod 4 (complex method); None if a is not a square"""
        P = self.P
        a0, a1 = a
        if a1 == 0:
            s = pow(a0, (P + 1) // 4, P)
            if s * s % P == a0 % P:
                return (s, 0)
            s = pow(-a0 % P, (P + 1) // 4, P)
            return (0, s) if s * s % P == -a0 % P else None
        norm = (a0 * a0 + a1 * a1) % P
        n = pow(norm, (P + 1) // 4, P)
        if n * n % P != norm:
            return None
        for sign in (1, -1):
            half = (a0 + sign * n) * pow(2, -1, P) % P
            x = pow(half, (P + 1) // 4, P)
            if x * x % P == half and x:
                y = a1 * pow(2 * x, -1, P) % P
                if self._fq2_mul((x, y), (x, y)) == (a0 % P, a1 % P):
                    return (x, y)
        return None

    def g2_generator(self):
        """A point of order r on the twist, found deterministically: smallest x = (k, 1) with a square right-hand
        side, cofactor-cleared.  (Any order-r point serves as `h` of a KZG SRS.)"""
        if self._g2 is not None:
            return self._g2
        assert self.P % 4 == 3
        xi = (self.alpha, 1)
        b = self.curve.b
        bt = self._fq2_mul((b, 0), xi) if self.twist == "M" else self._fq2_mul((b, 0), self._fq2_inv(xi))
        k = 0
        while True:
            x = (k, 1)
            x3 = self._fq2_mul(self._fq2_mul(x, x), x)
            rhs = ((x3[0] + bt[0]) % self.P, (x3[1] + bt[1]) % self.P)
            y = self._fq2_sqrt(rhs)
            if y is not None:
                Q = self.e12_mul(self.g2_cofactor, self.untwist(x, y))
                if Q is not None:
                    assert self.e12_on_curve(Q) and self.e12_mul(self.R, Q) is None
                    self._g2 = Q
                    return Q
            k += 1

    def g1_to_e12(self, Pt):
        return None if Pt is None else (self.Fq12.from_fq(Pt[0]), self.Fq12.from_fq(Pt[1]))

    def miller_loop(self, Pt, Q):
        """f_{r,P}(Q) for P in E(Fq) (affine ints), Q in E(Fq12); lines have slopes in Fq."""
        Fq12, P = self.Fq12, self.P
        if Pt is None or Q is None:
            return Fq12.one()
        xq, yq = Q
        xp, yp = Pt
        f = Fq12.one()
        tx, ty = xp, yp
        bits = bin(self.R)[3:]
        t_inf = False
        for bit in bits:
            # doubling step: tangent at T
            lam = 3 * tx * tx * pow(2 * ty, -1, P) % P
            line = (yq - Fq12.from_fq(ty)) - (xq - Fq12.from_fq(tx)).scale(lam)
            f = f.square() * line
            nx = (lam * lam - 2 * tx) % P
            ty = (lam * (tx - nx) - ty) % P
            tx = nx
            if bit == "1":
                if tx == xp:
                    # T = -P: vertical line, eliminated by the final exponentiation; T + P = infinity (last step only)
                    t_inf = True
                    continue
                lam = (ty - yp) * pow(tx - xp, -1, P) % P
                line = (yq - Fq12.from_fq(ty)) - (xq - Fq12.from_fq(tx)).scale(lam)
                f = f * line
                nx = (lam * lam - tx - xp) % P
                ty = (lam * (tx - nx) - ty) % P
                tx = nx
        assert t_inf, "Miller loop did not end at infinity (P not of order r?)"
        return f

    def pairing(self, Pt, Q):
        return self.miller_loop(Pt, Q).pow(self.final_exp)

    def pairing_product_is_one(self, pairs):
        """prod e(P_i, Q_i) == 1 with a single final exponentiation"""
        f = self.Fq12.one()
        for Pt, Q in pairs:
            f = f * self.miller_loop(Pt, Q)
        return f.pow(self.final_exp) == self.Fq12.one()


_BLS_G2_COFACTOR = 0x5d543a95414e7f1091d50792876a202cd91de4547085abaa68a205b2e5a7ddfa628f1cb4d9e82ef21537e293a6691ae1616ec6e786f0c70cf1c38e31c7238e5
_ENGINES = {}


def for_curve(curve):
    """the (cached) engine of a curve from oracle.params"""
    if curve.name not in _ENGINES:
        if curve.name == "bls12_381":
            _ENGINES[curve.name] = PairingEngine(BLS12_381, 1, "M", _BLS_G2_COFACTOR)
        elif curve.name == "bn254":
            # #E'(Fq2) = r (2p - r) for a BN curve
            _ENGINES[curve.name] = PairingEngine(BN254, 9, "D", 2 * BN254.fq.p - BN254.fr.p)
        else:
            raise ValueError("no pairing for " + curve.name)
    return _ENGINES[curve.name]


# ---- module-level BLS12-381 facade (the names tests/test_oracle.py uses) ---------------------------------------------
_bls = for_curve(BLS12_381)
P, R = _bls.P, _bls.R
Fq12 = _bls.Fq12
e12_add, e12_neg, e12_mul, e12_on_curve = _bls.e12_add, _bls.e12_neg, _bls.e12_mul, _bls.e12_on_curve
g2_generator, miller_loop, pairing, pairing_product_is_one = _bls.g2_generator, _bls.miller_loop, _bls.pairing, _bls.pairing_product_is_one
