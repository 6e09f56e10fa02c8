"""BLS12-381 pairing for the oracle's verifier: the reference's
`PC::check_combinations` ends in `KZG10::check`, a two-pairing product [U ark-poly-commit kzg10::check],
and this module lets the oracle run that check without the SRS trapdoor.

Deliberately simple rather than fast: Fq12 is the single extension Fq[w] / (w^12 - 2 w^6 + 2) (w^6 = 1 + u,
u^2 = -1), G2 points are carried straight in E(Fq12) coordinates through the untwisting map
(x', y') -> (x' / w^2, y' / w^3), and the pairing is the reduced Tate pairing
    e(P, Q) = f_{r,P}(psi(Q)) ^ ((p^12 - 1) / r)
with lines through multiples of P (slopes in Fq) and denominator elimination.  Bilinearity and
non-degeneracy are asserted by tests/test_oracle.py.  TEST INFRASTRUCTURE only.
"""
from .params import BLS12_381

P = BLS12_381.fq.p
R = BLS12_381.fr.p
DEG = 12
# w^12 = 2 w^6 - 2


class Fq12:
    __slots__ = ("c",)

    def __init__(self, coeffs):
        self.c = [x % P for x in coeffs] + [0] * (DEG - len(coeffs))

    @staticmethod
    def from_fq(a):
        return Fq12([a])

    @staticmethod
    def from_fq2(a, b):
        """a + b u with u = w^6 - 1"""
        return Fq12([a - b, 0, 0, 0, 0, 0, b])

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
        # reduce with w^12 = 2 w^6 - 2, highest degree first
        for k in range(2 * DEG - 2, DEG - 1, -1):
            v = t[k]
            if v:
                t[k - 6] += 2 * v
                t[k - 12] -= 2 * v
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
        mod = [2, 0, 0, 0, 0, 0, -2 % P, 0, 0, 0, 0, 0, 1]

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


W = Fq12([0, 1])
W2_INV = (W * W).inv()
W3_INV = (W * W * W).inv()
B12 = Fq12.from_fq(4)


# ---- E(Fq12): y^2 = x^3 + 4, affine, None = infinity ------------------------------------------------------------
def e12_add(A, B):
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


def e12_neg(A):
    return None if A is None else (A[0], -A[1])


def e12_mul(k, A):
    acc = None
    for bit in bin(k)[2:]:
        acc = e12_add(acc, acc)
        if bit == "1":
            acc = e12_add(acc, A)
    return acc


def e12_on_curve(A):
    return A is None or A[1].square() == A[0].square() * A[0] + B12


def untwist(x2, y2):
    """psi: E'(Fq2): y^2 = x^3 + 4(1 + u)  ->  E(Fq12);  x2, y2 are (a, b) pairs meaning a + b u."""
    return (Fq12.from_fq2(*x2) * W2_INV, Fq12.from_fq2(*y2) * W3_INV)


def _fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _fq2_sqrt(a):
    """square root in Fq2 = Fq[u]/(u^2+1), p = 3 mod 4 (complex method); None if a is not a square"""
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
            if _fq2_mul((x, y), (x, y)) == (a0 % P, a1 % P):
                return (x, y)
    return None


_G2_COFACTOR = 0x5d543a95414e7f1091d50792876a202cd91de4547085abaa68a205b2e5a7ddfa628f1cb4d9e82ef21537e293a6691ae1616ec6e786f0c70cf1c38e31c7238e5


def g2_generator():
    """A point of order r on the twist, found deterministically: smallest x = (k, 1) with a square right-hand side,
    cofactor-cleared.  (Any order-r point serves as `h` of a KZG SRS; no external constant is trusted.)"""
    k = 0
    while True:
        x = (k, 1)
        x3 = _fq2_mul(_fq2_mul(x, x), x)
        rhs = ((x3[0] + 4) % P, (x3[1] + 4) % P)  # + 4 (1 + u)
        y = _fq2_sqrt(rhs)
        if y is not None:
            Q = e12_mul(_G2_COFACTOR, untwist(x, y))
            if Q is not None:
                assert e12_on_curve(Q) and e12_mul(R, Q) is None
                return Q
        k += 1


def g1_to_e12(Pt):
    return None if Pt is None else (Fq12.from_fq(Pt[0]), Fq12.from_fq(Pt[1]))


def miller_loop(Pt, Q):
    """f_{r,P}(Q) for P in E(Fq) (affine ints), Q in E(Fq12); lines have slopes in Fq."""
    if Pt is None or Q is None:
        return Fq12.one()
    xq, yq = Q
    xp, yp = Pt
    f = Fq12.one()
    tx, ty = xp, yp
    bits = bin(R)[3:]
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


FINAL_EXP = (P ** 12 - 1) // R


def pairing(Pt, Q):
    return miller_loop(Pt, Q).pow(FINAL_EXP)


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 with a single final exponentiation"""
    f = Fq12.one()
    for Pt, Q in pairs:
        f = f * miller_loop(Pt, Q)
    return f.pow(FINAL_EXP) == Fq12.one()
