"""G1 group law and multi-scalar multiplication, Python big integers.

Points are affine tuples (x, y) of canonical integers, or None for the point at infinity.
`msm_naive` is the definition (sum of scalar multiples).  `msm_pippenger_arkworks` restates
`VariableBaseMSM::multi_scalar_mul` [U ark-ec 0.3 src/msm/variable_base.rs] -- window
c = 3 if n < 32 else ln(n)+2 (ark-std `ln_without_floats` = floor(log2(n)*69/100)), unsigned digits,
2^c - 1 buckets per window, running-sum bucket reduction, Horner over windows -- so the CPU
baseline does the same work as the reference.  Both return the same (unique) group element.
"""


def _inv(x, p):
    return pow(x, -1, p)


def affine_add(curve, P, Q):
    p = curve.fq.p
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = 3 * x1 * x1 * _inv(2 * y1, p) % p
    else:
        lam = (y2 - y1) * _inv(x2 - x1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    y3 = (lam * (x1 - x3) - y1) % p
    return (x3, y3)


def affine_neg(curve, P):
    if P is None:
        return None
    return (P[0], (-P[1]) % curve.fq.p)


def on_curve(curve, P):
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - curve.b) % curve.fq.p == 0


# --- Jacobian arithmetic (a = 0), used for speed inside scalar_mul / MSM ------------------
def jac_double(p, P):
    X1, Y1, Z1 = P
    if Z1 == 0 or Y1 == 0:
        return (1, 1, 0)
    A = X1 * X1 % p
    B = Y1 * Y1 % p
    C = B * B % p
    D = 2 * ((X1 + B) * (X1 + B) - A - C) % p
    E = 3 * A % p
    F = E * E % p
    X3 = (F - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y1 * Z1 % p
    return (X3, Y3, Z3)


def jac_add(p, P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % p
    Z2Z2 = Z2 * Z2 % p
    U1 = X1 * Z2Z2 % p
    U2 = X2 * Z1Z1 % p
    S1 = Y1 * Z2 * Z2Z2 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if U1 == U2:
        if S1 == S2:
            return jac_double(p, P)
        return (1, 1, 0)
    H = (U2 - U1) % p
    R = (S2 - S1) % p
    HH = H * H % p
    HHH = H * HH % p
    V = U1 * HH % p
    X3 = (R * R - HHH - 2 * V) % p
    Y3 = (R * (V - X3) - S1 * HHH) % p
    Z3 = Z1 * Z2 * H % p
    return (X3, Y3, Z3)


def jac_add_mixed(p, P, Q):
    """P Jacobian + Q affine (x, y)."""
    X1, Y1, Z1 = P
    if Z1 == 0:
        return (Q[0], Q[1], 1)
    X2, Y2 = Q
    Z1Z1 = Z1 * Z1 % p
    U2 = X2 * Z1Z1 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if X1 == U2:
        if Y1 == S2:
            return jac_double(p, P)
        return (1, 1, 0)
    H = (U2 - X1) % p
    R = (S2 - Y1) % p
    HH = H * H % p
    HHH = H * HH % p
    V = X1 * HH % p
    X3 = (R * R - HHH - 2 * V) % p
    Y3 = (R * (V - X3) - Y1 * HHH) % p
    Z3 = Z1 * H % p
    return (X3, Y3, Z3)


def jac_to_affine(p, P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = _inv(Z, p)
    zi2 = zi * zi % p
    return (X * zi2 % p, Y * zi2 * zi % p)


JAC_INF = (1, 1, 0)


def scalar_mul(curve, k, P):
    """k * P (affine in, affine out)."""
    p = curve.fq.p
    k %= curve.fr.p
    if P is None or k == 0:
        return None
    acc = JAC_INF
    for bit in bin(k)[2:]:
        acc = jac_double(p, acc)
        if bit == "1":
            acc = jac_add_mixed(p, acc, P)
    return jac_to_affine(p, acc)


def msm_naive(curve, bases, scalars):
    """sum_i scalars[i] * bases[i] straight from the definition."""
    p = curve.fq.p
    acc = JAC_INF
    for P, k in zip(bases, scalars):
        Q = scalar_mul(curve, k, P)
        if Q is not None:
            acc = jac_add_mixed(p, acc, Q)
    return jac_to_affine(p, acc)


def ln_without_floats(a):
    """[U ark-std ln_without_floats]: log2(a) * ln(2), ln(2) ~ 0.69, integer arithmetic."""
    return (a.bit_length() - 1) * 69 // 100 if a > 1 else 0


def arkworks_window(n):
    return 3 if n < 32 else ln_without_floats(n) + 2


def msm_pippenger_arkworks(curve, bases, scalars):
    """Restatement of ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul` (see module docstring)."""
    p = curve.fq.p
    size = min(len(bases), len(scalars))
    pairs = [(bases[i], scalars[i]) for i in range(size) if scalars[i] != 0 and bases[i] is not None]
    c = arkworks_window(size)
    num_bits = curve.fr.bits
    window_sums = []
    for w_start in range(0, num_bits, c):
        res = JAC_INF
        buckets = [JAC_INF] * ((1 << c) - 1)
        for base, scalar in pairs:
            if scalar == 1:
                if w_start == 0:
                    res = jac_add_mixed(p, res, base)
            else:
                d = (scalar >> w_start) % (1 << c)
                if d != 0:
                    buckets[d - 1] = jac_add_mixed(p, buckets[d - 1], base)
        running = JAC_INF
        for b in reversed(buckets):
            running = jac_add(p, running, b)
            res = jac_add(p, res, running)
        window_sums.append(res)
    lowest = window_sums[0]
    total = JAC_INF
    for ws in reversed(window_sums[1:]):
        total = jac_add(p, total, ws)
        for _ in range(c):
            total = jac_double(p, total)
    total = jac_add(p, lowest, total)
    return jac_to_affine(p, total)


def fixed_base_powers(curve, g, beta, n):
    """[beta^i * g for i in range(n)] (what KZG10::setup computes with FixedBaseMSM)."""
    p = curve.fq.p
    r = curve.fr.p
    # 8-bit fixed-base table
    wbits = 8
    nwin = (curve.fr.bits + wbits - 1) // wbits
    table = []
    base = g
    for _ in range(nwin):
        row = [None]
        acc = None
        for _ in range((1 << wbits) - 1):
            acc = affine_add(curve, acc, base)
            row.append(acc)
        table.append(row)
        for _ in range(wbits):
            base = affine_add(curve, base, base)
    out = []
    cur = 1
    for _ in range(n):
        acc = JAC_INF
        k = cur
        for w in range(nwin):
            d = (k >> (w * wbits)) & ((1 << wbits) - 1)
            if d:
                acc = jac_add_mixed(p, acc, table[w][d])
        out.append(jac_to_affine(p, acc))
        cur = cur * beta % r
    return out
