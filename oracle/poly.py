"""Radix-2 evaluation domains and dense polynomial helpers over a prime field (Python ints).

Restates the parts of ark-poly 0.3 the prover calls [U ark-poly src/domain/radix2/*,
src/polynomial/univariate/dense.rs] (call sites: reference src/ahp/prover.rs,
src/ahp/mod.rs:301-328, src/ahp/constraint_systems.rs:234-239).  Polynomials are coefficient
lists, lowest degree first, trailing zeros stripped exactly where `from_coefficients_vec` does.
"""


class Domain:
    """`Radix2EvaluationDomain::new(num_coeffs)`: size = next power of two,
    group_gen = TWO_ADIC_ROOT^(2^(TWO_ADICITY - log_size))."""

    def __init__(self, field, num_coeffs):
        size = 1
        log = 0
        while size < num_coeffs:
            size *= 2
            log += 1
        if log > field.two_adicity:
            raise ValueError("PolynomialDegreeTooLarge")
        self.f = field
        self.p = field.p
        self.size = size
        self.log_size = log
        g = field.two_adic_root
        for _ in range(field.two_adicity - log):
            g = g * g % self.p
        self.group_gen = g
        self.group_gen_inv = pow(g, -1, self.p)
        self.size_as_field_element = size % self.p
        self.size_inv = pow(size, -1, self.p)
        self._elems = None

    def elements(self):
        if self._elems is None:
            out = [1] * self.size
            for i in range(1, self.size):
                out[i] = out[i - 1] * self.group_gen % self.p
            self._elems = out
        return self._elems

    def element(self, i):
        return pow(self.group_gen, i, self.p)

    def evaluate_vanishing_polynomial(self, tau):
        return (pow(tau, self.size, self.p) - 1) % self.p

    # -- transforms -------------------------------------------------------------------------
    def _ntt(self, a, root):
        p = self.p
        n = self.size
        a = list(a) + [0] * (n - len(a))
        # bit reversal then iterative DIT; natural order in, natural order out
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            w_len = pow(root, n // length, p)
            half = length // 2
            ws = [1] * half
            for k in range(1, half):
                ws[k] = ws[k - 1] * w_len % p
            for start in range(0, n, length):
                for k in range(half):
                    u = a[start + k]
                    v = a[start + k + half] * ws[k] % p
                    a[start + k] = (u + v) % p
                    a[start + k + half] = (u - v) % p
            length *= 2
        return a

    def fft(self, coeffs):
        """evals[i] = sum_j coeffs[j] * group_gen^(i*j); input zero-padded / must fit the domain."""
        assert len(coeffs) <= self.size
        return self._ntt(coeffs, self.group_gen)

    def ifft(self, evals):
        assert len(evals) <= self.size
        out = self._ntt(evals, self.group_gen_inv)
        return [x * self.size_inv % self.p for x in out]

    def coset_fft(self, coeffs):
        g = self.f.generator
        p = self.p
        cur = 1
        sc = []
        for c in coeffs:
            sc.append(c * cur % p)
            cur = cur * g % p
        return self.fft(sc)

    def coset_ifft(self, evals):
        p = self.p
        ginv = pow(self.f.generator, -1, p)
        out = self.ifft(evals)
        cur = 1
        for i in range(len(out)):
            out[i] = out[i] * cur % p
            cur = cur * ginv % p
        return out

    def reindex_by_subdomain(self, other, index):
        """[U ark-poly EvaluationDomain::reindex_by_subdomain]; see SURVEY.md A.5."""
        assert self.size >= other.size
        period = self.size // other.size
        if index < other.size:
            return index * period
        i = index - other.size
        x = period - 1
        return i + (i // x) + 1

    def evaluate_all_lagrange_coefficients(self, tau):
        """[U ark-poly]: L_i(tau) for all i; special-cases tau in the domain."""
        p = self.p
        n = self.size
        t_size = pow(tau, n, p)
        if t_size == 1:
            u = [0] * n
            omega_i = 1
            for i in range(n):
                if omega_i == tau:
                    u[i] = 1
                    break
                omega_i = omega_i * self.group_gen % p
            return u
        # v_0 = (tau^n - 1)/n ; L_i = v_0 * g^i / (tau - g^i)
        l = (t_size - 1) * self.size_inv % p
        r = 1
        u = [0] * n
        ls = [0] * n
        for i in range(n):
            u[i] = (tau - r) % p
            ls[i] = l
            l = l * self.group_gen % p
            r = r * self.group_gen % p
        inv = batch_inversion(u, p)
        return [ls[i] * inv[i] % p for i in range(n)]


def batch_inversion(v, p):
    """Montgomery's trick; zeros are left untouched [U ark-ff batch_inversion]."""
    out = list(v)
    prod = []
    acc = 1
    for x in v:
        if x != 0:
            acc = acc * x % p
            prod.append(acc)
    inv = pow(acc, -1, p)
    nz = [i for i, x in enumerate(v) if x != 0]
    for k in range(len(nz) - 1, -1, -1):
        i = nz[k]
        prev = prod[k - 1] if k > 0 else 1
        out[i] = inv * prev % p
        inv = inv * v[i] % p
    return out


def strip(c):
    """`DensePolynomial::from_coefficients_vec`: drop trailing zero coefficients."""
    n = len(c)
    while n > 0 and c[n - 1] == 0:
        n -= 1
    return c[:n]


def degree(c):
    c = strip(c)
    return len(c) - 1 if c else 0


def evaluate(c, x, p):
    acc = 0
    for a in reversed(c):
        acc = (acc * x + a) % p
    return acc


def poly_add(a, b, p):
    n = max(len(a), len(b))
    out = [0] * n
    for i, x in enumerate(a):
        out[i] = x
    for i, x in enumerate(b):
        out[i] = (out[i] + x) % p
    return strip(out)


def poly_sub(a, b, p):
    n = max(len(a), len(b))
    out = [0] * n
    for i, x in enumerate(a):
        out[i] = x
    for i, x in enumerate(b):
        out[i] = (out[i] - x) % p
    return strip(out)


def poly_scale(a, k, p):
    return strip([x * k % p for x in a])


def poly_mul(field, a, b):
    """`&a * &b` for DensePolynomial: FFT on the domain of size len(a)+len(b)-1 [U dense.rs Mul]."""
    if not a or not b:
        return []
    d = Domain(field, len(a) + len(b) - 1)
    ea = d.fft(a)
    eb = d.fft(b)
    p = field.p
    return strip(d.ifft([x * y % p for x, y in zip(ea, eb)]))


def divide_by_vanishing_poly(c, domain):
    """(q, r) with c = q * (X^n - 1) + r  [U dense.rs divide_by_vanishing_poly]."""
    p = domain.p
    n = domain.size
    if len(c) < n:
        return [], strip(list(c))
    q = list(c[n:])
    for i in range(len(q) - n - 1, -1, -1):
        q[i] = (q[i] + q[i + n]) % p
    r = list(c[:n])
    for i in range(min(n, len(q))):
        r[i] = (r[i] + q[i]) % p
    return strip(q), strip(r)


def mul_by_vanishing_poly(c, domain):
    """c * (X^n - 1)  [U dense.rs mul_by_vanishing_poly]."""
    p = domain.p
    n = domain.size
    out = [0] * n + list(c)
    for i, x in enumerate(c):
        out[i] = (out[i] - x) % p
    return strip(out)


def divide_by_linear(c, z, p):
    """(q, rem) with c = q * (X - z) + rem: the `p / &divisor` of KZG10::compute_witness_polynomial."""
    if len(c) <= 1:
        return [], (c[0] if c else 0)
    q = [0] * (len(c) - 1)
    acc = 0
    for i in range(len(c) - 1, 0, -1):
        acc = (c[i] + acc * z) % p
        q[i - 1] = acc
    rem = (c[0] + acc * z) % p
    return strip(q), rem
