"""The algebraic holographic proof for R1CS: indexer, the three prover rounds, the verifier's
challenges and the linear combinations -- a line-by-line restatement of the reference's
src/ahp/{indexer,constraint_systems,prover,verifier,mod}.rs in Python integers (canonical field
values; polynomials are coefficient lists, low degree first)."""
from . import poly as P
from .kzg import LabeledPoly, LinearCombination
from .rng import field_rand, poly_rand


class IndexInfo:
    def __init__(self, num_variables, num_constraints, num_non_zero, num_instance_variables):
        self.num_variables = num_variables
        self.num_constraints = num_constraints
        self.num_non_zero = num_non_zero
        self.num_instance_variables = num_instance_variables


def max_degree(field, num_constraints, num_variables, num_non_zero):
    """[R src/ahp/mod.rs:71-93]"""
    padded = max(num_variables, num_constraints)
    zk_bound = 1
    h = P.Domain(field, padded).size
    k = P.Domain(field, num_non_zero).size
    return max(2 * h + zk_bound - 2, 3 * h + 2 * zk_bound - 3, h, h, k - 1)


def get_degree_bounds(field, info):
    """[R src/ahp/mod.rs:96-106]"""
    return [P.Domain(field, info.num_constraints).size - 2, P.Domain(field, info.num_non_zero).size - 2]


def sum_matrices(a, b, c):
    """[R src/ahp/indexer.rs:83-102] sorted union of column indices per row."""
    return [sorted({i for _, i in ra} | {i for _, i in rb} | {i for _, i in rc}) for ra, rb, rc in zip(a, b, c)]


def u_h_same_inputs(domain):
    """batch_eval_unnormalized_bivariate_lagrange_poly_with_same_inputs [R src/ahp/mod.rs:320-327]"""
    p = domain.p
    elems = [e * domain.size_as_field_element % p for e in domain.elements()]
    return [elems[0]] + elems[1:][::-1]


def u_h_diff_inputs(domain, x):
    """batch_eval_unnormalized_bivariate_lagrange_poly_with_diff_inputs [R src/ahp/mod.rs:311-318]"""
    p = domain.p
    vanish_x = domain.evaluate_vanishing_polynomial(x)
    inv = P.batch_inversion([(x - y) % p for y in domain.elements()], p)
    return [d * vanish_x % p for d in inv]


def u_h(domain, x, y):
    """eval_unnormalized_bivariate_lagrange_poly [R src/ahp/mod.rs:302-309]"""
    p = domain.p
    if x != y:
        return (domain.evaluate_vanishing_polynomial(x) - domain.evaluate_vanishing_polynomial(y)) * pow(x - y, -1, p) % p
    return domain.size_as_field_element * pow(x, domain.size - 1, p) % p


class Index:
    pass


def index(field, cs):
    """`AHPForR1CS::index` [R src/ahp/indexer.rs:151-234] + `arithmetize_matrix`
    [R src/ahp/constraint_systems.rs:125-262].  `cs` is a synthesized, padded, squared system."""
    p = field.p
    a, b, c = cs.to_matrices()
    joint = sum_matrices(a, b, c)
    nnz = sum(len(r) for r in joint)
    n_in, n_w, n_c = len(cs.instance), len(cs.witness), cs.num_constraints
    if n_c != n_in + n_w:
        raise ValueError("NonSquareMatrix")
    if n_in & (n_in - 1):
        raise ValueError("InvalidPublicInputLength")
    info = IndexInfo(n_in + n_w, n_c, nnz, n_in)
    dom_h = P.Domain(field, n_c)
    dom_k = P.Domain(field, nnz)
    dom_x = P.Domain(field, n_in)

    elems = dom_h.elements()
    am = {(r, i): f for r, row in enumerate(a) for f, i in row}
    bm = {(r, i): f for r, row in enumerate(b) for f, i in row}
    cm = {(r, i): f for r, row in enumerate(c) for f, i in row}
    eq_poly_vals = dict(zip(elems, u_h_same_inputs(dom_h)))
    row_vec, col_vec, va, vb, vc, inverses = [], [], [], [], [], []
    for r, row in enumerate(joint):
        for i in row:
            row_val = elems[r]
            col_val = elems[dom_h.reindex_by_subdomain(dom_x, i)]
            row_vec.append(col_val)  # transpose
            col_vec.append(row_val)
            va.append(am.get((r, i), 0))
            vb.append(bm.get((r, i), 0))
            vc.append(cm.get((r, i), 0))
            inverses.append(eq_poly_vals[col_val])
    inverses = P.batch_inversion(inverses, p)
    va = [x * y % p for x, y in zip(va, inverses)]
    vb = [x * y % p for x, y in zip(vb, inverses)]
    vc = [x * y % p for x, y in zip(vc, inverses)]
    pad = dom_k.size - len(row_vec)
    row_vec += [elems[0]] * pad
    col_vec += [elems[0]] * pad
    va += [0] * pad
    vb += [0] * pad
    vc += [0] * pad
    row_col = [x * y % p for x, y in zip(row_vec, col_vec)]

    idx = Index()
    idx.field = field
    idx.info = info
    idx.a, idx.b, idx.c = a, b, c
    idx.evals = {"row": row_vec, "col": col_vec, "row_col": row_col, "val_a": va, "val_b": vb, "val_c": vc}
    lab = [("row", row_vec), ("col", col_vec), ("a_val", va), ("b_val", vb), ("c_val", vc), ("row_col", row_col)]
    idx.polys = [LabeledPoly(l, dom_k.ifft(v), None, None) for l, v in lab]  # order of Index::iter
    return idx


class ProverState:
    pass


def prover_init(field, idx, cs):
    """[R src/ahp/prover.rs:211-306]"""
    info = idx.info
    if info.num_constraints != cs.num_constraints or len(cs.instance) + len(cs.witness) != info.num_variables:
        raise ValueError("InstanceDoesNotMatchIndex")
    n_in = len(cs.instance)
    if n_in & (n_in - 1):
        raise ValueError("InvalidPublicInputLength")
    p = field.p
    z = cs.instance + cs.witness

    def mat_vec(m):
        return [sum(coeff * z[i] for coeff, i in row) % p for row in m]

    st = ProverState()
    st.field = field
    st.index = idx
    st.formatted_input = list(cs.instance)
    st.witness = list(cs.witness)
    st.z_a = mat_vec(idx.a)
    st.z_b = mat_vec(idx.b)
    st.zk_bound = 1
    st.domain_h = P.Domain(field, cs.num_constraints)
    st.domain_k = P.Domain(field, info.num_non_zero)
    st.domain_x = P.Domain(field, n_in)
    return st


def prover_first_round(st, rng):
    """[R src/ahp/prover.rs:309-409]"""
    f = st.field
    p = f.p
    dom_h, dom_x = st.domain_h, st.domain_x
    nh = dom_h.size
    x_poly = P.strip(dom_x.ifft(st.formatted_input))
    x_evals = dom_h.fft(x_poly)
    ratio = nh // dom_x.size
    w_ext = st.witness + [0] * (nh - dom_x.size - len(st.witness))
    w_evals = [0 if k % ratio == 0 else (w_ext[k - k // ratio - 1] - x_evals[k]) % p for k in range(nh)]

    def blind(coeffs, rho):  # + rho * v_H
        out = list(coeffs) + [0] * (nh + 1 - len(coeffs))
        out[0] = (out[0] - rho) % p
        out[nh] = (out[nh] + rho) % p
        return P.strip(out)

    w_poly = blind(dom_h.ifft(w_evals), field_rand(f, rng))
    w_poly, rem = P.divide_by_vanishing_poly(w_poly, dom_x)
    assert not rem
    z_a_poly = blind(dom_h.ifft(st.z_a), field_rand(f, rng))
    z_b_poly = blind(dom_h.ifft(st.z_b), field_rand(f, rng))

    mask_degree = 3 * nh + 2 * st.zk_bound - 3
    mask = poly_rand(f, mask_degree, rng)
    mask = mask + [0] * (mask_degree + 1 - len(mask))
    r0 = sum(mask[nh * i] for i in range(mask_degree // nh + 1)) % p
    mask[0] = (mask[0] - r0) % p
    mask = P.strip(mask)

    assert P.degree(w_poly) < nh - dom_x.size + st.zk_bound
    assert P.degree(z_a_poly) < nh + st.zk_bound and P.degree(z_b_poly) < nh + st.zk_bound
    st.w_poly = LabeledPoly("w", w_poly, None, 1)
    st.z_a_poly = LabeledPoly("z_a", z_a_poly, None, 1)
    st.z_b_poly = LabeledPoly("z_b", z_b_poly, None, 1)
    st.mask_poly = LabeledPoly("mask_poly", mask, None, None)
    return [st.w_poly, st.z_a_poly, st.z_b_poly, st.mask_poly]


def calculate_t(st, etas, r_alpha_x_on_h):
    """[R src/ahp/prover.rs:411-428]"""
    p = st.field.p
    dom_h, dom_x = st.domain_h, st.domain_x
    t = [0] * dom_h.size
    for m, eta in zip((st.index.a, st.index.b, st.index.c), etas):
        for r, row in enumerate(m):
            for coeff, c in row:
                j = dom_h.reindex_by_subdomain(dom_x, c)
                t[j] = (t[j] + eta * coeff % p * r_alpha_x_on_h[r]) % p
    return P.strip(dom_h.ifft(t))


def prover_second_round(st, alpha, eta_a, eta_b, eta_c):
    """[R src/ahp/prover.rs:443-570]"""
    f = st.field
    p = f.p
    dom_h = st.domain_h
    nh = dom_h.size
    za, zb = st.z_a_poly.coeffs, st.z_b_poly.coeffs
    zc = P.poly_mul(f, za, zb)
    summed = [x * eta_c % p for x in zc]
    for i, (x, y) in enumerate(zip(za, zb)):
        if i < len(summed):
            summed[i] = (summed[i] + eta_a * x + eta_b * y) % p
    summed = P.strip(summed)
    r_alpha_x_evals = u_h_diff_inputs(dom_h, alpha)
    r_alpha_poly = P.strip(dom_h.ifft(r_alpha_x_evals))
    t_poly = calculate_t(st, (eta_a, eta_b, eta_c), r_alpha_x_evals)
    dom_x = P.Domain(f, len(st.formatted_input))
    x_poly = P.strip(dom_x.ifft(st.formatted_input))
    z_poly = P.mul_by_vanishing_poly(st.w_poly.coeffs, dom_x)
    z_poly = z_poly + [0] * (len(x_poly) - len(z_poly))
    for i, x in enumerate(x_poly):
        z_poly[i] = (z_poly[i] + x) % p
    z_poly = P.strip(z_poly)
    assert P.degree(z_poly) < nh + st.zk_bound
    mul_size = max(len(st.mask_poly.coeffs), len(r_alpha_poly) + len(summed), len(t_poly) + len(z_poly))
    mul_dom = P.Domain(f, mul_size)
    ra = mul_dom.fft(r_alpha_poly)
    sz = mul_dom.fft(summed)
    ze = mul_dom.fft(z_poly)
    te = mul_dom.fft(t_poly)
    rhs = P.strip(mul_dom.ifft([(a * b - c * d) % p for a, b, c, d in zip(ra, sz, ze, te)]))
    q_1 = P.poly_add(st.mask_poly.coeffs, rhs, p)
    h_1, x_g_1 = P.divide_by_vanishing_poly(q_1, dom_h)
    assert not x_g_1 or x_g_1[0] == 0
    g_1 = P.strip(x_g_1[1:])
    assert P.degree(g_1) <= nh - 2
    assert P.degree(h_1) <= 2 * nh + 2 * st.zk_bound - 2
    st.alpha, st.eta = alpha, (eta_a, eta_b, eta_c)
    return [LabeledPoly("t", t_poly, None, None), LabeledPoly("g_1", g_1, nh - 2, 1), LabeledPoly("h_1", h_1, None, None)]


def prover_third_round(st, beta):
    """[R src/ahp/prover.rs:588-706]"""
    f = st.field
    p = f.p
    dom_h, dom_k = st.domain_h, st.domain_k
    alpha = st.alpha
    eta_a, eta_b, eta_c = st.eta
    v = dom_h.evaluate_vanishing_polynomial(alpha) * dom_h.evaluate_vanishing_polynomial(beta) % p
    ea, eb, ec_ = eta_a * v % p, eta_b * v % p, eta_c * v % p
    polys = {pl.label: pl.coeffs for pl in st.index.polys}
    n = max(len(polys["a_val"]), len(polys["b_val"]), len(polys["c_val"]))
    # zip() truncates to the shortest coefficient vector [R prover.rs:629-637]
    n = min(len(polys["a_val"]), len(polys["b_val"]), len(polys["c_val"]))
    a_poly = P.strip([(ea * polys["a_val"][i] + eb * polys["b_val"][i] + ec_ * polys["c_val"][i]) % p for i in range(n)])
    ev = st.index.evals
    ab = alpha * beta % p
    b_poly = P.strip(dom_k.ifft([(ab - alpha * r - beta * c + rc) % p for r, c, rc in zip(ev["row"], ev["col"], ev["row_col"])]))
    inverses = P.batch_inversion([(beta - r) * (alpha - c) % p for r, c in zip(ev["row"], ev["col"])], p)
    f_evals = [inv * (ea * x + eb * y + ec_ * z) % p for inv, x, y, z in zip(inverses, ev["val_a"], ev["val_b"], ev["val_c"])]
    f_poly = P.strip(dom_k.ifft(f_evals))
    h_2, _ = P.divide_by_vanishing_poly(P.poly_sub(a_poly, P.poly_mul(f, b_poly, f_poly), p), dom_k)
    g_2 = P.strip(f_poly[1:])
    assert P.degree(h_2) <= dom_k.size - 2 and P.degree(g_2) <= dom_k.size - 2
    return [LabeledPoly("g_2", g_2, dom_k.size - 2, None), LabeledPoly("h_2", h_2, None, None)]


# ---- verifier side ------------------------------------------------------------------------------
def sample_outside_domain(field, domain, rng):
    t = field_rand(field, rng)
    while domain.evaluate_vanishing_polynomial(t) == 0:
        t = field_rand(field, rng)
    return t


class VerifierState:
    pass


def verifier_first_round(field, info, rng):
    """[R src/ahp/verifier.rs:44-79]"""
    if info.num_constraints != info.num_variables:
        raise ValueError("NonSquareMatrix")
    vs = VerifierState()
    vs.field = field
    vs.domain_h = P.Domain(field, info.num_constraints)
    vs.domain_k = P.Domain(field, info.num_non_zero)
    vs.alpha = sample_outside_domain(field, vs.domain_h, rng)
    vs.eta_a = field_rand(field, rng)
    vs.eta_b = field_rand(field, rng)
    vs.eta_c = field_rand(field, rng)
    return vs


def verifier_second_round(vs, rng):
    vs.beta = sample_outside_domain(vs.field, vs.domain_h, rng)
    return vs


def verifier_third_round(vs, rng):
    vs.gamma = field_rand(vs.field, rng)
    return vs


def verifier_query_set(vs):
    """[R src/ahp/verifier.rs:103-188] as a sorted list (BTreeSet order)."""
    qs = [("g_1", ("beta", vs.beta)), ("z_b", ("beta", vs.beta)), ("t", ("beta", vs.beta)),
          ("outer_sumcheck", ("beta", vs.beta)), ("g_2", ("gamma", vs.gamma)), ("inner_sumcheck", ("gamma", vs.gamma))]
    return sorted(qs)


LC_WITH_ZERO_EVAL = ("inner_sumcheck", "outer_sumcheck")


def construct_linear_combinations(field, public_input, evals, vs):
    """[R src/ahp/mod.rs:110-221].  `evals(label, point)` returns the evaluation of a single-polynomial LC."""
    p = field.p
    dom_h, dom_k = vs.domain_h, vs.domain_k
    formatted = [1] + list(public_input)
    if len(formatted) & (len(formatted) - 1):
        raise ValueError("InvalidPublicInputLength")
    x_dom = P.Domain(field, len(formatted))
    alpha, beta, gamma = vs.alpha, vs.beta, vs.gamma
    eta_a, eta_b, eta_c = vs.eta_a, vs.eta_b, vs.eta_c
    z_b = LinearCombination("z_b", [(1, "z_b")])
    g_1 = LinearCombination("g_1", [(1, "g_1")])
    t = LinearCombination("t", [(1, "t")])
    r_alpha_at_beta = u_h(dom_h, alpha, beta)
    v_h_alpha = dom_h.evaluate_vanishing_polynomial(alpha)
    v_h_beta = dom_h.evaluate_vanishing_polynomial(beta)
    v_x_beta = x_dom.evaluate_vanishing_polynomial(beta)
    z_b_at_beta = evals("z_b", beta)
    t_at_beta = evals("t", beta)
    g_1_at_beta = evals("g_1", beta)
    x_at_beta = sum(l * x for l, x in zip(x_dom.evaluate_all_lagrange_coefficients(beta), formatted)) % p
    outer = LinearCombination("outer_sumcheck", [
        (1, "mask_poly"),
        (r_alpha_at_beta * (eta_a + eta_c * z_b_at_beta) % p, "z_a"),
        (r_alpha_at_beta * eta_b % p * z_b_at_beta % p, None),
        ((-t_at_beta * v_x_beta) % p, "w"),
        ((-t_at_beta * x_at_beta) % p, None),
        ((-v_h_beta) % p, "h_1"),
        ((-beta * g_1_at_beta) % p, None),
    ])
    lcs = [z_b, g_1, t, outer]
    g_2 = LinearCombination("g_2", [(1, "g_2")])
    g_2_at_gamma = evals("g_2", gamma)
    v_k_gamma = dom_k.evaluate_vanishing_polynomial(gamma)
    a = LinearCombination("a_poly", [(eta_a, "a_val"), (eta_b, "b_val"), (eta_c, "c_val")])
    a.scale(v_h_alpha * v_h_beta % p, p)
    b = LinearCombination("denom", [(beta * alpha % p, None), ((-alpha) % p, "row"), ((-beta) % p, "col"), (1, "row_col")])
    b.scale((gamma * g_2_at_gamma + t_at_beta * pow(dom_k.size_as_field_element, -1, p)) % p, p)
    inner = a
    inner.sub(b, p)
    inner.sub(LinearCombination("h_2", [(v_k_gamma, "h_2")]), p)
    inner.label = "inner_sumcheck"
    lcs += [g_2, inner]
    lcs.sort(key=lambda l: l.label)
    return lcs
