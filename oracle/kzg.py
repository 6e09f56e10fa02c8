"""KZG10 and the two polynomial-commitment schemes Marlin is instantiated with, restated from
ark-poly-commit 0.3 [U ark-poly-commit src/kzg10/mod.rs, src/marlin/marlin_pc/mod.rs,
src/sonic_pc/mod.rs, src/lib.rs (batch_open / open_combinations defaults)] as recalled in
SURVEY.md App. B; call sites in the reference: src/lib.rs:93,115,125,172,193,213,292,413.

Verification uses the SRS trapdoor instead of a pairing: e(C - v*G - rv*gammaG, H) == e(W, (beta - z)*H)
holds iff  C - v*G - rv*gammaG == (beta - z)*W  in G1, which needs no G2 arithmetic.  The test SRS
is insecure by construction (beta known), exactly like the reference's `universal_setup(.., test_rng)`.
"""
from . import ec
from .poly import divide_by_linear, evaluate, poly_add, poly_scale, strip
from .rng import poly_rand

MARLIN = "marlin_kzg10"
SONIC = "sonic_kzg10"


class LabeledPoly:
    def __init__(self, label, coeffs, degree_bound=None, hiding_bound=None):
        self.label = label
        self.coeffs = strip(list(coeffs))
        self.degree_bound = degree_bound
        self.hiding_bound = hiding_bound

    def degree(self):
        return len(self.coeffs) - 1 if self.coeffs else 0


class Commitment:
    """marlin_pc::Commitment {comm, shifted_comm}; for Sonic shifted is always None."""

    def __init__(self, comm, shifted=None):
        self.comm = comm
        self.shifted = shifted


class Randomness:
    """marlin_pc::Randomness {rand, shifted_rand}; a kzg10::Randomness is its blinding polynomial."""

    def __init__(self, rand=None, shifted_rand=None):
        self.rand = rand if rand is not None else []
        self.shifted_rand = shifted_rand  # None <=> no degree bound


class UniversalParams:
    """Insecure test SRS: powers_of_g[i] = beta^i g, powers_of_gamma_g[i] = beta^i gamma g."""

    def __init__(self, curve, max_degree, beta, g, gamma, powers_of_g=None):
        self.curve = curve
        self.max_degree = max_degree
        self.beta = beta % curve.fr.p
        self.g = g
        self.gamma = gamma % curve.fr.p
        self.gamma_g = ec.scalar_mul(curve, gamma, g)
        if powers_of_g == "lazy":  # verifier-only use: individual powers on demand, O(1) each through the trapdoor
            self.powers_of_g = _LazyPowers(curve, g, self.beta)
        else:
            self.powers_of_g = powers_of_g if powers_of_g is not None else ec.fixed_base_powers(curve, g, beta, max_degree + 1)
        self._gamma_cache = {}

    def power_of_gamma_g(self, i):
        if i not in self._gamma_cache:
            self._gamma_cache[i] = ec.scalar_mul(self.curve, pow(self.beta, i, self.curve.fr.p), self.gamma_g)
        return self._gamma_cache[i]


class _LazyPowers:
    def __init__(self, curve, g, beta):
        self.curve, self.g, self.beta = curve, g, beta

    def __getitem__(self, i):
        return ec.scalar_mul(self.curve, pow(self.beta, i, self.curve.fr.p), self.g)


class CommitterKey:
    def __init__(self, pp, supported_degree, supported_hiding_bound, enforced_degree_bounds, scheme):
        """`PC::trim` [U marlin_pc / sonic_pc trim]: all base sets are contiguous slices of pp.powers_of_g."""
        assert supported_degree <= pp.max_degree
        self.pp = pp
        self.curve = pp.curve
        self.scheme = scheme
        self.supported_degree = supported_degree
        self.max_degree = pp.max_degree
        self.hiding_bound = supported_hiding_bound
        self.enforced_degree_bounds = sorted(set(enforced_degree_bounds)) if enforced_degree_bounds else None
        self.powers_of_gamma_g = [pp.power_of_gamma_g(i) for i in range(supported_hiding_bound + 2)]

    def powers(self):
        return 0, self.supported_degree + 1  # (offset into pp.powers_of_g, length)

    def shifted_offset(self, degree_bound):
        """offset of `shifted_powers(degree_bound)`: powers_of_g[max_degree - degree_bound ..]"""
        if degree_bound is None:
            degree_bound = self.enforced_degree_bounds[-1]
        assert degree_bound in self.enforced_degree_bounds
        return self.max_degree - degree_bound

    def shifted_gamma_powers(self, degree_bound):
        """sonic_pc: shifted_powers_of_gamma_g[bound][i] = powers_of_gamma_g[max_degree - bound + i]"""
        off = self.max_degree - degree_bound
        return [self.pp.power_of_gamma_g(off + i) for i in range(self.hiding_bound + 2)]


class Engine:
    """Where the group work happens.  `use_trapdoor=False`: a real MSM over the SRS slice (the
    reference's algorithm); `True`: (sum c_i beta^(off+i)) * g, the same group element computed
    through the trapdoor -- used to keep large oracle runs fast; tests check both agree."""

    def __init__(self, use_trapdoor=False):
        self.use_trapdoor = use_trapdoor
        self.msm_log = []  # (offset, length) of every MSM issued, for the size model in DESIGN.md

    def msm_powers(self, pp, offset, scalars):
        self.msm_log.append((offset, len(scalars)))
        curve = pp.curve
        if self.use_trapdoor:
            r = curve.fr.p
            acc = 0
            cur = pow(pp.beta, offset, r)
            for s in scalars:
                acc = (acc + s * cur) % r
                cur = cur * pp.beta % r
            return ec.scalar_mul(curve, acc, pp.g)
        return ec.msm_pippenger_arkworks(curve, pp.powers_of_g[offset:offset + len(scalars)], scalars)

    def msm_bases(self, curve, bases, scalars):
        return ec.msm_naive(curve, bases, scalars)


def _skip_leading_zeros(coeffs):
    """`skip_leading_zeros_and_convert_to_bigints`: number of low-order zero coefficients skipped."""
    k = 0
    while k < len(coeffs) and coeffs[k] == 0:
        k += 1
    return k, coeffs[k:]


def kzg_commit(engine, ck, base_offset, gamma_powers, coeffs, hiding_bound, rng):
    """`KZG10::commit(powers, p, hiding_bound, rng)` -> (affine commitment, blinding polynomial)."""
    curve = ck.curve
    lz, plain = _skip_leading_zeros(coeffs)
    commitment = engine.msm_powers(ck.pp, base_offset + lz, plain)
    blinding = []
    if hiding_bound is not None:
        if rng is None:
            raise ValueError("MissingRng")
        blinding = poly_rand(curve.fr, hiding_bound + 1, rng)  # Randomness::rand: degree hiding_bound + 1
        assert len(blinding) - 1 <= len(gamma_powers) - 1
    random_commitment = engine.msm_bases(curve, gamma_powers[:len(blinding)], blinding)
    return ec.affine_add(curve, commitment, random_commitment), blinding


def commit(engine, ck, polys, rng):
    """`PC::commit` for MarlinKZG10 / SonicKZG10: polynomials sequentially, rng shared."""
    comms, rands = [], []
    for p in polys:
        assert p.degree() <= ck.supported_degree
        if ck.scheme == MARLIN:
            c, r = kzg_commit(engine, ck, 0, ck.powers_of_gamma_g, p.coeffs, p.hiding_bound, rng)
            sc, sr = None, None
            if p.degree_bound is not None:
                off = ck.shifted_offset(p.degree_bound)
                sc, sr = kzg_commit(engine, ck, off, ck.powers_of_gamma_g, p.coeffs, p.hiding_bound, rng)
            comms.append(Commitment(c, sc))
            rands.append(Randomness(r, sr))
        else:
            if p.degree_bound is not None:
                c, r = kzg_commit(engine, ck, ck.shifted_offset(p.degree_bound), ck.shifted_gamma_powers(p.degree_bound),
                                  p.coeffs, p.hiding_bound, rng)
            else:
                c, r = kzg_commit(engine, ck, 0, ck.powers_of_gamma_g, p.coeffs, p.hiding_bound, rng)
            comms.append(Commitment(c, None))
            rands.append(Randomness(r, None))
    return comms, rands


def _open_with_witness(engine, ck, base_offset, gamma_powers, point, rand_poly, witness, hiding_witness):
    """`KZG10::open_with_witness_polynomial` -> (w affine, random_v or None)"""
    curve = ck.curve
    p = curve.fr.p
    lz, wc = _skip_leading_zeros(witness)
    w = engine.msm_powers(ck.pp, base_offset + lz, wc)
    random_v = None
    if hiding_witness is not None:
        random_v = evaluate(rand_poly, point, p)
        w = ec.affine_add(curve, w, engine.msm_bases(curve, gamma_powers[:len(hiding_witness)], hiding_witness))
    return w, random_v


def _witness_polys(coeffs, point, rand_poly, p):
    """`KZG10::compute_witness_polynomial`: (p / (X - z), r / (X - z) if hiding)"""
    witness, _ = divide_by_linear(coeffs, point, p)
    hiding = None
    if strip(list(rand_poly)):
        hiding, _ = divide_by_linear(rand_poly, point, p)
    return witness, hiding


def open_at_point(engine, ck, polys, rands, point, opening_challenges):
    """`PC::open_individual_opening_challenges` for one point -> kzg10::Proof (w, random_v)."""
    curve = ck.curve
    p = curve.fr.p
    if ck.scheme == SONIC:
        comb, comb_rand = [], []
        counter = 0
        for poly, rand in zip(polys, rands):
            ch = opening_challenges(counter)
            counter += 1
            comb = poly_add(comb, poly_scale(poly.coeffs, ch, p), p)
            comb_rand = poly_add(comb_rand, poly_scale(rand.rand, ch, p), p)
        witness, hiding = _witness_polys(comb, point, comb_rand, p)
        return _open_with_witness(engine, ck, 0, ck.powers_of_gamma_g, point, comb_rand, witness, hiding)

    comb, r = [], []
    shifted_w, shifted_r, shifted_r_witness = [], [], []
    enforce = False
    counter = 0
    for poly, rand in zip(polys, rands):
        assert (poly.degree_bound is not None) == (rand.shifted_rand is not None)
        ch = opening_challenges(counter)
        counter += 1
        comb = poly_add(comb, poly_scale(poly.coeffs, ch, p), p)
        r = poly_add(r, poly_scale(rand.rand, ch, p), p)
        if poly.degree_bound is not None:
            enforce = True
            witness, srw = _witness_polys(poly.coeffs, point, rand.shifted_rand, p)
            ch1 = opening_challenges(counter)
            counter += 1
            # shift_polynomial: prepend (largest bound - this bound) zero coefficients
            pad = ck.enforced_degree_bounds[-1] - poly.degree_bound
            shifted = ([0] * pad + witness) if witness else []
            shifted_w = poly_add(shifted_w, poly_scale(shifted, ch1, p), p)
            shifted_r = poly_add(shifted_r, poly_scale(rand.shifted_rand, ch1, p), p)
            if srw is not None:
                shifted_r_witness = poly_add(shifted_r_witness, poly_scale(srw, ch1, p), p)
    witness, hiding = _witness_polys(comb, point, r, p)
    w, random_v = _open_with_witness(engine, ck, 0, ck.powers_of_gamma_g, point, r, witness, hiding)
    if enforce:
        sw, srv = _open_with_witness(engine, ck, ck.shifted_offset(None), ck.powers_of_gamma_g, point, shifted_r, shifted_w,
                                     shifted_r_witness)
        w = ec.affine_add(curve, w, sw)
        if srv is not None and random_v is not None:
            random_v = (random_v + srv) % p
    return w, random_v


class LinearCombination:
    def __init__(self, label, terms):
        self.label = label
        self.terms = [(c, t) for c, t in terms]  # t = polynomial label or None for LCTerm::One

    def scale(self, k, p):
        self.terms = [(c * k % p, t) for c, t in self.terms]

    def sub(self, other, p):
        self.terms += [((-c) % p, t) for c, t in other.terms]


def open_combinations(engine, ck, lcs, polys, rands, query_set, opening_challenge):
    """`PC::open_combinations` -> BatchLCProof.proof = [kzg10::Proof per point label in BTreeMap order].
    query_set: iterable of (lc_label, (point_label, point))."""
    curve = ck.curve
    p = curve.fr.p
    by_label = {pl.label: (pl, r) for pl, r in zip(polys, rands)}
    lc_polys, lc_rands = {}, {}
    for lc in lcs:
        poly, rnd, srnd = [], [], None
        degree_bound, hiding_bound = None, None
        num_polys = len(lc.terms)
        for coeff, label in lc.terms:
            if label is None:
                continue
            cur, cur_rand = by_label[label]
            if num_polys == 1 and cur.degree_bound is not None:
                assert coeff == 1, "Coefficient must be one for degree-bounded equations"
                degree_bound = cur.degree_bound
            elif cur.degree_bound is not None:
                raise ValueError("EquationHasDegreeBounds")
            if cur.hiding_bound is not None:
                hiding_bound = cur.hiding_bound if hiding_bound is None else max(hiding_bound, cur.hiding_bound)
            poly = poly_add(poly, poly_scale(cur.coeffs, coeff, p), p)
            rnd = poly_add(rnd, poly_scale(cur_rand.rand, coeff, p), p)
            if degree_bound is not None:
                srnd = poly_scale(cur_rand.shifted_rand, coeff, p) if cur_rand.shifted_rand is not None else None
        lc_polys[lc.label] = LabeledPoly(lc.label, poly, degree_bound, hiding_bound)
        # Marlin PC: a degree-bounded LC keeps its shifted randomness (it is the single polynomial itself)
        lc_rands[lc.label] = Randomness(rnd, srnd if (ck.scheme == MARLIN and degree_bound is not None) else None)
        if ck.scheme == MARLIN and degree_bound is not None and lc_rands[lc.label].shifted_rand is None:
            lc_rands[lc.label].shifted_rand = []
    # group by point label (BTreeMap order), labels within a point in BTreeSet order
    by_point = {}
    for label, (point_label, point) in query_set:
        by_point.setdefault(point_label, (point, set()))[1].add(label)
    challenges = lambda k: pow(opening_challenge, k, p)
    proofs = []
    for point_label in sorted(by_point):
        point, labels = by_point[point_label]
        ls = sorted(labels)
        proofs.append(open_at_point(engine, ck, [lc_polys[l] for l in ls], [lc_rands[l] for l in ls], point, challenges))
    return proofs


def check_combinations(ck, lcs, commitments, degree_bounds, query_set, evaluations, proofs, opening_challenge, g2=None):
    """`PC::check_combinations`.  g2 = None: the pairing is replaced by the trapdoor identity (module docstring).
    g2 = G2Key: the reference's own check, a product of pairings [U kzg10::check / sonic_pc check]:
        e(C - v g - rv gamma_g, h) * prod_d e(C_d, beta^-(D-d) h) * e(-W, beta h - z h) == 1
    (the middle factors only for SonicKZG10's degree-bounded commitments).
    commitments: {label: Commitment}; degree_bounds: {label: bound or None}; evaluations: {(lc_label, point): value}."""
    curve = ck.curve
    pp = ck.pp
    p = curve.fr.p
    evals = dict(evaluations)
    lc_comm, lc_bound = {}, {}
    for lc in lcs:
        num_polys = len(lc.terms)
        comm, shifted, bound = None, None, None
        for coeff, label in lc.terms:
            if label is None:
                for key in list(evals):
                    if key[0] == lc.label:
                        evals[key] = (evals[key] - coeff) % p
                continue
            c = commitments[label]
            if num_polys == 1 and degree_bounds.get(label) is not None:
                assert coeff == 1
                bound = degree_bounds[label]
                shifted = c.shifted
            elif degree_bounds.get(label) is not None:
                raise ValueError("EquationHasDegreeBounds")
            comm = ec.affine_add(curve, comm, ec.scalar_mul(curve, coeff, c.comm))
        lc_comm[lc.label] = Commitment(comm, shifted)
        lc_bound[lc.label] = bound
    by_point = {}
    for label, (point_label, point) in query_set:
        by_point.setdefault(point_label, (point, set()))[1].add(label)
    ok = True
    for (point_label, (w, random_v)) in zip(sorted(by_point), proofs):
        point, labels = by_point[point_label]
        plain, combined_value = None, 0
        by_bound = {}  # sonic: bound -> combined commitment made with shifted powers
        counter = 0
        for label in sorted(labels):
            ch = pow(opening_challenge, counter, p)
            counter += 1
            c = lc_comm[label]
            v = evals[(label, point)]
            combined_value = (combined_value + v * ch) % p
            if ck.scheme == SONIC and lc_bound[label] is not None:
                d = lc_bound[label]
                by_bound[d] = ec.affine_add(curve, by_bound.get(d), ec.scalar_mul(curve, ch, c.comm))
                continue
            plain = ec.affine_add(curve, plain, ec.scalar_mul(curve, ch, c.comm))
            if ck.scheme == MARLIN and lc_bound[label] is not None:
                ch1 = pow(opening_challenge, counter, p)
                counter += 1
                shift_power = pp.powers_of_g[pp.max_degree - lc_bound[label]]
                adjusted = ec.affine_add(curve, c.shifted, ec.affine_neg(curve, ec.scalar_mul(curve, v, shift_power)))
                plain = ec.affine_add(curve, plain, ec.scalar_mul(curve, ch1, adjusted))
        # the combined value is taken out of the unshifted side: C - v g - rv gamma_g
        plain = ec.affine_add(curve, plain, ec.affine_neg(curve, ec.scalar_mul(curve, combined_value, pp.g)))
        if random_v is not None:
            plain = ec.affine_add(curve, plain, ec.affine_neg(curve, ec.scalar_mul(curve, random_v, pp.gamma_g)))
        if g2 is None:
            # trapdoor: unshift the bounded parts (multiply by beta^-(D-d)) and compare with (beta - z) W
            lhs = plain
            for d, cd in by_bound.items():
                unshift = pow(pow(pp.beta, pp.max_degree - d, p), -1, p)
                lhs = ec.affine_add(curve, lhs, ec.scalar_mul(curve, unshift, cd))
            ok = ok and (lhs == ec.scalar_mul(curve, (pp.beta - point) % p, w))
        else:
            ok = ok and g2.check(curve, plain, by_bound, w, point)
    return ok


class G2Key:
    """The G2 half of the verifier key: h, beta h and (SonicKZG10) beta^-(D-d) h per enforced bound, produced
    at setup.  `check` runs the pairing product of the reference's KZG10 / Sonic checks (oracle/pairing.py)."""

    def __init__(self, pp, bounds=()):
        from . import pairing
        self.pairing = pg = pairing.for_curve(pp.curve)  # BLS12-381 or BN254
        self.h = pg.g2_generator()
        self.beta_h = pg.e12_mul(pp.beta, self.h)
        p = pp.curve.fr.p
        self.neg_powers = {d: pg.e12_mul(pow(pow(pp.beta, pp.max_degree - d, p), -1, p), self.h) for d in bounds}

    def check(self, curve, plain, by_bound, w, z):
        pg = self.pairing
        rhs_g2 = pg.e12_add(self.beta_h, pg.e12_neg(pg.e12_mul(z % curve.fr.p, self.h)))  # beta h - z h
        pairs = [(plain, self.h), (ec.affine_neg(curve, w), rhs_g2)]
        for d, cd in by_bound.items():
            pairs.append((cd, self.neg_powers[d]))
        return pg.pairing_product_is_one(pairs)
