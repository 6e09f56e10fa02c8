"""Minimal R1CS builder restating what the reference needs from ark-relations 0.3
[U ark-relations src/r1cs/constraint_system.rs] plus the reference's own padding rules
[R src/ahp/constraint_systems.rs:45-81, 282-290].

Variable numbering in the matrices: One -> 0, instance k -> k, witness k -> num_instance + k
(SURVEY.md A.5).  `finalize()` with OptimizationGoal::Weight only outlines linear combinations
that are reused; none of the circuits below reuse any, so it is the identity here.
"""
from .poly import Domain


class ConstraintSystem:
    def __init__(self, field):
        self.f = field
        self.instance = [1]  # formatted input assignment (leading one)
        self.witness = []
        self.constraints = []  # (a_lc, b_lc, c_lc), each lc = [(coeff, ('i'|'w', idx))]

    def new_input_variable(self, v):
        self.instance.append(v % self.f.p)
        return ("i", len(self.instance) - 1)

    def new_witness_variable(self, v):
        self.witness.append(v % self.f.p)
        return ("w", len(self.witness) - 1)

    def enforce_constraint(self, a, b, c):
        self.constraints.append((list(a), list(b), list(c)))

    @property
    def num_constraints(self):
        return len(self.constraints)

    def pad_input(self):
        """[R constraint_systems.rs:45-58] pad the instance to a power of two with zeros."""
        size = Domain(self.f, len(self.instance)).size
        while len(self.instance) < size:
            self.new_input_variable(0)

    def make_square(self):
        """[R constraint_systems.rs:60-81]"""
        nv = len(self.instance) + len(self.witness)
        nc = self.num_constraints
        if nv > nc:
            for _ in range(nv - nc):
                self.enforce_constraint([], [], [])
        else:
            for _ in range(nc - nv):
                self.new_witness_variable(1)

    def to_matrices(self):
        ni = len(self.instance)

        def row(lc):
            out = []
            for coeff, (kind, idx) in lc:
                coeff %= self.f.p
                if coeff:
                    out.append((coeff, idx if kind == "i" else ni + idx))
            return out

        a = [row(x[0]) for x in self.constraints]
        b = [row(x[1]) for x in self.constraints]
        c = [row(x[2]) for x in self.constraints]
        return a, b, c


def synthesize(field, circuit):
    """What `index` / `prover_init` do before touching polynomials:
    generate_constraints -> pad_input_for_indexer_and_prover -> finalize -> make_matrices_square."""
    cs = ConstraintSystem(field)
    circuit(cs)
    cs.pad_input()
    cs.make_square()
    return cs


def dummy_circuit(field, a, b, num_variables, num_constraints):
    """`DummyCircuit` of the reference's bench [R benches/bench.rs:25-67]."""
    p = field.p

    def gen(cs):
        va = cs.new_witness_variable(a)
        vb = cs.new_witness_variable(b)
        vc = cs.new_input_variable(a * b % p)
        for _ in range(num_variables - 3):
            cs.new_witness_variable(a)
        for _ in range(num_constraints - 1):
            cs.enforce_constraint([(1, va)], [(1, vb)], [(1, vc)])
        cs.enforce_constraint([], [], [])

    return gen


def test_circuit(field, a, b, num_constraints, num_variables):
    """`Circuit` of the reference's tests [R src/test.rs:8-50]: public inputs c = a*b, d = c*b."""
    p = field.p

    def gen(cs):
        va = cs.new_witness_variable(a)
        vb = cs.new_witness_variable(b)
        vc = cs.new_input_variable(a * b % p)
        vd = cs.new_input_variable(a * b % p * b % p)
        for _ in range(num_variables - 3):
            cs.new_witness_variable(a)
        for _ in range(num_constraints - 1):
            cs.enforce_constraint([(1, va)], [(1, vb)], [(1, vc)])
        cs.enforce_constraint([(1, vc)], [(1, vb)], [(1, vd)])

    return gen


def dense_circuit(field, seed, num_constraints, num_variables, per_row):
    """Seeded random satisfiable R1CS with `per_row` non-zeros per row per matrix (BASELINE.json
    config 1's "dense" case; generator documented in DESIGN.md): z random, A and B rows random,
    C row = single fresh witness holding (A z)(B z)."""
    import random
    p = field.p

    def gen(cs):
        rnd = random.Random(seed)  # re-seeded per synthesis: index() and prove() must see the same system
        pub = cs.new_input_variable(rnd.randrange(p))
        free = [pub]
        vals = {pub: cs.instance[1]}
        nfree = max(per_row, num_variables - num_constraints - 2)
        for _ in range(nfree):
            v = rnd.randrange(p)
            w = cs.new_witness_variable(v)
            free.append(w)
            vals[w] = v
        for _ in range(num_constraints):
            ra = [(rnd.randrange(1, p), v) for v in rnd.sample(free, min(per_row, len(free)))]
            rb = [(rnd.randrange(1, p), v) for v in rnd.sample(free, min(per_row, len(free)))]
            az = sum(c * vals[v] for c, v in ra) % p
            bz = sum(c * vals[v] for c, v in rb) % p
            out = cs.new_witness_variable(az * bz % p)
            cs.enforce_constraint(ra, rb, [(1, out)])

    return gen
