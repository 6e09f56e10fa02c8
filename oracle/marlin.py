"""`Marlin::<F, PC, FS>::{universal_setup, index, prove, verify}` [R src/lib.rs:79-434] restated over
the oracle's AHP (ahp.py), PC schemes (kzg.py) and Fiat-Shamir RNG (transcript.py), plus the
`CanonicalSerialize` byte layout of `Proof` [R src/data_structures.rs:100-110].

`universal_setup` here takes the trapdoor explicitly: the SRS is an INPUT of the prover path and
the reference's own RNG consumption in `KZG10::setup` (G2 sampling etc.) is out of scope."""
import struct

from . import ahp, kzg, r1cs
from . import transcript as T
from .poly import evaluate
from .rng import u128_rand

PROTOCOL_NAME = b"MARLIN-2019"


def universal_setup(curve, num_constraints, num_variables, num_non_zero, beta, g_scalar=1, gamma=7, powers_of_g=None):
    """[R src/lib.rs:79-96] with an explicit trapdoor (see module docstring)."""
    from . import ec
    md = ahp.max_degree(curve.fr, num_constraints, num_variables, num_non_zero)
    g = ec.scalar_mul(curve, g_scalar, curve.g)
    return kzg.UniversalParams(curve, md, beta, g, gamma, powers_of_g)


class IndexProverKey:
    pass


def commitment_to_bytes(curve, scheme, c):
    """ToBytes of PC::Commitment."""
    if scheme == kzg.SONIC:
        return T.g1_affine_bytes(curve, c.comm)
    out = T.g1_affine_bytes(curve, c.comm)
    out += b"\x01" if c.shifted is not None else b"\x00"
    out += T.g1_affine_bytes(curve, c.shifted)  # identity when absent
    return out


def index_vk_bytes(curve, scheme, info, index_comms):
    """`IndexVerifierKey::write` [R src/data_structures.rs:36-43]: index_info (3 x u64) || index_comms."""
    out = T.u64_bytes(info.num_variables) + T.u64_bytes(info.num_constraints) + T.u64_bytes(info.num_non_zero)
    for c in index_comms:
        out += commitment_to_bytes(curve, scheme, c)
    return out


def index(srs, circuit, scheme=kzg.MARLIN, engine=None):
    """[R src/lib.rs:100-148]"""
    curve = srs.curve
    f = curve.fr
    engine = engine or kzg.Engine()
    cs = r1cs.synthesize(f, circuit)
    idx = ahp.index(f, cs)
    md = ahp.max_degree(f, idx.info.num_constraints, idx.info.num_variables, idx.info.num_non_zero)
    if srs.max_degree < md:
        raise ValueError("IndexTooLarge")
    ck = kzg.CommitterKey(srs, md, 1, ahp.get_degree_bounds(f, idx.info), scheme)
    comms, rands = kzg.commit(engine, ck, idx.polys, None)
    pk = IndexProverKey()
    pk.curve, pk.scheme, pk.index, pk.ck = curve, scheme, idx, ck
    pk.index_comms, pk.index_rands = comms, rands
    pk.vk_bytes = index_vk_bytes(curve, scheme, idx.info, comms)
    return pk


class Proof:
    def __init__(self, commitments, evaluations, pc_proof):
        self.commitments = commitments  # [[Commitment]*4, [..]*3, [..]*2]
        self.evaluations = evaluations  # [g_1(beta), g_2(gamma), t(beta), z_b(beta)]
        self.pc_proof = pc_proof        # [(w, random_v or None)] for beta, gamma


def prove(pk, circuit, zk_rng, engine=None):
    """[R src/lib.rs:151-311]"""
    curve, scheme, ck = pk.curve, pk.scheme, pk.ck
    f = curve.fr
    p = f.p
    engine = engine or kzg.Engine()
    cs = r1cs.synthesize(f, circuit)
    st = ahp.prover_init(f, pk.index, cs)
    public_input = st.formatted_input[1:]
    fs = T.FiatShamirRng(PROTOCOL_NAME + pk.vk_bytes + b"".join(T.fe_bytes(f, x) for x in public_input))

    def comms_bytes(comms):
        return b"".join(commitment_to_bytes(curve, scheme, c) for c in comms)

    first_oracles = ahp.prover_first_round(st, zk_rng)
    first_comms, first_rands = kzg.commit(engine, ck, first_oracles, zk_rng)
    fs.absorb(comms_bytes(first_comms))  # prover message is EmptyMessage: writes nothing
    vs = ahp.verifier_first_round(f, pk.index.info, fs)

    second_oracles = ahp.prover_second_round(st, vs.alpha, vs.eta_a, vs.eta_b, vs.eta_c)
    second_comms, second_rands = kzg.commit(engine, ck, second_oracles, zk_rng)
    fs.absorb(comms_bytes(second_comms))
    vs = ahp.verifier_second_round(vs, fs)

    third_oracles = ahp.prover_third_round(st, vs.beta)
    third_comms, third_rands = kzg.commit(engine, ck, third_oracles, zk_rng)
    fs.absorb(comms_bytes(third_comms))
    vs = ahp.verifier_third_round(vs, fs)

    polynomials = pk.index.polys + first_oracles + second_oracles + third_oracles
    rands = pk.index_rands + first_rands + second_rands + third_rands
    by_label = {pl.label: pl for pl in polynomials}
    query_set = ahp.verifier_query_set(vs)

    def poly_eval(label, point):
        return evaluate(by_label[label].coeffs, point, p)

    lcs = ahp.construct_linear_combinations(f, public_input, poly_eval, vs)
    lc_by_label = {lc.label: lc for lc in lcs}
    evaluations = []
    for label, (_, point) in query_set:
        lc = lc_by_label[label]
        ev = sum(c * (1 if t is None else poly_eval(t, point)) for c, t in lc.terms) % p
        if label in ahp.LC_WITH_ZERO_EVAL:
            assert ev == 0, f"{label} does not vanish"
        else:
            evaluations.append((label, ev))
    evaluations.sort(key=lambda x: x[0])
    evaluations = [e for _, e in evaluations]
    fs.absorb(b"".join(T.fe_bytes(f, e) for e in evaluations))
    opening_challenge = u128_rand(fs) % p
    pc_proof = kzg.open_combinations(engine, ck, lcs, polynomials, rands, query_set, opening_challenge)
    proof = Proof([first_comms, second_comms, third_comms], evaluations, pc_proof)
    proof.debug = {"alpha": vs.alpha, "eta": (vs.eta_a, vs.eta_b, vs.eta_c), "beta": vs.beta, "gamma": vs.gamma,
                   "xi": opening_challenge, "oracles": {pl.label: pl.coeffs for pl in first_oracles + second_oracles + third_oracles},
                   "zk_rng_word_pos": getattr(zk_rng, "word_pos", None)}
    return proof


def verify(pk, public_input, proof, g2=None):
    """[R src/lib.rs:315-433].  g2 = None: `PC::check_combinations` through the SRS trapdoor (kzg.py);
    g2 = kzg.G2Key: through the pairing product the reference computes (oracle/pairing.py: BLS12-381, BN254)."""
    curve, scheme, ck = pk.curve, pk.scheme, pk.ck
    f = curve.fr
    p = f.p
    info = pk.index.info
    from .poly import Domain
    dom_x = Domain(f, len(public_input) + 1)
    public_input = list(public_input) + [0] * (max(len(public_input), dom_x.size - 1) - len(public_input))
    fs = T.FiatShamirRng(PROTOCOL_NAME + pk.vk_bytes + b"".join(T.fe_bytes(f, x) for x in public_input))

    def comms_bytes(comms):
        return b"".join(commitment_to_bytes(curve, scheme, c) for c in comms)

    fs.absorb(comms_bytes(proof.commitments[0]))
    vs = ahp.verifier_first_round(f, info, fs)
    fs.absorb(comms_bytes(proof.commitments[1]))
    vs = ahp.verifier_second_round(vs, fs)
    fs.absorb(comms_bytes(proof.commitments[2]))
    vs = ahp.verifier_third_round(vs, fs)
    h_size, k_size = vs.domain_h.size, vs.domain_k.size
    labels = ["row", "col", "a_val", "b_val", "c_val", "row_col", "w", "z_a", "z_b", "mask_poly", "t", "g_1", "h_1", "g_2", "h_2"]
    bounds = [None] * 6 + [None] * 4 + [None, h_size - 2, None] + [k_size - 2, None]
    all_comms = pk.index_comms + proof.commitments[0] + proof.commitments[1] + proof.commitments[2]
    commitments = dict(zip(labels, all_comms))
    degree_bounds = dict(zip(labels, bounds))
    query_set = ahp.verifier_query_set(vs)
    fs.absorb(b"".join(T.fe_bytes(f, e) for e in proof.evaluations))
    opening_challenge = u128_rand(fs) % p
    evaluations = {}
    eval_labels = []
    for label, (_, point) in query_set:
        if label in ahp.LC_WITH_ZERO_EVAL:
            evaluations[(label, point)] = 0
        else:
            eval_labels.append((label, point))
    eval_labels.sort(key=lambda x: x[0])
    for q, e in zip(eval_labels, proof.evaluations):
        evaluations[q] = e
    lcs = ahp.construct_linear_combinations(f, public_input, lambda l, pt: evaluations[(l, pt)], vs)
    return kzg.check_combinations(ck, lcs, commitments, degree_bounds, query_set, evaluations, proof.pc_proof, opening_challenge, g2)


# ---- CanonicalSerialize -----------------------------------------------------------------------
def serialize_proof(curve, scheme, proof):
    f = curve.fr
    out = struct.pack("<Q", len(proof.commitments))
    for rnd in proof.commitments:
        out += struct.pack("<Q", len(rnd))
        for c in rnd:
            out += T.g1_compressed(curve, c.comm)
            if scheme == kzg.MARLIN:
                if c.shifted is None:
                    out += b"\x00"
                else:
                    out += b"\x01" + T.g1_compressed(curve, c.shifted)
    out += struct.pack("<Q", len(proof.evaluations))
    for e in proof.evaluations:
        out += e.to_bytes(f.nbytes, "little")
    out += struct.pack("<Q", 3) + b"\x00\x00\x00"  # three ProverMsg::EmptyMessage = Option::None
    out += struct.pack("<Q", len(proof.pc_proof))
    for w, rv in proof.pc_proof:
        out += T.g1_compressed(curve, w)
        out += b"\x00" if rv is None else b"\x01" + rv.to_bytes(f.nbytes, "little")
    out += b"\x00"  # BatchLCProof.evals = None
    return out


# ---- CanonicalDeserialize + verification from public data only ------------------------------------------------
def _g1_decompress(curve, data):
    """Inverse of transcript.g1_compressed (ark-serialize SWFlags): x little-endian, bit 7 = y is the larger root,
    bit 6 = infinity.  y = sqrt(x^3 + b) via exponent (p + 1) / 4 (both base fields are 3 mod 4)."""
    fq = curve.fq
    b = bytearray(data)
    flags = b[-1] & 0xc0
    b[-1] &= 0x3f
    if flags & 0x40:
        return None
    x = int.from_bytes(bytes(b), "little")
    p = fq.p
    assert p % 4 == 3
    rhs = (x * x * x + curve.b) % p
    y = pow(rhs, (p + 1) // 4, p)
    if y * y % p != rhs:
        raise ValueError("compressed point is not on the curve")
    larger = y > (p - y) % p
    if bool(flags & 0x80) != larger:
        y = (p - y) % p
    return (x, y)


def deserialize_proof(curve, scheme, data):
    """`Proof::deserialize` (CanonicalDeserialize) for the layout serialize_proof writes."""
    f = curve.fr
    nq = curve.fq.nbytes
    off = 0

    def u64():
        nonlocal off
        v = struct.unpack_from("<Q", data, off)[0]
        off += 8
        return v

    def point():
        nonlocal off
        P = _g1_decompress(curve, data[off:off + nq])
        off += nq
        return P

    def byte():
        nonlocal off
        v = data[off]
        off += 1
        return v

    commitments = []
    for _ in range(u64()):
        rnd = []
        for _ in range(u64()):
            c = point()
            sh = None
            if scheme == kzg.MARLIN and byte():
                sh = point()
            rnd.append(kzg.Commitment(c, sh))
        commitments.append(rnd)
    evaluations = []
    for _ in range(u64()):
        evaluations.append(int.from_bytes(data[off:off + f.nbytes], "little"))
        off += f.nbytes
    for _ in range(u64()):
        assert byte() == 0  # ProverMsg::EmptyMessage
    pc_proof = []
    for _ in range(u64()):
        w = point()
        rv = None
        if byte():
            rv = int.from_bytes(data[off:off + f.nbytes], "little")
            off += f.nbytes
        pc_proof.append((w, rv))
    assert byte() == 0 and off == len(data)
    return Proof(commitments, evaluations, pc_proof)


def verifier_key_from_public(curve, scheme, srs, num_constraints, num_variables, num_non_zero, index_comms):
    """What `Marlin::verify` needs, built from public data only: index_info, the six index commitments and the
    (trimmed) SRS.  `srs` is a kzg.UniversalParams (the trapdoor replaces the pairing, see kzg.py)."""
    f = curve.fr
    pk = IndexProverKey()
    pk.curve, pk.scheme = curve, scheme

    class _Idx:
        pass

    pk.index = _Idx()
    pk.index.info = ahp.IndexInfo(num_variables, num_constraints, num_non_zero, None)
    md = ahp.max_degree(f, num_constraints, num_variables, num_non_zero)
    pk.ck = kzg.CommitterKey(srs, md, 1, ahp.get_degree_bounds(f, pk.index.info), scheme)
    pk.index_comms = [kzg.Commitment(c, None) for c in index_comms]
    pk.vk_bytes = index_vk_bytes(curve, scheme, pk.index.info, pk.index_comms)
    return pk
