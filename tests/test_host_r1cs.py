"""CPU: the product's array-form R1CS builders (marlin_b200/r1cs.py) produce exactly the matrices and
assignments the oracle's restatement of ark-relations + the reference's padding rules produces."""
import pytest

from marlin_b200 import _lib, fields, r1cs as gr1cs
from oracle import r1cs as or1cs
from oracle import rng as orng
from oracle.params import BLS12_381


def rows_of(mat, curve_id):
    row_ptr, col, coeff = mat
    out = []
    for r in range(len(row_ptr) - 1):
        lo, hi = int(row_ptr[r]), int(row_ptr[r + 1])
        vals = [fields.fr_from_mont(curve_id, v) for v in _lib.limbs_to_ints(coeff[lo:hi])] if hi > lo else []
        out.append([(v, int(c)) for v, c in zip(vals, col[lo:hi])])
    return out


def check(g, ocirc):
    f = BLS12_381.fr
    cs = or1cs.synthesize(f, ocirc)
    a, b, c = cs.to_matrices()
    assert rows_of(g.a, 0) == a and rows_of(g.b, 0) == b and rows_of(g.c, 0) == c
    assert [fields.fr_from_mont(0, v) for v in _lib.limbs_to_ints(g.instance)] == cs.instance
    assert [fields.fr_from_mont(0, v) for v in _lib.limbs_to_ints(g.witness)] == cs.witness
    assert g.num_constraints == cs.num_constraints == g.num_variables


@pytest.mark.parametrize("nc,nv", [(100, 25), (26, 25), (25, 100), (25, 26), (25, 25), (8, 6)])
def test_test_circuit(nc, nv):
    f = BLS12_381.fr
    rng = orng.test_rng()
    a, b = orng.field_rand(f, rng), orng.field_rand(f, rng)
    check(gr1cs.test_circuit(0, a, b, nc, nv), or1cs.test_circuit(f, a, b, nc, nv))


@pytest.mark.parametrize("n,nv", [(16, 10), (64, 10), (8, 20), (1024, 10)])
def test_dummy_circuit(n, nv):
    f = BLS12_381.fr
    check(gr1cs.dummy_circuit(0, 12345, 67890, nv, n), or1cs.dummy_circuit(f, 12345, 67890, nv, n))
