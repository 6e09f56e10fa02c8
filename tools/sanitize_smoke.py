"""Small prove + MSM/NTT calls for compute-sanitizer runs (memcheck / initcheck / racecheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from marlin_b200 import _lib, api, r1cs

m = api.Marlin("bls12_381", "marlin_kzg10", device=0)
n = 256
circ = r1cs.dummy_circuit(m.curve_id, 1234567, 7654321, 10, n)
srs = m.universal_setup(n, n, 3 * n, beta=0x1234567, gamma=7, degree_bounds=(n - 2, 4 * n - 2))
pk = m.index(srs, circ)
proof = m.prove(pk, circ, api.ZkRng.test_rng())
m2 = api.Marlin("bls12_381", "sonic_kzg10", ctx=m.ctx)
pk2 = m2.index(srs, circ)
proof2 = m2.prove(pk2, circ, api.ZkRng.test_rng())
buf = np.arange(4 * 1024, dtype=np.uint64).reshape(1024, 4)
_lib.check(_lib.lib().b2m_ntt(m.ctx.handle, 0, _lib.ptr(buf), 10, 0, 1))
print("sanitize_smoke ok", len(proof), len(proof2))
