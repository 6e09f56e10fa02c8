"""CPU oracle for the Marlin prover hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import or execute anything under oracle/.  The product (marlin_b200/) never does.

PARITY UNPINNED: the reference (arkworks-rs/marlin, pure Rust) cannot be built in this
environment and its tests hold no golden vectors for this path; its
arithmetic lives in un-vendored crates (ark-ff/ark-ec/ark-poly/ark-poly-commit ^0.3.0).  This
package restates those algorithms from their published definitions; it is pinned by
mathematics (unique MSM / DFT / quotient values), by the reference's own algebraic unit
tests re-expressed in tests/, and by KZG soundness checks (trapdoor and pairing-free
verification of every proof), not by reference-produced bytes.
"""
