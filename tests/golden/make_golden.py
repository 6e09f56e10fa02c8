#!/usr/bin/env python3
"""Generates tests/golden/marlin_proofs.json from the oracle (the reference itself cannot be run
here: see DESIGN.md section 5).  Each case fixes curve, PC scheme, circuit, SRS trapdoor and RNG seeds and
records sha256(index_vk ToBytes) and the serialized proof.  Both the oracle tests (CPU) and the GPU
parity tests compare against these bytes.   Usage: python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from tests_golden import CASES, regenerate_case  # noqa: E402

if __name__ == "__main__":
    out = {"generator": "tests/golden/make_golden.py (oracle/, parity unpinned by the reference)", "cases": [regenerate_case(c) for c in CASES]}
    with open(os.path.join(HERE, "marlin_proofs.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out["cases"]), "cases")
