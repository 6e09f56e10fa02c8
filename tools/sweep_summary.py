#!/usr/bin/env python3
"""Summarise a tools/sweep_bench.sh output file: one row per run with the step time and the per-kernel spans."""
import json
import sys

KEYS = ["msm_affine_levels", "msm_aff_level0", "msm_aff_level1", "msm_aff_level2", "msm_aff_level3", "msm_accumulate_kernel", "msm_stitch", "msm_reduce", "msm_sort", "ntt", "msm_allgather"]
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
print(f"{'label':22s} {'ms/step':>8s} {'e2e M/s':>8s} " + " ".join(f"{k.replace('msm_', '')[:10]:>10s}" for k in KEYS) + "  hash verified")
for r in rows:
    ln = r["line"]
    if not ln:
        print(f"{r['label']:22s} FAILED")
        continue
    steps = ln["steps"]
    k = ln.get("kernels", {})
    cells = " ".join(f"{k[x]['ms'] / steps:10.2f}" if x in k else f"{'-':>10s}" for x in KEYS)
    print(f"{r['label']:22s} {ln['ms_per_step']:8.2f} {ln['e2e']['value'] / 1e6:8.3f} {cells}  {ln.get('proof_matches_pinned_1gpu_hash')} {ln.get('proof_verified')}")
