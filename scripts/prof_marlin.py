#!/usr/bin/env python3
"""Timing target: index + warm proves of a synthetic R1CS with 2^LOGN rows (argv[1], default 14) through
kzg_snark_b200.marlin (device Indexer / Prover)."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import marlin                                  # noqa: E402

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 14
n = 1 << logn
t0 = time.perf_counter()
A, B, C, x, w = marlin.synthetic_r1cs(n, 8, R, seed=logn)
t1 = time.perf_counter()
idx = marlin.Indexer("bn254")
m = 1 << (2 * n - 1).bit_length()
ipk, _ = idx.preprocess(A, B, C, max_degree=6 * m, tau=0x1234567890abcdef)
t2 = time.perf_counter()
print(f"rows 2^{logn}: generate {t1 - t0:.2f} s, index {t2 - t1:.2f} s (|H| = {ipk['subgroups']['n']}, |K| = {ipk['subgroups']['m']})")
pr = marlin.Prover("bn254")
xs = [idx.kzg.Fq(v) for v in x]
from kzg_snark_b200.limbs import ints_to_limbs, random_scalars   # noqa: E402
draws = random_scalars(8 + 2 * n + 1, R, seed=1)                  # limb arrays: no per-element conversion inside prove
wl = ints_to_limbs(w, R)
for i in range(3):
    t0 = time.perf_counter()
    pr.prove(ipk, xs, wl, draws=draws)
    print(f"prove {i}: {1e3 * (time.perf_counter() - t0):.1f} ms", {k: round(1e3 * v, 1) for k, v in getattr(pr, 'timings', {}).items()})
assert set(pr.checks.values()) == {0}
