#!/usr/bin/env python3
"""Profiling target: one warm PLONK prove of a synthetic 2^LOGN-gate circuit (argv[1], default 18)
after two warm-up proves.  Used under `ncu --metrics gpu__time_duration.sum` for the per-kernel
launch list of the prover (profiles/)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi                                   # noqa: E402
from kzg_snark_b200.limbs import ints_to_limbs                     # noqa: E402
from kzg_snark_b200.plonk import Indexer, Prover                   # noqa: E402
from kzg_snark_b200.plonk_synth import synthetic_circuit           # noqa: E402

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << logn
qM, qL, qR, qO, qC, perm, w = synthetic_circuit(n, 16, R, seed=logn)
idx = Indexer("bn254")
ipk, _ = idx.preprocess(qM, qL, qR, qO, qC, perm, max_degree=n + 5, tau=12345678901234567890, k1=7, k2=13)
pin = _ffi.PinnedArray((3 * n - 16, 4))
pin.array[:] = ints_to_limbs(w[16:], R)
xs = [idx.kzg.Fq(v) for v in w[:16]]
pr = Prover("bn254")
for i in range(3):
    t0 = time.perf_counter()
    pr.prove(ipk, xs, pin.array)
    print(f"prove {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", {k: round(1e3 * v, 2) for k, v in pr.timings.items()})
assert pr.last_r_zeta == 0
