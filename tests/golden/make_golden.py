#!/usr/bin/env python3
"""Writes the committed golden fixtures.  Run in the BUILD container (where /root/reference is
mounted):   python tests/golden/make_golden.py

  plonk_instance.json / r1cs_instance.json
      Sage-free decodes of the reference's pickled fixtures
      (constraint-system/*.pkl, main.py:43-48,68-79) via oracle/fixtures.py.
  oracle_vectors.json
      Seeded inputs and the ORACLE's outputs for commit / open / fft_ff / ifft_ff / coset at the
      reference's demo scale (main.py:21-35) -- regression pins for the CUDA path.  They are
      outputs of the restated oracle, not of the reference itself (SageMath / py_ecc cannot be
      imported in this image): "parity unpinned" in the sense of SURVEY.md 8c.
  public_kat.json is hand-written (public constants), not generated.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import fixtures, fft_ff as off          # noqa: E402
from oracle.curve import get_curve                  # noqa: E402
from oracle.kzg import KZGOracle, poly_eval         # noqa: E402
from oracle.params import CURVES, root_of_unity     # noqa: E402

REF = "/root/reference/constraint-system"


def hexs(v):
    return [hex(x) for x in v]


def main():
    if os.path.isdir(REF):
        p = fixtures.load_plonk_instance(os.path.join(REF, "PLONK_ARITHMETIZATION_INSTANCE.pkl"))
        json.dump({k: hexs(v) for k, v in p.items()}, open(os.path.join(HERE, "plonk_instance.json"), "w"), indent=0)
        r = fixtures.load_r1cs_instance(os.path.join(REF, "R1CS_INSTANCE.pkl"))
        json.dump({k: ([hexs(row) for row in v] if k in "ABC" else hexs(v)) for k, v in r.items()},
                  open(os.path.join(HERE, "r1cs_instance.json"), "w"), indent=0)
    out = {}
    for name in ("bn254", "bls12_381"):
        cv = get_curve(name)
        ko = KZGOracle(name)
        rng = random.Random(0xC0FFEE)
        r = cv.r
        tau = rng.randrange(1, r)
        d = 21                                              # PLONK fixture SRS: 22 points (main.py:84-85)
        ck = ko.setup(d, tau)
        polys = [[rng.randrange(r) for _ in range(m)] for m in (22, 17, 1, 5)]
        polys.append([0, 0, 3, 0, r - 1])                  # sparse, with zero coefficients (kzg.py:113)
        polys.append([])                                    # zero polynomial -> Z1 (kzg.py:109)
        comm = [cv.normalize(c) for c in ko.commit(ck, polys)]
        z, xi = rng.randrange(r), rng.randrange(r)
        proof = cv.normalize(ko.open(ck, polys[:4], z, xi))
        evals = [poly_eval(p, z, r) for p in polys[:4]]
        assert ko.check_with_tau(tau, ko.commit(ck, polys[:4]), z, evals, ko.open(ck, polys[:4], z, xi), xi)
        n = 16
        w = root_of_unity(CURVES[name], n)
        x = [rng.randrange(r) for _ in range(n)]
        out[name] = {
            "tau": hex(tau),
            "ck_affine": [hexs(cv.normalize(p)) for p in ck],
            "polys": [hexs(p) for p in polys],
            "commitments": [None if c is None else hexs(c) for c in comm],
            "open": {"k": 4, "z": hex(z), "xi": hex(xi), "proof": hexs(proof), "evals": hexs(evals)},
            "ntt": {"n": n, "w": hex(w), "x": hexs(x), "fft": hexs(off.fft_ff_int(x, w, r)),
                    "ifft": hexs(off.ifft_ff_int(x, w, r)), "shift": "0x7",
                    "coset_fft": hexs(off.coset_fft_ff_int(x, w, 7, r)),
                    "coset_ifft": hexs(off.coset_ifft_ff_int(x, w, 7, r))},
        }
    json.dump(out, open(os.path.join(HERE, "oracle_vectors.json"), "w"), indent=0)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
