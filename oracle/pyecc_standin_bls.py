"""Stand-in for `py_ecc.optimized_bls12_381` (reference kzg.py:32-35), the BLS12-381 instance of
oracle/pyecc_standin.py: same names (G1, G2, multiply, add, curve_order, pairing, neg, Z1, Z2, eq), G1 arithmetic from
oracle/curve.py, G2 over Fp2 = Fp[i]/(i^2 + 1) on the twist y^2 = x^3 + 4(1 + i), ate pairing into
Fp12 = Fp[w]/(w^12 - 2 w^6 + 2).  Oracle / test infrastructure only (see oracle/__init__.py); tests/test_oracle.py
checks the generator, its order, the twist map and bilinearity."""
from .pyecc_standin import make

globals().update(make("bls12_381"))
