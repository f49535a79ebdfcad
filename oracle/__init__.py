"""CPU oracle for the kzg-snark hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on plain CPython integers, the algorithms of the
reference's two hot-path modules:

    reference kzg.py:56-159   (KZG.setup / commit / open)  -> oracle/kzg.py
    reference fft_ff.py:3-85  (fft_ff / ifft_ff / fft_ff_interpolation) -> oracle/fft_ff.py

and the third-party arithmetic those modules delegate to and that is NOT
vendored in /root/reference:

    py_ecc (PyPI, version un-pinned by the reference: README.md:15,24-25)
        optimized_bn128 / optimized_bls12_381: FQ, add, double, multiply,
        neg, eq, normalize, G1, Z1           -> oracle/curve.py
    SageMath (un-pinned: README.md:14)  GF(r) elements and PolynomialRing
        arithmetic used by kzg.py:144-154    -> oracle/field.py, oracle/poly.py

PARITY STATUS: "parity unpinned".  The reference holds no golden vectors or
known-answer tests for this path and neither SageMath
nor py_ecc can be imported in this image, so the oracle cannot be checked
against the reference's own outputs.  It is pinned instead by
  (i)   the mathematical definitions (DFT definition; group law on the fixed
        curves -> the normalised affine result is canonical),
  (ii)  public curve constants / known multiples of the generators
        (tests/golden/public_kat.json, see tests/test_oracle.py),
  (iii) the tau-identity  commit(ck, p) == p(tau)*G1  (kzg.py:108).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (kzg_snark_b200/) never
does; it fails loudly when the CUDA library is missing.
"""
