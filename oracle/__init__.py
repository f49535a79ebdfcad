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

PARITY STATUS: pinned against the reference's own code, run here.  oracle/refrun.py imports
the UNMODIFIED kzg.py, fft_ff.py, transcript.py, plonk/*.py and marlin/*.py from
/root/reference (with stand-ins for the two uninstallable third-party packages:
kzg_snark_b200/sageshim.py for `sage.all`, oracle/pyecc_standin.py / oracle/pyecc_standin_bls.py --
restated G1/G2 arithmetic and ate pairings -- for `py_ecc.optimized_bn128` / `py_ecc.optimized_bls12_381`), runs main.py's three demos (both
verifiers accept, tampered proofs are rejected) and configs[0], and records every boundary call
(KZG.commit / open, fft_ff / ifft_ff / fft_ff_interpolation) with the reference's outputs in
tests/golden/ref_trace_*.json (generator: tests/golden/make_traces.py).  The oracle reproduces
all of them bit for bit (tests/test_reference_traces.py), and so does the CUDA path
(tests/test_gpu_dropin.py).  What stays restated rather than pinned is the inside of py_ecc and
SageMath (exact integer arithmetic with canonical results); for it the oracle relies on
  (i)   the mathematical definitions (DFT definition; group law on the fixed curves),
  (ii)  public curve constants / known multiples of the generators
        (tests/golden/public_kat.json), bilinearity of the restated pairing,
  (iii) the tau-identity  commit(ck, p) == p(tau)*G1  (kzg.py:108).
BLS12-381: the reference's demos and fixtures are BN254 only, so its code is also run with
curve_type="bls12_381" -- kzg.py's setup / commit / open / check / batch_check, fft_ff.py over GF(r_bls), and the PLONK and
Marlin indexers / provers / verifiers on the bundled instances (residues read as signed integers; both verifiers accept) --
and recorded the same way (ref_trace_kzg_bls.json, ref_trace_fft_bls.json, ref_{plonk,marlin}_normalized_bls.json).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (kzg_snark_b200/) never
does; it fails loudly when the CUDA library is missing.
"""
