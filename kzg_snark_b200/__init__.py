"""kzg_snark_b200 -- B200 (sm_100a) KZG-MSM / NTT kernels behind the reference's kzg.py and
fft_ff.py call signatures (swusjask/kzg-snark).  See DESIGN.md.

    from kzg_snark_b200.kzg import KZG
    from kzg_snark_b200.fft_ff import fft_ff, ifft_ff, fft_ff_interpolation
    from kzg_snark_b200.plonk import Indexer, Prover          # device-resident PLONK

Importing the package does not touch the GPU; the first call does, and raises if
libkzgpu.so is missing or no sm_100 device is present (no CPU fallback).
"""

__all__ = ["_ffi", "device", "limbs"]
