"""Drop-in replacement for the reference's fft_ff.py: same three names, same signatures
(fft_ff.py:3, 39, 60), computed by the sm_100a NTT kernels through libkzgpu.so.

    from kzg_snark_b200.fft_ff import fft_ff, ifft_ff, fft_ff_interpolation
or put `kzg_snark_b200/dropin` ahead of the reference on sys.path so that
`from fft_ff import ...` (plonk/encoder.py:3, marlin/prover.py:4) resolves here.

`F` is the caller's field (Sage GF(r) or the shim): used as a coercion `F(x)` exactly as the
reference uses it (fft_ff.py:29-30,57); its order selects the curve.  No CPU path.
"""

import numpy as np

from . import device
from ._ffi import BN254, BLS12_381
from .limbs import ints_to_limbs, int_to_limbs, limbs_to_ints

_FIELD_BY_ORDER = {device.FR[BN254]: BN254, device.FR[BLS12_381]: BLS12_381}


def _field_id(F, sample=None):
    q = None
    for attr in ("order", "cardinality", "characteristic"):
        f = getattr(F, attr, None)
        if callable(f):
            try:
                q = int(f())
                break
            except Exception:
                pass
    if q is None:
        q = int(getattr(F, "q", 0)) or None
    if q is None and sample is not None:
        par = getattr(sample, "parent", None)
        if callable(par):
            return _field_id(par())
    if q not in _FIELD_BY_ORDER:
        raise ValueError(f"unsupported scalar field (order {q}); supported: BN254 r, BLS12-381 r")
    return _FIELD_BY_ORDER[q], q


def _transform(values, w, F, inverse, coset=None):
    fid, q = _field_id(F, w)
    data = ints_to_limbs(values, q)
    wl = int_to_limbs(w, q)
    cl = None if coset is None else int_to_limbs(coset, q)
    device.ntt(fid, data, wl, inverse=inverse, coset_limbs=cl)
    return [F(v) for v in limbs_to_ints(data)]


def fft_ff(coeffs, w, F):
    """out[k] = sum_j coeffs[j] * w^(j*k), natural order in and out (fft_ff.py:3-37)."""
    n = len(coeffs)
    if n == 1:
        return coeffs                                   # fft_ff.py:16-17: the same list object
    if n == 0 or n & (n - 1):
        # the reference silently mis-computes odd lengths; refuse instead
        raise ValueError("fft_ff: length must be a power of two")
    return _transform(coeffs, w, F, inverse=False)


def ifft_ff(values, w, F):
    """fft_ff with w^-1, then scaled by n^-1 (fft_ff.py:39-58); the scale is fused on the device."""
    n = len(values)
    if n == 1:
        ninv = F(1) ** (-1)
        return [x * ninv for x in values]
    if n == 0 or n & (n - 1):
        raise ValueError("ifft_ff: length must be a power of two")
    return _transform(values, w, F, inverse=True)


def coset_fft_ff(coeffs, w, shift, F):
    """north_star extension (no reference counterpart):
    out[k] = sum_j coeffs[j] * shift^j * w^(j*k)  ==  fft_ff([c_j * shift**j], w, F)."""
    n = len(coeffs)
    if n == 1:
        return list(coeffs)
    if n == 0 or n & (n - 1):
        raise ValueError("coset_fft_ff: length must be a power of two")
    return _transform(coeffs, w, F, inverse=False, coset=shift)


def coset_ifft_ff(values, w, shift, F):
    """Inverse of coset_fft_ff (ifft_ff followed by multiplying coefficient j by shift^-j)."""
    n = len(values)
    if n == 1:
        return list(values)
    if n == 0 or n & (n - 1):
        raise ValueError("coset_ifft_ff: length must be a power of two")
    return _transform(values, w, F, inverse=True, coset=shift)


def _poly_ring(F):
    try:
        from sage.all import PolynomialRing          # the reference's own type when Sage exists
        return PolynomialRing(F, "X")
    except ImportError:
        from .sageshim import PolynomialRing
        return PolynomialRing(F, "X")


def fft_ff_interpolation(values, g, F):
    """Polynomial through (g^i, values[i]) (fft_ff.py:60-85): same assertions, same return type."""
    n = len(values)
    assert (n & (n - 1)) == 0, "Length of values must be a power of 2"          # fft_ff.py:74
    order = g.multiplicative_order()                                            # fft_ff.py:77
    assert order >= n, f"Order of g ({order}) must be at least n ({n})"         # fft_ff.py:78
    coeffs = ifft_ff(values, g, F)
    return _poly_ring(F)(coeffs)                                                # fft_ff.py:84-85
