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


def _elements(F, ints):
    """canonical residues -> the caller's element type (bulk constructor of the shim field when it has one)."""
    bulk = getattr(F, "from_canonical_ints", None)
    return bulk(ints) if bulk is not None else [F(v) for v in ints]


def _transform(values, w, F, inverse, coset=None):
    fid, q = _field_id(F, w)
    data = ints_to_limbs(values, q)
    wl = int_to_limbs(w, q)
    cl = None if coset is None else int_to_limbs(coset, q)
    device.ntt(fid, data, wl, inverse=inverse, coset_limbs=cl)
    return _elements(F, limbs_to_ints(data))


def _ragged(data, w, fid, q):
    """fft_ff.py:14-37 for a length that is NOT a power of two, on (n, 4) limb rows.  The reference never validates n
    (only fft_ff_interpolation does, :74): its recursion halves ragged lists, lets zip-style indexing drop the last even
    entry and leaves result[n-1] = F(0) for odd n.  The same recursion is followed here level by level;
    every sub-list whose length IS a power of two goes through the NTT kernel, and the combining butterflies
    (fft_ff.py:32-35: e[i] +- w^i o[i]) run as element-wise device operations.  marlin/prover.py:439 can reach this
    (an un-padded `list(row_A)` whose top coefficient is zero)."""
    n = data.shape[0]
    if n == 1:
        return data
    if n & (n - 1) == 0:
        out = np.ascontiguousarray(data).copy()
        device.ntt(fid, out, int_to_limbs(w, q), inverse=False)
        return out
    w2 = w * w % q
    e, o = _ragged(data[0::2], w2, fid, q), _ragged(data[1::2], w2, fid, q)
    h = n // 2
    tw = device.powers(fid, w, h)                                              # w^0 .. w^(h-1), computed on the device
    t = device.field_op(fid, 1, 0, tw, np.ascontiguousarray(o[:h]))            # w^i * o[i]
    out = np.zeros((n, 4), dtype=np.uint64)                                    # result = [F(0)] * n   (fft_ff.py:29)
    out[:h] = device.field_op(fid, 1, 1, np.ascontiguousarray(e[:h]), t)
    out[h:2 * h] = device.field_op(fid, 1, 2, np.ascontiguousarray(e[:h]), t)
    return out


def fft_ff(coeffs, w, F):
    """out[k] = sum_j coeffs[j] * w^(j*k), natural order in and out (fft_ff.py:3-37).  Like the reference, no check of n or
    w: lengths that are not powers of two reproduce what its recursion returns (see _ragged)."""
    n = len(coeffs)
    if n == 1:
        return coeffs                                   # fft_ff.py:16-17: the same list object
    if n == 0:
        raise RecursionError("maximum recursion depth exceeded")        # fft_ff.py:20-26 on an empty list never terminates
    if n & (n - 1):
        fid, q = _field_id(F, w)
        return _elements(F, limbs_to_ints(_ragged(ints_to_limbs(coeffs, q), int(w) % q, fid, q)))
    return _transform(coeffs, w, F, inverse=False)


def ifft_ff(values, w, F):
    """fft_ff with w^-1, then scaled by n^-1 (fft_ff.py:39-58); the scale is fused on the device."""
    n = len(values)
    if n == 1:
        ninv = F(1) ** (-1)
        return [x * ninv for x in values]
    if n == 0:
        raise RecursionError("maximum recursion depth exceeded")
    if n & (n - 1):                                      # fft_ff.py:53-58 around the ragged recursion
        fid, q = _field_id(F, w)
        res = _ragged(ints_to_limbs(values, q), pow(int(w) % q, -1, q), fid, q)
        ninv = np.ascontiguousarray(np.broadcast_to(int_to_limbs(pow(n % q, -1, q), q), (n, 4)))
        return _elements(F, limbs_to_ints(device.field_op(fid, 1, 0, res, ninv)))
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
