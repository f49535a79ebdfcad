"""G1 point values at the drop-in boundary.

The reference's callers hold commitments as py_ecc "optimized" points: homogeneous projective
triples (x, y, z) of FQ with Z1 = (1, 1, 0) (bound at kzg.py:27-35,40-49).  The
GPU returns the canonical affine point; it is presented as (FQ(x), FQ(y), FQ(1)), or Z1 for the
identity, using py_ecc's own FQ class when py_ecc is importable and the value class below when
it is not (py_ecc is absent from this image).
"""

import numpy as np

from .device import FP, FP_LIMBS
from .limbs import limbs_to_ints


class FQ:
    """Minimal stand-in for py_ecc's optimized FQ: `.n`, int(), ==, printable as the integer."""
    __slots__ = ("n", "field_modulus")

    def __init__(self, n, field_modulus):
        self.n = int(n) % field_modulus
        self.field_modulus = field_modulus

    def __int__(self):
        return self.n

    __index__ = __int__

    def __eq__(self, o):
        if isinstance(o, FQ):
            return self.n == o.n
        if isinstance(o, int):
            return self.n == o % self.field_modulus
        return NotImplemented

    def __hash__(self):
        return hash(self.n)

    def __repr__(self):
        return repr(self.n)


def fq_class(curve_name):
    """py_ecc's FQ for the curve if available, else a factory for the stand-in."""
    try:
        if curve_name == "bn254":
            from py_ecc.fields import optimized_bn128_FQ as _FQ
        else:
            from py_ecc.fields import optimized_bls12_381_FQ as _FQ
        return _FQ
    except ImportError:
        return None


class PointCodec:
    def __init__(self, curve_name, curve_id):
        self.curve_name = curve_name
        self.cid = curve_id
        self.p = FP[curve_id]
        self.nl = FP_LIMBS[curve_id]
        self._fq = fq_class(curve_name)

    def fq(self, v):
        return self._fq(v) if self._fq is not None else FQ(v, self.p)

    @property
    def Z1(self):
        return (self.fq(1), self.fq(1), self.fq(0))

    def from_device(self, limbs, is_inf):
        """(2*nl,) uint64 affine -> py_ecc-shaped projective triple."""
        if is_inf:
            return self.Z1
        x, y = limbs_to_ints(np.asarray(limbs).reshape(2, self.nl))
        return (self.fq(x), self.fq(y), self.fq(1))

    def to_affine_ints(self, pt):
        """py_ecc triple / affine pair / int tuples -> (x, y) ints, (0, 0) for the identity."""
        p = self.p
        if len(pt) == 2:
            return int(pt[0]) % p, int(pt[1]) % p
        x, y, z = (int(c) % p for c in pt)
        if z == 0:
            return 0, 0
        if z == 1:
            return x, y
        zi = pow(z, -1, p)
        return x * zi % p, y * zi % p

    def points_to_limbs(self, pts):
        """list of points -> (n, 2*nl) uint64 canonical affine rows.  One batched inversion
        for the projective representatives the reference's setup produces (kzg.py:72)."""
        p, nl = self.p, self.nl
        n = len(pts)
        zs, trip = [], []
        for pt in pts:
            if len(pt) == 2:
                trip.append((int(pt[0]) % p, int(pt[1]) % p, 1))
            else:
                trip.append((int(pt[0]) % p, int(pt[1]) % p, int(pt[2]) % p))
        # Montgomery batch inversion over the non-trivial z
        idx = [i for i, t in enumerate(trip) if t[2] not in (0, 1)]
        pref, acc = [], 1
        for i in idx:
            pref.append(acc)
            acc = acc * trip[i][2] % p
        inv = pow(acc, -1, p) if idx else 1
        zinv = {}
        for k in range(len(idx) - 1, -1, -1):
            i = idx[k]
            zinv[i] = inv * pref[k] % p
            inv = inv * trip[i][2] % p
        nbytes = nl * 8
        buf = bytearray(n * 2 * nbytes)
        off = 0
        for i, (x, y, z) in enumerate(trip):
            if z == 0:
                x = y = 0
            elif z != 1:
                x = x * zinv[i] % p
                y = y * zinv[i] % p
            buf[off:off + nbytes] = x.to_bytes(nbytes, "little")
            buf[off + nbytes:off + 2 * nbytes] = y.to_bytes(nbytes, "little")
            off += 2 * nbytes
        return np.frombuffer(bytes(buf), dtype="<u8").reshape(n, 2 * nl).copy()
