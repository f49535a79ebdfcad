"""Drop-in replacement for the reference's kzg.py: class `KZG` with the same constructor,
methods and attributes (kzg.py:18-288), whose prover-side hot path -- `commit` (G1 MSM) and
`open` (xi-combination, synthetic division, MSM) -- runs on the sm_100a kernels behind
libkzgpu.so.  `setup` builds the G1 powers on the device as well.

Verifier side: the G1 linear combinations inside `check` / `batch_check`
(kzg.py:183-205, 252-281) and the `multiply` / `add` / `neg` / `eq` attributes the verifiers read run
on the device too (`kzgpu_g1_lincomb`); the two pairings and G2 stay on py_ecc as in the reference, so
`check` / `batch_check` raise ImportError at the pairing when py_ecc is not installed.

The GPU context is a process-wide singleton (six KZG instances exist in one PLONK run,
SURVEY.md section 8b); device copies of commitment keys are cached per `ck` list.
There is no CPU fallback for commit/open/setup.
"""

import collections
import weakref

import numpy as np

from . import device
from ._ffi import CURVE_IDS
from .limbs import ints_to_limbs, int_to_limbs
from .points import PointCodec

try:                                                   # the reference's own algebra types
    from sage.all import GF, PolynomialRing            # kzg.py:1
    HAVE_SAGE = True
except ImportError:                                    # not installed here: caller-side shim
    from .sageshim import GF, PolynomialRing
    HAVE_SAGE = False

_G1 = {"bn254": (1, 2), "bls12_381": (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)}


class CommitmentKey(list):
    """ck = [tau^i * G1] as a plain list of py_ecc-shaped points (what kzg.py:69-72 returns),
    carrying the handle of its device-resident copy so commit/open never re-upload it.  `_snapshot`
    is the list of point objects the device copy was built from: a key edited in place (ck[i] = ...)
    no longer equals it and is uploaded again, as the reference would read the new ck[i] (kzg.py:115).
    The device copy is released when the key is garbage collected."""
    srs = None
    _snapshot = None

    def attach(self, srs):
        self.srs = srs
        self._snapshot = list(self)
        weakref.finalize(self, srs.destroy)


def _needs_py_ecc(name):
    def stub(*a, **k):
        raise ImportError(f"py_ecc is required for the verifier-side operation `{name}` "
                          "(the pairing and G2 arithmetic of KZG.check / batch_check stay on py_ecc)")
    stub.missing = True
    return stub


# Device copies of commitment keys handed in as plain lists: id(ck) -> (snapshot, Srs).  The snapshot is a shallow copy of
# the list, so (i) it keeps every point object alive -- a recycled id() cannot alias an entry -- and (ii) `snapshot == ck`
# (C-level list comparison: identity first, value equality of the coordinates otherwise) detects any in-place edit of the
# caller's list at ~5 ns per unchanged point.  The cache owns its Srs objects; eviction releases the device memory.
_SRS_CACHE = collections.OrderedDict()
_SRS_CACHE_MAX = 8


class KZG:
    def __init__(self, curve_type="bn254"):
        if curve_type not in CURVE_IDS:
            raise ValueError(f"Unsupported curve type: {curve_type}")          # kzg.py:37
        self.curve_type = curve_type
        self._cid = CURVE_IDS[curve_type]
        self._codec = PointCodec(curve_type, self._cid)
        try:                                                                   # kzg.py:26-35
            if curve_type == "bn254":
                from py_ecc.optimized_bn128 import (G1, G2, multiply, add, curve_order, pairing,
                                                    neg, Z1, Z2, eq)
            else:
                from py_ecc.optimized_bls12_381 import (G1, G2, multiply, add, curve_order, pairing,
                                                        neg, Z1, Z2, eq)
            self.have_py_ecc = True
        except ImportError:
            self.have_py_ecc = False
            curve_order = device.FR[self._cid]
            gx, gy = _G1[curve_type]
            G1 = (self._codec.fq(gx), self._codec.fq(gy), self._codec.fq(1))
            Z1 = self._codec.Z1
            G2 = Z2 = None
            # G1 group operations without py_ecc: device-backed (kzgpu_g1_lincomb); G2 and the pairing stay py_ecc's
            multiply, add, neg, eq = self._g1_multiply, self._g1_add, self._g1_neg, self._g1_eq
            pairing = _needs_py_ecc("pairing")
        self.G1, self.G2, self.Z1, self.Z2 = G1, G2, Z1, Z2                    # kzg.py:40-49
        self.multiply, self.add, self.neg = multiply, add, neg
        self.pairing, self.eq = pairing, eq
        self.curve_order = curve_order
        self.Fq = GF(curve_order)                                              # kzg.py:52-54
        self.R = PolynomialRing(self.Fq, "X")
        self.X = self.R.gen()

    # ------------------------------------------------------------------ G1 on the device
    def g1_lincomb(self, points, scalars):
        """sum_i scalars[i] * points[i] for a few arbitrary G1 points (py_ecc-shaped triples): the
        verifier-side combinations of kzg.py:183-205 and :252-281 as one kernel launch."""
        q = self.curve_order
        if not points:
            return self.Z1
        out, inf = device.g1_lincomb(self._cid, self._codec.points_to_limbs(points), ints_to_limbs([int(s) % q for s in scalars], q))
        return self._codec.from_device(out, inf)

    def _g1_multiply(self, pt, n):
        return self.g1_lincomb([pt], [n])

    def _g1_add(self, p1, p2):
        return self.g1_lincomb([p1, p2], [1, 1])

    def _g1_neg(self, pt):
        x, y, z = pt
        return (x, self._codec.fq(-int(y)), z)

    def _g1_eq(self, p1, p2):
        p = self._codec.p
        x1, y1, z1 = (int(c) for c in p1)
        x2, y2, z2 = (int(c) for c in p2)
        return (x1 * z2 - x2 * z1) % p == 0 and (y1 * z2 - y2 * z1) % p == 0

    # ------------------------------------------------------------------ helpers
    def _coeff_limbs(self, poly):
        """polynomial (Sage / shim / list of int-likes) -> (len, 4) uint64, trailing zeros kept out."""
        q = self.curve_order
        c = getattr(poly, "c", None)
        if isinstance(c, list):                         # shim Poly: ints already
            coeffs = c
        else:
            coeffs = poly.list()                        # kzg.py:110
        if not coeffs:
            return np.zeros((0, 4), dtype=np.uint64)
        return ints_to_limbs(coeffs, q)

    def _coerce_polys(self, polynomials):
        out = []
        for poly in polynomials:                        # kzg.py:92-97 / 136-141
            out.append(self.R(poly) if isinstance(poly, list) else poly)
        return out

    def _device_srs(self, ck):
        srs = getattr(ck, "srs", None)
        if srs is not None and srs.handle and srs.curve == self._cid:
            if ck._snapshot == ck:
                return srs
            srs.destroy()                                 # edited in place since setup(): falls through to a fresh upload
            ck.srs = None
        key = (id(ck), self._cid)
        hit = _SRS_CACHE.get(key)
        if hit is not None and hit[1].handle and hit[0] == ck:
            _SRS_CACHE.move_to_end(key)
            return hit[1]
        if hit is not None:                               # same list object, edited since: drop the stale copy
            _SRS_CACHE.pop(key)[1].destroy()
        srs = device.Srs.from_affine(self._cid, self._codec.points_to_limbs(ck))
        while len(_SRS_CACHE) >= _SRS_CACHE_MAX:
            _SRS_CACHE.popitem(last=False)[1][1].destroy()
        _SRS_CACHE[key] = (list(ck), srs)
        return srs

    # ------------------------------------------------------------------ setup (kzg.py:56-78)
    def setup(self, max_degree, tau=None):
        """(ck, rk).  ck[i] = tau^i * G1 is computed on the device and kept resident; rk = tau*G2
        needs py_ecc's G2 arithmetic (None without it).  `tau` may be supplied for reproducible
        tests; by default it is drawn like the reference does (kzg.py:67)."""
        if tau is None:
            tau = self.Fq.random_element()
        t = int(tau) % self.curve_order
        srs = device.Srs.generate(self._cid, t, max_degree + 1)
        rows = srs.read(0, max_degree + 1)
        ck = CommitmentKey(self._codec.from_device(r, not r.any()) for r in rows)
        ck.attach(srs)
        rk = self.multiply(self.G2, t) if self.have_py_ecc else None
        return (ck, rk)

    # ------------------------------------------------------------------ commit (kzg.py:80-120)
    def commit(self, ck, polynomials):
        polys = self._coerce_polys(polynomials)
        max_degree = len(ck) - 1
        for poly in polys:
            if poly.degree() > max_degree:              # kzg.py:103-106, same message
                raise ValueError(
                    f"Polynomial degree {poly.degree()} exceeds maximum allowed degree {max_degree}"
                )
        if not polys:
            return []
        srs = self._device_srs(ck)
        out, infs = device.msm_batch(srs, [self._coeff_limbs(p) for p in polys])
        return [self._codec.from_device(o, f) for o, f in zip(out, infs)]

    # ------------------------------------------------------------------ open (kzg.py:122-159)
    def open(self, ck, polynomials, z, xi):
        polys = self._coerce_polys(polynomials)
        q = self.curve_order
        z = self.Fq(z)
        xi = self.Fq(xi)
        srs = self._device_srs(ck)
        # the degree check is commit's (kzg.py:103-106 via :157): the library applies it to the quotient and reports it with the
        # reference's message (KZGPU_ERANGE -> ValueError in device.open_proof)
        out, inf = device.open_proof(srs, [self._coeff_limbs(p) for p in polys],
                                     int_to_limbs(z, q), int_to_limbs(xi, q))
        return self._codec.from_device(out, inf)

    # ------------------------------------------------------------------ verifier side (py_ecc)
    def check(self, rk, commitments, z, evaluations, proof, xi):
        """e(C - v*G1, G2) == e(proof, tau*G2 - z*G2) with C, v the xi-combinations
        (kzg.py:161-211).  Runs on py_ecc like the reference."""
        if getattr(self.pairing, "missing", False):
            self.pairing()                              # ImportError before any device work
        z = self.Fq(z)
        xi = self.Fq(xi)
        v = self.Fq(0)
        for i, e in enumerate(evaluations):
            v += xi ** (i + 1) * self.Fq(e)
        # C - v*G1 = sum_i xi^(i+1) C_i - v G1: one device combination (kzg.py:183-201)
        lhs_pt = self.g1_lincomb(list(commitments) + [self.G1], [int(xi ** (i + 1)) for i in range(len(commitments))] + [-int(v)])
        rhs_g2 = self.add(rk, self.neg(self.multiply(self.G2, int(z))))
        return self.pairing(self.G2, lhs_pt) == self.pairing(rhs_g2, proof)

    def batch_check(self, rk, commitments_list, z_list, evaluations_list, proof_list, xi_list, r=None):
        """One pairing equation for several openings (kzg.py:213-288):
        e(sum r^i (C_i - v_i G1 + z_i pi_i), G2) == e(sum r^i pi_i, tau G2)."""
        if getattr(self.pairing, "missing", False):
            self.pairing()
        if r is None:
            r = self.Fq.random_element()
        r = self.Fq(r)
        # left = sum_i r^(i+1) (sum_j xi_i^(j+1) C_ij - v_i G1 + z_i pi_i), right = sum_i r^(i+1) pi_i  (kzg.py:252-281):
        # every term is scalar * point, so each side is one device combination
        lp, ls, rp_, rs = [], [], [], []
        for i, (commitments, z, evaluations, proof, xi) in enumerate(
                zip(commitments_list, z_list, evaluations_list, proof_list, xi_list)):
            z = self.Fq(z)
            xi = self.Fq(xi)
            rp = r ** (i + 1)
            v = self.Fq(0)
            for j, comm in enumerate(commitments):
                xp = xi ** (j + 1)
                lp.append(comm); ls.append(int(rp * xp))
                v += xp * self.Fq(evaluations[j])
            lp += [self.G1, proof]; ls += [-int(rp * v), int(rp * z)]
            rp_.append(proof); rs.append(int(rp))
        left = self.g1_lincomb(lp, ls)
        right = self.g1_lincomb(rp_, rs)
        return self.pairing(self.G2, left) == self.pairing(rk, right)
