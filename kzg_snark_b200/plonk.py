"""Device-resident PLONK indexer and prover with the reference's class and method names
(plonk/indexer.py:20 `Indexer.preprocess`, plonk/prover.py:24 `Prover.prove`) -- SURVEY.md
section 8f N3: the callers on either side of the commit/open/fft_ff hot path.

The reference's prover holds Sage polynomials and calls kzg.commit / kzg.open /
fft_ff_interpolation between stretches of Sage polynomial arithmetic (products, long division by
v_H).  Here every polynomial lives in HBM from the wire values to the opening proofs; only
scalars (challenges, evaluations, blinding factors) and the 9 proof points cross PCIe:

  round 1  wire values -> iNTT -> blinding -> 3 MSMs                     plonk/prover.py:78-93
  round 2  grand product z (kzgpu_plonk_permutation_dev) -> iNTT -> MSM  plonk/prover.py:101-117
  round 3  13 coset NTTs (size 4n), point-wise quotient, 1 inverse coset NTT, split, 3 MSMs
                                                                          plonk/prover.py:124-141
  round 4  6 evaluations (parallel Horner)                               plonk/prover.py:147-158
  round 5  r(X) as one linear combination, two KZG openings              plonk/prover.py:162-185

The Fiat-Shamir transcript is the reference's (transcript.py:18-100, restated), fed with the
same messages in the same order; given the same SRS, circuit, blinding factors and k1/k2 the
proof equals, bit for bit, the one plonk/prover.py produces when its KZG returns normalised
points (tests/golden/ref_plonk_normalized.json; tests/test_gpu_plonk.py).

There is no CPU path: everything below goes through libkzgpu.so.
"""

import ctypes
import os
import hashlib
import random
import struct
import time

import numpy as np

from . import _ffi, device
from ._ffi import check, ptr
from .kzg import KZG
from .limbs import ints_to_limbs, int_to_limbs, limbs_to_int, limbs_to_ints

_GEN = {_ffi.BN254: 5, _ffi.BLS12_381: 7}          # least primitive roots of the two scalar fields


class Transcript:
    """transcript.py:18-100: SHA-256 chain over (state || label || data)."""

    def __init__(self, label, F):
        self.F = F
        self.state = hashlib.sha256(label.encode()).digest()

    def _serialize(self, data):
        if isinstance(data, str):
            return data.encode()
        if isinstance(data, int):
            return struct.pack(">q", data)              # transcript.py:69-70
        if isinstance(data, bytes):
            return data
        if isinstance(data, list):
            return b"".join(self._serialize(item) for item in data)
        return str(data).encode()                       # field elements and point tuples: str()

    def _update(self, label, data):
        h = hashlib.sha256()
        h.update(self.state)
        h.update(label.encode())
        h.update(data)
        self.state = h.digest()

    def append_message(self, label, data):
        self._update(label, self._serialize(data))

    def get_challenge(self, label):
        cs = hashlib.sha256(self.state + label.encode()).digest()
        c = self.F(int.from_bytes(cs, byteorder="big"))
        self._update(label, cs)
        return c


_POOL = {}          # capacity in bytes -> idle DeviceBuffers (cudaMalloc / cudaFree stay out of the prove loop)
_POOL_BYTES = 0     # bytes parked in the pool
# Idle buffers above this many bytes go back to the driver instead of being parked: a long-lived process that proves circuits of
# many different sizes must not grow monotonically.  Default 8 GiB (a 2^20-gate prove parks ~3 GiB); KZGPU_POOL_GIB overrides.
_POOL_CAP = int(float(os.environ.get("KZGPU_POOL_GIB", "8")) * (1 << 30))


def release_pool():
    """Return every pooled device buffer to the driver."""
    global _POOL_BYTES
    for bufs in _POOL.values():
        for b in bufs:
            b.free()
    _POOL.clear()
    _POOL_BYTES = 0


class DVec:
    """A vector of scalar-field elements in HBM (canonical limbs, 32 B each).  Buffers are
    recycled through a size-keyed pool: `free()` parks the allocation, the next vector of the
    same size reuses it."""

    def __init__(self, n, zero=False):
        self.n = int(n)
        cap = max(self.n, 1) * 32
        global _POOL_BYTES
        idle = _POOL.get(cap)
        if idle:
            self.buf = idle.pop()
            _POOL_BYTES -= cap
        else:
            self.buf = _ffi.DeviceBuffer(cap)
        if zero:
            check(_ffi._lib.kzgpu_memset(self.buf.ptr, 0, self.n * 32))

    @property
    def ptr(self):
        return self.buf.ptr

    def at(self, i):
        return ctypes.c_void_p(self.buf.ptr.value + 32 * i)

    @classmethod
    def from_limbs(cls, arr, n=None):
        arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
        v = cls(n or arr.shape[0], zero=bool(n and n > arr.shape[0]))
        if arr.shape[0]:
            check(_ffi._lib.kzgpu_h2d(v.ptr, ptr(arr), arr.nbytes))
        return v

    def write(self, i, ints, r):
        a = ints_to_limbs(ints, r)
        check(_ffi._lib.kzgpu_h2d(self.at(i), ptr(a), a.nbytes))

    def read(self, i=0, count=None):
        count = self.n - i if count is None else count
        out = np.zeros((count, 4), dtype=np.uint64)
        if count:
            check(_ffi._lib.kzgpu_d2h(ptr(out), self.at(i), out.nbytes))
        return out

    def read_ints(self, i=0, count=None):
        return limbs_to_ints(self.read(i, count)) if (count is None or count) else []

    def copy_from(self, src, count, dst_off=0, src_off=0):
        check(_ffi._lib.kzgpu_d2d(self.at(dst_off), src.at(src_off), count * 32))

    def free(self):
        global _POOL_BYTES
        if self.buf is not None:
            if _POOL_BYTES + self.buf.nbytes > _POOL_CAP:
                self.buf.free()                                   # the pool is full: back to the driver
            else:
                _POOL.setdefault(self.buf.nbytes, []).append(self.buf)
                _POOL_BYTES += self.buf.nbytes
            self.buf = None

    def __del__(self):
        try:
            self.free()
        except Exception:                      # interpreter shutdown: module globals may be gone
            pass


class _View:
    """A window of a DVec starting at element `off` (same interface, no ownership)."""

    def __init__(self, base, off):
        self.base, self.off = base, off

    @property
    def ptr(self):
        return self.base.at(self.off)

    def at(self, i):
        return self.base.at(self.off + i)

    def write(self, i, ints, r):
        self.base.write(self.off + i, ints, r)

    def read_ints(self, i=0, count=None):
        return self.base.read_ints(self.off + i, count)

    def copy_from(self, src, count, dst_off=0, src_off=0):
        self.base.copy_from(src, count, self.off + dst_off, src_off)


def _voidp_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[p.value if isinstance(p, ctypes.c_void_p) else p for p in ptrs])


class _Field:
    """Scalar-field helpers bound to one curve id."""

    def __init__(self, cid):
        self.cid = cid
        self.r = device.FR[cid]
        self.lib = _ffi.init()

    def L(self, v):
        return int_to_limbs(int(v) % self.r, self.r)

    def root(self, n):
        assert n & (n - 1) == 0 and (self.r - 1) % n == 0
        return pow(_GEN[self.cid], (self.r - 1) // n, self.r)

    def intt(self, vec, n, w):
        check(self.lib.kzgpu_ntt_dev(self.cid, vec.ptr, n, ptr(self.L(w)), 1, None))

    def coset_ntt(self, vec, n, w, shift, inverse=False):
        check(self.lib.kzgpu_ntt_dev(self.cid, vec.ptr, n, ptr(self.L(w)), 1 if inverse else 0, ptr(self.L(shift))))

    def powers(self, vec, n, base, scale=None):
        check(self.lib.kzgpu_powers_dev(self.cid, vec.ptr, n, ptr(self.L(base)), None if scale is None else ptr(self.L(scale))))

    def eval(self, vec, length, x):
        out = np.zeros(4, dtype=np.uint64)
        check(self.lib.kzgpu_poly_eval_dev(self.cid, vec.ptr, length, ptr(self.L(x)), ptr(out)))
        return limbs_to_int(out)

    def lincomb(self, out, out_len, terms, constant=None):
        """terms: [(pointer, length, scalar)]"""
        k = len(terms)
        ptrs = _voidp_array([t[0] for t in terms])
        lens = (ctypes.c_size_t * max(k, 1))(*[t[1] for t in terms])
        sc = ints_to_limbs([t[2] for t in terms], self.r) if k else np.zeros((1, 4), np.uint64)
        check(self.lib.kzgpu_poly_lincomb_dev(self.cid, out.ptr, out_len, ptrs, lens, ptr(sc), k,
                                              None if constant is None else ptr(self.L(constant))))


class Indexer:
    """plonk/indexer.py:8-118 on the device.  `preprocess` returns (ipk, ivk) with the reference's
    keys; polynomials are `DVec`s and ipk["ck"] is the device-resident key (`device.Srs`)."""

    def __init__(self, curve_type="bn254"):
        self.kzg = KZG(curve_type=curve_type)

    def preprocess(self, qM, qL, qR, qO, qC, perm, max_degree, *, tau=None, k1=None, k2=None, ck=None, rng=None):
        """Selector values (length n = 2^k; lists of ints / field elements or (n,4) limb arrays),
        the wire permutation `perm` (3n indices) and the SRS degree bound.  tau / k1 / k2 are drawn
        from `rng` when not given (kzg.py:67, plonk/encoder.py:80-91); `ck` may be an existing
        commitment key (list of points or device.Srs)."""
        kzg = self.kzg
        cid, r = kzg._cid, kzg.curve_order
        f = _Field(cid)
        rng = rng or random.SystemRandom()
        n = len(qM)
        assert n >= 8 and n & (n - 1) == 0, "the number of gates must be a power of two >= 8"
        assert max_degree >= n + 5, "PLONK needs an SRS of degree n + 5 (t_hi has n + 6 coefficients)"
        assert len(perm) == 3 * n
        g = f.root(n)
        if k1 is None or k2 is None:                                   # plonk/encoder.py:80-91
            while True:
                k1, k2 = rng.randrange(r), rng.randrange(r)
                if k1 and k2 and pow(k1, n, r) != 1 and pow(k2, n, r) != 1 and pow(k1 * pow(k2, -1, r), n, r) != 1:
                    break
        k1, k2 = int(k1) % r, int(k2) % r
        if ck is None:
            tau = rng.randrange(1, r) if tau is None else int(tau) % r
            srs = device.Srs.generate(cid, tau, max_degree + 1)
        elif isinstance(ck, device.Srs):
            srs = ck
        else:
            srs = device.Srs.from_affine(cid, kzg._codec.points_to_limbs(ck))
        assert srs.n >= n + 6

        def to_limbs(v):
            return v if isinstance(v, np.ndarray) else ints_to_limbs(v, r)

        # subgroup H and the two cosets (plonk/encoder.py:45-48)
        H = DVec(n)
        f.powers(H, n, g)
        sigma_src = DVec(3 * n)
        f.powers(sigma_src, n, g)
        check(f.lib.kzgpu_powers_dev(cid, sigma_src.at(n), n, ptr(f.L(g)), ptr(f.L(k1))))
        check(f.lib.kzgpu_powers_dev(cid, sigma_src.at(2 * n), n, ptr(f.L(g)), ptr(f.L(k2))))
        ident = sigma_src.read()                                        # [H | k1 H | k2 H]
        sigma_vals = np.ascontiguousarray(ident[np.asarray(perm, dtype=np.int64)])   # plonk/encoder.py:126-137
        sigma_src.free()
        sigma_star = DVec.from_limbs(sigma_vals)

        # selector and permutation polynomials: iNTT of the values (plonk/encoder.py:99-104,139-141)
        polys, commitments = {}, {}
        names = ["qM", "qL", "qR", "qO", "qC", "S_sigma1", "S_sigma2", "S_sigma3"]
        values = [to_limbs(qM), to_limbs(qL), to_limbs(qR), to_limbs(qO), to_limbs(qC),
                  sigma_vals[:n], sigma_vals[n:2 * n], sigma_vals[2 * n:]]
        for name, val in zip(names, values):
            v = DVec.from_limbs(val)
            f.intt(v, n, g)
            polys[name] = v
            out, inf = device.msm_dev(srs, v, n)                        # plonk/indexer.py:76
            commitments[name] = kzg._codec.from_device(out, inf)

        # prover-key precomputation: evaluations on the coset s*<w_4n> used by the quotient
        n4 = 4 * n
        w4 = f.root(n4)
        shift = _GEN[cid]
        # (kept in Montgomery form, value * 2^256 mod r: the NTT is linear, so scaling the n coefficients once saves the
        # quotient kernel a conversion per operand and point)
        rm = pow(2, 256, r)
        coset = {}
        for name in names:
            e = DVec(n4, zero=True)
            f.lincomb(e, n, [(polys[name].ptr, n, rm)])
            f.coset_ntt(e, n4, w4, shift)
            coset[name] = e
        l1 = DVec(n4, zero=True)                                        # L1 = (X^n - 1)/(n (X - 1)) = (1/n) sum X^i
        f.powers(l1, n, 1, pow(n, -1, r) * rm % r)
        f.coset_ntt(l1, n4, w4, shift)
        xs = DVec(n4)
        f.powers(xs, n4, w4, shift * rm % r)
        sn = pow(shift, n, r)
        iota = pow(w4, n, r)                                            # primitive 4th root of unity
        zh_inv = [pow((sn * pow(iota, k, r) - 1) % r, -1, r) for k in range(4)]

        sub = {"n": n, "g": kzg.Fq(g), "k1": kzg.Fq(k1), "k2": kzg.Fq(k2)}
        ipk = {
            "ck": srs, "polynomials": polys, "commitments": commitments,
            "subgroups": {**sub, "H": H}, "sigma_star": sigma_star,
            "vanishing_poly": ("X^n - 1", n),
            "coset": {"n4": n4, "w4": w4, "shift": shift, "evals": coset, "L1": l1, "X": xs, "zh_inv": zh_inv, "mont": rm},
        }
        # rk = tau * G2 needs G2 arithmetic, which stays with py_ecc as in the reference (kzg.py:75); None without it
        rk = kzg.multiply(kzg.G2, tau) if (kzg.have_py_ecc and tau is not None) else None
        # the trapdoor never leaves this function: the verifier key carries rk = tau * G2 only (plonk/indexer.py:109-110)
        ivk = {"rk": rk, "commitments": commitments, "subgroups": sub}
        del tau
        return ipk, ivk


class Prover:
    """plonk/prover.py:7-206 on the device."""

    def __init__(self, curve_type="bn254"):
        self.kzg = KZG(curve_type=curve_type)
        self.timings = {}

    def prove(self, ipk, x, w, blinders=None):
        """ipk from `Indexer.preprocess`; x public inputs, w the remaining wire values so that
        x + w = a-values | b-values | c-values (plonk/prover.py:66-81).  `blinders` = the 11
        blinding scalars b1..b11 (drawn at plonk/prover.py:72-75,346); random when omitted."""
        kzg = self.kzg
        cid, r, Fq = kzg._cid, kzg.curve_order, kzg.Fq
        f = _Field(cid)
        lib = f.lib
        srs = ipk["ck"]
        P = ipk["polynomials"]
        sub = ipk["subgroups"]
        n, g, k1, k2 = sub["n"], int(sub["g"]), int(sub["k1"]), int(sub["k2"])
        cs = ipk["coset"]
        n4, w4, shift = cs["n4"], cs["w4"], cs["shift"]
        if blinders is None:
            sr = random.SystemRandom()
            blinders = [sr.randrange(r) for _ in range(11)]
        b1, b2, b3, b4, b5, b6, b7, b8, b9, b10, b11 = (int(b) % r for b in blinders)

        def commit(vec, length):
            out, inf = device.msm_dev(srs, vec, length)
            return kzg._codec.from_device(out, inf)

        self.timings = tm = {}
        clock = [time.perf_counter()]

        def lap(name):
            check(lib.kzgpu_sync())
            now = time.perf_counter()
            tm[name] = tm.get(name, 0.0) + now - clock[0]
            clock[0] = now

        transcript = Transcript("plonk-proof", Fq)
        transcript.append_message("public-inputs", x)                   # plonk/prover.py:57

        wires = DVec(3 * n)                                             # a | b | c values on H
        xl = ints_to_limbs(x, r)
        wl = w.reshape(-1, 4) if isinstance(w, np.ndarray) else ints_to_limbs(w, r)   # a (pinned) limb array is copied as is
        assert xl.shape[0] + wl.shape[0] == 3 * n, "x + w must hold 3n wire values"
        check(lib.kzgpu_h2d(wires.at(0), ptr(xl), xl.nbytes))
        check(lib.kzgpu_h2d(wires.at(xl.shape[0]), ptr(np.ascontiguousarray(wl, dtype=np.uint64)), wl.nbytes))

        # PI(X) = -sum x_i L_i(X): values -x_i on the first len(x) points of H (plonk/encoder.py:218-223)
        pi = DVec.from_limbs(ints_to_limbs([-int(v) for v in x], r), n)
        f.intt(pi, n, g)

        # ---- round 1 (plonk/prover.py:78-93): wire polynomials with degree-1 blinding
        wire_buf = DVec(3 * (n + 2), zero=True)                         # a | b | c coefficients, one batched MSM
        wire_polys = [_View(wire_buf, j * (n + 2)) for j in range(3)]
        for j, (bh, bl) in enumerate(((b1, b2), (b3, b4), (b5, b6))):
            p = wire_polys[j]
            p.copy_from(wires, n, 0, j * n)
            f.intt(p, n, g)
            lo = p.read_ints(0, 2)                                      # (bh X + bl)(X^n - 1) + interp
            p.write(0, [lo[0] - bl, lo[1] - bh], r)
            p.write(n, [bl, bh], r)
        a_poly, b_poly, c_poly = wire_polys
        lap("round1_upload_intt")
        outs, infs = device.msm_batch_dev(srs, wire_buf, n + 2, 3)
        wire_commitments = [kzg._codec.from_device(o, i) for o, i in zip(outs, infs)]
        lap("round1_msm")
        transcript.append_message("round1-commitments", wire_commitments)
        beta = transcript.get_challenge("beta")
        gamma = transcript.get_challenge("gamma")

        # ---- round 2 (plonk/prover.py:101-117, 245-261): permutation polynomial
        z_poly = DVec(n + 3, zero=True)
        zero_den = ctypes.c_int(0)
        check(lib.kzgpu_plonk_permutation_dev(cid, n, wires.at(0), wires.at(n), wires.at(2 * n), ipk["sigma_star"].ptr,
                                              sub["H"].ptr, ptr(f.L(k1)), ptr(f.L(k2)), ptr(f.L(beta)), ptr(f.L(gamma)),
                                              z_poly.ptr, ctypes.byref(zero_den)))
        if zero_den.value:
            raise ValueError("Denominator is zero in permutation polynomial calculation")   # plonk/prover.py:255
        f.intt(z_poly, n, g)
        lo = z_poly.read_ints(0, 3)                                     # (b7 X^2 + b8 X + b9)(X^n - 1) + interp
        z_poly.write(0, [lo[0] - b9, lo[1] - b8, lo[2] - b7], r)
        z_poly.write(n, [b9, b8, b7], r)
        lap("round2_grand_product_intt")
        z_commit = commit(z_poly, n + 3)
        lap("round2_msm")
        transcript.append_message("round2-commitment", z_commit)
        alpha = transcript.get_challenge("alpha")

        # ---- round 3 (plonk/prover.py:124-141, 297-351): quotient on the coset, split in three
        ev = {}
        for name, vec, length in (("a", a_poly, n + 2), ("b", b_poly, n + 2), ("c", c_poly, n + 2),
                                  ("z", z_poly, n + 3), ("PI", pi, n)):
            e = DVec(n4, zero=True)
            f.lincomb(e, length, [(vec.ptr, length, cs["mont"])])       # copy scaled by 2^256: Montgomery-form evaluations
            f.coset_ntt(e, n4, w4, shift)
            ev[name] = e
        ce = cs["evals"]
        order = [ev["a"], ev["b"], ev["c"], ev["z"], ce["qM"], ce["qL"], ce["qR"], ce["qO"], ce["qC"],
                 ce["S_sigma1"], ce["S_sigma2"], ce["S_sigma3"], ev["PI"], cs["L1"], cs["X"]]
        params = ints_to_limbs([int(alpha), int(beta), int(gamma), k1, k2] + cs["zh_inv"], r)
        t = DVec(n4)
        check(lib.kzgpu_plonk_quotient_dev(cid, n4, _voidp_array([v.ptr for v in order]), ptr(params), 1, t.ptr))
        f.coset_ntt(t, n4, w4, shift, inverse=True)
        for e in ev.values():
            e.free()
        t_buf = DVec(3 * (n + 6), zero=True)                            # t_lo | t_mid | t_hi, padded to n + 6 each
        t_lo, t_mid, t_hi = (_View(t_buf, j * (n + 6)) for j in range(3))
        f.lincomb(t_lo, n + 1, [(t.at(0), n, 1)])                       # t_lo + b10 X^n
        t_lo.write(n, [b10], r)
        f.lincomb(t_mid, n + 1, [(t.at(n), n, 1)], constant=-b10)       # t_mid - b10 + b11 X^n
        t_mid.write(n, [b11], r)
        f.lincomb(t_hi, n + 6, [(t.at(2 * n), n + 6, 1)], constant=-b11)  # t_hi - b11
        lap("round3_quotient")
        outs, infs = device.msm_batch_dev(srs, t_buf, n + 6, 3)
        t_commitments = [kzg._codec.from_device(o, i) for o, i in zip(outs, infs)]
        lap("round3_msm")
        transcript.append_message("round3-commitments", t_commitments)
        zeta = transcript.get_challenge("zeta")

        # ---- round 4 (plonk/prover.py:147-158)
        zi = int(zeta)
        a_z, b_z, c_z = f.eval(a_poly, n + 2, zi), f.eval(b_poly, n + 2, zi), f.eval(c_poly, n + 2, zi)
        s1_z, s2_z = f.eval(P["S_sigma1"], n, zi), f.eval(P["S_sigma2"], n, zi)
        zw_z = f.eval(z_poly, n + 3, zi * g % r)
        evaluations = [Fq(v) for v in (a_z, b_z, c_z, s1_z, s2_z, zw_z)]
        lap("round4_evaluations")
        transcript.append_message("round4-evaluations", evaluations)
        v = transcript.get_challenge("v")

        # ---- round 5 (plonk/prover.py:162-185, 383-407): linearisation polynomial and openings
        al, be, ga = int(alpha), int(beta), int(gamma)
        zn = pow(zi, n, r)
        zh_z = (zn - 1) % r
        l1_z = zh_z * pow(n * (zi - 1) % r, -1, r) % r
        pi_z = f.eval(pi, n, zi)
        perm1 = al * (a_z + be * zi + ga) % r * (b_z + be * k1 * zi + ga) % r * (c_z + be * k2 * zi + ga) % r
        ab2 = al * (a_z + be * s1_z + ga) % r * (b_z + be * s2_z + ga) % r * zw_z % r
        terms = [
            (P["qM"].ptr, n, a_z * b_z % r), (P["qL"].ptr, n, a_z), (P["qR"].ptr, n, b_z), (P["qO"].ptr, n, c_z),
            (P["qC"].ptr, n, 1),
            (z_poly.ptr, n + 3, (perm1 + al * al % r * l1_z) % r),
            (P["S_sigma3"].ptr, n, -ab2 * be % r),
            (t_lo.ptr, n + 1, -zh_z % r), (t_mid.ptr, n + 1, -zh_z * zn % r), (t_hi.ptr, n + 6, -zh_z * zn % r * zn % r),
        ]
        const = (pi_z - ab2 * (c_z + ga) - al * al % r * l1_z) % r
        r_poly = DVec(n + 6)
        f.lincomb(r_poly, n + 6, terms, constant=const)

        def quotient_into(dst, polys, point):
            """(sum_j v^(j+1) p_j - value) / (X - point) left on the device (kzg.py:147-154)."""
            k = len(polys)
            lens = (ctypes.c_size_t * k)(*[ln for _, ln in polys])
            ql = ctypes.c_size_t(0)
            check(lib.kzgpu_open_quotient_dev(cid, _voidp_array([p.ptr for p, _ in polys]), lens, k, ptr(f.L(point)), ptr(f.L(v)),
                                              dst.ptr, ctypes.byref(ql), None))
            return ql.value

        lap("round5_linearisation")
        # the two opening proofs are the commitments of two quotients: formed on the device, committed in one MSM pass
        q_buf = DVec(2 * (n + 5), zero=True)
        quotient_into(_View(q_buf, 0), [(r_poly, n + 6), (a_poly, n + 2), (b_poly, n + 2), (c_poly, n + 2),
                                        (P["S_sigma1"], n), (P["S_sigma2"], n)], zi)            # n + 5 coefficients
        quotient_into(_View(q_buf, n + 5), [(z_poly, n + 3)], zi * g % r)                         # n + 2 coefficients
        outs, infs = device.msm_batch_dev(srs, q_buf, n + 5, 2)
        W_z, W_zw = (kzg._codec.from_device(o, i) for o, i in zip(outs, infs))
        q_buf.free()
        lap("round5_openings")
        self.last_r_zeta = f.eval(r_poly, n + 6, zi)                    # plonk/prover.py:171 asserts this is 0
        self.last_t_top = t.read_ints(3 * n + 6, min(8, n4 - 3 * n - 6))   # deg t <= 3n+5: must be zeros

        for vec in (wires, pi, t, r_poly, wire_buf, z_poly, t_buf):
            vec.free()
        return {
            "commitments": {"a": wire_commitments[0], "b": wire_commitments[1], "c": wire_commitments[2], "z": z_commit,
                            "t_lo": t_commitments[0], "t_mid": t_commitments[1], "t_hi": t_commitments[2]},
            "evaluations": {"a": evaluations[0], "b": evaluations[1], "c": evaluations[2],
                            "s_sigma1": evaluations[3], "s_sigma2": evaluations[4], "z_omega": evaluations[5]},
            "kzg_proofs": {"W_z": W_z, "W_zw": W_zw},
        }


class Verifier:
    """plonk/verifier.py:8-205 with the reference's interface (`verify(ivk, x, proof)`).  The transcript replay and
    the scalar arithmetic run on the host; the linearisation commitment and the two sides of the batched opening
    check are G1 combinations on the device (`KZG.batch_check` -> `kzgpu_g1_lincomb`); the final two pairings are
    py_ecc's, exactly as in the reference (kzg.py:283-286), so `verify` raises ImportError at that point when py_ecc
    is not installed."""

    def __init__(self, curve_type="bn254"):
        self.kzg = KZG(curve_type=curve_type)

    def verify(self, ivk, x, proof):
        kzg = self.kzg
        r, Fq = kzg.curve_order, kzg.Fq
        C = ivk["commitments"]
        sub = ivk["subgroups"]
        n, g, k1, k2 = sub["n"], int(sub["g"]), int(sub["k1"]), int(sub["k2"])
        pc, ev = proof["commitments"], proof["evaluations"]
        a, b, c = (int(ev[k]) % r for k in ("a", "b", "c"))
        s1, s2, zw = (int(ev[k]) % r for k in ("s_sigma1", "s_sigma2", "z_omega"))
        t = Transcript("plonk-proof", Fq)                                           # plonk/verifier.py:91-110
        t.append_message("public-inputs", x)
        t.append_message("round1-commitments", [pc["a"], pc["b"], pc["c"]])
        beta, gamma = int(t.get_challenge("beta")), int(t.get_challenge("gamma"))
        t.append_message("round2-commitment", pc["z"])
        alpha = int(t.get_challenge("alpha"))
        t.append_message("round3-commitments", [pc["t_lo"], pc["t_mid"], pc["t_hi"]])
        zeta = int(t.get_challenge("zeta"))
        t.append_message("round4-evaluations", [ev["a"], ev["b"], ev["c"], ev["s_sigma1"], ev["s_sigma2"], ev["z_omega"]])
        v = t.get_challenge("v")
        u = t.get_challenge("u")
        zn = pow(zeta, n, r)
        zh = (zn - 1) % r
        l1 = zh * pow(n * (zeta - 1) % r, -1, r) % r
        pi = 0                                                                      # PI(zeta), plonk/encoder.py:196-223
        for i, xi in enumerate(x):
            gi = pow(g, i, r)
            pi = (pi - int(xi) * gi % r * zh % r * pow(n * (zeta - gi) % r, -1, r)) % r
        # r(X) commitment (plonk/verifier.py:117-157): every term is scalar * point -> one device combination
        f1 = alpha * (a + beta * zeta + gamma) % r * (b + beta * k1 * zeta + gamma) % r * (c + beta * k2 * zeta + gamma) % r
        f2 = alpha * (a + beta * s1 + gamma) % r * (b + beta * s2 + gamma) % r * zw % r
        al2l1 = alpha * alpha % r * l1 % r
        r_comm = kzg.g1_lincomb(
            [C["qM"], C["qL"], C["qR"], C["qO"], C["qC"], kzg.G1, pc["z"], C["S_sigma3"], pc["t_lo"], pc["t_mid"], pc["t_hi"]],
            [a * b, a, b, c, 1, pi - f2 * (c + gamma) - al2l1, f1 + al2l1, -f2 * beta, -zh, -zh * zn, -zh * zn % r * zn])
        return kzg.batch_check(ivk["rk"],
                               [[r_comm, pc["a"], pc["b"], pc["c"], C["S_sigma1"], C["S_sigma2"]], [pc["z"]]],
                               [zeta, zeta * g % r], [[0, a, b, c, s1, s2], [zw]],
                               [proof["kzg_proofs"]["W_z"], proof["kzg_proofs"]["W_zw"]], [v, v], u)
