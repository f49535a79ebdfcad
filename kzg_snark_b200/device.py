"""Buffer-level host API over the C ABI: numpy limb arrays in, numpy limb arrays out.

These are the calls bench.py times and the drop-in modules (kzg.py / fft_ff.py) build on.
Every function goes through libkzgpu.so; none has a CPU path.
"""

import ctypes
import numpy as np

from . import _ffi
from ._ffi import check, ptr, CURVE_IDS

# scalar-field and base-field moduli (public curve parameters; same values as py_ecc's
# curve_order / field_modulus that reference kzg.py:27-35 imports)
FR = {
    _ffi.BN254: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    _ffi.BLS12_381: 52435875175126190479447740508185965837690552500527637822603658699938581184513,
}
FP = {
    _ffi.BN254: 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    _ffi.BLS12_381: 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
}
FP_LIMBS = {_ffi.BN254: 4, _ffi.BLS12_381: 6}


def curve_id(curve):
    if isinstance(curve, str):
        if curve not in CURVE_IDS:
            raise ValueError(f"Unsupported curve type: {curve}")     # kzg.py:37
        return CURVE_IDS[curve]
    return int(curve)


def _scalars(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1:
        a = a.reshape(-1, 4)
    assert a.shape[-1] == 4
    return a


# ---------------------------------------------------------------------------- NTT
def ntt(field, data, w_limbs, inverse=False, coset_limbs=None, batch=1):
    """In-place NTT of `data` ((batch*n, 4) uint64, canonical).  Returns the same array."""
    lib = _ffi.init()
    data = _scalars(data)
    n = data.shape[0] // batch
    w = np.ascontiguousarray(w_limbs, dtype=np.uint64)
    cs = None if coset_limbs is None else np.ascontiguousarray(coset_limbs, dtype=np.uint64)
    check(lib.kzgpu_ntt_batch(curve_id(field), ptr(data), n, batch, ptr(w), 1 if inverse else 0, ptr(cs)))
    return data


def ntt_dev(field, dbuf, n, w_limbs, inverse=False, coset_limbs=None, batch=1):
    lib = _ffi.init()
    w = np.ascontiguousarray(w_limbs, dtype=np.uint64)
    cs = None if coset_limbs is None else np.ascontiguousarray(coset_limbs, dtype=np.uint64)
    check(lib.kzgpu_ntt_batch_dev(curve_id(field), dbuf.ptr, n, batch, ptr(w), 1 if inverse else 0, ptr(cs)))


# ---------------------------------------------------------------------------- SRS + MSM
class Srs:
    """Device-resident commitment key (ck of kzg.py:69-72)."""

    def __init__(self, curve, handle, n):
        self.curve = curve
        self.handle = handle
        self.n = n

    @classmethod
    def from_affine(cls, curve, affine_xy):
        """affine_xy: (n, 2*fp_limbs) uint64 canonical; (0,0) rows = infinity."""
        lib = _ffi.init()
        cid = curve_id(curve)
        a = np.ascontiguousarray(affine_xy, dtype=np.uint64)
        n = a.shape[0]
        assert a.size == n * 2 * FP_LIMBS[cid]
        h = ctypes.c_uint64(0)
        check(lib.kzgpu_srs_create(cid, ptr(a), n, ctypes.byref(h)))
        return cls(cid, h.value, n)

    @classmethod
    def generate(cls, curve, tau, n, start=0):
        """ck[i] = tau^(start+i) * G1 computed on the device (kzg.py:69-72 with the secret
        supplied); start > 0 gives one GPU's shard of a point-sharded key."""
        lib = _ffi.init()
        cid = curve_id(curve)
        t = np.frombuffer((int(tau) % FR[cid]).to_bytes(32, "little"), dtype="<u8").copy()
        h = ctypes.c_uint64(0)
        check(lib.kzgpu_srs_generate_range(cid, ptr(t), start, n, ctypes.byref(h)))
        return cls(cid, h.value, n)

    def info(self):
        """{'c': window bits of the fixed-base tables (0 = plain key), 'tables': W, 'bytes': device footprint}"""
        c, w, b = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_size_t(0)
        check(_ffi.load_library().kzgpu_srs_info(self.handle, ctypes.byref(c), ctypes.byref(w), ctypes.byref(b)))
        return {"c": c.value, "tables": w.value, "bytes": b.value}

    def read(self, first, count):
        out = np.zeros((count, 2 * FP_LIMBS[self.curve]), dtype=np.uint64)
        check(_ffi.load_library().kzgpu_srs_read(self.handle, first, count, ptr(out)))
        return out

    def destroy(self):
        """Release the device copy (idempotent; also run when the object is garbage collected)."""
        h, self.handle = self.handle, 0
        if h and _ffi._lib is not None and _ffi._inited:
            _ffi._lib.kzgpu_srs_destroy(h)

    def __del__(self):
        try:
            self.destroy()
        except Exception:                                 # interpreter shutdown: the library may already be gone
            pass


def _point_out(cid):
    return np.zeros(2 * FP_LIMBS[cid], dtype=np.uint64)


def msm(srs, scalars, first=0):
    """sum_i scalars[i] * ck[first+i] -> (affine limbs (2*L,), is_inf)."""
    lib = _ffi.init()
    s = _scalars(scalars)
    out = _point_out(srs.curve)
    inf = ctypes.c_int(0)
    rc = lib.kzgpu_msm(srs.handle, first, ptr(s), s.shape[0], ptr(out), ctypes.byref(inf))
    if rc == _ffi.E_RANGE:
        raise ValueError(_ffi.last_error())          # degree overflow, kzg.py:103-106
    check(rc)
    return out, bool(inf.value)


def msm_dev(srs, dbuf, n, first=0):
    lib = _ffi.init()
    out = _point_out(srs.curve)
    inf = ctypes.c_int(0)
    rc = lib.kzgpu_msm_dev(srs.handle, first, dbuf.ptr, n, ptr(out), ctypes.byref(inf))
    if rc == _ffi.E_RANGE:
        raise ValueError(_ffi.last_error())
    check(rc)
    return out, bool(inf.value)


def msm_batch(srs, scalar_arrays):
    """One commit() call: k polynomials -> k points (kzg.py:102)."""
    lib = _ffi.init()
    arrs = [_scalars(a) for a in scalar_arrays]
    k = len(arrs)
    L = FP_LIMBS[srs.curve]
    if k == 0:
        return np.zeros((0, 2 * L), dtype=np.uint64), []
    lens = (ctypes.c_size_t * k)(*[a.shape[0] for a in arrs])
    cat = np.ascontiguousarray(np.concatenate(arrs, axis=0)) if sum(a.shape[0] for a in arrs) else np.zeros((1, 4), np.uint64)
    out = np.zeros((k, 2 * L), dtype=np.uint64)
    inf = (ctypes.c_int * k)()
    rc = lib.kzgpu_msm_batch(srs.handle, ptr(cat), lens, k, ptr(out), inf)
    if rc == _ffi.E_RANGE:
        raise ValueError(_ffi.last_error())
    check(rc)
    return out, [bool(x) for x in inf]


def msm_batch_packed(srs, packed, lens):
    """The same call with the k polynomials already concatenated in one (sum(lens), 4) uint64 array (e.g. page-locked memory):
    no Python-side copy between the caller's buffer and kzgpu_msm_batch."""
    lib = _ffi.init()
    k = len(lens)
    a = _scalars(packed)
    assert a.shape[0] == sum(lens)
    out = np.zeros((k, 2 * FP_LIMBS[srs.curve]), dtype=np.uint64)
    inf = (ctypes.c_int * k)()
    rc = lib.kzgpu_msm_batch(srs.handle, ptr(a), (ctypes.c_size_t * k)(*lens), k, ptr(out), inf)
    if rc == _ffi.E_RANGE:
        raise ValueError(_ffi.last_error())
    check(rc)
    return out, [bool(x) for x in inf]


def msm_batch_dev(srs, dbuf, poly_len, k):
    """k polynomials of `poly_len` scalars each, back to back on the device (zero padded), in one
    sort / accumulate / reduce pass -> ((k, 2*L) affine limbs, [is_inf])."""
    lib = _ffi.init()
    out = np.zeros((k, 2 * FP_LIMBS[srs.curve]), dtype=np.uint64)
    inf = (ctypes.c_int * k)()
    rc = lib.kzgpu_msm_batch_dev(srs.handle, dbuf.ptr, poly_len, k, ptr(out), inf)
    if rc == _ffi.E_RANGE:
        raise ValueError(_ffi.last_error())
    check(rc)
    return out, [bool(x) for x in inf]


def msm_partial_dev(srs, dbuf, n, dout, first=0):
    """Un-normalised XYZZ partial sum left on the device (multi-GPU shard)."""
    check(_ffi.init().kzgpu_msm_partial_dev(srs.handle, first, dbuf.ptr, n, dout.ptr))


def msm_partial(srs, scalars, dout, first=0):
    """Same from host scalars (numpy limb array, ideally pinned): the upload overlaps the compute inside the call."""
    a = _scalars(scalars)
    check(_ffi.init().kzgpu_msm_partial(srs.handle, first, ptr(a), a.shape[0], dout.ptr))


def g1_fold(curve, dbuf, count):
    cid = curve_id(curve)
    out = _point_out(cid)
    inf = ctypes.c_int(0)
    check(_ffi.init().kzgpu_g1_fold(cid, dbuf.ptr, count, ptr(out), ctypes.byref(inf)))
    return out, bool(inf.value)


def g1_lincomb(curve, affine_xy, scalars):
    """sum_i scalars[i] * P_i over a few arbitrary points (verifier-side combinations, kzg.py:183-205):
    affine_xy (k, 2*L) canonical limbs ((0,0) rows = identity), scalars (k, 4) -> (affine limbs, is_inf)."""
    cid = curve_id(curve)
    a = np.ascontiguousarray(affine_xy, dtype=np.uint64).reshape(-1, 2 * FP_LIMBS[cid])
    s = _scalars(scalars)
    assert a.shape[0] == s.shape[0]
    out = _point_out(cid)
    inf = ctypes.c_int(0)
    check(_ffi.init().kzgpu_g1_lincomb(cid, ptr(a), ptr(s), a.shape[0], ptr(out), ctypes.byref(inf)))
    return out, bool(inf.value)


# ---------------------------------------------------------------------------- open
def open_proof(srs, poly_arrays, z_limbs, xi_limbs, want_eval=False):
    lib = _ffi.init()
    arrs = [_scalars(a) for a in poly_arrays]
    k = len(arrs)
    lens = (ctypes.c_size_t * max(k, 1))(*[a.shape[0] for a in arrs])
    total = sum(a.shape[0] for a in arrs)
    cat = np.ascontiguousarray(np.concatenate(arrs, axis=0)) if total else np.zeros((1, 4), np.uint64)
    out = _point_out(srs.curve)
    inf = ctypes.c_int(0)
    ev = np.zeros(4, dtype=np.uint64)
    z = np.ascontiguousarray(z_limbs, dtype=np.uint64)
    xi = np.ascontiguousarray(xi_limbs, dtype=np.uint64)
    rc = lib.kzgpu_open(srs.handle, ptr(cat), lens, k, ptr(z), ptr(xi), ptr(out), ctypes.byref(inf), ptr(ev))
    if rc == _ffi.E_RANGE:
        raise ValueError(_ffi.last_error())
    check(rc)
    return (out, bool(inf.value), ev) if want_eval else (out, bool(inf.value))


def open_quotient(field, poly_arrays, z_limbs, xi_limbs):
    """(quotient coefficients (m,4), P(z) limbs) -- the polynomial half of KZG.open."""
    lib = _ffi.init()
    arrs = [_scalars(a) for a in poly_arrays]
    k = len(arrs)
    lens = (ctypes.c_size_t * max(k, 1))(*[a.shape[0] for a in arrs])
    total = sum(a.shape[0] for a in arrs)
    maxlen = max([a.shape[0] for a in arrs], default=0)
    cat = np.ascontiguousarray(np.concatenate(arrs, axis=0)) if total else np.zeros((1, 4), np.uint64)
    quot = np.zeros((max(maxlen, 1), 4), dtype=np.uint64)
    qlen = ctypes.c_size_t(0)
    ev = np.zeros(4, dtype=np.uint64)
    z = np.ascontiguousarray(z_limbs, dtype=np.uint64)
    xi = np.ascontiguousarray(xi_limbs, dtype=np.uint64)
    check(lib.kzgpu_open_quotient(curve_id(field), ptr(cat), lens, k, ptr(z), ptr(xi), ptr(quot), ctypes.byref(qlen), ptr(ev)))
    return quot[:qlen.value], ev


def powers(field, base, n, scale=1):
    """[scale * base^i for i < n] as (n, 4) limbs, computed on the device (kzgpu_powers_dev)."""
    lib = _ffi.init()
    fid = curve_id(field)
    out = np.zeros((n, 4), dtype=np.uint64)
    if n == 0:
        return out
    d = _ffi.DeviceBuffer(n * 32)
    b = np.frombuffer((int(base) % FR[fid]).to_bytes(32, "little"), dtype="<u8").copy()
    s = np.frombuffer((int(scale) % FR[fid]).to_bytes(32, "little"), dtype="<u8").copy()
    check(lib.kzgpu_powers_dev(fid, d.ptr, n, ptr(b), ptr(s)))
    d.download(out)
    d.free()
    return out


# ---------------------------------------------------------------------------- diagnostics
def field_op(curve, which, op, a, b=None):
    """Elementwise Montgomery-core check (tests): which 0=Fp 1=Fr; op 0 mul 1 add 2 sub 3 inv."""
    lib = _ffi.init()
    cid = curve_id(curve)
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.zeros_like(a)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    check(lib.kzgpu_field_op(cid, which, op, ptr(a), ptr(bb), ptr(out), a.shape[0]))
    return out
