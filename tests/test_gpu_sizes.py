"""GPU: the largest sizes of BASELINE.json's sweeps (configs[1] NTT up to 2^26, configs[2] MSM up to 2^26) and the second curve
at scale, checked through size-independent properties -- the oracle cannot run there:
  NTT  out[k] == p(w^k) at sampled k, with p evaluated by the device Horner kernel on an untouched copy of the input
       (fft_ff.py:3-37: out[k] = sum_j c[j] w^(jk)); inverse(forward(x)) == x on sampled windows (fft_ff.py:39-58);
  MSM  commit(ck, p) == p(tau) * G1 (kzg.py:108) with p(tau) from the device Horner kernel and the scalar multiplication of the
       generator from the (separately tested) verifier-side combination kernel."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TAU = 0x1D2C3B4A5F6E7D8C9BA55AA55
GEN = {"bn254": 5, "bls12_381": 7}
G1 = {"bn254": (1, 2), "bls12_381": (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)}


class _At:
    def __init__(self, base, off):
        self.ptr = ctypes.c_void_p(base.ptr.value + off)


def _horner(cid, dbuf, n, x):
    from kzg_snark_b200 import _ffi, device
    from kzg_snark_b200.limbs import ints_to_limbs, limbs_to_int
    out = np.zeros(4, dtype=np.uint64)
    _ffi.check(_ffi._lib.kzgpu_poly_eval_dev(cid, dbuf.ptr, n, _ffi.ptr(ints_to_limbs([x], device.FR[cid])[0]), _ffi.ptr(out)))
    return limbs_to_int(out)


@pytest.mark.parametrize("curve,logn", [("bn254", 26), ("bls12_381", 24)])
def test_ntt_at_the_top_of_the_sweep(curve, logn):
    from kzg_snark_b200 import _ffi, device
    from kzg_snark_b200.limbs import random_scalars, ints_to_limbs, limbs_to_ints
    cid = device.curve_id(curve)
    r = device.FR[cid]
    n = 1 << logn
    x = random_scalars(n, r, seed=logn)
    w = pow(GEN[curve], (r - 1) // n, r)
    wl = ints_to_limbs([w], r)[0]
    keep = _ffi.DeviceBuffer(n * 32).upload(x)
    d = _ffi.DeviceBuffer(n * 32)
    _ffi.check(_ffi._lib.kzgpu_d2d(d.ptr, keep.ptr, n * 32))
    device.ntt_dev(cid, d, n, wl)
    row = np.zeros((1, 4), dtype=np.uint64)
    for k in (0, 1, 2, n // 2, n // 2 + 12345, n - 1):
        _At(d, 32 * k)
        _ffi.check(_ffi._lib.kzgpu_d2h(_ffi.ptr(row), _At(d, 32 * k).ptr, 32))
        assert limbs_to_ints(row)[0] == _horner(cid, keep, n, pow(w, k, r)), f"out[{k}] != p(w^{k})"
    device.ntt_dev(cid, d, n, wl, inverse=True)
    win = np.zeros((1 << 12, 4), dtype=np.uint64)
    for start in (0, n // 3, n - (1 << 12)):
        _ffi.check(_ffi._lib.kzgpu_d2h(_ffi.ptr(win), _At(d, 32 * start).ptr, win.nbytes))
        assert (win == x[start:start + (1 << 12)]).all(), f"inverse(forward(x)) != x near {start}"
    # coset variant (north_star extension): out[k] == p(s * w^k)
    _ffi.check(_ffi._lib.kzgpu_d2d(d.ptr, keep.ptr, n * 32))
    device.ntt_dev(cid, d, n, wl, coset_limbs=ints_to_limbs([7], r)[0])
    for k in (0, 5, n - 3):
        _ffi.check(_ffi._lib.kzgpu_d2h(_ffi.ptr(row), _At(d, 32 * k).ptr, 32))
        assert limbs_to_ints(row)[0] == _horner(cid, keep, n, 7 * pow(w, k, r) % r)
    d.free(); keep.free()


@pytest.mark.parametrize("curve,logn", [("bn254", 26), ("bls12_381", 22)])
def test_msm_tau_identity_at_the_top_of_the_sweep(curve, logn):
    from kzg_snark_b200 import _ffi, device
    from kzg_snark_b200.limbs import random_scalars, ints_to_limbs
    cid = device.curve_id(curve)
    r, p, L = device.FR[cid], device.FP[cid], device.FP_LIMBS[cid]
    n = 1 << logn
    tau = TAU % r
    srs = device.Srs.generate(cid, tau, n)
    sc = random_scalars(n, r, seed=900 + logn)
    d = _ffi.DeviceBuffer(n * 32).upload(sc)
    out, inf = device.msm_dev(srs, d, n)
    e = _horner(cid, d, n, tau)
    g = np.concatenate([ints_to_limbs([c], p, L)[0] for c in G1[curve]]).reshape(1, 2 * L)
    exp, einf = device.g1_lincomb(cid, g, ints_to_limbs([e], r))
    assert not inf and not einf and (out == exp).all(), "commit(ck, p) != p(tau) * G1"
    # the host-buffer entry point (chunked upload) must agree, also from pageable memory
    out2, inf2 = device.msm(srs, sc)
    assert (out2 == out).all() and not inf2
    d.free(); srs.destroy()
