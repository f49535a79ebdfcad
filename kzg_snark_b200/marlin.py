"""Device versions of the two evaluation loops of the reference's Marlin prover that SURVEY.md
section 8f (N4) names: `Prover._compute_t_polynomial` (marlin/prover.py:248-301) and
`Prover._compute_f2_polynomial` (marlin/prover.py:404-470).

Both sum eta_M * val_M(kappa) / ((x - row_M(kappa)) (alpha - col_M(kappa))) over the entries kappa in K of
the three index matrices; the reference does one rational-function division per entry in an
interpreted loop.  Here the 3m denominators are inverted in one batched pass
(`kzgpu_marlin_*_evals_dev`) and the polynomial is recovered with one inverse NTT.

Inputs are the K-domain evaluations of the index polynomials (what `fft_ff(list(row_A), g_K, Fq)`
yields at marlin/prover.py:439-449) as (3m, 4) limb arrays, matrices A, B, C back to back.
No CPU path: everything goes through libkzgpu.so.
"""

import numpy as np

from . import _ffi, device
from ._ffi import check, ptr
from .limbs import ints_to_limbs
from .plonk import DVec, _Field


def _up(arr):
    return DVec.from_limbs(np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4))


def compute_f2_polynomial(curve, row, col, val, eta, alpha, beta1, n, g_K):
    """f_2 coefficients ((m, 4) limbs, low -> high).  row / col / val: (3m, 4) limb arrays; eta: the three
    matrix challenges; n = |H| (for v_H); g_K generates K."""
    cid = device.curve_id(curve)
    f = _Field(cid)
    r = f.r
    m = np.asarray(row).reshape(-1, 4).shape[0] // 3
    scale = (pow(int(beta1), n, r) - 1) * (pow(int(alpha), n, r) - 1) % r            # v_H(beta_1) v_H(alpha), :430-431
    d_row, d_col, d_val, out = _up(row), _up(col), _up(val), DVec(m)
    check(f.lib.kzgpu_marlin_f2_evals_dev(cid, m, d_row.ptr, d_col.ptr, d_val.ptr, ptr(ints_to_limbs(eta, r)), ptr(f.L(alpha)),
                                          ptr(f.L(beta1)), ptr(f.L(scale)), out.ptr))
    f.intt(out, m, int(g_K))                                                         # fft_ff_interpolation, :469
    res = out.read()
    for v in (d_row, d_col, d_val, out):
        v.free()
    return res


def compute_t_polynomial(curve, row_index, col, val, eta, alpha, n, g_H):
    """t coefficients ((n, 4) limbs).  row_index: (3m,) indices i with row_M(kappa) = g_H^i, ascending per matrix,
    -1 for the padding entries of the index (marlin/encoder.py:105-107 leaves them 0)."""
    cid = device.curve_id(curve)
    f = _Field(cid)
    r = f.r
    ri = np.ascontiguousarray(np.asarray(row_index, dtype=np.int64).astype(np.uint32))
    m = ri.shape[0] // 3
    scale = n * (pow(int(alpha), n, r) - 1) % r                                      # n * v_H(alpha)
    d_ri = _ffi.DeviceBuffer(ri.nbytes).upload(ri)
    d_col, d_val, H, out = _up(col), _up(val), DVec(n), DVec(n)
    f.powers(H, n, int(g_H))
    check(f.lib.kzgpu_marlin_t_evals_dev(cid, n, m, d_ri.ptr, d_col.ptr, d_val.ptr, H.ptr, ptr(ints_to_limbs(eta, r)),
                                         ptr(f.L(alpha)), ptr(f.L(scale)), out.ptr))
    f.intt(out, n, int(g_H))
    res = out.read()
    d_ri.free()
    for v in (d_col, d_val, H, out):
        v.free()
    return res
