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

import random

import numpy as np

from . import _ffi, device
from ._ffi import check, ptr
from .kzg import KZG
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


def _entries(M):
    """(nrows, ncols, [(i, j, value)]) of a matrix given as a dense list of rows, as {"shape": (r, c), "entries":
    [(i, j, v), ...]}, or as an object with Sage's nrows() / ncols() / nonzero_positions() / [i, j]."""
    if isinstance(M, dict):
        return M["shape"][0], M["shape"][1], [(int(i), int(j), int(v)) for i, j, v in M["entries"] if int(v)]
    if hasattr(M, "nonzero_positions"):
        return M.nrows(), M.ncols(), [(i, j, int(M[i, j])) for i, j in M.nonzero_positions()]
    rows = [list(r) for r in M]
    return len(rows), (len(rows[0]) if rows else 0), [(i, j, int(v)) for i, r in enumerate(rows) for j, v in enumerate(r) if int(v)]


class Indexer:
    """marlin/indexer.py:8-112 with the polynomial work on the device: the R1CS matrices are encoded as the row / col /
    val polynomials over K of their "star" forms (marlin/indexer.py:47-54, marlin/encoder.py:98-125), all nine are
    committed in ONE batched MSM pass (marlin/indexer.py:69), and the K-domain evaluations plus row indices stay resident
    for the prover's evaluation loops (`compute_t_polynomial`, `compute_f2_polynomial`)."""

    def __init__(self, curve_type="bn254"):
        self.kzg = KZG(curve_type=curve_type)

    def preprocess(self, A, B, C, max_degree, *, tau=None, ck=None, rng=None):
        kzg = self.kzg
        cid, r = kzg._cid, kzg.curve_order
        f = _Field(cid)
        mats = {"A": _entries(A), "B": _entries(B), "C": _entries(C)}
        pow2 = lambda v: 1 << max(v - 1, 0).bit_length()                       # noqa: E731  (encoder.find_subgroup_size)
        n = pow2(max(mats["A"][0], mats["A"][1]))                               # marlin/encoder.py:37
        m = pow2(max(len(e[2]) for e in mats.values()))                         # :38-44
        assert max_degree + 1 >= m, "the SRS must cover the index polynomials (degree m - 1)"
        g_H, g_K = f.root(n), f.root(m)
        if ck is None:
            tau = (rng or random.SystemRandom()).randrange(1, r) if tau is None else int(tau) % r
            srs = device.Srs.generate(cid, tau, max_degree + 1)
        elif isinstance(ck, device.Srs):
            srs = ck
        else:
            srs = device.Srs.from_affine(cid, kzg._codec.points_to_limbs(ck))
        Hs = [1] * n
        for i in range(1, n):
            Hs[i] = Hs[i - 1] * g_H % r
        ninv = pow(n, -1, r)
        # star matrix M* = M^T with column c scaled by u_H(h_c, h_c) = n / h_c; entry (r_, c_) of M* encodes
        # row = h_r, col = h_c, val = M*[r_, c_] / (u_H(h_r, h_r) u_H(h_c, h_c)) = M[c_, r_] * h_r / n, in row-major order
        evals, row_index = {}, {}
        for name, (_, _, ent) in mats.items():
            star = sorted((j, i, v) for i, j, v in ent)                         # (row of M*, column of M*, M[c_, r_])
            row = [Hs[rr] for rr, _, _ in star] + [0] * (m - len(star))
            col = [Hs[cc] for _, cc, _ in star] + [0] * (m - len(star))
            val = [v % r * Hs[rr] % r * ninv % r for rr, _, v in star] + [0] * (m - len(star))
            evals[name] = (row, col, val)
            row_index[name] = [rr for rr, _, _ in star] + [-1] * (m - len(star))
        names = [f"{kind}_{M}" for M in "ABC" for kind in ("row", "col", "val")]  # commit order of marlin/indexer.py:63-69
        coeff = DVec(9 * m)
        flat = [v for M in "ABC" for k in range(3) for v in evals[M][k]]
        coeff.write(0, flat, r)
        check(f.lib.kzgpu_ntt_batch_dev(cid, coeff.ptr, m, 9, ptr(f.L(g_K)), 1, None))      # 9 interpolations over K
        outs, infs = device.msm_batch_dev(srs, coeff, m, 9)
        commitments = {nm: kzg._codec.from_device(o, i) for nm, o, i in zip(names, outs, infs)}
        # resident K-domain evaluations, kind-major (row_A row_B row_C | col ... | val ...) as the loop kernels take them
        kd = {kind: DVec.from_limbs(ints_to_limbs([v for M in "ABC" for v in evals[M][k]], r))
              for k, kind in enumerate(("row", "col", "val"))}
        ridx = np.ascontiguousarray(np.array([i for M in "ABC" for i in row_index[M]], dtype=np.int64).astype(np.uint32))
        sub = {"n": n, "m": m, "g_H": kzg.Fq(g_H), "g_K": kzg.Fq(g_K)}
        ipk = {"ck": srs, "A": A, "B": B, "C": C, "commitments": commitments, "subgroups": sub,
               "polynomials": {"buffer": coeff, "names": names, "length": m},
               "evals": kd, "row_index": _ffi.DeviceBuffer(ridx.nbytes).upload(ridx),
               "vanishing_polys": {"v_H": ("X^n - 1", n), "v_K": ("X^m - 1", m)}}
        rk = kzg.multiply(kzg.G2, tau) if (kzg.have_py_ecc and tau is not None) else None
        ivk = {"rk": rk, "commitments": commitments, "subgroups": {"n": n, "m": m, "g_H": kzg.Fq(g_H)},
               "vanishing_polys": ipk["vanishing_polys"], "tau": tau}
        return ipk, ivk

    @staticmethod
    def polynomial(ipk, name):
        """Coefficients (ints, low -> high, length m) of one index polynomial."""
        p = ipk["polynomials"]
        return p["buffer"].read_ints(p["names"].index(name) * p["length"], p["length"])
