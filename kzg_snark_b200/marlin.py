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

import ctypes
import random
import time

import numpy as np

from . import _ffi, device
from ._ffi import check, ptr
from .kzg import KZG
from .limbs import ints_to_limbs, limbs_to_ints
from .plonk import DVec, Transcript, _Field, _View, _voidp_array


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
        # third-round precomputation: the nine polynomials on the coset 5 * <w_8m> (kind-major), where h_2 is formed
        m8, shift = 8 * m, {_ffi.BN254: 5, _ffi.BLS12_381: 7}[cid]
        w8 = f.root(m8)
        cos = DVec(9 * m8, zero=True)
        for k in range(3):
            for j in range(3):                                                  # names[] is matrix-major, the coset kind-major
                cos.copy_from(coeff, m, (k * 3 + j) * m8, (j * 3 + k) * m)
        check(f.lib.kzgpu_ntt_batch_dev(cid, cos.ptr, m8, 9, ptr(f.L(w8)), 0, ptr(f.L(shift))))
        sm = pow(shift, m, r)
        z8 = pow(w8, m, r)                                                      # primitive 8th root of unity
        vk_inv = [pow((sm * pow(z8, i, r) - 1) % r, -1, r) for i in range(8)]
        ridx = np.ascontiguousarray(np.array([i for M in "ABC" for i in row_index[M]], dtype=np.int64).astype(np.uint32))
        # CSR copies of A, B, C for the prover's z_M = M z
        csr = {}
        for name, (nr, _, ent) in mats.items():
            ent = sorted(ent)
            rp = np.zeros(n + 1, dtype=np.uint32)
            if ent:
                np.add.at(rp, np.array([e[0] for e in ent], dtype=np.int64) + 1, 1)
            rp = np.ascontiguousarray(np.cumsum(rp, dtype=np.uint64).astype(np.uint32))
            ci = np.ascontiguousarray(np.array([e[1] for e in ent], dtype=np.uint32))
            csr[name] = (_ffi.DeviceBuffer(rp.nbytes).upload(rp), _ffi.DeviceBuffer(max(ci.nbytes, 4)).upload(ci),
                         DVec.from_limbs(ints_to_limbs([e[2] for e in ent], r)) if ent else DVec(1))
        sub = {"n": n, "m": m, "g_H": kzg.Fq(g_H), "g_K": kzg.Fq(g_K)}
        ipk = {"ck": srs, "A": A, "B": B, "C": C, "commitments": commitments, "subgroups": sub,
               "polynomials": {"buffer": coeff, "names": names, "length": m},
               "evals": kd, "row_index": _ffi.DeviceBuffer(ridx.nbytes).upload(ridx), "matrices": mats, "csr": csr,
               "coset8": {"m8": m8, "w8": w8, "shift": shift, "evals": cos, "vk_inv": vk_inv},
               "vanishing_polys": {"v_H": ("X^n - 1", n), "v_K": ("X^m - 1", m)}}
        rk = kzg.multiply(kzg.G2, tau) if (kzg.have_py_ecc and tau is not None) else None
        ivk = {"rk": rk, "commitments": commitments, "subgroups": {"n": n, "m": m, "g_H": kzg.Fq(g_H)},
               "vanishing_polys": ipk["vanishing_polys"]}      # rk only, never the trapdoor (marlin/indexer.py:109-110)
        del tau
        return ipk, ivk

    @staticmethod
    def polynomial(ipk, name):
        """Coefficients (ints, low -> high, length m) of one index polynomial."""
        p = ipk["polynomials"]
        return p["buffer"].read_ints(p["names"].index(name) * p["length"], p["length"])


class Prover:
    """marlin/prover.py:8-246 on the device, with the reference's interface: `prove(ipk, x, w)` -> the same proof dictionary.
    The Sage polynomial arithmetic between the reference's kzg / fft_ff calls becomes: iNTTs for the encodings
    (marlin/encoder.py:133-229), exact divisions by linear factors and by X^n - 1 as device recurrences / foldings,
    polynomial products through 4n-point NTTs (:96, :131), the two evaluation loops as batched-inversion kernels
    (:248-301, :404-470), h_2 formed point-wise on an 8m coset (:166-171), every linear combination as one kernel, and
    each round's commitments as one batched MSM pass.  Same transcript, same draws in the same order."""

    def __init__(self, curve_type="bn254"):
        self.kzg = KZG(curve_type=curve_type)
        self.capture = False

    def prove(self, ipk, x, w, zero_knowledge_bound=2, draws=None):
        kzg = self.kzg
        cid, r, Fq = kzg._cid, kzg.curve_order, kzg.Fq
        f = _Field(cid)
        lib = f.lib
        srs = ipk["ck"]
        sub = ipk["subgroups"]
        n, m, g_H, g_K = sub["n"], sub["m"], int(sub["g_H"]), int(sub["g_K"])
        b = zero_knowledge_bound
        ell = len(x)
        if draws is None:
            sr = random.SystemRandom()
            draws = [sr.randrange(r) for _ in range(4 * b + 2 * n + b - 1)]
        # `w` and `draws` may be (k, 4) limb arrays (large instances: no per-element Python conversion)
        d_limbs = np.ascontiguousarray(draws, dtype=np.uint64).reshape(-1, 4) if isinstance(draws, np.ndarray) else ints_to_limbs(draws, r)
        head = limbs_to_ints(d_limbs[:4 * b])                                            # the 4b masking scalars
        w_limbs = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4) if isinstance(w, np.ndarray) else ints_to_limbs(w, r)
        xs = [int(v) % r for v in x]
        nz = ell + w_limbs.shape[0]
        Hs = [pow(g_H, i, r) for i in range(ell)]                                        # only the public part of H is needed here

        def new(length, zero=True):
            return DVec(length, zero=zero)

        def lin(out, out_len, terms, constant=None):
            f.lincomb(out, out_len, [(p.ptr if hasattr(p, "ptr") else p, ln, int(sc) % r) for p, ln, sc in terms], constant)

        def div_linear(dst, src, length, root):
            """dst <- src // (X - root) (src(root) == 0): length - 1 coefficients."""
            ql = ctypes.c_size_t(0)
            lens = (ctypes.c_size_t * 1)(length)
            check(lib.kzgpu_open_quotient_dev(cid, _voidp_array([src.ptr]), lens, 1, ptr(f.L(root)), ptr(f.L(1)), dst.ptr,
                                              ctypes.byref(ql), None))
            return ql.value

        def mul(a, la, bb, lb):
            """a * b through a power-of-two NTT; returns (DVec, la + lb - 1)."""
            N = 1 << (la + lb - 2).bit_length()
            wN = f.root(N)
            ea, eb = new(N), new(N)
            ea.copy_from(a, la)
            eb.copy_from(bb, lb)
            check(lib.kzgpu_ntt_dev(cid, ea.ptr, N, ptr(f.L(wN)), 0, None))
            check(lib.kzgpu_ntt_dev(cid, eb.ptr, N, ptr(f.L(wN)), 0, None))
            check(lib.kzgpu_poly_mul_pointwise_dev(cid, ea.ptr, ea.ptr, eb.ptr, N))
            check(lib.kzgpu_ntt_dev(cid, ea.ptr, N, ptr(f.L(wN)), 1, None))
            eb.free()
            return ea, la + lb - 1

        def commit(polys):
            """One batched MSM pass for the polynomials [(vec, length)] of a round."""
            L = max(ln for _, ln in polys)
            buf = new(len(polys) * L)
            for j, (v, ln) in enumerate(polys):
                buf.copy_from(v, ln, j * L)
            outs, infs = device.msm_batch_dev(srs, buf, L, len(polys))
            buf.free()
            return [kzg._codec.from_device(o, i) for o, i in zip(outs, infs)]

        self.captured = cap = {}
        self.timings = tm = {}
        clock = [time.perf_counter()]

        def lap(name):
            check(lib.kzgpu_sync())
            now = time.perf_counter()
            tm[name] = tm.get(name, 0.0) + now - clock[0]
            clock[0] = now

        def keep(name, vec, length):                                                     # tests compare every round's polynomials
            if self.capture:
                cap[name] = vec.read_ints(0, length)

        transcript = Transcript("marlin-proof", Fq)
        transcript.append_message("public-inputs", x)                                    # marlin/prover.py:56

        lap("setup")
        # ---- witness and linear-combination encodings (marlin/encoder.py:133-229)
        # x_poly through (h_i, x_i), i < ell, and v_H_x = prod (X - h_i): ell is the public-input size (small), host
        vhx = [1]
        for i in range(ell):
            vhx = [(-Hs[i] * vhx[0]) % r] + [(vhx[k - 1] - Hs[i] * vhx[k]) % r for k in range(1, len(vhx))] + [vhx[-1]]
        xpoly = [0] * ell
        for i in range(ell):
            quo = [0] * ell                                                               # v_H_x // (X - h_i)
            quo[ell - 1] = vhx[ell]
            for k in range(ell - 1, 0, -1):
                quo[k - 1] = (vhx[k] + Hs[i] * quo[k]) % r
            den = 0
            for c in reversed(quo):
                den = (den * Hs[i] + c) % r
            sc = xs[i] * pow(den, -1, r) % r
            xpoly = [(a + sc * q_) % r for a, q_ in zip(xpoly, quo)]
        while xpoly and xpoly[-1] == 0:
            xpoly.pop()
        x_poly = new(n)
        if xpoly:
            x_poly.write(0, xpoly, r)
        xe = new(n)
        xe.copy_from(x_poly, n)
        check(lib.kzgpu_ntt_dev(cid, xe.ptr, n, ptr(f.L(g_H)), 0, None))                 # x_poly on H
        zv = new(n)
        if ell:
            zv.write(0, xs, r)
        check(lib.kzgpu_h2d(zv.at(ell), ptr(w_limbs), w_limbs.nbytes))
        vals = new(n)
        lin(vals, n, [(zv, n, 1), (xe, n, -1)])                                          # w_i - x_poly(h_i) ...
        check(lib.kzgpu_memset(vals.ptr, 0, ell * 32))                                   # ... zero on the public part
        if nz < n:
            check(lib.kzgpu_memset(vals.at(nz), 0, (n - nz) * 32))                       # and on the padding (:150-154)
        f.intt(vals, n, g_H)
        w_poly, tmp, wl = vals, new(n), n
        for i in range(ell):                                                             # w_poly = f // v_H_x (:156)
            wl = div_linear(tmp, w_poly, wl, Hs[i])
            w_poly, tmp = tmp, w_poly
        # v_H_w = (X^n - 1) // v_H_x, kept behind b zero slots so that X^i * v_H_w is a shifted view
        vhw_a, vhw_b = new(n + 1 + b), new(n + 1 + b)
        vhw_a.write(b, [r - 1], r)
        vhw_a.write(b + n, [1], r)
        cur, oth, vl = _View(vhw_a, b), _View(vhw_b, b), n + 1
        for i in range(ell):
            vl = div_linear(oth, cur, vl, Hs[i])
            cur, oth = oth, cur
        w_rand, zr = head[0:b], [head[b * (j + 1): b * (j + 2)] for j in range(3)]
        w_masked, wml = new(n + b), vl + b - 1                                           # w_poly + w_random * v_H_w (:89)
        lin(w_masked, wml, [(w_poly, wl, 1)] + [(_View(cur.base, cur.off - i), vl + i, w_rand[i]) for i in range(b)])
        lap("witness_encoding")
        # z_M = M z (CSR product on the device) -> interpolation -> masking with z_M_random * (X^n - 1) (:90-92)
        zm = []
        for j, M in enumerate("ABC"):
            rp, ci, vv_ = ipk["csr"][M]
            p = new(n + b)
            check(lib.kzgpu_spmv_dev(cid, n, rp.ptr, ci.ptr, vv_.ptr, zv.ptr, p.ptr))
            f.intt(p, n, g_H)
            lo = p.read_ints(0, b)
            p.write(0, [(lo[k] - zr[j][k]) % r for k in range(b)], r)
            p.write(n, zr[j], r)
            zm.append(p)
        zA, zB, zC = zm
        lap("matvec_encoding")
        # z_masked = w_masked * v_H_x + x_poly (:93): v_H_x has ell + 1 coefficients -> shifted views of w_masked
        wpad = new(ell + wml)
        wpad.copy_from(w_masked, wml, ell)
        z_masked, zml = new(n + b), wml + ell
        lin(z_masked, zml, [(_View(wpad, ell - k), wml + k, vhx[k]) for k in range(ell + 1)] + [(x_poly, n, 1)])
        # h_0 = (z_A z_B - z_C) // v_H (:94): product through a 4n-point NTT, exact division by X^n - 1 by folding
        prod, pl = mul(zA, n + b, zB, n + b)
        lin(prod, pl, [(prod, pl, 1), (zC, n + b, -1)])
        h_0, h0l = new(pl - n), pl - n
        lin(h_0, h0l, [(prod.at(n), pl - n, 1), (prod.at(2 * n), max(pl - 2 * n, 0), 1)])
        prod.free()
        lap("h_0")
        # s: the draws as coefficients, constant term adjusted so that the sum over H vanishes (:100-102)
        sl = 2 * n + b - 1
        assert d_limbs.shape[0] >= 4 * b + sl, "not enough random draws"
        s_poly = new(sl, zero=False)
        check(lib.kzgpu_h2d(s_poly.ptr, ptr(np.ascontiguousarray(d_limbs[4 * b: 4 * b + sl])), sl * 32))
        fold = limbs_to_ints(d_limbs[[4 * b + k for k in range(0, sl, n)]])              # c_0, c_n, c_2n
        s_poly.write(0, [(fold[0] - sum(fold)) % r], r)
        for nm, v_, ln in (("w_masked", w_masked, wml), ("zA", zA, n + b), ("zB", zB, n + b), ("zC", zC, n + b), ("h_0", h_0, h0l), ("s", s_poly, sl)):
            keep(nm, v_, ln)
        lap("s_upload")
        first = commit([(w_masked, wml), (zA, n + b), (zB, n + b), (zC, n + b), (h_0, h0l), (s_poly, sl)])
        lap("round1_msm")
        transcript.append_message("round1-commitments", first)
        eta = [int(transcript.get_challenge(k)) for k in ("eta_A", "eta_B", "eta_C")]
        alpha = int(transcript.get_challenge("alpha"))
        while pow(alpha, n, r) == 1:                                                      # alpha in H (:119-120)
            alpha = int(transcript.get_challenge("alpha-retry"))

        # ---- first sumcheck (:123-138)
        vHa = (pow(alpha, n, r) - 1) % r
        t_poly = new(n, zero=False)
        check(lib.kzgpu_marlin_t_evals_dev(cid, n, m, ipk["row_index"].ptr, ipk["evals"]["col"].ptr, ipk["evals"]["val"].ptr,
                                           self._H(f, n, g_H).ptr, ptr(ints_to_limbs(eta, r)), ptr(f.L(alpha)),
                                           ptr(f.L(n * vHa % r)), t_poly.ptr))
        f.intt(t_poly, n, g_H)
        r_alpha = new(n, zero=False)                                                      # u_H(alpha, X) = sum alpha^(n-1-i) X^i
        f.powers(r_alpha, n, pow(alpha, -1, r), pow(alpha, n - 1, r))
        S = new(n + b)
        lin(S, n + b, [(zA, n + b, eta[0]), (zB, n + b, eta[1]), (zC, n + b, eta[2])])
        p1, p1l = mul(r_alpha, n, S, n + b)
        p2, p2l = mul(t_poly, n, z_masked, zml)
        pol, pll = new(3 * n), max(sl, p1l, p2l)
        lin(pol, pll, [(s_poly, sl, 1), (p1, p1l, 1), (p2, p2l, -1)])
        h_1, h1l = new(2 * n), pll - n                                                    # poly // (X^n - 1)
        lin(h_1, h1l, [(pol.at(n), 2 * n, 1), (pol.at(2 * n), n, 1)])
        rem = new(n)
        lin(rem, n, [(pol, n, 1), (pol.at(n), n, 1), (pol.at(2 * n), n, 1)])             # poly % (X^n - 1); g_1 = rem // X
        assert rem.read_ints(0, 1) == [0], "Sum over H is not 0"                          # :134
        g_1 = _View(rem, 1)
        for nm, v_, ln in (("t", t_poly, n), ("g_1", g_1, n - 1), ("h_1", h_1, h1l)):
            keep(nm, v_, ln)
        lap("sumcheck1")
        second = commit([(t_poly, n), (g_1, n - 1), (h_1, h1l)])
        lap("round2_msm")
        transcript.append_message("round2-commitments", second)
        beta1 = int(transcript.get_challenge("beta_1"))
        while pow(beta1, n, r) == 1:
            beta1 = int(transcript.get_challenge("beta_1-retry"))

        # ---- second sumcheck (:149-171)
        vHb = (pow(beta1, n, r) - 1) % r
        vv = vHb * vHa % r
        t_b1 = f.eval(t_poly, n, beta1)
        f_2 = new(m, zero=False)
        check(lib.kzgpu_marlin_f2_evals_dev(cid, m, ipk["evals"]["row"].ptr, ipk["evals"]["col"].ptr, ipk["evals"]["val"].ptr,
                                            ptr(ints_to_limbs(eta, r)), ptr(f.L(alpha)), ptr(f.L(beta1)), ptr(f.L(vv)), f_2.ptr))
        f.intt(f_2, m, g_K)
        g_2 = _View(f_2, 1)                                                               # f_2 // X (:163)
        cs = ipk["coset8"]
        m8 = cs["m8"]
        f2c = new(m8)
        f2c.copy_from(f_2, m)
        f.coset_ntt(f2c, m8, cs["w8"], cs["shift"])
        params = ints_to_limbs(eta + [alpha, beta1, vv] + cs["vk_inv"], r)
        ev9 = cs["evals"]
        check(lib.kzgpu_marlin_h2_evals_dev(cid, m8, ev9.at(0), ev9.at(3 * m8), ev9.at(6 * m8), f2c.ptr, ptr(params), f2c.ptr))
        f.coset_ntt(f2c, m8, cs["w8"], cs["shift"], inverse=True)
        h_2, h2l = f2c, 6 * m - 6                                                         # deg h_2 = 6(m-1) + (m-1) - m
        keep("g_2", g_2, m - 1)
        keep("h_2", h_2, h2l)
        lap("sumcheck2_h2")
        third = commit([(g_2, m - 1), (h_2, h2l)])
        lap("round3_msm")
        transcript.append_message("round3-commitments", third)
        beta2 = int(transcript.get_challenge("beta_2"))

        # ---- linearisations and openings (:178-227)
        names = ipk["polynomials"]["names"]
        coeff = ipk["polynomials"]["buffer"]
        ipoly = {nm: _View(coeff, names.index(nm) * m) for nm in names}
        zA_b1 = f.eval(zA, n + b, beta1)
        vhx_b1 = 0
        for c in reversed(vhx):
            vhx_b1 = (vhx_b1 * beta1 + c) % r
        xp_b1 = 0
        for c in reversed(xpoly):
            xp_b1 = (xp_b1 * beta1 + c) % r
        f_1 = new(h0l)
        lin(f_1, max(h0l, n + b), [(zB, n + b, zA_b1), (zC, n + b, -1), (h_0, h0l, -vHb)])
        r_ab = ((vHa - vHb) * pow((alpha - beta1) % r, -1, r) if alpha != beta1 else n * pow(alpha, n - 1, r)) % r   # u_H(alpha, beta_1)
        f2l = max(sl, h1l)
        f_2p = new(f2l)
        lin(f_2p, f2l, [(s_poly, sl, 1), (zB, n + b, r_ab * eta[1]), (zC, n + b, r_ab * eta[2]), (w_masked, wml, -t_b1 * vhx_b1),
                        (h_1, h1l, -vHb), (g_1, n - 1, -beta1)], constant=(r_ab * eta[0] % r * zA_b1 - t_b1 * xp_b1) % r)
        rc = {nm: f.eval(ipoly[nm], m, beta2) for nm in names if not nm.startswith("val")}
        fac = [(beta1 - rc[f"row_{M}"]) * (alpha - rc[f"col_{M}"]) % r for M in "ABC"]
        b_lin = fac[0] * fac[1] % r * fac[2] % r
        op = [fac[1] * fac[2] % r, fac[0] * fac[2] % r, fac[0] * fac[1] % r]
        vKb2 = (pow(beta2, m, r) - 1) % r
        f_3 = new(h2l)
        lin(f_3, h2l, [(h_2, h2l, vKb2)] + [(ipoly[f"val_{M}"], m, -eta[j] * vv % r * op[j]) for j, M in enumerate("ABC")] +
            [(g_2, m - 1, b_lin * beta2)], constant=b_lin * t_b1 % r * pow(m, -1, r) % r)
        keep("f_1", f_1, max(h0l, n + b))
        keep("f_2", f_2p, f2l)
        keep("f_3", f_3, h2l)
        evals_b1 = [Fq(zA_b1), Fq(t_b1)]
        evals_b2 = [Fq(rc[f"{kind}_{M}"]) for M in "ABC" for kind in ("row", "col")]
        transcript.append_message("evaluations-beta1", evals_b1)
        transcript.append_message("evaluations-beta2", evals_b2)
        xi1, xi2 = int(transcript.get_challenge("xi_1")), int(transcript.get_challenge("xi_2"))

        def open_dev(polys, point, xi):
            k = len(polys)
            out = np.zeros(2 * device.FP_LIMBS[cid], dtype=np.uint64)
            inf = ctypes.c_int(0)
            lens = (ctypes.c_size_t * k)(*[ln for _, ln in polys])
            rc_ = lib.kzgpu_open_dev(srs.handle, _voidp_array([p.ptr for p, _ in polys]), lens, k, ptr(f.L(point)), ptr(f.L(xi)),
                                     ptr(out), ctypes.byref(inf), None)
            if rc_ == _ffi.E_RANGE:
                raise ValueError(_ffi.last_error())
            check(rc_)
            return kzg._codec.from_device(out, bool(inf.value))

        lap("linearisations")
        proof_b1 = open_dev([(f_1, max(h0l, n + b)), (f_2p, f2l), (zA, n + b), (t_poly, n)], beta1, xi1)
        proof_b2 = open_dev([(f_3, h2l)] + [(ipoly[f"{kind}_{M}"], m) for M in "ABC" for kind in ("row", "col")], beta2, xi2)
        lap("openings")
        self.checks = {"f_1(beta_1)": f.eval(f_1, max(h0l, n + b), beta1), "f_2(beta_1)": f.eval(f_2p, f2l, beta1),
                       "f_3(beta_2)": f.eval(f_3, h2l, beta2)}                            # the reference asserts all three are 0
        return {"commitments": {"first_round": first, "second_round": second, "third_round": third},
                "evaluations": {"beta1": evals_b1, "beta2": evals_b2},
                "kzg_proofs": {"beta1": proof_b1, "beta2": proof_b2}}

    _h_cache = {}

    @classmethod
    def _H(cls, f, n, g_H):
        key = (f.cid, n)
        if key not in cls._h_cache:
            H = DVec(n, zero=False)
            f.powers(H, n, g_H)
            cls._h_cache[key] = H
        return cls._h_cache[key]


def synthetic_r1cs(n_rows, n_pub, r, nnz_per_row=2, seed=0):
    """A satisfied R1CS instance of arbitrary size in the reference's format (constraint-system/R1CS_INSTANCE.pkl: square
    matrices A, B, C and an assignment z with (A z) o (B z) = C z, z[0] = 1, the first n_pub entries public): sparse
    descriptions {"shape", "entries"} plus z.  Row i of A and B holds `nnz_per_row` random entries; row i of C holds the one
    entry that makes the constraint true."""
    rng = random.Random(seed)
    z = [1] + [rng.randrange(1, r) for _ in range(n_rows - 1)]
    ents = {"A": [], "B": [], "C": []}
    for i in range(n_rows):
        acc = {}
        for M in "AB":
            cols = rng.sample(range(n_rows), nnz_per_row)
            vals = [rng.randrange(1, 1 << 16) for _ in cols]
            ents[M] += [(i, c, v) for c, v in zip(cols, vals)]
            acc[M] = sum(v * z[c] for c, v in zip(cols, vals)) % r
        j = rng.randrange(n_rows)
        ents["C"].append((i, j, acc["A"] * acc["B"] % r * pow(z[j], -1, r) % r))
    mats = [{"shape": (n_rows, n_rows), "entries": ents[M]} for M in "ABC"]
    return mats[0], mats[1], mats[2], z[:n_pub], z[n_pub:]
