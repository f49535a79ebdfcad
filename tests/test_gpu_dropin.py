"""GPU: the drop-in modules (same names/signatures as the reference's kzg.py and fft_ff.py)
against the oracle and the committed golden vectors.  Reads like the reference's own self-test
(kzg.py:291-380): commit, open, check, tamper."""
import json
import os
import random

import numpy as np
import pytest

from oracle.curve import get_curve
from oracle.kzg import KZGOracle, poly_eval
from oracle.params import CURVES, root_of_unity
from oracle import fft_ff as off

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CURVE_NAMES = ["bn254", "bls12_381"]


def H(v):
    return int(v, 16)


def aff(cv, pt):
    """py_ecc-shaped triple from the drop-in -> affine ints or None."""
    x, y, z = (int(c) for c in pt)
    if z == 0:
        return None
    assert z == 1
    return (x, y)


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_golden_vectors_through_the_dropin(curve):
    from kzg_snark_b200.kzg import KZG
    from kzg_snark_b200.fft_ff import fft_ff, ifft_ff, coset_fft_ff, coset_ifft_ff
    g = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))[curve]
    cv = get_curve(curve)
    kzg = KZG(curve)
    # a plain Python list of affine-as-projective points, like a ck built elsewhere
    ck = [(kzg._codec.fq(H(p[0])), kzg._codec.fq(H(p[1])), kzg._codec.fq(1)) for p in g["ck_affine"]]
    polys = [[H(v) for v in p] for p in g["polys"]]
    comm = kzg.commit(ck, polys)
    exp = [None if v is None else tuple(H(t) for t in v) for v in g["commitments"]]
    assert [aff(cv, c) for c in comm] == exp
    o = g["open"]
    proof = kzg.open(ck, polys[:o["k"]], H(o["z"]), H(o["xi"]))
    assert aff(cv, proof) == tuple(H(v) for v in o["proof"])
    t = g["ntt"]
    F = kzg.Fq
    x = [F(H(v)) for v in t["x"]]
    w = F(H(t["w"]))
    assert [int(v) for v in fft_ff(x, w, F)] == [H(v) for v in t["fft"]]
    assert [int(v) for v in ifft_ff(x, w, F)] == [H(v) for v in t["ifft"]]
    assert [int(v) for v in coset_fft_ff(x, w, F(7), F)] == [H(v) for v in t["coset_fft"]]
    assert [int(v) for v in coset_ifft_ff(x, w, F(7), F)] == [H(v) for v in t["coset_ifft"]]


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_kzg_self_test_like_the_reference(curve):
    """kzg.py:291-380 restated: three lists of two low-degree polynomials, commit, open,
    accept, then tamper one evaluation and reject (pairing-free check with the known tau)."""
    from kzg_snark_b200.kzg import KZG
    cv = get_curve(curve); ko = KZGOracle(curve)
    kzg = KZG(curve_type=curve)
    rng = random.Random(42)
    tau = rng.randrange(1, cv.r)
    ck, rk = kzg.setup(5, tau=tau)
    assert len(ck) == 6 and aff(cv, ck[0]) == cv.normalize(cv.G1)
    assert [aff(cv, p) for p in ck] == [cv.normalize(p) for p in ko.setup(5, tau)]
    X = kzg.X
    poly_list = [
        [1 + 2 * X + 3 * X ** 2, 4 + 5 * X ** 3],
        [7 - 2 * X ** 2 + X ** 3, 3 + 4 * X + 2 * X ** 2],
        [2 * X + 5 * X ** 2, 1 + X + X ** 2 + X ** 3],
    ]
    for polys in poly_list:
        comms = kzg.commit(ck, polys)
        z, xi = kzg.Fq.random_element(), kzg.Fq.random_element()
        evals = [int(p(z)) for p in polys]
        proof = kzg.open(ck, polys, z, xi)
        o_comms = [(int(c[0]), int(c[1]), int(c[2])) for c in comms]
        o_proof = (int(proof[0]), int(proof[1]), int(proof[2]))
        assert ko.check_with_tau(tau, o_comms, int(z), evals, o_proof, int(xi))
        evals[0] = (evals[0] + 1) % cv.r
        assert not ko.check_with_tau(tau, o_comms, int(z), evals, o_proof, int(xi))
        # element-wise equality with the oracle's own commit/open on the same inputs
        oc = ko.commit([(int(p[0]), int(p[1]), int(p[2])) for p in ck], [[int(c) for c in p.list()] for p in polys])
        assert [aff(cv, c) for c in comms] == [cv.normalize(c) for c in oc]
    # commit() accepts coefficient lists (kzg.py:94-95) and the zero polynomial gives Z1
    c1 = kzg.commit(ck, [[1, 2, 3]])[0]
    c2 = kzg.commit(ck, [1 + 2 * X + 3 * X ** 2])[0]
    assert aff(cv, c1) == aff(cv, c2)
    assert int(kzg.commit(ck, [kzg.R(0)])[0][2]) == 0
    with pytest.raises(ValueError, match="exceeds maximum allowed degree 5"):
        kzg.commit(ck, [X ** 6])


def test_fft_interpolation_dropin_plonk_sizes():
    """The 24 call sites use n = 16 / 32 on BN254: interpolate and re-evaluate."""
    from kzg_snark_b200.kzg import KZG
    from kzg_snark_b200.fft_ff import fft_ff, fft_ff_interpolation
    kzg = KZG("bn254"); F = kzg.Fq
    rng = random.Random(1)
    for n in (16, 32):
        g = F(1).nth_root(n)
        vals = [F(rng.randrange(kzg.curve_order)) for _ in range(n)]
        poly = fft_ff_interpolation(vals, g, F)
        assert poly.degree() <= n - 1
        assert [int(poly(g ** i)) for i in range(n)] == [int(v) for v in vals]
        coeffs = list(poly) + [F(0)] * (n - len(list(poly)))
        assert [int(v) for v in fft_ff(coeffs, g, F)] == [int(v) for v in vals]
    # witness-like input from the PLONK fixture (tiny values, zeros)
    p = json.load(open(os.path.join(GOLD, "plonk_instance.json")))
    w = [F(H(x)) for x in p["w"][:16]]
    g = F(1).nth_root(16)
    poly = fft_ff_interpolation(w, g, F)
    exp = off.ifft_ff_int([int(x) for x in w], int(g), kzg.curve_order)
    while exp and exp[-1] == 0:
        exp.pop()
    assert [int(c) for c in poly.list()] == exp


def test_commit_key_cache_and_plain_list_ck():
    from kzg_snark_b200.kzg import KZG, _SRS_CACHE
    kzg = KZG("bn254"); cv = get_curve("bn254")
    ck, _ = kzg.setup(8, tau=12345)
    plain = list(ck)                              # callers carry ck as a plain list in ipk (plonk/indexer.py:92-93)
    a = kzg.commit(plain, [[1, 2, 3, 4]])[0]
    b = KZG("bn254").commit(plain, [[1, 2, 3, 4]])[0]      # a second KZG instance, same list -> cached SRS
    assert aff(cv, a) == aff(cv, b) == aff(cv, kzg.commit(ck, [[1, 2, 3, 4]])[0])
    assert (id(plain), 0) in _SRS_CACHE
    exp = cv.normalize(cv.multiply(cv.G1, poly_eval([1, 2, 3, 4], 12345, cv.r)))
    assert aff(cv, a) == exp
    # the reference reads ck[i] afresh on every call (kzg.py:115): a key edited in place must not reuse the stale device copy
    handle = _SRS_CACHE[(id(plain), 0)][1].handle
    plain[1] = plain[2]                                        # ck = [G, t^2 G, t^2 G, t^3 G, ...]
    c = kzg.commit(plain, [[1, 2, 3, 4]])[0]
    t = 12345
    exp2 = cv.normalize(cv.multiply(cv.G1, (1 + 2 * t ** 2 + 3 * t ** 2 + 4 * t ** 3) % cv.r))
    assert aff(cv, c) == exp2 != exp
    assert _SRS_CACHE[(id(plain), 0)][1].handle != handle, "stale device key reused after an in-place edit"
    ck[1] = ck[3]                                              # the same for a CommitmentKey returned by setup()
    d = kzg.commit(ck, [[1, 2, 3, 4]])[0]
    assert aff(cv, d) == cv.normalize(cv.multiply(cv.G1, (1 + 2 * t ** 3 + 3 * t ** 2 + 4 * t ** 3) % cv.r))
    # cache eviction releases device keys nobody else holds, and never one that a live CommitmentKey still uses
    keys = [list(kzg.setup(4, tau=100 + i)[0]) for i in range(10)]
    for k in keys:
        kzg.commit(k, [[1, 1]])
    assert len(_SRS_CACHE) <= 8
    assert aff(cv, kzg.commit(ck, [[5]])[0]) == cv.normalize(cv.multiply(cv.G1, 5))


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_fft_ff_ragged_lengths_like_the_reference(curve):
    """fft_ff / ifft_ff never validate the length (fft_ff.py:14-37, 39-58; only fft_ff_interpolation asserts, :74): for
    lengths that are not powers of two the recursion drops entries and leaves result[n-1] = 0.  The drop-in reproduces
    exactly that (marlin/prover.py:439 can hand over such a list), checked against the restated recursion."""
    from kzg_snark_b200.kzg import KZG
    from kzg_snark_b200.fft_ff import fft_ff, ifft_ff, fft_ff_interpolation
    from oracle.field import GFp
    kzg = KZG(curve)
    F = kzg.Fq
    q = kzg.curve_order
    Fo = GFp(q)
    rng = random.Random(5)
    for n in (2, 3, 5, 6, 7, 12, 24, 31, 33, 100):
        vals = [rng.randrange(q) for _ in range(n)]
        for w in (root_of_unity(CURVES[curve], 1 << max(n - 1, 1).bit_length()), rng.randrange(1, q)):
            exp = [int(v) for v in off.fft_ff([Fo(v) for v in vals], Fo(w), Fo)]
            assert [int(v) for v in fft_ff([F(v) for v in vals], F(w), F)] == exp, (n, "fft")
            expi = [int(v) for v in off.ifft_ff([Fo(v) for v in vals], Fo(w), Fo)]
            assert [int(v) for v in ifft_ff([F(v) for v in vals], F(w), F)] == expi, (n, "ifft")
        if n & 1 and n > 1:
            assert exp[-1] == 0
    with pytest.raises(AssertionError, match="power of 2"):
        fft_ff_interpolation([F(1)] * 3, F(root_of_unity(CURVES[curve], 4)), F)
    with pytest.raises(RecursionError):
        fft_ff([], F(1), F)
    one = [F(7)]
    assert fft_ff(one, F(1), F) is one                                  # fft_ff.py:16-17


def test_open_degree_overflow_is_commits_error():
    """kzg.py:157 -> :103-106: open() of a polynomial whose QUOTIENT exceeds the key raises commit's ValueError and message."""
    from kzg_snark_b200.kzg import KZG
    kzg = KZG("bn254")
    ck, _ = kzg.setup(4, tau=99)
    X = kzg.X
    with pytest.raises(ValueError, match="Polynomial degree 5 exceeds maximum allowed degree 4"):
        kzg.open(ck, [X ** 6 + 1], 3, 5)
    kzg.open(ck, [X ** 5 + 1], 3, 5)                                    # quotient of degree 4 fits


@pytest.mark.parametrize("logn", [20])
def test_tau_identity_large(logn):
    """Full-size style check: commit(ck, p) == p(tau) * G1 at 2^20 points."""
    from kzg_snark_b200 import device
    from kzg_snark_b200.limbs import random_scalars, limbs_to_ints
    cv = get_curve("bn254")
    n = 1 << logn
    tau = 0x1234567 ** 5 % cv.r
    srs = device.Srs.generate("bn254", tau, n)
    sc = random_scalars(n, cv.r, seed=77)
    out, inf = device.msm(srs, sc)
    exp = cv.normalize(cv.multiply(cv.G1, poly_eval(limbs_to_ints(sc), tau, cv.r)))
    assert (None if inf else tuple(limbs_to_ints(out.reshape(2, 4)))) == exp
    # sharded: two index ranges + fold == whole (the multi-GPU decomposition on one device)
    import numpy as np
    from kzg_snark_b200 import _ffi
    h = n // 2
    s0 = device.Srs.generate("bn254", tau, h, start=0)
    s1 = device.Srs.generate("bn254", tau, h, start=h)
    d0 = _ffi.DeviceBuffer(h * 32).upload(sc[:h]); d1 = _ffi.DeviceBuffer(h * 32).upload(sc[h:])
    parts = _ffi.DeviceBuffer(2 * 128)
    class Off:
        def __init__(self, base, off):
            import ctypes
            self.ptr = ctypes.c_void_p(base.ptr.value + off)
    device.msm_partial_dev(s0, d0, h, Off(parts, 0))
    device.msm_partial_dev(s1, d1, h, Off(parts, 128))
    out2, inf2 = device.g1_fold("bn254", parts, 2)
    assert tuple(limbs_to_ints(out2.reshape(2, 4))) == exp
    for s in (srs, s0, s1):
        s.destroy()


@pytest.mark.parametrize("name", ["ref_trace_kzg.json", "ref_trace_kzg_bls.json", "ref_trace_fft.json", "ref_trace_fft_bls.json", "ref_trace_plonk.json",
                                  "ref_trace_marlin.json"])
def test_dropin_reproduces_reference_trace(name):
    """Every commit / open / fft_ff / ifft_ff / fft_ff_interpolation call the reference's own
    kzg.py, plonk and marlin provers made (recorded in the build container by
    tests/golden/make_traces.py) replayed through the GPU drop-in: identical affine points and
    identical residues, call by call (configs[0], configs[3] and configs[4] call sites)."""
    from trace_replay import load, replay
    from kzg_snark_b200.kzg import KZG
    from kzg_snark_b200.fft_ff import fft_ff, ifft_ff, fft_ff_interpolation
    trace = load(name)
    curve = trace.get("curve", "bn254")                  # ref_trace_kzg_bls.json: the reference's KZG("bls12_381") run
    cv = get_curve(curve)
    kzg = KZG(curve)
    F, fq = kzg.Fq, kzg._codec.fq

    def poly_ints(p):
        return [int(c) for c in (p.list() if hasattr(p, "list") else p)]

    n = replay(
        trace,
        make_key=lambda pts: [kzg.Z1 if p is None else (fq(p[0]), fq(p[1]), fq(1)) for p in pts],
        commit=kzg.commit, open_=kzg.open,
        fft=lambda v, w: fft_ff([F(x) for x in v], F(w), F),
        ifft=lambda v, w: ifft_ff([F(x) for x in v], F(w), F),
        interp=lambda v, w: fft_ff_interpolation([F(x) for x in v], F(w), F),
        to_affine=lambda pt: aff(cv, pt), to_ints=poly_ints)
    assert n == len(trace["calls"]) > 0


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_g1_lincomb_matches_oracle(curve):
    """kzgpu_g1_lincomb: sum_i s_i * P_i over arbitrary points, incl. the identity,
    repeated points, P + (-P), scalars 0 / 1 / r-1, and the empty sum."""
    from kzg_snark_b200 import device
    from kzg_snark_b200.limbs import ints_to_limbs, limbs_to_ints
    cv = get_curve(curve)
    nl = CURVES[curve]["fp_limbs32"] // 2
    rng = random.Random(12)
    base = [cv.multiply(cv.G1, rng.randrange(1, cv.r)) for _ in range(9)]
    cases = [([], []), ([base[0]], [0]), ([base[0]], [1]), ([base[1]], [cv.r - 1]), ([cv.Z1, base[2]], [5, 7]),
             ([base[3], base[3]], [3, 4]), ([base[4], cv.neg(base[4])], [11, 11]),
             (base, [rng.randrange(cv.r) for _ in base]),
             ([base[i % 9] for i in range(150)], [rng.randrange(cv.r) for _ in range(150)])]
    for pts, sc in cases:
        flat = []
        for p in pts:
            a = cv.normalize(p)
            flat += [0, 0] if a is None else [a[0], a[1]]
        arr = ints_to_limbs(flat, cv.p, nl).reshape(len(pts), 2 * nl) if pts else np.zeros((0, 2 * nl), np.uint64)
        out, inf = device.g1_lincomb(curve, arr, ints_to_limbs(sc, cv.r) if sc else np.zeros((0, 4), np.uint64))
        acc = cv.Z1
        for p, k in zip(pts, sc):
            acc = cv.add(acc, cv.multiply(p, k))
        exp = cv.normalize(acc)
        assert (None if inf else tuple(limbs_to_ints(out.reshape(2, nl)))) == exp


def test_check_and_batch_check_with_device_combinations():
    """KZG.check / batch_check (kzg.py:161-288) with their G1 combinations on the device; the pairing
    and G2 arithmetic -- py_ecc's in the reference, absent here -- are supplied by the oracle's
    stand-in, exactly the role py_ecc plays for the drop-in.  Accepts honest openings, rejects
    tampered ones, like the reference's self-test (kzg.py:337-380)."""
    from kzg_snark_b200.kzg import KZG
    from oracle import pyecc_standin as E
    kzg = KZG("bn254")
    if kzg.have_py_ecc:
        pytest.skip("py_ecc present: the drop-in already uses it")
    g1 = lambda P: tuple(E.FQ(int(c)) for c in P)                              # noqa: E731
    is_g2 = lambda P: isinstance(P[0], E.FQ2)                                  # noqa: E731
    kzg.G2 = E.G2
    kzg.pairing = lambda Q, P: E.pairing(Q, g1(P))
    dev_mul, dev_add, dev_neg = kzg.multiply, kzg.add, kzg.neg
    kzg.multiply = lambda P, n: E.multiply(P, n) if is_g2(P) else dev_mul(P, n)
    kzg.add = lambda P, Q: E.add(P, Q) if is_g2(P) else dev_add(P, Q)
    kzg.neg = lambda P: E.neg(P) if is_g2(P) else dev_neg(P)
    rng = random.Random(21)
    r = kzg.curve_order
    tau = rng.randrange(1, r)
    ck, _ = kzg.setup(40, tau=tau)
    rk = E.multiply(E.G2, tau)
    polys = [kzg.R([rng.randrange(r) for _ in range(m)]) for m in (41, 17, 1)]
    comm = kzg.commit(ck, polys)
    z, xi = rng.randrange(r), rng.randrange(r)
    proof = kzg.open(ck, polys, z, xi)
    evals = [p(z) for p in polys]
    assert kzg.check(rk, comm, z, evals, proof, xi)
    assert not kzg.check(rk, comm, z, [evals[0] + 1] + evals[1:], proof, xi)
    z2, xi2 = rng.randrange(r), rng.randrange(r)
    proof2 = kzg.open(ck, polys[:2], z2, xi2)
    ok = kzg.batch_check(rk, [comm, comm[:2]], [z, z2], [evals, [p(z2) for p in polys[:2]]], [proof, proof2], [xi, xi2], r=rng.randrange(r))
    assert ok
    bad = kzg.batch_check(rk, [comm, comm[:2]], [z, z2], [evals, [polys[0](z2) + 1, polys[1](z2)]], [proof, proof2], [xi, xi2], r=rng.randrange(r))
    assert not bad
    # the group-operation attributes the verifiers read (plonk/verifier.py:117-157) agree with the oracle
    cv = get_curve("bn254")
    P = kzg.multiply(kzg.G1, 12345)
    assert aff(cv, P) == cv.normalize(cv.multiply(cv.G1, 12345))
    assert aff(cv, kzg.add(P, kzg.neg(P))) is None and kzg.eq(kzg.add(P, P), kzg.multiply(P, 2))
