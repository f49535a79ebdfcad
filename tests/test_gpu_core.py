"""GPU parity tests proper: every check goes through the C ABI (libkzgpu.so) and compares
with the CPU oracle (oracle/) on the same seeded inputs.  Bit-exact (integer arithmetic)."""
import random

import numpy as np
import pytest

from oracle.params import CURVES, root_of_unity
from oracle.curve import get_curve
from oracle.kzg import KZGOracle, poly_eval, poly_div_linear
from oracle import fft_ff as off

pytestmark = pytest.mark.gpu

CURVE_NAMES = ["bn254", "bls12_381"]


@pytest.fixture(scope="module")
def dev():
    from kzg_snark_b200 import device, _ffi
    _ffi.init()
    return device


def L(vals, mod, n=4):
    from kzg_snark_b200.limbs import ints_to_limbs
    return ints_to_limbs(vals, mod, n)


def I(arr):
    from kzg_snark_b200.limbs import limbs_to_ints
    return limbs_to_ints(arr)


def affine_arr(cv, pts, nl):
    """oracle projective points -> (n, 2*nl) canonical limbs, infinity = (0,0)."""
    flat = []
    for p in pts:
        a = cv.normalize(p)
        flat += [0, 0] if a is None else [a[0], a[1]]
    return L(flat, cv.p, nl).reshape(len(pts), 2 * nl)


def point_of(cv, out, inf, nl):
    if inf:
        return None
    v = I(out.reshape(2, nl))
    return (v[0], v[1])


# ------------------------------------------------------------------ field core
@pytest.mark.parametrize("curve", CURVE_NAMES)
@pytest.mark.parametrize("which", [0, 1])
def test_field_ops(dev, curve, which):
    cv = CURVES[curve]
    mod = cv["p"] if which == 0 else cv["r"]
    nl = (cv["fp_limbs32"] if which == 0 else cv["fr_limbs32"]) // 2
    rng = random.Random(100 + which)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, 1 << 32, (1 << 64) - 1]
    a = edge + [rng.randrange(mod) for _ in range(3000)]
    b = list(reversed(edge)) + [rng.randrange(mod) for _ in range(3000)]
    A, Bm = L(a, mod, nl), L(b, mod, nl)
    assert I(dev.field_op(curve, which, 0, A, Bm)) == [x * y % mod for x, y in zip(a, b)]
    assert I(dev.field_op(curve, which, 1, A, Bm)) == [(x + y) % mod for x, y in zip(a, b)]
    assert I(dev.field_op(curve, which, 2, A, Bm)) == [(x - y) % mod for x, y in zip(a, b)]
    inv = I(dev.field_op(curve, which, 3, A[:64]))
    assert inv == [pow(x, -1, mod) if x else 0 for x in a[:64]]


# ------------------------------------------------------------------ NTT
@pytest.mark.parametrize("curve", CURVE_NAMES)
@pytest.mark.parametrize("logn", list(range(0, 15)))
def test_ntt_matches_oracle(dev, curve, logn):
    cv = CURVES[curve]
    r = cv["r"]
    n = 1 << logn
    rng = random.Random(1000 + logn)
    x = [rng.randrange(r) for _ in range(n)]
    if n >= 4:
        x[1] = 0; x[2] = r - 1
    w = root_of_unity(cv, n)
    wl = L([w], r)[0]
    got = I(dev.ntt(curve, L(x, r), wl))
    assert got == off.fft_ff_int(x, w, r)
    got = I(dev.ntt(curve, L(x, r), wl, inverse=True))
    assert got == off.ifft_ff_int(x, w, r)
    sh = L([7], r)[0]
    got = I(dev.ntt(curve, L(x, r), wl, coset_limbs=sh))
    assert got == off.coset_fft_ff_int(x, w, 7, r)
    got = I(dev.ntt(curve, L(x, r), wl, inverse=True, coset_limbs=sh))
    assert got == off.coset_ifft_ff_int(x, w, 7, r)


def test_ntt_batch_and_nonprimitive_root(dev):
    cv = CURVES["bn254"]; r = cv["r"]
    rng = random.Random(5)
    n, batch = 32, 9                              # the Marlin prover's 9 forward FFTs of size m=32
    w = root_of_unity(cv, n)
    xs = [[rng.randrange(r) for _ in range(n)] for _ in range(batch)]
    flat = [v for x in xs for v in x]
    got = I(dev.ntt("bn254", L(flat, r), L([w], r)[0], batch=batch))
    exp = [v for x in xs for v in off.fft_ff_int(x, w, r)]
    assert got == exp
    # fft_ff never validates w: any w gives sum_j c_j w^(jk)
    w2 = rng.randrange(2, r)
    x = xs[0][:16]
    assert I(dev.ntt("bn254", L(x, r), L([w2], r)[0])) == off.fft_ff_int(x, w2, r)


def test_ntt_rejects_non_power_of_two(dev):
    from kzg_snark_b200._ffi import KzgpuError
    cv = CURVES["bn254"]; r = cv["r"]
    with pytest.raises(KzgpuError):
        dev.ntt("bn254", L([1, 2, 3], r), L([5], r)[0])


@pytest.mark.parametrize("logn", [16, 18, 20, 22, 24])
def test_ntt_large_properties(dev, logn):
    """Full-size checks (up to BASELINE.json's 2^24) that need no O(n log n) CPU work: Horner spot
    checks (on the host up to 2^18, by the device Horner kernel above), round trip, delta."""
    import ctypes
    from kzg_snark_b200 import _ffi
    from kzg_snark_b200.limbs import random_scalars, int_to_limbs, limbs_to_int
    cv = CURVES["bn254"]; r = cv["r"]
    n = 1 << logn
    w = root_of_unity(cv, n); wl = L([w], r)[0]
    x = random_scalars(n, r, seed=logn)
    y = dev.ntt("bn254", x.copy(), wl)
    xi = I(x) if logn <= 18 else None
    if xi is not None:
        ks = [0, 1, 2, n // 2, n - 1, 12345 % n]
        yi = I(y[ks])
        assert yi == off.dft_definition(xi, w, r, ks)
    else:
        dx = _ffi.DeviceBuffer(n * 32).upload(x)
        for k in (0, 1, n // 2, n - 1, 987654321 % n):            # y[k] = x(w^k), fft_ff.py:32-35
            out = np.zeros(4, dtype=np.uint64)
            _ffi.check(_ffi.load_library().kzgpu_poly_eval_dev(_ffi.BN254, dx.ptr, n, _ffi.ptr(int_to_limbs(pow(w, k, r), r)),
                                                               _ffi.ptr(out)))
            assert limbs_to_int(out) == I(y[k:k + 1])[0]
        dx.free()
    back = dev.ntt("bn254", y.copy(), wl, inverse=True)
    assert np.array_equal(back, x)
    d = np.zeros((n, 4), dtype=np.uint64); d[1, 0] = 1          # delta_1 -> [w^k]
    yd = dev.ntt("bn254", d, wl)
    ks = [0, 1, 2, 3, n // 2 + 1, n - 1]
    assert I(yd[ks]) == [pow(w, k, r) for k in ks]
    # linearity checksum: sum_k y[k] = n * x[0]
    if xi is not None:
        assert sum(I(y)) % r == n * xi[0] % r


# ------------------------------------------------------------------ MSM / commit
@pytest.mark.parametrize("curve", CURVE_NAMES)
@pytest.mark.parametrize("n", [0, 1, 2, 3, 22, 201, 1025])
def test_msm_matches_oracle(dev, curve, n):
    cv = get_curve(curve); k = KZGOracle(curve)
    nl = CURVES[curve]["fp_limbs32"] // 2
    rng = random.Random(n + 7)
    tau = rng.randrange(1, cv.r)
    ck = k.setup_fast(max(n - 1, 0), tau)[:max(n, 1)]
    srs = dev.Srs.from_affine(curve, affine_arr(cv, ck, nl))
    coeffs = [rng.randrange(cv.r) for _ in range(n)]
    if n >= 3:
        coeffs[0] = 0; coeffs[1] = 1; coeffs[2] = cv.r - 1
    out, inf = dev.msm(srs, L(coeffs, cv.r) if n else np.zeros((0, 4), np.uint64))
    exp = cv.normalize(cv.multiply(cv.G1, poly_eval(coeffs, tau, cv.r)))      # tau-identity, kzg.py:108
    assert point_of(cv, out, inf, nl) == exp
    if 0 < n <= 201:
        exp2 = cv.normalize(k.commit(ck, [coeffs])[0])                       # the reference's own loop
        assert point_of(cv, out, inf, nl) == exp2
    srs.destroy()


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_msm_edge_points(dev, curve):
    """duplicate points, P + (-P), infinity entries, all-equal scalars, zero polynomial."""
    cv = get_curve(curve)
    nl = CURVES[curve]["fp_limbs32"] // 2
    rng = random.Random(3)
    P = cv.multiply(cv.G1, 1234567)
    Q = cv.multiply(cv.G1, 7654321)
    pts = [P, P, cv.neg(P), cv.Z1, Q, Q, P, cv.G1]
    srs = dev.Srs.from_affine(curve, affine_arr(cv, pts, nl))
    cases = [
        [1, 1, 1, 5, 0, 0, 0, 0],
        [1, 0, 1, 0, 0, 0, 0, 0],          # P + (-P) = O
        [0] * 8,                           # zero polynomial -> Z1 (kzg.py:109)
        [3, 3, 3, 3, 3, 3, 3, 3],
        [cv.r - 1] * 8,
        [rng.randrange(cv.r) for _ in range(8)],
        [2, cv.r - 2, 0, 9, 5, cv.r - 5, 0, 0],
    ]
    for sc in cases:
        acc = cv.Z1
        for p, s in zip(pts, sc):
            acc = cv.add(acc, cv.multiply(p, s))
        out, inf = dev.msm(srs, L(sc, cv.r))
        assert point_of(cv, out, inf, nl) == cv.normalize(acc), sc
    srs.destroy()


@pytest.mark.parametrize("mode", ["0", "c=5", "c=11"])
def test_msm_key_layouts_agree(dev, mode, monkeypatch):
    """Plain key (per-call windows + Horner) and fixed-base window tables give the same point."""
    cv = get_curve("bn254"); k = KZGOracle("bn254")
    rng = random.Random(17)
    tau = rng.randrange(1, cv.r)
    n = 300
    ck = k.setup_fast(n - 1, tau)
    monkeypatch.setenv("KZGPU_SRS_TABLES", mode)
    srs = dev.Srs.from_affine("bn254", affine_arr(cv, ck, 4))
    info = srs.info()
    assert (info["c"] == 0 and info["tables"] == 1) if mode == "0" else info["c"] == int(mode[2:])
    for coeffs in ([rng.randrange(cv.r) for _ in range(n)], [1] * n, [cv.r - 1] * n, [0] * n,
                   [rng.randrange(1 << 16) for _ in range(n)], [rng.randrange(cv.r) for _ in range(7)]):
        out, inf = dev.msm(srs, L(coeffs, cv.r))
        exp = cv.normalize(cv.multiply(cv.G1, poly_eval(coeffs, tau, cv.r)))
        assert point_of(cv, out, inf, 4) == exp
    # offset into the key (first > 0): sum_i s_i * ck[first + i]
    out, inf = dev.msm(srs, L([3, 5, 7], cv.r), first=10)
    acc = cv.Z1
    for j, s_ in enumerate([3, 5, 7]):
        acc = cv.add(acc, cv.multiply(ck[10 + j], s_))
    assert point_of(cv, out, inf, 4) == cv.normalize(acc)
    srs.destroy()


def test_msm_degree_check(dev):
    cv = get_curve("bn254")
    srs = dev.Srs.from_affine("bn254", affine_arr(cv, [cv.G1] * 4, 4))
    with pytest.raises(ValueError, match="exceeds maximum allowed degree 3"):   # kzg.py:103-106
        dev.msm(srs, L([1] * 5, cv.r))
    srs.destroy()


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_srs_generate_and_tau_identity(dev, curve):
    from kzg_snark_b200.limbs import random_scalars
    cv = get_curve(curve)
    nl = CURVES[curve]["fp_limbs32"] // 2
    tau = 0xC0FFEE1234567 % cv.r
    n = 1 << 14
    srs = dev.Srs.generate(curve, tau, n)
    for i in (0, 1, 2, n - 1):
        got = I(srs.read(i, 1).reshape(2, nl))
        assert tuple(got) == cv.normalize(cv.multiply(cv.G1, pow(tau, i, cv.r)))
    sc = random_scalars(n, cv.r, seed=11)
    out, inf = dev.msm(srs, sc)
    exp = cv.normalize(cv.multiply(cv.G1, poly_eval(I(sc), tau, cv.r)))
    assert point_of(cv, out, inf, nl) == exp
    # skewed, witness-like scalars: half zeros, a quarter small
    s2 = sc.copy(); s2[::2] = 0; s2[1::4, 1:] = 0; s2[1::4, 0] &= np.uint64(0xFFFF)
    out, inf = dev.msm(srs, s2)
    exp = cv.normalize(cv.multiply(cv.G1, poly_eval(I(s2), tau, cv.r)))
    assert point_of(cv, out, inf, nl) == exp
    srs.destroy()


def test_msm_full_size_chunked_upload_and_heavy_buckets(dev):
    """Sizes at which the two-level sort runs its full geometry and the host-scalar entry point
    uploads in four overlapped chunks that accumulate into the same buckets: tau-identity
    (kzg.py:108) for uniform, all-equal (one huge bucket per window, split + merged across
    chunks) and witness-like scalars; host-buffer and device-resident calls agree."""
    from kzg_snark_b200 import _ffi
    from kzg_snark_b200.limbs import random_scalars, ints_to_limbs
    cv = get_curve("bn254")
    tau = 0x1D2C3B4A5F6E7D8C9BA % cv.r
    n = (1 << 20) + 77
    srs = dev.Srs.generate("bn254", tau, n)
    rng = random.Random(5)
    uniform = random_scalars(n, cv.r, seed=21)
    same = np.tile(ints_to_limbs([rng.randrange(cv.r)], cv.r), (n, 1))
    skew = uniform.copy(); skew[::2] = 0; skew[1::4, 1:] = 0; skew[1::4, 0] &= np.uint64(0xFFFF)
    for sc in (uniform, same, skew):
        exp = cv.normalize(cv.multiply(cv.G1, poly_eval(I(sc), tau, cv.r)))
        out, inf = dev.msm(srs, sc)                                   # chunked H2D inside
        assert point_of(cv, out, inf, 4) == exp
        d = _ffi.DeviceBuffer(n * 32).upload(np.ascontiguousarray(sc))
        out2, inf2 = dev.msm_dev(srs, d, n)                           # single pass, resident
        d.free()
        assert point_of(cv, out2, inf2, 4) == exp
    srs.destroy()


# ------------------------------------------------------------------ open
@pytest.mark.parametrize("curve", CURVE_NAMES)
@pytest.mark.parametrize("lens", [[5], [1], [4, 9, 2], [70, 64, 65, 1, 130], [5000, 4097]])
def test_open_matches_oracle(dev, curve, lens):
    cv = get_curve(curve); k = KZGOracle(curve)
    nl = CURVES[curve]["fp_limbs32"] // 2
    rng = random.Random(sum(lens))
    tau = rng.randrange(1, cv.r)
    polys = [[rng.randrange(cv.r) for _ in range(m)] for m in lens]
    z, xi = rng.randrange(cv.r), rng.randrange(cv.r)
    wit = k.witness(polys, z, xi)
    comb = k.combine(polys, xi)
    quot, ev = dev.open_quotient(curve, [L(p, cv.r) for p in polys], L([z], cv.r)[0], L([xi], cv.r)[0])
    q = I(quot) if len(quot) else []
    while q and q[-1] == 0:
        q.pop()
    assert q == wit
    assert I(ev)[0] == poly_eval(comb, z, cv.r)
    n = max(lens)
    srs = dev.Srs.generate(curve, tau, n)
    out, inf = dev.open_proof(srs, [L(p, cv.r) for p in polys], L([z], cv.r)[0], L([xi], cv.r)[0])
    exp = cv.normalize(cv.multiply(cv.G1, poly_eval(wit, tau, cv.r)))
    assert point_of(cv, out, inf, nl) == exp
    srs.destroy()


def test_msm_tau_identity_at_baseline_size(dev):
    """BASELINE.json's size: a 2^24-point BN254 MSM with uniform scalars must equal p(tau)*G1
    (kzg.py:108).  p(tau) comes from the device Horner kernel (kzgpu_poly_eval_dev), which is
    checked here against CPython on a 2^16 slice and by the split p = p_lo + tau^(n/2) p_hi."""
    import ctypes
    from kzg_snark_b200 import _ffi
    from kzg_snark_b200.limbs import random_scalars, int_to_limbs, limbs_to_int
    cv = get_curve("bn254")
    tau = 0x5A17C0FFEE5EED1234567 % cv.r
    n = 1 << 24
    srs = dev.Srs.generate("bn254", tau, n)
    sc = random_scalars(n, cv.r, seed=24)
    d = _ffi.DeviceBuffer(n * 32).upload(sc)
    lib = _ffi.load_library()

    def horner(offset, count):
        out = np.zeros(4, dtype=np.uint64)
        _ffi.check(lib.kzgpu_poly_eval_dev(_ffi.BN254, ctypes.c_void_p(d.ptr.value + 32 * offset), count,
                                           _ffi.ptr(int_to_limbs(tau, cv.r)), _ffi.ptr(out)))
        return limbs_to_int(out)

    assert horner(12345, 1 << 16) == poly_eval(I(sc[12345:12345 + (1 << 16)]), tau, cv.r)
    p_tau = horner(0, n)
    assert p_tau == (horner(0, n // 2) + pow(tau, n // 2, cv.r) * horner(n // 2, n // 2)) % cv.r
    exp = cv.normalize(cv.multiply(cv.G1, p_tau))
    out, inf = dev.msm_dev(srs, d, n)
    assert point_of(cv, out, inf, 4) == exp
    out, inf = dev.msm(srs, sc)                                        # host scalars, chunked upload
    assert point_of(cv, out, inf, 4) == exp
    d.free(); srs.destroy()


@pytest.mark.parametrize("curve", CURVE_NAMES)
@pytest.mark.parametrize("mode", ["auto", "0"])
def test_msm_batch_single_pass_matches_per_polynomial(dev, curve, mode, monkeypatch):
    """KZG.commit(ck, [k polys]) (kzg.py:102): the k MSMs share one sort/accumulate/reduce pass;
    every result must equal the tau-identity of its own polynomial -- tabled and plain keys, ragged
    lengths, a zero polynomial and an empty one, zero coefficients, device-resident equal-length form."""
    from kzg_snark_b200 import _ffi
    cv = get_curve(curve)
    nl = CURVES[curve]["fp_limbs32"] // 2
    if mode != "auto":
        monkeypatch.setenv("KZGPU_SRS_TABLES", mode)
    rng = random.Random(33)
    tau = rng.randrange(1, cv.r)
    n = 700
    srs = dev.Srs.generate(curve, tau, n)
    polys = [[rng.randrange(cv.r) for _ in range(m)] for m in (700, 1, 333, 64)]
    polys += [[0] * 50, [], [0, 0, 5, 0, cv.r - 1], [rng.randrange(1 << 20) for _ in range(699)]]
    out, infs = dev.msm_batch(srs, [L(p, cv.r) if p else np.zeros((0, 4), np.uint64) for p in polys])
    for p, o, f in zip(polys, out, infs):
        t = poly_eval(p, tau, cv.r)
        exp = cv.normalize(cv.multiply(cv.G1, t)) if t else None
        assert point_of(cv, o, f, nl) == exp
    # device-resident, equal length (what the PLONK prover uses for a | b | c and t_lo | t_mid | t_hi)
    m, k = 513, 5
    eq = [[rng.randrange(cv.r) for _ in range(m)] for _ in range(k)]
    eq[2] = [0] * m
    d = _ffi.DeviceBuffer(k * m * 32).upload(np.concatenate([L(p, cv.r) for p in eq]))
    out, infs = dev.msm_batch_dev(srs, d, m, k)
    for p, o, f in zip(eq, out, infs):
        t = poly_eval(p, tau, cv.r)
        assert point_of(cv, o, f, nl) == (cv.normalize(cv.multiply(cv.G1, t)) if t else None)
    with pytest.raises(ValueError, match="exceeds maximum allowed degree 699"):
        dev.msm_batch_dev(srs, d, 701, 2)
    d.free(); srs.destroy()


def test_msm_batch_large_matches_single(dev):
    """2^18-point polynomials: the batched pass and three single MSMs give identical points."""
    from kzg_snark_b200 import _ffi
    from kzg_snark_b200.limbs import random_scalars
    cv = get_curve("bn254")
    n, k = (1 << 18) + 6, 3
    srs = dev.Srs.generate("bn254", 0xABCDEF123456789, n)
    sc = random_scalars(n * k, cv.r, seed=77)
    sc[n + 5:2 * n] = 0                                       # a mostly-zero polynomial (padding case)
    d = _ffi.DeviceBuffer(n * k * 32).upload(sc)
    out, infs = dev.msm_batch_dev(srs, d, n, k)
    for j in range(k):
        o1, f1 = dev.msm(srs, sc[j * n:(j + 1) * n])
        assert f1 == infs[j] and (o1 == out[j]).all()
    d.free(); srs.destroy()


def test_point_sharded_open_equals_single_open(dev):
    """The multi-GPU form of KZG.open: the quotient is formed on the device
    (kzgpu_open_quotient_dev), index ranges of it go through kzgpu_msm_partial_dev, and the XYZZ
    partials are folded -- here the 3 'ranks' run one after another on one GPU.  Must equal
    kzgpu_open and the oracle."""
    import ctypes
    from kzg_snark_b200 import _ffi
    from kzg_snark_b200.parallel import shard_range
    cv = get_curve("bn254"); ko = KZGOracle("bn254")
    rng = random.Random(8)
    tau = rng.randrange(1, cv.r)
    lens = [3000, 2999, 1, 1500]
    polys = [[rng.randrange(cv.r) for _ in range(m)] for m in lens]
    z, xi = rng.randrange(cv.r), rng.randrange(cv.r)
    srs = dev.Srs.generate("bn254", tau, 3000)
    ref, inf = dev.open_proof(srs, [L(p, cv.r) for p in polys], L([z], cv.r)[0], L([xi], cv.r)[0])
    bufs = [_ffi.DeviceBuffer(max(m, 1) * 32).upload(L(p, cv.r)) for p, m in zip(polys, lens)]
    dq = _ffi.DeviceBuffer(3000 * 32)
    ptrs = (ctypes.c_void_p * 4)(*[b.ptr.value for b in bufs])
    ls = (ctypes.c_size_t * 4)(*lens)
    ql = ctypes.c_size_t(0)
    ev = np.zeros(4, dtype=np.uint64)
    lib = _ffi.load_library()
    _ffi.check(lib.kzgpu_open_quotient_dev(_ffi.BN254, ptrs, ls, 4, _ffi.ptr(L([z], cv.r)[0]), _ffi.ptr(L([xi], cv.r)[0]),
                                           dq.ptr, ctypes.byref(ql), _ffi.ptr(ev)))
    assert ql.value == 2999 and I(ev)[0] == poly_eval(ko.combine(polys, xi), z, cv.r)
    world = 3
    parts = _ffi.DeviceBuffer(128 * world)

    class At:
        def __init__(self, base, off):
            self.ptr = ctypes.c_void_p(base.ptr.value + off)

    for rank in range(world):
        s0, cnt = shard_range(ql.value, world, rank)
        dev.msm_partial_dev(srs, At(dq, 32 * s0), cnt, At(parts, 128 * rank), first=s0)
    out, finf = dev.g1_fold("bn254", parts, world)
    assert not inf and not finf and (out == ref).all()
    wit = ko.witness(polys, z, xi)
    assert point_of(cv, out, finf, 4) == cv.normalize(cv.multiply(cv.G1, poly_eval(wit, tau, cv.r)))
    srs.destroy()


def test_point_sharded_msm_from_host_scalars(dev):
    """kzgpu_msm_partial (each rank's slice of the scalars in host memory, upload chunked and overlapped inside the call:
    2^21 scalars per 'rank' take the 4-chunk pipelined path) against kzgpu_msm_partial_dev and the single-GPU MSM; the
    two 'ranks' run one after another on one GPU."""
    import ctypes
    from kzg_snark_b200 import _ffi
    from kzg_snark_b200.limbs import random_scalars
    from kzg_snark_b200.parallel import shard_range
    cv = get_curve("bn254")
    n, world = (1 << 22) + 12345, 2
    srs = dev.Srs.generate("bn254", 0x1234ABCD5678EF, n)
    sc = random_scalars(n, cv.r, seed=91)
    sc[1000:5000] = 0
    ref, rinf = dev.msm(srs, sc)

    class At:
        def __init__(self, base, off):
            self.ptr = ctypes.c_void_p(base.ptr.value + off)

    parts_h, parts_d = _ffi.DeviceBuffer(128 * world), _ffi.DeviceBuffer(128 * world)
    d = _ffi.DeviceBuffer(n * 32).upload(sc)
    for rank in range(world):
        s0, cnt = shard_range(n, world, rank)
        dev.msm_partial(srs, sc[s0:s0 + cnt], At(parts_h, 128 * rank), first=s0)
        dev.msm_partial_dev(srs, At(d, 32 * s0), cnt, At(parts_d, 128 * rank), first=s0)
    out_h, inf_h = dev.g1_fold("bn254", parts_h, world)
    out_d, inf_d = dev.g1_fold("bn254", parts_d, world)
    assert not rinf and not inf_h and not inf_d
    assert (out_h == ref).all() and (out_d == ref).all()
    d.free(); srs.destroy()


def test_msm_batch_falls_back_per_polynomial_when_the_pass_does_not_fit(dev):
    """More polynomials than one pass takes (> 64 bucket sets): the batch entry point runs them one by one;
    same results.  Also the argument checks of the verifier-side combination."""
    from kzg_snark_b200 import _ffi
    cv = get_curve("bn254")
    rng = random.Random(3)
    tau = rng.randrange(1, cv.r)
    srs = dev.Srs.generate("bn254", tau, 40)
    k, m = 70, 33
    polys = [[rng.randrange(cv.r) for _ in range(m)] for _ in range(k)]
    d = _ffi.DeviceBuffer(k * m * 32).upload(np.concatenate([L(p, cv.r) for p in polys]))
    out, infs = dev.msm_batch_dev(srs, d, m, k)
    for j in (0, 1, 37, 69):
        assert point_of(cv, out[j], infs[j], 4) == cv.normalize(cv.multiply(cv.G1, poly_eval(polys[j], tau, cv.r)))
    with pytest.raises(_ffi.KzgpuError):
        dev.g1_lincomb("bn254", np.zeros((70000, 8), np.uint64), np.zeros((70000, 4), np.uint64))
    d.free(); srs.destroy()
