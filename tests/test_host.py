"""CPU: host-side logic -- limb marshalling, the Sage-free shim, the drop-in modules' error
behaviour, Montgomery constants, and the field/curve formulas of csrc/ compiled for the host."""
import os
import random
import subprocess

import numpy as np
import pytest

from oracle.params import CURVES
from oracle.curve import get_curve

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_limb_round_trip_and_random_scalars():
    from kzg_snark_b200.limbs import ints_to_limbs, limbs_to_ints, int_to_limbs, limbs_to_int, random_scalars
    r = CURVES["bn254"]["r"]
    vals = [0, 1, r - 1, r, r + 5, 1 << 200, -3]
    a = ints_to_limbs(vals, r)
    assert a.shape == (7, 4) and a.dtype == np.uint64
    assert limbs_to_ints(a) == [v % r for v in vals]
    assert limbs_to_int(int_to_limbs(12345678901234567890123, r)) == 12345678901234567890123
    p = CURVES["bls12_381"]["p"]
    assert limbs_to_ints(ints_to_limbs([p - 1], p, 6)) == [p - 1]
    s = random_scalars(5000, r, seed=3)
    ints = limbs_to_ints(s)
    assert all(0 <= v < r for v in ints) and len(set(ints)) == 5000
    assert np.array_equal(s, random_scalars(5000, r, seed=3))          # seeded, reproducible


def test_montgomery_constants_match_moduli():
    import re
    src = open(os.path.join(ROOT, "kzg_snark_b200", "csrc", "params_gen.cuh")).read()
    mods = {"FpBN254": (CURVES["bn254"]["p"], 8), "FrBN254": (CURVES["bn254"]["r"], 8),
            "FpBLS381": (CURVES["bls12_381"]["p"], 12), "FrBLS381": (CURVES["bls12_381"]["r"], 8)}
    for name, (m, n) in mods.items():
        body = src[src.index(f"struct {name} "):]
        body = body[:body.index("\n};")]
        def arr(fn):
            t = body[body.index(f"uint32_t {fn}(int i)"):]
            words = re.findall(r"0x([0-9a-f]{8})u", t[:t.index("return")])
            return sum(int(w, 16) << (32 * i) for i, w in enumerate(words))
        R = 1 << (32 * n)
        assert arr("mod") == m
        assert arr("one") == R % m
        assert arr("r2") == R * R % m
        inv = int(re.search(r"INV32 = 0x([0-9a-f]{8})u", body).group(1), 16)
        assert (inv * m + 1) % (1 << 32) == 0
    c = get_curve("bls12_381")
    assert c.is_on_curve(c.G1)


@pytest.fixture(scope="module")
def host_arith(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host") / "host_arith")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "host_arith_test.cpp")], check=True)
    return exe


def test_csrc_field_and_curve_formulas_on_host(host_arith):
    """Same Montgomery row algorithm and XYZZ formulas as the device build (the carry-chain
    primitives are swapped for their portable mirror) against Python big-int arithmetic."""
    rng = random.Random(7)
    lines, expect = [], []
    F = {"fp_bn": CURVES["bn254"]["p"], "fr_bn": CURVES["bn254"]["r"], "fp_bls": CURVES["bls12_381"]["p"], "fr_bls": CURVES["bls12_381"]["r"]}
    for f, q in F.items():
        vals = [0, 1, 2, q - 1, q - 2, (q - 1) // 2] + [rng.randrange(q) for _ in range(20)]
        for _ in range(150):
            a, b = rng.choice(vals), rng.choice(vals)
            for op, fn in (("mul", lambda a, b: a * b % q), ("sqr", lambda a, b: a * a % q), ("add", lambda a, b: (a + b) % q),
                           ("sub", lambda a, b: (a - b) % q)):
                lines.append(f"{f} {op} {a:x} {b:x}"); expect.append("%x" % fn(a, b))
        for _ in range(3):
            a = rng.randrange(1, q); lines.append(f"{f} inv {a:x}"); expect.append("%x" % pow(a, -1, q))
    for cn, tag in (("bn254", "bn"), ("bls12_381", "bls")):
        c = get_curve(cn)
        aff = lambda pt: (0, 0) if c.normalize(pt) is None else c.normalize(pt)
        pts = [aff(c.multiply(c.G1, rng.randrange(1, c.r))) for _ in range(4)]
        for _ in range(10):
            a, b = rng.choice(pts), rng.choice(pts)
            for A, B in ((a, b), (a, a), (a, (a[0], (-a[1]) % c.p)), (a, (0, 0)), ((0, 0), b)):
                pa = (A[0], A[1], 1) if A != (0, 0) else c.Z1
                pb = (B[0], B[1], 1) if B != (0, 0) else c.Z1
                e = aff(c.add(pa, pb))
                lines.append(f"{tag} madd {A[0]:x} {A[1]:x} {B[0]:x} {B[1]:x}"); expect.append("%x %x" % e)
                if A != (0, 0) and B != (0, 0):
                    lines.append(f"{tag} add3 {A[0]:x} {A[1]:x} {B[0]:x} {B[1]:x}"); expect.append("%x %x" % e)
            k = rng.randrange(0, 1 << 20)
            lines.append(f"{tag} smul {a[0]:x} {a[1]:x} {k}"); expect.append("%x %x" % aff(c.multiply((a[0], a[1], 1), k)))
    out = subprocess.run([host_arith], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.split("\n")
    bad = [(l, e, o) for l, e, o in zip(lines, expect, out) if e != o]
    assert not bad, bad[:3]


def test_csrc_semi_reduced_arithmetic_on_host(host_arith):
    """The [0, 2p) primitives (field.cuh: fe_mul_lz, fe_add_lz, fe_sub_lz, *_nr) and the semi-reduced mixed addition
    (curve.cuh: xyzz_madd_lz) against Python integers: bounds, residues, zero test, fold -- for the three moduli
    with 4p <= 2^(32N), at the extremes of the ranges."""
    rng = random.Random(11)
    lines, expect = [], []
    F = {"fp_bn": (CURVES["bn254"]["p"], 256), "fr_bn": (CURVES["bn254"]["r"], 256), "fp_bls": (CURVES["bls12_381"]["p"], 384)}
    for f, (q, bits) in F.items():
        assert 4 * q <= 1 << bits
        Rinv = pow(1 << bits, -1, q)
        lo = [0, 1, q - 1, q, q + 1, 2 * q - 1] + [rng.randrange(2 * q) for _ in range(30)]
        hi = [2 * q, 3 * q, 4 * q - 1] + [rng.randrange(4 * q) for _ in range(20)]
        canon = [0, 1, q - 1] + [rng.randrange(q) for _ in range(20)]
        pairs = [(rng.choice(lo), rng.choice(lo)) for _ in range(200)] + [(rng.choice(hi), rng.choice(canon)) for _ in range(200)]
        pairs += [(2 * q - 1, 2 * q - 1), (4 * q - 1, q - 1), (q, q), (q, 0)]
        for a, b in pairs:
            lines.append(f"{f} lzmul {a:x} {b:x}"); expect.append(("mul", q, a * b * Rinv % q, None))
        for a in lo + [(1 << (bits - 2)) - 1, 2 * q - 2] + [rng.randrange(2 * q) for _ in range(200)]:
            lines.append(f"{f} lzsqr {a:x} 0"); expect.append(("mul", q, a * a * Rinv % q, None))
        # a*b - c*d with one Montgomery reduction (fe_mulsub_lz -> fe_mul2_nofinal): all four operands semi-reduced, extremes included
        quads = [tuple(rng.choice(lo) for _ in range(4)) for _ in range(300)]
        quads += [(2 * q - 1,) * 4, (2 * q - 1, 2 * q - 1, 0, 2 * q - 1), (2 * q - 1, 2 * q - 1, 1, 2 * q - 1), (0, 0, 2 * q - 1, 2 * q - 1),
                  (q, q, q, q), (1, 1, 1, 1), (5, 7, 7, 5)]
        for a, b, c, d in quads:
            lines.append(f"{f} lzmulsub {a:x} {b:x} {c:x} {d:x}"); expect.append(("fold", q, (a * b - c * d) * Rinv % q, None))
        for _ in range(200):
            a, b = rng.choice(lo), rng.choice(lo)
            lines.append(f"{f} lzadd {a:x} {b:x}"); expect.append(("fold", q, (a + b) % q, None))
            lines.append(f"{f} lzsub {a:x} {b:x}"); expect.append(("fold", q, (a - b) % q, None))
            lines.append(f"{f} nradd {a:x} {b:x}"); expect.append(("exact", q, (a + b) % q, a + b))
            lines.append(f"{f} nrsub {a:x} {b:x}"); expect.append(("exact", q, (a - b) % q, a - b + 2 * q))
    n_field = len(lines)
    for cn, tag in (("bn254", "bn"), ("bls12_381", "bls")):
        c = get_curve(cn)
        aff = lambda pt: (0, 0) if c.normalize(pt) is None else c.normalize(pt)
        pts = [aff(c.multiply(c.G1, rng.randrange(1, c.r))) for _ in range(4)]
        for _ in range(10):
            a, b = rng.choice(pts), rng.choice(pts)
            for A, B in ((a, b), (a, a), (a, (a[0], (-a[1]) % c.p)), (a, (0, 0)), ((0, 0), b)):
                pa = (A[0], A[1], 1) if A != (0, 0) else c.Z1
                pb = (B[0], B[1], 1) if B != (0, 0) else c.Z1
                e = aff(c.add(pa, pb))
                lines.append(f"{tag} lzmadd {A[0]:x} {A[1]:x} {B[0]:x} {B[1]:x}"); expect.append("%x %x" % e)
                if A != (0, 0) and B != (0, 0):
                    lines.append(f"{tag} lzchain {A[0]:x} {A[1]:x} {B[0]:x} {B[1]:x}"); expect.append("%x %x" % e)
                    lines.append(f"{tag} lzadd3 {A[0]:x} {A[1]:x} {B[0]:x} {B[1]:x}"); expect.append("%x %x" % e)
    out = subprocess.run([host_arith], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.split("\n")
    for i, (l, e, o) in enumerate(zip(lines, expect, out)):
        if i >= n_field:
            assert o == e, (l, e, o)
            continue
        kind, q, residue, exact = e
        raw, is_zero, folded = o.split()
        raw, folded = int(raw, 16), int(folded, 16)
        assert raw % q == residue, l
        if kind == "exact":
            assert raw == exact, l
        else:
            assert raw < 2 * q, l                                     # stays semi-reduced
            assert int(is_zero) == (1 if residue == 0 else 0), l
            assert folded == residue, l


def test_sageshim_field_and_polynomials():
    from kzg_snark_b200.sageshim import GF, PolynomialRing
    r = CURVES["bn254"]["r"]
    F = GF(r); R = PolynomialRing(F, "X"); X = R.gen()
    a, b = F(5), F(r - 2)
    assert int(a * b) == (5 * (r - 2)) % r and int(a / b * b) == 5 and int(a ** -1 * a) == 1
    assert str(F(r + 3)) == "3"                                         # prints like a Sage residue
    g = F(1).nth_root(16)
    assert g.multiplicative_order() == 16 and int(g ** 16) == 1 and int(g ** 8) == r - 1
    p = 1 + 2 * X + 3 * X ** 2
    assert [int(c) for c in p.list()] == [1, 2, 3] and p.degree() == 2
    assert R(0).degree() == -1 and R(0).list() == [] and R([0, 0]).degree() == -1
    assert int(p(F(2))) == 17
    q = (p - p(F(7))) // (X - 7)
    assert (q * (X - 7) + p(F(7))) == p
    vH = X ** 8 - 1
    assert ((X - 1) * vH) % vH == 0
    assert R(((X ** 8 - 1) * p) / vH) == p                                 # exact fraction -> polynomial
    with pytest.raises(TypeError):
        R(p / vH)
    assert p(g * X).degree() == 2                                         # composition z_poly(g*X)
    L = R.lagrange_polynomial([(F(1), F(3)), (F(2), F(5)), (F(4), F(9))])
    assert [int(L(F(x))) for x in (1, 2, 4)] == [3, 5, 9]


def test_dropin_error_behaviour_before_any_gpu_call():
    from kzg_snark_b200.kzg import KZG
    from kzg_snark_b200 import fft_ff as ff
    with pytest.raises(ValueError, match="Unsupported curve type: secp256k1"):            # kzg.py:37
        KZG("secp256k1")
    k = KZG("bn254")
    assert k.curve_order == CURVES["bn254"]["r"]
    assert int(k.G1[0]) == 1 and int(k.G1[1]) == 2 and int(k.Z1[2]) == 0
    ck = [k.G1] * 4
    with pytest.raises(ValueError, match="Polynomial degree 5 exceeds maximum allowed degree 3"):   # kzg.py:103-106
        k.commit(ck, [[1, 2, 3, 4, 5, 6]])
    g = k.Fq(1).nth_root(8)
    with pytest.raises(AssertionError, match="power of 2"):                                # fft_ff.py:74
        ff.fft_ff_interpolation([k.Fq(1)] * 6, g, k.Fq)
    with pytest.raises(AssertionError, match="must be at least"):                          # fft_ff.py:78
        ff.fft_ff_interpolation([k.Fq(1)] * 16, g, k.Fq)
    one = [k.Fq(3)]
    assert ff.fft_ff(one, g, k.Fq) is one                                                  # fft_ff.py:16-17
    if not k.have_py_ecc:
        with pytest.raises(ImportError, match="py_ecc"):
            k.check(None, [k.G1], 1, [1], k.G1, 2)


def test_point_codec_round_trip():
    from kzg_snark_b200.points import PointCodec
    from kzg_snark_b200.limbs import limbs_to_ints
    c = get_curve("bls12_381")
    codec = PointCodec("bls12_381", 1)
    P = c.multiply(c.G1, 99)                         # projective, z != 1
    arr = codec.points_to_limbs([P, c.Z1, c.G1, c.normalize(P)])
    rows = [tuple(limbs_to_ints(r.reshape(2, 6))) for r in arr]
    assert rows[0] == c.normalize(P) and rows[1] == (0, 0) and rows[2] == c.normalize(c.G1) and rows[3] == rows[0]
    pt = codec.from_device(arr[0], False)
    assert (int(pt[0]), int(pt[1]), int(pt[2])) == (*c.normalize(P), 1)
    z = codec.from_device(arr[1], True)
    assert int(z[2]) == 0


def test_ints_to_limbs_fast_paths_and_edge_values():
    """The Python-object boundary (kzg.py:110,115): plain ints, shim field elements (`.n`), arbitrary int()-able objects, and
    values outside [0, q) -- negative, == q, > q, wider than 256 bits -- all land as canonical little-endian limbs."""
    from kzg_snark_b200.limbs import ints_to_limbs, limbs_to_ints, int_to_limbs
    from kzg_snark_b200.sageshim import GF
    q = CURVES["bn254"]["r"]
    rng = random.Random(3)
    vals = [rng.randrange(q) for _ in range(1000)] + [0, 1, q - 1]
    a = ints_to_limbs(vals, q)
    assert a.shape == (1003, 4) and a.dtype == np.uint64 and a.flags["WRITEABLE"] and limbs_to_ints(a) == vals
    F = GF(q)
    assert (ints_to_limbs([F(v) for v in vals], q) == a).all()
    edge = [-1, q, q + 5, q - 1, 0, 2 ** 256 - 1, 3 * q + 7, -q, 2 ** 300 + 1]
    assert limbs_to_ints(ints_to_limbs(edge, q)) == [v % q for v in edge]
    mixed = vals[:10] + [q + 1] + vals[10:20]
    assert limbs_to_ints(ints_to_limbs(mixed, q)) == vals[:10] + [1] + vals[10:20]

    class Odd:
        def __init__(self, v):
            self.v = v

        def __int__(self):
            return self.v

    assert limbs_to_ints(ints_to_limbs([Odd(5), Odd(q + 2), Odd(-3)], q)) == [5, 2, q - 3]
    assert ints_to_limbs([], q).shape == (0, 4)
    p = CURVES["bls12_381"]["p"]
    big = [rng.randrange(p) for _ in range(50)] + [p, p + 1]
    assert limbs_to_ints(ints_to_limbs(big, p, 6)) == [v % p for v in big]
    assert limbs_to_ints(int_to_limbs(q + 3, q).reshape(1, 4)) == [3]
