"""CPU: pin the oracle (oracle/) against definitions, public constants and the committed
golden fixtures.  No GPU."""
import json
import os
import random

import pytest

from oracle import fft_ff as off
from oracle.curve import get_curve
from oracle.field import GFp
from oracle.kzg import KZGOracle, poly_eval, poly_div_linear
from oracle.params import CURVES, root_of_unity

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CURVE_NAMES = ["bn254", "bls12_381"]


def H(v):
    return int(v, 16)


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_public_constants(curve):
    from sympy import isprime
    cv = CURVES[curve]
    assert isprime(cv["p"]) and isprime(cv["r"])
    c = get_curve(curve)
    kat = json.load(open(os.path.join(GOLD, "public_kat.json")))[curve]
    g = tuple(H(x) for x in kat["G1"])
    assert g == cv["G1"] and c.is_on_curve(c.G1)
    assert c.normalize(c.double(c.G1)) == tuple(H(x) for x in kat["2G1"])
    assert c.normalize(c.add(c.G1, c.G1)) == tuple(H(x) for x in kat["2G1"])
    assert c.normalize(c.multiply(c.G1, 2)) == tuple(H(x) for x in kat["2G1"])
    assert c.is_inf(c.multiply(c.G1, cv["r"]))
    # 2-adicity and primitive root used for synthetic NTT inputs
    s = cv["fr_two_adicity"]
    assert (cv["r"] - 1) % (1 << s) == 0 and (cv["r"] - 1) % (1 << (s + 1)) != 0
    w = root_of_unity(cv, 1 << s)
    assert pow(w, 1 << (s - 1), cv["r"]) == cv["r"] - 1


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_group_law_properties(curve):
    c = get_curve(curve)
    rng = random.Random(1)
    a, b = rng.randrange(c.r), rng.randrange(c.r)
    P, Q = c.multiply(c.G1, a), c.multiply(c.G1, b)
    assert c.eq(c.add(P, Q), c.multiply(c.G1, (a + b) % c.r))
    assert c.is_inf(c.add(P, c.neg(P)))
    assert c.eq(c.add(P, c.Z1), P) and c.eq(c.add(c.Z1, P), P)
    assert c.eq(c.add(P, P), c.double(P))
    assert c.is_on_curve(c.add(P, Q))
    assert c.multiply(P, 0) == (1, 1, 0) and c.multiply(P, 1) is P


@pytest.mark.parametrize("curve", CURVE_NAMES)
@pytest.mark.parametrize("n", [1, 2, 4, 8, 64, 256])
def test_fft_matches_definition(curve, n):
    cv = CURVES[curve]; r = cv["r"]
    rng = random.Random(n)
    w = root_of_unity(cv, n)
    x = [rng.randrange(r) for _ in range(n)]
    y = off.fft_ff_int(x, w, r)
    assert y == off.dft_definition(x, w, r)
    assert off.ifft_ff_int(y, w, r) == x
    F = GFp(r)
    xe = [F(v) for v in x]
    assert [int(v) for v in off.fft_ff(xe, F(w), F)] == y                   # generic flavour == int flavour
    assert [int(v) for v in off.ifft_ff([F(v) for v in y], F(w), F)] == x
    assert off.coset_ifft_ff_int(off.coset_fft_ff_int(x, w, 7, r), w, 7, r) == x
    if n == 1:
        assert off.fft_ff(xe, F(w), F) is xe                                 # fft_ff.py:16-17 aliasing


def test_fft_interpolation_asserts():
    cv = CURVES["bn254"]; r = cv["r"]; F = GFp(r)
    g = F(root_of_unity(cv, 8))
    vals = [F(i * i + 1) for i in range(8)]
    coeffs = off.fft_ff_interpolation(vals, g, F)
    for i in range(8):
        assert poly_eval([int(c) for c in coeffs], pow(int(g), i, r), r) == int(vals[i])
    with pytest.raises(AssertionError, match="power of 2"):
        off.fft_ff_interpolation(vals[:6], g, F)
    with pytest.raises(AssertionError, match="must be at least"):
        off.fft_ff_interpolation(vals, F(root_of_unity(cv, 4)), F)


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_kzg_oracle_tau_identity_and_check(curve):
    c = get_curve(curve); k = KZGOracle(curve)
    rng = random.Random(9)
    tau = rng.randrange(1, c.r)
    ck = k.setup(6, tau)
    assert all(c.eq(a, b) for a, b in zip(ck, k.setup_fast(6, tau)))
    polys = [[rng.randrange(c.r) for _ in range(m)] for m in (7, 3, 1)]
    comms = k.commit(ck, polys)
    for p, C in zip(polys, comms):
        assert c.eq(C, c.multiply(c.G1, poly_eval(p, tau, c.r)))            # kzg.py:108
    assert c.is_inf(k.commit(ck, [[]])[0]) and c.is_inf(k.commit(ck, [[0, 0]])[0])
    with pytest.raises(ValueError, match="Polynomial degree 7 exceeds maximum allowed degree 6"):
        k.commit(ck, [[1] * 8])
    z, xi = rng.randrange(c.r), rng.randrange(c.r)
    proof = k.open(ck, polys, z, xi)
    evals = [poly_eval(p, z, c.r) for p in polys]
    assert k.check_with_tau(tau, comms, z, evals, proof, xi)
    evals[0] = (evals[0] + 1) % c.r                                          # tamper test, kzg.py:361-380
    assert not k.check_with_tau(tau, comms, z, evals, proof, xi)
    # quotient identity: W(X) * (X - z) + P(z) == P(X)
    P = k.combine(polys, xi)
    Wq = poly_div_linear(P, z, c.r)
    x0 = rng.randrange(c.r)
    assert (poly_eval(Wq, x0, c.r) * (x0 - z) + poly_eval(P, z, c.r) - poly_eval(P, x0, c.r)) % c.r == 0


@pytest.mark.parametrize("curve", CURVE_NAMES)
def test_oracle_reproduces_golden_vectors(curve):
    g = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))[curve]
    c = get_curve(curve); k = KZGOracle(curve); r = c.r
    tau = H(g["tau"])
    ck = k.setup(len(g["ck_affine"]) - 1, tau)
    assert [list(c.normalize(p)) for p in ck] == [[H(v) for v in p] for p in g["ck_affine"]]
    polys = [[H(v) for v in p] for p in g["polys"]]
    comm = [c.normalize(x) for x in k.commit(ck, polys)]
    assert comm == [None if v is None else tuple(H(t) for t in v) for v in g["commitments"]]
    o = g["open"]
    assert c.normalize(k.open(ck, polys[:o["k"]], H(o["z"]), H(o["xi"]))) == tuple(H(v) for v in o["proof"])
    t = g["ntt"]
    x, w = [H(v) for v in t["x"]], H(t["w"])
    assert off.fft_ff_int(x, w, r) == [H(v) for v in t["fft"]]
    assert off.ifft_ff_int(x, w, r) == [H(v) for v in t["ifft"]]
    assert off.coset_fft_ff_int(x, w, 7, r) == [H(v) for v in t["coset_fft"]]


def test_fixture_json_decodes():
    """The two reference fixtures (decoded Sage-free) have the documented shape and
    satisfy their own constraint systems."""
    r = CURVES["bn254"]["r"]
    p = {k: [H(x) for x in v] for k, v in json.load(open(os.path.join(GOLD, "plonk_instance.json"))).items()}
    n = 16
    assert all(len(p[k]) == n for k in ("qM", "qL", "qR", "qO", "qC"))
    assert sorted(p["perm"]) == list(range(3 * n)) and len(p["w"]) == 3 * n
    assert p["w"][:5] == [7, 11, 13, 17, 19]                                 # public inputs, main.py:79
    a, b, c = p["w"][:n], p["w"][n:2 * n], p["w"][2 * n:]
    for i in range(5, n):                                                    # gate equation off the public rows
        assert (p["qM"][i] * a[i] * b[i] + p["qL"][i] * a[i] + p["qR"][i] * b[i] + p["qO"][i] * c[i] + p["qC"][i]) % r == 0
    for i, j in enumerate(p["perm"]):                                        # copy constraints
        assert p["w"][i] == p["w"][j]
    d = json.load(open(os.path.join(GOLD, "r1cs_instance.json")))
    A, B, C = ([[H(x) for x in row] for row in d[k]] for k in "ABC")
    z = [H(x) for x in d["z"]]
    mv = lambda M: [sum(x * y for x, y in zip(row, z)) % r for row in M]
    assert all((x * y - w) % r == 0 for x, y, w in zip(mv(A), mv(B), mv(C)))
    assert [sum(1 for row in M for x in row if x) for M in (A, B, C)] == [20, 16, 16]


@pytest.mark.skipif(not os.path.isdir("/root/reference/constraint-system"), reason="reference checkout not mounted")
def test_fixture_json_matches_pickles():
    from oracle import fixtures
    p = fixtures.load_plonk_instance("/root/reference/constraint-system/PLONK_ARITHMETIZATION_INSTANCE.pkl")
    j = json.load(open(os.path.join(GOLD, "plonk_instance.json")))
    assert {k: [hex(x) for x in v] for k, v in p.items()} == j
