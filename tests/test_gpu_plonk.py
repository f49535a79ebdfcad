"""GPU: the device-resident PLONK prover (kzg_snark_b200/plonk.py) and the
polynomial kernels under it, through the C ABI.

  * bit-exact against the reference's own prover on the bundled instance
    (tests/golden/ref_plonk_normalized.json, produced by plonk/prover.py in the build container);
  * at sizes the fixture does not cover: accepted by the test-side restatement of the reference's
    verifier (pairing check), tampered proofs rejected -- the reference's own test strategy
    (plonk/verifier.py:272-290);
  * every new kernel against Python integers."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

from oracle.params import CURVES
from oracle.kzg import KZGOracle, poly_eval
from oracle.curve import get_curve
from trace_replay import H, load

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
R = CURVES["bn254"]["r"]


def aff(pt):
    x, y, z = (int(c) for c in pt)
    return None if z == 0 else (x, y)


@pytest.mark.parametrize("fixture", ["ref_plonk_normalized.json", "ref_plonk_normalized_bls.json"])
def test_prover_matches_the_reference_prover_bit_for_bit(fixture):
    """Both curves of kzg.py:26-35: the reference's plonk/{indexer,prover}.py were run with curve_type "bn254" and "bls12_381" on
    the bundled circuit (tests/golden/make_traces.py); key, index polynomials and the whole proof must be identical."""
    from kzg_snark_b200.plonk import Indexer, Prover
    d = load(fixture)
    curve = d["curve"]
    rq, r_bn = CURVES[curve]["r"], CURVES["bn254"]["r"]
    nl = CURVES[curve]["fp_limbs32"] // 2                                  # 64-bit limbs per base-field coordinate
    inst = json.load(open(os.path.join(GOLD, "plonk_instance.json")))
    signed = lambda v: (v if v <= r_bn // 2 else v - r_bn) % rq          # noqa: E731  (the pickle's residues as signed integers)
    sel = [[signed(H(v)) for v in inst[k]] for k in ("qM", "qL", "qR", "qO", "qC")]
    perm = [H(v) for v in inst["perm"]]
    n = d["n"]
    idx = Indexer(curve)
    ipk, ivk = idx.preprocess(*sel, perm, max_degree=n + 5, tau=H(d["index_draws"][0]), k1=H(d["k1"]), k2=H(d["k2"]))
    # the device-generated key is the reference's key
    key = ipk["ck"].read(0, n + 6)
    assert [tuple(int.from_bytes(row[nl * j:nl * j + nl].tobytes(), "little") for j in (0, 1)) for row in key] == \
        [(H(p[0]), H(p[1])) for p in d["keys"][0]]
    # index polynomials (plonk/encoder.py:99-141) and sigma_star
    for name, coeffs in d["index_polys"].items():
        got = ipk["polynomials"][name].read_ints()
        while got and got[-1] == 0:
            got.pop()
        assert got == [H(c) for c in coeffs], name
    assert ipk["sigma_star"].read_ints() == [H(s) for s in d["sigma_star"]]
    Fq = idx.kzg.Fq
    x = [Fq(H(v)) for v in d["x"]]
    w = [H(v) for v in d["w"]]
    prover = Prover(curve)
    proof = prover.prove(ipk, x, w, blinders=[H(b) for b in d["prover_draws"][-11:]])
    assert prover.last_r_zeta == 0 and not any(prover.last_t_top)
    exp = d["proof"]
    for k, v in exp["commitments"].items():
        assert aff(proof["commitments"][k]) == (H(v[0]), H(v[1])), k
    for k, v in exp["evaluations"].items():
        assert int(proof["evaluations"][k]) == H(v), k
    for k, v in exp["kzg_proofs"].items():
        assert aff(proof["kzg_proofs"][k]) == (H(v[0]), H(v[1])), k
    # the same with the commitment key handed in as a list of points, the way the reference's ipk carries it
    # (plonk/indexer.py:92-93), and witness values as field elements
    fq = idx.kzg._codec.fq
    ck = [(fq(H(p[0])), fq(H(p[1])), fq(1)) for p in d["keys"][0]]
    ipk2, _ = idx.preprocess(*sel, perm, max_degree=n + 5, ck=ck, k1=H(d["k1"]), k2=H(d["k2"]))
    proof2 = prover.prove(ipk2, x, [Fq(v) for v in w], blinders=[H(b) for b in d["prover_draws"][-11:]])
    assert {k: aff(v) for k, v in proof2["commitments"].items()} == {k: aff(v) for k, v in proof["commitments"].items()}
    assert {k: aff(v) for k, v in proof2["kzg_proofs"].items()} == {k: aff(v) for k, v in proof["kzg_proofs"].items()}


@pytest.mark.parametrize("logn,n_pub", [(3, 2), (6, 5), (10, 7), (14, 16), (20, 16)])
def test_synthetic_circuit_proof_is_accepted_and_tamper_rejected(logn, n_pub):
    import plonk_verifier
    from kzg_snark_b200.plonk import Indexer, Prover
    from kzg_snark_b200.plonk_synth import synthetic_circuit
    n = 1 << logn
    qM, qL, qR, qO, qC, perm, w = synthetic_circuit(n, n_pub, R, seed=logn)
    rng = random.Random(100 + logn)
    idx = Indexer("bn254")
    tau = random.Random(500 + logn).randrange(1, R)                      # the test's own trapdoor (never part of ivk)
    ipk, ivk = idx.preprocess(qM, qL, qR, qO, qC, perm, max_degree=n + 5, rng=rng, tau=tau)
    assert "tau" not in ivk
    x, wit = w[:n_pub], w[n_pub:]
    prover = Prover("bn254")
    proof = prover.prove(ipk, [idx.kzg.Fq(v) for v in x], wit)
    assert prover.last_r_zeta == 0 and not any(prover.last_t_top)
    vk = {"commitments": {k: aff(c) for k, c in ivk["commitments"].items()}, "n": n, "g": int(ivk["subgroups"]["g"]),
          "k1": int(ivk["subgroups"]["k1"]), "k2": int(ivk["subgroups"]["k2"]), "tau": tau}
    pf = {"commitments": {k: aff(c) for k, c in proof["commitments"].items()},
          "evaluations": {k: int(v) for k, v in proof["evaluations"].items()},
          "kzg_proofs": {k: aff(c) for k, c in proof["kzg_proofs"].items()}}
    assert plonk_verifier.verify(vk, x, pf)
    bad = {**pf, "evaluations": {**pf["evaluations"], "c": (pf["evaluations"]["c"] + 1) % R}}
    assert not plonk_verifier.verify(vk, x, bad)
    # a witness that violates one gate must not yield an accepting proof
    wbad = list(w)
    wbad[2 * n + n_pub + 1] = (wbad[2 * n + n_pub + 1] + 1) % R          # c value of a non-public gate
    proof2 = prover.prove(ipk, [idx.kzg.Fq(v) for v in x], wbad[n_pub:])
    pf2 = {"commitments": {k: aff(c) for k, c in proof2["commitments"].items()},
           "evaluations": {k: int(v) for k, v in proof2["evaluations"].items()},
           "kzg_proofs": {k: aff(c) for k, c in proof2["kzg_proofs"].items()}}
    assert not plonk_verifier.verify(vk, x, pf2)


@pytest.mark.parametrize("logn", [4, 11])
def test_bls12_381_prover_accepted_by_the_trapdoor_verifier(logn):
    """The second curve of KZG.__init__ (kzg.py:31-35): same prover, BLS12-381 scalar field and G1;
    checked with the trapdoor form of the reference's verifier equation."""
    import plonk_verifier
    from kzg_snark_b200.plonk import Indexer, Prover
    from kzg_snark_b200.plonk_synth import synthetic_circuit
    rq = CURVES["bls12_381"]["r"]
    n, n_pub = 1 << logn, 3
    qM, qL, qR, qO, qC, perm, w = synthetic_circuit(n, n_pub, rq, seed=40 + logn)
    idx = Indexer("bls12_381")
    tau = random.Random(600 + logn).randrange(1, rq)
    ipk, ivk = idx.preprocess(qM, qL, qR, qO, qC, perm, max_degree=n + 5, rng=random.Random(7 + logn), tau=tau)
    prover = Prover("bls12_381")
    proof = prover.prove(ipk, [idx.kzg.Fq(v) for v in w[:n_pub]], w[n_pub:])
    assert prover.last_r_zeta == 0 and not any(prover.last_t_top)
    vk = {"commitments": {k: aff(c) for k, c in ivk["commitments"].items()}, "n": n, "g": int(ivk["subgroups"]["g"]),
          "k1": int(ivk["subgroups"]["k1"]), "k2": int(ivk["subgroups"]["k2"]), "tau": tau}
    pf = {"commitments": {k: aff(c) for k, c in proof["commitments"].items()},
          "evaluations": {k: int(v) for k, v in proof["evaluations"].items()},
          "kzg_proofs": {k: aff(c) for k, c in proof["kzg_proofs"].items()}}
    assert plonk_verifier.verify_trapdoor(vk, w[:n_pub], pf, "bls12_381")
    bad = {**pf, "evaluations": {**pf["evaluations"], "z_omega": (pf["evaluations"]["z_omega"] + 1) % rq}}
    assert not plonk_verifier.verify_trapdoor(vk, w[:n_pub], bad, "bls12_381")


# ----------------------------------------------------------------------------- kernels
def _vec(ints):
    from kzg_snark_b200.plonk import DVec
    from kzg_snark_b200.limbs import ints_to_limbs
    return DVec.from_limbs(ints_to_limbs(ints, R))


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 4096, 4097, 300000])
def test_poly_eval_and_powers(n):
    from kzg_snark_b200.plonk import _Field, DVec
    from kzg_snark_b200 import _ffi
    f = _Field(_ffi.BN254)
    rng = random.Random(n)
    c = [rng.randrange(R) for _ in range(n)]
    x = rng.randrange(R)
    vc = _vec(c)
    assert f.eval(vc, n, x) == poly_eval(c, x, R)
    base, scale = rng.randrange(R), rng.randrange(R)
    out = DVec(n)
    f.powers(out, n, base, scale)
    got = out.read_ints()
    idxs = sorted({0, n - 1, n // 2, min(n - 1, 17), min(n - 1, 16), min(n - 1, 15)})
    for i in idxs:
        assert got[i] == scale * pow(base, i, R) % R
    if n <= 4097:
        acc = scale
        for i in range(n):
            assert got[i] == acc
            acc = acc * base % R


def test_lincomb_and_open_dev_match_host_paths():
    from kzg_snark_b200.plonk import _Field, DVec, _voidp_array
    from kzg_snark_b200 import _ffi, device
    from kzg_snark_b200.limbs import ints_to_limbs, int_to_limbs
    f = _Field(_ffi.BN254)
    rng = random.Random(5)
    lens = [100, 37, 1, 64, 0]
    polys = [[rng.randrange(R) for _ in range(m)] for m in lens]
    sc = [rng.randrange(R) for _ in lens]
    const = rng.randrange(R)
    vecs = [_vec(p) for p in polys]
    out = DVec(120)
    f.lincomb(out, 120, [(v.ptr, m, s) for v, m, s in zip(vecs, lens, sc)], constant=const)
    exp = [0] * 120
    exp[0] = const
    for p, s in zip(polys, sc):
        for i, cf in enumerate(p):
            exp[i] = (exp[i] + s * cf) % R
    assert out.read_ints() == exp
    # open_dev == open (host polys) == oracle
    tau = rng.randrange(R)
    srs = device.Srs.generate("bn254", tau, 128)
    z, xi = rng.randrange(R), rng.randrange(R)
    host, inf = device.open_proof(srs, [ints_to_limbs(p, R) for p in polys[:4]], int_to_limbs(z, R), int_to_limbs(xi, R))
    o = np.zeros(8, dtype=np.uint64)
    fl = ctypes.c_int(0)
    ls = (ctypes.c_size_t * 4)(*lens[:4])
    _ffi.check(f.lib.kzgpu_open_dev(srs.handle, _voidp_array([v.ptr for v in vecs[:4]]), ls, 4, _ffi.ptr(f.L(z)), _ffi.ptr(f.L(xi)),
                                    _ffi.ptr(o), ctypes.byref(fl), None))
    assert not inf and not fl.value and (o == host).all()
    ko = KZGOracle("bn254")
    cv = get_curve("bn254")
    w = ko.witness(polys[:4], z, xi)
    t = sum(cf * pow(tau, i, R) for i, cf in enumerate(w)) % R
    from kzg_snark_b200.limbs import limbs_to_ints
    assert tuple(limbs_to_ints(o.reshape(2, 4))) == cv.normalize(cv.multiply(cv.G1, t))


@pytest.mark.parametrize("n", [8, 64, 128, 8192])
def test_permutation_grand_product(n):
    from kzg_snark_b200.plonk import _Field, DVec
    from kzg_snark_b200 import _ffi
    f = _Field(_ffi.BN254)
    rng = random.Random(n)
    a, b, c = ([rng.randrange(R) for _ in range(n)] for _ in range(3))
    sig = [rng.randrange(R) for _ in range(3 * n)]
    g = f.root(n)
    Hs = [pow(g, i, R) for i in range(n)]
    k1, k2, beta, gamma = (rng.randrange(1, R) for _ in range(4))
    z = DVec(n)
    flag = ctypes.c_int(0)
    L = lambda v: _ffi.ptr(f.L(v))                                       # noqa: E731
    va, vb, vc, vs, vh = _vec(a), _vec(b), _vec(c), _vec(sig), _vec(Hs)     # keep the buffers alive across the call
    _ffi.check(f.lib.kzgpu_plonk_permutation_dev(_ffi.BN254, n, va.ptr, vb.ptr, vc.ptr, vs.ptr, vh.ptr,
                                                 L(k1), L(k2), L(beta), L(gamma), z.ptr, ctypes.byref(flag)))
    exp = [1]                                                             # plonk/prover.py:245-258
    for i in range(n - 1):
        num = (a[i] + beta * Hs[i] + gamma) * (b[i] + beta * k1 * Hs[i] + gamma) * (c[i] + beta * k2 * Hs[i] + gamma) % R
        den = (a[i] + beta * sig[i] + gamma) * (b[i] + beta * sig[i + n] + gamma) * (c[i] + beta * sig[i + 2 * n] + gamma) % R
        exp.append(exp[-1] * num % R * pow(den, -1, R) % R)
    assert not flag.value and z.read_ints() == exp
    # a zero denominator is reported (the reference raises ValueError, plonk/prover.py:254-255)
    a2 = list(a)
    a2[3] = (-(beta * sig[3] + gamma)) % R
    va2 = _vec(a2)
    _ffi.check(f.lib.kzgpu_plonk_permutation_dev(_ffi.BN254, n, va2.ptr, vb.ptr, vc.ptr, vs.ptr, vh.ptr,
                                                 L(k1), L(k2), L(beta), L(gamma), z.ptr, ctypes.byref(flag)))
    assert flag.value == 1


def test_marlin_loops_match_the_reference_functions():
    """`_compute_t_polynomial` and `_compute_f2_polynomial` of marlin/prover.py, run by the reference
    itself on the bundled R1CS index (tests/golden/ref_marlin_loops.json), against the batched device versions."""
    from kzg_snark_b200 import marlin
    from kzg_snark_b200.limbs import ints_to_limbs, limbs_to_ints
    d = load("ref_marlin_loops.json")
    n, m = d["n"], d["m"]
    cat = lambda kind: ints_to_limbs([H(v) for M in "ABC" for v in d["evals"][f"{kind}_{M}"]], R)    # noqa: E731
    eta = [H(v) for v in d["eta"]]
    alpha, beta1 = H(d["alpha"]), H(d["beta1"])
    g_H, g_K = pow(5, (R - 1) // n, R), H(d["g_K"])
    ridx = [i for M in "ABC" for i in d["row_index"][M]]

    def strip(c):
        c = list(c)
        while c and c[-1] == 0:
            c.pop()
        return c

    f2 = limbs_to_ints(marlin.compute_f2_polynomial("bn254", cat("row"), cat("col"), cat("val"), eta, alpha, beta1, n, g_K))
    assert strip(f2) == [H(v) for v in d["f2"]]
    t = limbs_to_ints(marlin.compute_t_polynomial("bn254", ridx, cat("col"), cat("val"), eta, alpha, n, g_H))
    assert strip(t) == [H(v) for v in d["t"]]
    # a zero denominator (alpha equal to a column value) drops that term, as `if denom != 0` does in the reference
    col0 = H(d["evals"]["col_A"][0])
    f2b = limbs_to_ints(marlin.compute_f2_polynomial("bn254", cat("row"), cat("col"), cat("val"), eta, col0, beta1, n, g_K))
    ev = {k: [H(v) for v in vals] for k, vals in d["evals"].items()}
    scale = (pow(beta1, n, R) - 1) * (pow(col0, n, R) - 1) % R
    exp = []
    for k in range(m):
        acc = 0
        for e, M in zip(eta, "ABC"):
            den = (beta1 - ev[f"row_{M}"][k]) * (col0 - ev[f"col_{M}"][k]) % R
            if den:
                acc += e * scale % R * ev[f"val_{M}"][k] % R * pow(den, -1, R)
        exp.append(acc % R)
    from oracle.fft_ff import ifft_ff_int
    assert f2b == ifft_ff_int(exp, g_K, R)


def test_product_verifier_with_the_oracle_pairing_in_py_eccs_role():
    """kzg_snark_b200.plonk.Verifier (plonk/verifier.py's interface; G1 combinations on the device) accepts the
    device prover's proof and rejects tampered ones; the two pairings -- py_ecc's in the reference, absent here --
    come from the oracle's stand-in, the role py_ecc plays for the drop-in."""
    from kzg_snark_b200.plonk import Indexer, Prover, Verifier
    from kzg_snark_b200.plonk_synth import synthetic_circuit
    from oracle import pyecc_standin as E
    n, n_pub = 1 << 9, 4
    qM, qL, qR, qO, qC, perm, w = synthetic_circuit(n, n_pub, R, seed=9)
    idx = Indexer("bn254")
    tau = random.Random(32).randrange(1, R)
    ipk, ivk = idx.preprocess(qM, qL, qR, qO, qC, perm, max_degree=n + 5, rng=random.Random(31), tau=tau)
    x = [idx.kzg.Fq(v) for v in w[:n_pub]]
    proof = Prover("bn254").prove(ipk, x, w[n_pub:])
    ver = Verifier("bn254")
    if ver.kzg.have_py_ecc:
        pytest.skip("py_ecc present: the verifier already uses it")
    ver.kzg.G2 = E.G2
    ver.kzg.pairing = lambda Q, P: E.pairing(Q, tuple(E.FQ(int(c)) for c in P))
    ivk = {**ivk, "rk": E.multiply(E.G2, tau)}
    assert ver.verify(ivk, x, proof)
    bad = {**proof, "evaluations": {**proof["evaluations"], "b": proof["evaluations"]["b"] + 1}}
    assert not ver.verify(ivk, x, bad)
    assert not ver.verify(ivk, [x[0] + 1] + x[1:], proof)
    swapped = {**proof, "kzg_proofs": {"W_z": proof["kzg_proofs"]["W_zw"], "W_zw": proof["kzg_proofs"]["W_z"]}}
    assert not ver.verify(ivk, x, swapped)


def test_marlin_indexer_matches_the_reference_indexer():
    """marlin/indexer.py:Indexer.preprocess on the bundled R1CS instance: the nine index polynomials and their
    commitments (ONE batched MSM pass here) against what the reference's indexer computed (the indexer's commit
    record in tests/golden/ref_trace_marlin.json); then the two prover loops on the indexer's resident data."""
    from kzg_snark_b200 import marlin
    tr = load("ref_trace_marlin.json")
    inst = json.load(open(os.path.join(GOLD, "r1cs_instance.json")))
    A, B, C = ([[H(v) for v in row] for row in inst[k]] for k in "ABC")
    rec = tr["calls"][tr["notes"]["index_calls"] - 1]
    assert rec["fn"] == "commit" and len(rec["polys"]) == 9
    idx = marlin.Indexer("bn254")
    fq = idx.kzg._codec.fq
    ck = [(fq(H(p[0])), fq(H(p[1])), fq(1)) for p in tr["keys"][rec["ck_id"]]]
    ipk, ivk = idx.preprocess(A, B, C, max_degree=200, ck=ck)
    names = ipk["polynomials"]["names"]
    for nm, exp_poly, exp_pt in zip(names, rec["polys"], rec["out"]):
        got = marlin.Indexer.polynomial(ipk, nm)
        while got and got[-1] == 0:
            got.pop()
        assert got == [H(c) for c in exp_poly], nm
        assert aff(ipk["commitments"][nm]) == (H(exp_pt[0]), H(exp_pt[1])), nm
    # a sparse description of the same matrices gives the same index
    sp = lambda M: {"shape": (len(M), len(M[0])), "entries": [(i, j, v) for i, r_ in enumerate(M) for j, v in enumerate(r_) if v]}   # noqa: E731
    ipk2, _ = idx.preprocess(sp(A), sp(B), sp(C), max_degree=200, ck=ck)
    assert {k: aff(v) for k, v in ipk2["commitments"].items()} == {k: aff(v) for k, v in ipk["commitments"].items()}
    # the loop fixtures were produced on the same index: its resident evaluations reproduce them
    d = load("ref_marlin_loops.json")
    for kind in ("row", "col", "val"):
        assert ipk["evals"][kind].read_ints() == [H(v) for M in "ABC" for v in d["evals"][f"{kind}_{M}"]]


@pytest.mark.parametrize("fixture", ["ref_marlin_normalized.json", "ref_marlin_normalized_bls.json"])
def test_marlin_prover_matches_the_reference_prover_bit_for_bit(fixture):
    """marlin/prover.py on the bundled R1CS instance (configs[4]), run by the reference itself with commitments normalised
    to (x, y, 1) before the transcript (tests/golden/ref_marlin_normalized*.json, curve_type "bn254" and "bls12_381"): the
    device prover, fed the same SRS and the same 41 random draws, must produce the same polynomials in every round and the
    same proof."""
    from kzg_snark_b200 import marlin
    d = load(fixture)
    curve = d["curve"]
    rq, r_bn = CURVES[curve]["r"], CURVES["bn254"]["r"]
    signed = lambda v: (v if v <= r_bn // 2 else v - r_bn) % rq          # noqa: E731  (the pickle's residues as signed integers)
    inst = json.load(open(os.path.join(GOLD, "r1cs_instance.json")))
    A, B, C = ([[signed(H(v)) for v in row] for row in inst[k]] for k in "ABC")
    idx = marlin.Indexer(curve)
    ipk, _ = idx.preprocess(A, B, C, max_degree=200, tau=H(d["index_draws"][0]))
    Fq = idx.kzg.Fq
    prover = marlin.Prover(curve)
    prover.capture = True
    proof = prover.prove(ipk, [Fq(H(v)) for v in d["x"]], [H(v) for v in d["w"]], draws=[H(v) for v in d["prover_draws"]])
    calls = [c for c in d["prover_calls"] if c["fn"] in ("commit", "open")]
    expected = dict(zip(["w_masked", "zA", "zB", "zC", "h_0", "s"], calls[0]["polys"]))
    expected.update(zip(["t", "g_1", "h_1"], calls[1]["polys"]))
    expected.update(zip(["g_2", "h_2"], calls[2]["polys"]))
    expected.update(zip(["f_1", "f_2"], calls[3]["polys"][:2]))
    expected["f_3"] = calls[4]["polys"][0]
    for nm, exp in expected.items():
        got = list(prover.captured[nm])
        while got and got[-1] == 0:
            got.pop()
        assert got == [H(c) for c in exp], nm
    assert set(prover.checks.values()) == {0}
    exp = d["proof"]
    for rnd, pts in exp["commitments"].items():
        assert [aff(p) for p in proof["commitments"][rnd]] == [(H(p[0]), H(p[1])) for p in pts], rnd
    for k, vals in exp["evaluations"].items():
        assert [int(v) for v in proof["evaluations"][k]] == [H(v) for v in vals], k
    for k, p in exp["kzg_proofs"].items():
        assert aff(proof["kzg_proofs"][k]) == (H(p[0]), H(p[1])), k


@pytest.mark.parametrize("curve,logn", [("bn254", 6), ("bn254", 11), ("bls12_381", 8)])
def test_marlin_prover_synthetic_r1cs_identities(curve, logn):
    """Sizes and the second curve the bundled instance does not cover: on a random satisfied R1CS the prover's three
    linearisation polynomials must vanish at their challenge points -- f_1(beta_1) = f_2(beta_1) = f_3(beta_2) = 0 are the
    polynomial identities the Marlin verifier checks (marlin/prover.py:198-200 asserts the same) -- the first sumcheck's
    remainder has no constant term, and an unsatisfied instance must break them."""
    from kzg_snark_b200 import marlin
    rq = CURVES[curve]["r"]
    n = 1 << logn
    A, B, C, x, w = marlin.synthetic_r1cs(n, 5, rq, seed=logn)
    idx = marlin.Indexer(curve)
    m_need = 1 << (2 * n - 1).bit_length()                     # nnz(A) = 2n -> |K| = 2n
    ipk, _ = idx.preprocess(A, B, C, max_degree=6 * m_need, rng=random.Random(logn))
    assert ipk["subgroups"]["m"] == m_need
    prover = marlin.Prover(curve)
    proof = prover.prove(ipk, [idx.kzg.Fq(v) for v in x], w)
    assert set(prover.checks.values()) == {0}
    assert len(proof["commitments"]["first_round"]) == 6 and all(int(p[2]) == 1 for p in proof["commitments"]["first_round"])
    bad = list(w)
    bad[7] = (bad[7] + 1) % rq
    with pytest.raises(AssertionError):                        # "Sum over H is not 0" (marlin/prover.py:134) or a failed identity
        prover.prove(ipk, [idx.kzg.Fq(v) for v in x], bad)
        assert set(prover.checks.values()) == {0}
