#!/usr/bin/env python3
"""Generates tests/golden/ref_trace_*.json by running the REFERENCE'S OWN code here.

Run in the BUILD container (where /root/reference is mounted):
    python tests/golden/make_traces.py

The reference's kzg.py, fft_ff.py, transcript.py, plonk/*.py and marlin/*.py are imported
unmodified from /root/reference by oracle/refrun.py (SageMath and py_ecc replaced by the
stand-ins described there) and every outermost call that crosses the hot-path boundary
(KZG.commit, KZG.open, fft_ff, ifft_ff, fft_ff_interpolation) is recorded with its inputs and the
reference's outputs (points normalised to affine):

  ref_trace_kzg.json     configs[0]: setup(2^10) -> commit -> open -> check of a random degree-2^10
                         polynomial, the main.py:17-36 demo, and kzg.py's edge behaviour (zero
                         polynomial, zero coefficients, list inputs, multi-poly open, degree
                         overflow message)
  ref_trace_kzg_bls.json the same KZG demo and edge inputs through KZG(curve_type="bls12_381") (kzg.py:32-35), degree 2^8
  ref_trace_fft.json     fft_ff / ifft_ff / fft_ff_interpolation from fft_ff.py on n = 1 .. 2^8 (fft_ff also 2^10)
  ref_trace_fft_bls.json the same over the BLS12-381 scalar field
  ref_trace_plonk.json   main.py:64-94: Indexer.preprocess, Prover.prove, Verifier.verify (accepts)
  ref_trace_marlin.json  main.py:39-61 likewise
  ref_marlin_loops.json  the reference's `_compute_t_polynomial` / `_compute_f2_polynomial` (marlin/prover.py:248-301,
                         404-470) on the bundled index with seeded challenges: inputs and outputs (for the N4 kernels)
  ref_plonk_normalized.json
                         the reference PLONK prover run once more with commitments NORMALISED to
                         (x, y, 1) before they reach the transcript -- what any drop-in returning
                         canonical affine points produces -- with the blinding scalars, challenges'
                         inputs and the full proof, for the bit-exact check of kzg_snark_b200.plonk

  ref_plonk_normalized_bls.json, ref_marlin_normalized_bls.json
                         the same two with curve_type="bls12_381" (kzg.py:32-35) on the bundled instances, residues read as signed integers

All fixtures are deterministic (seeded) and small (< 1 MB together).
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import fixtures, refrun                 # noqa: E402
from kzg_snark_b200 import sageshim                 # noqa: E402

REF_CS = "/root/reference/constraint-system"
SEED = 20261018


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(obj.get('calls', []))} calls")


def trace_kzg():
    with refrun.ReferenceRun(seed=SEED) as rr:
        kzg = rr.kzg.KZG(curve_type="bn254")
        Fq, R, X = kzg.Fq, kzg.R, kzg.X
        notes = {}
        # main.py:17-36
        ck, rk = kzg.setup(max_degree=10)
        polys = [1 + 2 * X + 3 * X**2, 4 + 5 * X**3]
        comm = kzg.commit(ck, polys)
        proof = kzg.open(ck, polys, 7, 42)
        notes["demo_check"] = bool(kzg.check(rk, comm, 7, [p(7) for p in polys], proof, 42))
        notes["demo_check_wrong_eval"] = bool(kzg.check(rk, comm, 7, [polys[0](7) + 1, polys[1](7)], proof, 42))
        # edge behaviour on the same key
        edge = [R(0), R([0, 0, 5, 0, Fq(-1)]), R([Fq.random_element() for _ in range(11)]), R(3)]
        kzg.commit(ck, edge)
        kzg.commit(ck, [[1, 2, 3], [0, 0, 0, 7]])                       # coefficient lists (kzg.py:94-95)
        kzg.commit(ck, [])
        kzg.open(ck, edge, Fq.random_element(), Fq.random_element())
        kzg.open(ck, [edge[3]], 5, 9)                                   # constant polynomial -> quotient 0 -> Z1
        try:
            kzg.commit(ck, [X**11])
            notes["overflow"] = None
        except ValueError as e:
            notes["overflow"] = str(e)
        notes["batch_check"] = bool(kzg.batch_check(
            rk, [comm, comm[:1]], [7, 3], [[p(7) for p in polys], [polys[0](3)]],
            [proof, kzg.open(ck, polys[:1], 3, 11)], [42, 11]))
        # configs[0]: degree 2^10
        t0 = time.time()
        ck2, rk2 = kzg.setup(max_degree=1 << 10)
        big = R([Fq.random_element() for _ in range((1 << 10) + 1)])
        z, xi = Fq.random_element(), Fq.random_element()
        c2 = kzg.commit(ck2, [big])
        pr2 = kzg.open(ck2, [big], z, xi)
        notes["config0_check"] = bool(kzg.check(rk2, c2, z, [big(z)], pr2, xi))
        notes["config0_seconds"] = round(time.time() - t0, 2)
        assert notes["demo_check"] and notes["config0_check"] and notes["batch_check"] and not notes["demo_check_wrong_eval"]
        dump("ref_trace_kzg.json", {"source": "reference kzg.py run by oracle/refrun.py", "seed": SEED, "curve": "bn254",
                                    "notes": notes, "keys": rr.keys, "calls": rr.trace})


def trace_kzg_bls():
    """The same demo on the reference's second curve (kzg.py:32-35): KZG("bls12_381") setup -> commit -> open -> check
    (accept / reject), edge inputs, batch_check, and a degree-2^8 polynomial."""
    with refrun.ReferenceRun(seed=SEED + 7) as rr:
        kzg = rr.kzg.KZG(curve_type="bls12_381")
        Fq, R, X = kzg.Fq, kzg.R, kzg.X
        notes = {}
        ck, rk = kzg.setup(max_degree=10)
        polys = [1 + 2 * X + 3 * X**2, 4 + 5 * X**3]
        comm = kzg.commit(ck, polys)
        proof = kzg.open(ck, polys, 7, 42)
        notes["demo_check"] = bool(kzg.check(rk, comm, 7, [p(7) for p in polys], proof, 42))
        notes["demo_check_wrong_eval"] = bool(kzg.check(rk, comm, 7, [polys[0](7) + 1, polys[1](7)], proof, 42))
        edge = [R(0), R([0, 0, 5, 0, Fq(-1)]), R([Fq.random_element() for _ in range(11)]), R(3)]
        kzg.commit(ck, edge)
        kzg.commit(ck, [[1, 2, 3], [0, 0, 0, 7]])
        kzg.open(ck, edge, Fq.random_element(), Fq.random_element())
        kzg.open(ck, [edge[3]], 5, 9)
        notes["batch_check"] = bool(kzg.batch_check(
            rk, [comm, comm[:1]], [7, 3], [[p(7) for p in polys], [polys[0](3)]],
            [proof, kzg.open(ck, polys[:1], 3, 11)], [42, 11]))
        ck2, rk2 = kzg.setup(max_degree=1 << 8)
        big = R([Fq.random_element() for _ in range((1 << 8) + 1)])
        z, xi = Fq.random_element(), Fq.random_element()
        c2 = kzg.commit(ck2, [big])
        pr2 = kzg.open(ck2, [big], z, xi)
        notes["deg256_check"] = bool(kzg.check(rk2, c2, z, [big(z)], pr2, xi))
        assert notes["demo_check"] and notes["deg256_check"] and notes["batch_check"] and not notes["demo_check_wrong_eval"]
        dump("ref_trace_kzg_bls.json", {"source": "reference kzg.py (curve_type='bls12_381') run by oracle/refrun.py", "seed": SEED + 7,
                                        "curve": "bls12_381", "notes": notes, "keys": rr.keys, "calls": rr.trace})


def trace_fft():
    with refrun.ReferenceRun(seed=SEED + 1) as rr:
        kzg = rr.kzg.KZG(curve_type="bn254")
        Fq = kzg.Fq
        q = Fq.order()
        notes = {}
        for logn in (0, 1, 2, 3, 4, 5, 6, 7, 8, 10):
            n = 1 << logn
            w = Fq(5) ** ((q - 1) // n)
            vals = [Fq.random_element() for _ in range(n)]
            rr.fft_ff.fft_ff(vals, w, Fq)
            if logn > 8:
                continue
            rr.fft_ff.ifft_ff(vals, w, Fq)
            if n >= 2:
                sparse = [Fq(0)] * n
                sparse[1] = Fq(1)
                rr.fft_ff.fft_ff(sparse, w, Fq)                          # delta_1 -> [w^k]
                rr.fft_ff.fft_ff_interpolation(vals, w, Fq)
        one = [Fq(7)]
        notes["n1_returns_same_object"] = rr.fft_ff.fft_ff(one, Fq(1), Fq) is one       # fft_ff.py:16-17
        for bad, key in (([Fq(1)] * 3, "assert_not_pow2"), ([Fq(1)] * 8, "assert_short_order")):
            try:
                rr.fft_ff.fft_ff_interpolation(bad, Fq(5) ** ((q - 1) // 4), Fq)
                notes[key] = None
            except AssertionError as e:
                notes[key] = str(e)
        dump("ref_trace_fft.json", {"source": "reference fft_ff.py run by oracle/refrun.py", "seed": SEED + 1,
                                    "field": "bn254_r", "notes": notes, "calls": rr.trace})


def trace_fft_bls():
    """fft_ff.py over the BLS12-381 scalar field (the field KZG("bls12_381") hands to its callers): n = 1 .. 2^8, and 2^10."""
    with refrun.ReferenceRun(seed=SEED + 8) as rr:
        kzg = rr.kzg.KZG(curve_type="bls12_381")
        Fq = kzg.Fq
        q = Fq.order()
        for logn in (0, 1, 2, 3, 4, 5, 6, 7, 8, 10):
            n = 1 << logn
            w = Fq(7) ** ((q - 1) // n)                               # 7 generates the multiplicative group of r_bls
            vals = [Fq.random_element() for _ in range(n)]
            rr.fft_ff.fft_ff(vals, w, Fq)
            if logn > 8:
                continue
            rr.fft_ff.ifft_ff(vals, w, Fq)
            if n >= 2:
                rr.fft_ff.fft_ff_interpolation(vals, w, Fq)
        dump("ref_trace_fft_bls.json", {"source": "reference fft_ff.py over GF(r_bls12_381) run by oracle/refrun.py", "seed": SEED + 8,
                                        "curve": "bls12_381", "field": "bls12_381_r", "calls": rr.trace})


def _plonk_inputs(Fq):
    inst = fixtures.load_plonk_instance(os.path.join(REF_CS, "PLONK_ARITHMETIZATION_INSTANCE.pkl"))
    r0 = inst["modulus"] if "modulus" in inst else 21888242871839275222246405745257275088548364400416034343698204186575808495617
    # the pickle holds residues mod r_bn254 (small values and -1); read them as signed integers so that the same circuit
    # is also satisfied over the BLS12-381 scalar field (identity for Fq = GF(r_bn254))
    f = lambda v: [Fq(x if x <= r0 // 2 else x - r0) for x in v]         # noqa: E731
    w = f(inst["w"])
    return [f(inst[k]) for k in ("qM", "qL", "qR", "qO", "qC")], list(inst["perm"]), w[:5], w[5:]


def enc_proof(rr, proof):
    out = {}
    for sec, d in proof.items():
        out[sec] = {k: (rr.enc_point(v) if isinstance(v, tuple) else rr.enc_scalar(v)) for k, v in d.items()}
    return out


def trace_plonk():
    with refrun.ReferenceRun(seed=SEED + 2) as rr:
        Fq = rr.kzg.KZG("bn254").Fq
        sel, perm, x, wit = _plonk_inputs(Fq)
        n = len(sel[0])
        t0 = time.time()
        ipk, ivk = rr.load("plonk.indexer").Indexer(curve_type="bn254").preprocess(*sel, perm, max_degree=n + 5)
        t1 = time.time()
        index_calls = len(rr.trace)
        proof = rr.load("plonk.prover").Prover(curve_type="bn254").prove(ipk, x, wit)
        t2 = time.time()
        V = rr.load("plonk.verifier").Verifier
        ok = V(curve_type="bn254").verify(ivk, x, proof)
        bad = {**proof, "evaluations": {**proof["evaluations"], "a": proof["evaluations"]["a"] + 1}}
        rejected = not V(curve_type="bn254").verify(ivk, x, bad)
        assert ok and rejected
        dump("ref_trace_plonk.json", {
            "source": "reference plonk/{indexer,prover,verifier}.py + kzg.py + fft_ff.py run by oracle/refrun.py",
            "seed": SEED + 2, "curve": "bn254",
            "notes": {"verify": bool(ok), "tampered_rejected": bool(rejected), "index_calls": index_calls,
                      "index_seconds": round(t1 - t0, 3), "prove_seconds": round(t2 - t1, 3)},
            "keys": rr.keys, "calls": rr.trace, "proof": enc_proof(rr, proof)})


def trace_marlin():
    with refrun.ReferenceRun(seed=SEED + 3) as rr:
        Fq = rr.kzg.KZG("bn254").Fq
        inst = fixtures.load_r1cs_instance(os.path.join(REF_CS, "R1CS_INSTANCE.pkl"))
        A, B, C = (sageshim.matrix(Fq, inst[k]) for k in "ABC")
        z = [Fq(v) for v in inst["z"]]
        x, w = z[:5], z[5:]
        t0 = time.time()
        ipk, ivk = rr.load("marlin.indexer").Indexer(curve_type="bn254").preprocess(A, B, C, max_degree=200)
        t1 = time.time()
        index_calls = len(rr.trace)
        proof = rr.load("marlin.prover").Prover(curve_type="bn254").prove(ipk, x, w)
        t2 = time.time()
        ok = rr.load("marlin.verifier").Verifier(curve_type="bn254").verify(ivk, x, proof)
        assert ok
        dump("ref_trace_marlin.json", {
            "source": "reference marlin/{indexer,prover,verifier}.py + kzg.py + fft_ff.py run by oracle/refrun.py",
            "seed": SEED + 3, "curve": "bn254",
            "notes": {"verify": bool(ok), "index_calls": index_calls, "index_seconds": round(t1 - t0, 3),
                      "prove_seconds": round(t2 - t1, 3)},
            "keys": rr.keys, "calls": rr.trace})


def marlin_loops():
    """The two evaluation loops of marlin/prover.py that SURVEY.md 8f (N4) names -- `_compute_t_polynomial`
    (:248-301) and `_compute_f2_polynomial` (:404-470) -- run by the reference itself on the bundled R1CS index
    with seeded challenges: inputs (K-domain row / col / val evaluations of the star matrices) and outputs."""
    import random
    with refrun.ReferenceRun(seed=SEED + 5, record=False) as rr:
        Fq = rr.kzg.KZG("bn254").Fq
        inst = fixtures.load_r1cs_instance(os.path.join(REF_CS, "R1CS_INSTANCE.pkl"))
        A, B, C = (sageshim.matrix(Fq, inst[k]) for k in "ABC")
        ipk, _ = rr.load("marlin.indexer").Indexer(curve_type="bn254").preprocess(A, B, C, max_degree=200)
        prover = rr.load("marlin.prover").Prover(curve_type="bn254")
        sub, polys = ipk["subgroups"], ipk["polynomials"]
        H, K, n, m, g_K = sub["H"], sub["K"], sub["n"], sub["m"], sub["g_K"]
        v_H, v_K = ipk["vanishing_polys"]["v_H"], ipk["vanishing_polys"]["v_K"]
        rng = random.Random(SEED + 5)
        q = Fq.order()
        eta = [Fq(rng.randrange(q)) for _ in range(3)]
        alpha, beta1 = Fq(rng.randrange(q)), Fq(rng.randrange(q))
        assert alpha not in H and beta1 not in H
        R = rr.kzg.KZG("bn254").R
        t = prover._compute_t_polynomial(polys, *eta, alpha, v_H, K, R)
        f2 = prover._compute_f2_polynomial(polys, *eta, beta1, alpha, v_H, v_K, K, g_K, Fq, R)
        hidx = {int(h): i for i, h in enumerate(H)}
        ev = {k: [int(polys[k](kap)) for kap in K] for k in polys}
        dump("ref_marlin_loops.json", {
            "source": "reference marlin/prover.py _compute_t_polynomial / _compute_f2_polynomial on the bundled R1CS index",
            "seed": SEED + 5, "n": n, "m": m, "g_H": rr.enc_scalar(sub["g_H"] if "g_H" in sub else H[1]), "g_K": rr.enc_scalar(g_K),
            "eta": [rr.enc_scalar(e) for e in eta], "alpha": rr.enc_scalar(alpha), "beta1": rr.enc_scalar(beta1),
            "evals": {k: [hex(v) for v in vals] for k, vals in ev.items()},
            "row_index": {M: [hidx.get(v, -1) for v in ev[f"row_{M}"]] for M in "ABC"},
            "t": rr.enc_poly(t), "f2": rr.enc_poly(f2)})


def trace_marlin_normalized(curve="bn254"):
    """Reference Marlin indexer + prover + verifier with commitments normalised to (x, y, 1) before the transcript (what a
    canonical-affine drop-in returns): random draws, the polynomials of every commit / open call and the proof -- the
    fixture for a device Marlin prover."""
    from oracle import pyecc_standin, pyecc_standin_bls
    E = pyecc_standin if curve == "bn254" else pyecc_standin_bls
    r_bn = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    sg = lambda v: v if v <= r_bn // 2 else v - r_bn               # noqa: E731  (the pickle's residues as signed integers)
    with refrun.ReferenceRun(seed=SEED + 6) as rr:
        KZG = rr.kzg.KZG

        def norm(pt):
            if E.is_inf(pt):
                return E.Z1
            x, y = E.normalize(pt)
            return (E.FQ(x.n), E.FQ(y.n), E.FQ(1))

        traced_commit, traced_open = KZG.commit, KZG.open
        KZG.commit = lambda s, ck, polys: [norm(c) for c in traced_commit(s, ck, polys)]
        KZG.open = lambda s, ck, polys, z, xi: norm(traced_open(s, ck, polys, z, xi))
        draws = []
        orig_rand = sageshim.GFShim.random_element

        def rand(self):
            v = orig_rand(self)
            draws.append(hex(int(v)))
            return v

        sageshim.GFShim.random_element = rand
        try:
            Fq = KZG(curve).Fq
            inst = fixtures.load_r1cs_instance(os.path.join(REF_CS, "R1CS_INSTANCE.pkl"))
            A, B, C = (sageshim.matrix(Fq, [[sg(v) for v in row] for row in inst[k]]) for k in "ABC")
            z = [Fq(sg(v)) for v in inst["z"]]
            x, w = z[:5], z[5:]
            ipk, ivk = rr.load("marlin.indexer").Indexer(curve_type=curve).preprocess(A, B, C, max_degree=200)
            n_index_draws, n_index_calls = len(draws), len(rr.trace)
            proof = rr.load("marlin.prover").Prover(curve_type=curve).prove(ipk, x, w)
            n_prove_draws = len(draws)
            ok = rr.load("marlin.verifier").Verifier(curve_type=curve).verify(ivk, x, proof)
        finally:
            sageshim.GFShim.random_element = orig_rand
        assert ok
        pr = {"commitments": {k: [rr.enc_point(c) for c in v] for k, v in proof["commitments"].items()},
              "evaluations": {k: [rr.enc_scalar(e) for e in v] for k, v in proof["evaluations"].items()},
              "kzg_proofs": {k: rr.enc_point(v) for k, v in proof["kzg_proofs"].items()}}
        dump("ref_marlin_normalized.json" if curve == "bn254" else "ref_marlin_normalized_bls.json", {
            "source": "reference marlin indexer + prover with commitments normalised to (x,y,1) before the transcript",
            "seed": SEED + 6, "curve": curve, "index_draws": draws[:n_index_draws],
            "prover_draws": draws[n_index_draws:n_prove_draws], "keys": rr.keys,
            "x": [rr.enc_scalar(v) for v in x], "w": [rr.enc_scalar(v) for v in w],
            "prover_calls": [{k: v for k, v in c.items() if k != "out" or c["fn"] != "x"} for c in rr.trace[n_index_calls:]],
            "proof": pr, "notes": {"verify": bool(ok)}})


def trace_marlin_normalized_bls():
    """The reference's Marlin indexer / prover / verifier with curve_type="bls12_381" on the bundled R1CS instance (residues read as
    signed integers): the bit-exact target of kzg_snark_b200.marlin on the second curve."""
    trace_marlin_normalized("bls12_381")


def trace_plonk_normalized(curve="bn254"):
    """Reference prover + indexer with KZG.commit / KZG.open outputs normalised to (x, y, 1): the
    transcript then hashes what a canonical-affine drop-in returns, so kzg_snark_b200.plonk can be
    compared bit for bit (commitments, evaluations, opening proofs)."""
    from oracle import pyecc_standin, pyecc_standin_bls
    E = pyecc_standin if curve == "bn254" else pyecc_standin_bls
    with refrun.ReferenceRun(seed=SEED + 4) as rr:
        KZG = rr.kzg.KZG

        def norm(pt):
            if E.is_inf(pt):
                return E.Z1
            x, y = E.normalize(pt)
            return (E.FQ(x.n), E.FQ(y.n), E.FQ(1))

        traced_commit, traced_open = KZG.commit, KZG.open
        KZG.commit = lambda s, ck, polys: [norm(c) for c in traced_commit(s, ck, polys)]
        KZG.open = lambda s, ck, polys, z, xi: norm(traced_open(s, ck, polys, z, xi))
        draws = []
        orig_rand = sageshim.GFShim.random_element

        def rand(self):
            v = orig_rand(self)
            draws.append(hex(int(v)))
            return v

        sageshim.GFShim.random_element = rand
        try:
            Fq = KZG(curve).Fq
            sel, perm, x, wit = _plonk_inputs(Fq)
            n = len(sel[0])
            ipk, ivk = rr.load("plonk.indexer").Indexer(curve_type=curve).preprocess(*sel, perm, max_degree=n + 5)
            n_index_draws = len(draws)
            proof = rr.load("plonk.prover").Prover(curve_type=curve).prove(ipk, x, wit)
            n_prove_draws = len(draws)
            ok = rr.load("plonk.verifier").Verifier(curve_type=curve).verify(ivk, x, proof)
        finally:
            sageshim.GFShim.random_element = orig_rand
        assert ok
        sub = ipk["subgroups"]
        dump("ref_plonk_normalized.json" if curve == "bn254" else "ref_plonk_normalized_bls.json", {
            "source": "reference plonk prover with commitments normalised to (x,y,1) before the transcript",
            "seed": SEED + 4, "curve": curve, "n": n,
            "index_draws": draws[:n_index_draws],                 # tau, k1, k2 ... (kzg.py:67, plonk/encoder.py:83-84)
            # the prover's Encoder.update_state draws a throw-away k1, k2 first (plonk/prover.py:63,
            # plonk/encoder.py:80-91); the last 11 draws are b1..b9 (:72-75) and b10, b11 (:346)
            "prover_draws": draws[n_index_draws:n_prove_draws],
            "g": rr.enc_scalar(sub["g"]), "k1": rr.enc_scalar(sub["k1"]), "k2": rr.enc_scalar(sub["k2"]),
            "sigma_star": [rr.enc_scalar(s) for s in ipk["sigma_star"]],
            "index_polys": {k: rr.enc_poly(v) for k, v in ipk["polynomials"].items()},
            "keys": rr.keys, "x": [rr.enc_scalar(v) for v in x], "w": [rr.enc_scalar(v) for v in wit],
            "proof": enc_proof(rr, proof), "notes": {"verify": bool(ok)}})


def trace_plonk_normalized_bls():
    """The reference's PLONK indexer / prover / verifier with curve_type="bls12_381" on the bundled circuit (its residues read as
    signed integers, see _plonk_inputs): the bit-exact target of kzg_snark_b200.plonk on the second curve."""
    trace_plonk_normalized("bls12_381")


if __name__ == "__main__":
    assert refrun.available(), "/root/reference not mounted"
    if len(sys.argv) > 1:                                  # regenerate selected fixtures only: make_traces.py trace_kzg_bls ...
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    trace_kzg()
    trace_kzg_bls()
    trace_fft()
    trace_fft_bls()
    trace_plonk()
    trace_marlin()
    marlin_loops()
    trace_marlin_normalized()
    trace_plonk_normalized()
    trace_plonk_normalized_bls()
    trace_marlin_normalized_bls()
