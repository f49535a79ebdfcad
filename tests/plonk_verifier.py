"""Test-side PLONK verifier: a restatement of the reference's plonk/verifier.py:25-205 and
kzg.py:213-288 (batch_check) on the oracle's py_ecc stand-in (G1/G2 arithmetic and the BN254
pairing of oracle/pyecc_standin.py).  Test infrastructure only.

It is itself pinned against the reference: tests/test_reference_traces.py requires it to accept
the proof the reference's own prover produced (tests/golden/ref_plonk_normalized.json) and to
reject a tampered one, as plonk/verifier.py:272-290 does."""
import hashlib
import struct

from oracle import pyecc_standin as E

R = E.curve_order


class _Transcript:                                     # transcript.py:18-100
    def __init__(self, label, modulus=None):
        self.state = hashlib.sha256(label.encode()).digest()
        self.modulus = modulus or R

    def _ser(self, d):
        if isinstance(d, str):
            return d.encode()
        if isinstance(d, int):
            return struct.pack(">q", d)
        if isinstance(d, bytes):
            return d
        if isinstance(d, list):
            return b"".join(self._ser(i) for i in d)
        return str(d).encode()

    def append(self, label, data):
        self._upd(label, self._ser(data))

    def _upd(self, label, data):
        self.state = hashlib.sha256(self.state + label.encode() + data).digest()

    def challenge(self, label):
        cs = hashlib.sha256(self.state + label.encode()).digest()
        self._upd(label, cs)
        return int.from_bytes(cs, "big") % self.modulus


class _Dec:
    """A residue that prints like a Sage / shim field element (its decimal representative)."""
    def __init__(self, v, modulus=None):
        self.v = int(v) % (modulus or R)

    def __repr__(self):
        return str(self.v)


def pt(p):
    """affine (x, y) ints / None, or any (x, y, z) triple of int-likes -> stand-in point (x, y, 1)."""
    if p is None:
        return E.Z1
    if len(p) == 3:
        x, y, z = (int(c) for c in p)
        if z == 0:
            return E.Z1
        assert z == 1, "the verifier hashes normalised points"
        return (E.FQ(x), E.FQ(y), E.FQ(1))
    return (E.FQ(int(p[0])), E.FQ(int(p[1])), E.FQ(1))


class _Group:
    """G1 operations for the verifier equation: BN254 on the py_ecc stand-in (final check = the
    reference's pairing equation), any curve on oracle.curve with the trapdoor form of the same
    equation (e(L, G2) == e(R, tau G2)  <=>  L == tau R)."""

    def __init__(self, curve):
        from oracle.curve import get_curve
        self.curve = curve
        self.cv = get_curve(curve)
        self.r = self.cv.r
        self.G1, self.Z1 = self.cv.G1, self.cv.Z1
        self.multiply = lambda p, k: self.cv.multiply(p, int(k) % self.r)
        self.add, self.neg = self.cv.add, self.cv.neg

    def pt(self, p):
        if p is None:
            return self.Z1
        if len(p) == 3:
            if int(p[2]) == 0:
                return self.Z1
            assert int(p[2]) == 1
        return (int(p[0]), int(p[1]), 1)

    def show(self, p):
        """str() of the point as the prover hashed it: (x, y, 1) / (1, 1, 0)."""
        return str(self.Z1 if self.cv.is_inf(p) else (p[0], p[1], 1))

    def final(self, left, right, tau):
        t = self.multiply(right, tau)
        if self.cv.is_inf(left) or self.cv.is_inf(t):
            return self.cv.is_inf(left) and self.cv.is_inf(t)
        return self.cv.eq(left, t)


def verify_trapdoor(ivk, x, proof, curve):
    """The same verifier equation on either curve, closed with the trapdoor instead of a pairing
    (BLS12-381 has no pairing stand-in here).  Same inputs as verify()."""
    G = _Group(curve)
    Rq = G.r
    n, g, k1, k2 = ivk["n"], ivk["g"] % Rq, ivk["k1"] % Rq, ivk["k2"] % Rq
    C = {k: G.pt(v) for k, v in ivk["commitments"].items()}
    pc = {k: G.pt(v) for k, v in proof["commitments"].items()}
    ev = {k: int(v) % Rq for k, v in proof["evaluations"].items()}
    W_z, W_zw = G.pt(proof["kzg_proofs"]["W_z"]), G.pt(proof["kzg_proofs"]["W_zw"])
    a, b, c, s1, s2, zw = ev["a"], ev["b"], ev["c"], ev["s_sigma1"], ev["s_sigma2"], ev["z_omega"]

    class _S:                                              # prints like the prover's point tuples
        def __init__(self, p):
            self.p = p

        def __repr__(self):
            return G.show(self.p)

    t = _Transcript("plonk-proof", Rq)
    t.append("public-inputs", [_Dec(v, Rq) for v in x])
    t.append("round1-commitments", [_S(pc["a"]), _S(pc["b"]), _S(pc["c"])])
    beta, gamma = t.challenge("beta"), t.challenge("gamma")
    t.append("round2-commitment", _S(pc["z"]))
    alpha = t.challenge("alpha")
    t.append("round3-commitments", [_S(pc["t_lo"]), _S(pc["t_mid"]), _S(pc["t_hi"])])
    zeta = t.challenge("zeta")
    t.append("round4-evaluations", [_Dec(v, Rq) for v in (a, b, c, s1, s2, zw)])
    v = t.challenge("v")
    u = t.challenge("u")
    zn = pow(zeta, n, Rq)
    zh = (zn - 1) % Rq
    l1 = zh * pow(n * (zeta - 1) % Rq, -1, Rq) % Rq
    pi = 0
    for i, xi in enumerate(x):
        gi = pow(g, i, Rq)
        pi = (pi - int(xi) * gi % Rq * zh % Rq * pow(n * (zeta - gi) % Rq, -1, Rq)) % Rq
    mul, add, neg = G.multiply, G.add, G.neg
    r_comm = mul(C["qM"], a * b % Rq)
    for P_, s in ((C["qL"], a), (C["qR"], b), (C["qO"], c), (G.G1, pi), (C["qC"], 1)):
        r_comm = add(r_comm, mul(P_, s))
    f1 = (a + beta * zeta + gamma) * (b + beta * k1 * zeta + gamma) % Rq * (c + beta * k2 * zeta + gamma) % Rq
    cterm = add(mul(C["S_sigma3"], beta), mul(G.G1, (c + gamma) % Rq))
    f2 = (a + beta * s1 + gamma) * (b + beta * s2 + gamma) % Rq * zw % Rq
    r_comm = add(r_comm, mul(add(mul(pc["z"], f1), neg(mul(cterm, f2))), alpha))
    r_comm = add(r_comm, mul(add(pc["z"], neg(G.G1)), alpha * alpha % Rq * l1 % Rq))
    tcomb = add(add(pc["t_lo"], mul(pc["t_mid"], zn)), mul(pc["t_hi"], zn * zn % Rq))
    r_comm = add(r_comm, neg(mul(tcomb, zh)))
    inst = [([r_comm, pc["a"], pc["b"], pc["c"], C["S_sigma1"], C["S_sigma2"]], zeta, [0, a, b, c, s1, s2], W_z),
            ([pc["z"]], zeta * g % Rq, [zw], W_zw)]
    left, right = G.Z1, G.Z1
    for i, (comms, z, evals, proof_pt) in enumerate(inst):
        cc, ce = G.Z1, 0
        for j, cm in enumerate(comms):
            xp = pow(v, j + 1, Rq)
            cc = add(cc, mul(cm, xp))
            ce = (ce + xp * evals[j]) % Rq
        cmv = add(cc, neg(mul(G.G1, ce)))
        left = add(left, mul(add(cmv, mul(proof_pt, z)), pow(u, i + 1, Rq)))
        right = add(right, mul(proof_pt, pow(u, i + 1, Rq)))
    return G.final(left, right, ivk["tau"] % Rq)


def verify(ivk, x, proof):
    """ivk: {"commitments": {name: point}, "n", "g", "k1", "k2", "tau"}; x: public inputs (ints);
    proof: {"commitments": {...}, "evaluations": {...}, "kzg_proofs": {...}} with points as
    (x, y) / (x, y, 1) / None and evaluations as ints."""
    n, g, k1, k2 = ivk["n"], ivk["g"] % R, ivk["k1"] % R, ivk["k2"] % R
    rk = E.multiply(E.G2, ivk["tau"] % R)                                   # kzg.py:75
    C = {k: pt(v) for k, v in ivk["commitments"].items()}
    pc = {k: pt(v) for k, v in proof["commitments"].items()}
    ev = {k: int(v) % R for k, v in proof["evaluations"].items()}
    W_z, W_zw = pt(proof["kzg_proofs"]["W_z"]), pt(proof["kzg_proofs"]["W_zw"])
    a, b, c, s1, s2, zw = ev["a"], ev["b"], ev["c"], ev["s_sigma1"], ev["s_sigma2"], ev["z_omega"]

    t = _Transcript("plonk-proof")                                          # plonk/verifier.py:91-110
    t.append("public-inputs", [_Dec(v) for v in x])
    t.append("round1-commitments", [pc["a"], pc["b"], pc["c"]])
    beta, gamma = t.challenge("beta"), t.challenge("gamma")
    t.append("round2-commitment", pc["z"])
    alpha = t.challenge("alpha")
    t.append("round3-commitments", [pc["t_lo"], pc["t_mid"], pc["t_hi"]])
    zeta = t.challenge("zeta")
    t.append("round4-evaluations", [_Dec(v) for v in (a, b, c, s1, s2, zw)])
    v = t.challenge("v")
    u = t.challenge("u")

    zn = pow(zeta, n, R)
    zh = (zn - 1) % R
    l1 = zh * pow(n * (zeta - 1) % R, -1, R) % R
    # PI(zeta) = -sum x_i L_i(zeta), L_i(X) = g^i (X^n - 1) / (n (X - g^i))     plonk/encoder.py:196-223
    pi = 0
    for i, xi in enumerate(x):
        gi = pow(g, i, R)
        pi = (pi - int(xi) * gi % R * zh % R * pow(n * (zeta - gi) % R, -1, R)) % R

    mul, add, neg = E.multiply, E.add, E.neg
    r_comm = mul(C["qM"], a * b % R)                                        # plonk/verifier.py:117-157
    for P_, s in ((C["qL"], a), (C["qR"], b), (C["qO"], c), (E.G1, pi), (C["qC"], 1)):
        r_comm = add(r_comm, mul(P_, s))
    f1 = (a + beta * zeta + gamma) * (b + beta * k1 * zeta + gamma) % R * (c + beta * k2 * zeta + gamma) % R
    term1 = mul(pc["z"], f1)
    cterm = add(mul(C["S_sigma3"], beta), mul(E.G1, (c + gamma) % R))
    f2 = (a + beta * s1 + gamma) * (b + beta * s2 + gamma) % R * zw % R
    term2 = mul(cterm, f2)
    r_comm = add(r_comm, mul(add(term1, neg(term2)), alpha))
    r_comm = add(r_comm, mul(add(pc["z"], neg(E.G1)), alpha * alpha % R * l1 % R))
    tcomb = add(add(pc["t_lo"], mul(pc["t_mid"], zn)), mul(pc["t_hi"], zn * zn % R))
    r_comm = add(r_comm, neg(mul(tcomb, zh)))

    # kzg.batch_check (kzg.py:213-288) over the two openings with xi = v and batching scalar u
    inst = [([r_comm, pc["a"], pc["b"], pc["c"], C["S_sigma1"], C["S_sigma2"]], zeta, [0, a, b, c, s1, s2], W_z),
            ([pc["z"]], zeta * g % R, [zw], W_zw)]
    left, right = E.Z1, E.Z1
    for i, (comms, z, evals, proof_pt) in enumerate(inst):
        cc, ce = E.Z1, 0
        for j, cm in enumerate(comms):
            xp = pow(v, j + 1, R)
            cc = add(cc, mul(cm, xp))
            ce = (ce + xp * evals[j]) % R
        cmv = add(cc, neg(mul(E.G1, ce)))
        tl = mul(add(cmv, mul(proof_pt, z)), pow(u, i + 1, R))
        left = add(left, tl)
        right = add(right, mul(proof_pt, pow(u, i + 1, R)))
    return E.pairing(E.G2, left) == E.pairing(rk, right)
