"""Restated oracle of the reference's KZG.setup / commit / open (kzg.py:56-159).

Oracle / test infrastructure only (see oracle/__init__.py).  Parity is PINNED against runs of the
reference's own kzg.py (oracle/refrun.py -> tests/golden/ref_trace_kzg*.json, reproduced bit for bit by
tests/test_reference_traces.py); the reference itself holds no golden vectors.
Also checked by the tau-identity commit(ck, p) == p(tau)*G1 (kzg.py:108) and the pairing-free form
of `check` (kzg.py:161-211) that a known tau allows:
    e(C - v*G1, G2) == e(pi, tau*G2 - z*G2)   <=>   C - v*G1 == (tau - z) * pi.

Polynomials are coefficient lists low->high of ints mod r with trailing zeros
stripped -- what Sage's `poly.list()` returns (kzg.py:110); the zero polynomial
is [] and has degree -1.
"""

from .curve import get_curve


def poly_strip(c, r):
    c = [int(x) % r for x in c]
    while c and c[-1] == 0:
        c.pop()
    return c


def poly_degree(c):
    return len(c) - 1            # Sage: degree of zero polynomial is -1


def poly_eval(c, z, r):
    acc = 0
    for x in reversed(c):
        acc = (acc * z + x) % r
    return acc


def poly_div_linear(c, z, r):
    """(P(X) - P(z)) // (X - z) by synthetic division (kzg.py:153-154): returns q with
    q[d-1] = c[d], q[i-1] = c[i] + z*q[i]."""
    d = len(c) - 1
    if d <= 0:
        return []
    q = [0] * d
    q[d - 1] = c[d]
    for i in range(d - 1, 0, -1):
        q[i - 1] = (c[i] + z * q[i]) % r
    return poly_strip(q, r)


class KZGOracle:
    """Mirror of reference `KZG` for the prover-side methods, on ints."""

    def __init__(self, curve_type="bn254"):
        self.cv = get_curve(curve_type)          # raises ValueError like kzg.py:37
        self.curve_order = self.cv.r
        self.G1 = self.cv.G1
        self.Z1 = self.cv.Z1
        self.multiply = self.cv.multiply
        self.add = self.cv.add
        self.neg = self.cv.neg
        self.eq = self.cv.eq

    def setup(self, max_degree, tau):
        """kzg.py:56-78 with the secret supplied (the reference draws it at :67).
        Returns ck only ([tau^i * G1]); rk = tau*G2 is verifier-side (out of scope).
        NB the reference recomputes tau**i from scratch each step (:72); same result."""
        r = self.curve_order
        ck = [self.G1]
        for i in range(1, max_degree + 1):
            ck.append(self.multiply(self.G1, pow(tau, i, r)))
        return ck

    def setup_fast(self, max_degree, tau):
        """Same ck as setup() up to projective representative (affine-equal), using
        one shared doubling table; for oracle-side SRS at sizes where setup() is too slow."""
        r = self.curve_order
        dbl = [self.G1]
        for _ in range(r.bit_length()):
            dbl.append(self.cv.double(dbl[-1]))
        ck, t = [], 1
        for _ in range(max_degree + 1):
            acc, k, e = self.Z1, 0, t
            while e:
                if e & 1:
                    acc = self.add(acc, dbl[k])
                e >>= 1
                k += 1
            ck.append(acc)
            t = t * tau % r
        return ck

    def commit(self, ck, polynomials):
        """kzg.py:80-120, loop for loop."""
        r = self.curve_order
        polys = [poly_strip(p, r) for p in polynomials]          # :92-97 coercion
        max_degree = len(ck) - 1                                  # :99
        commitments = []
        for poly in polys:
            if poly_degree(poly) > max_degree:                    # :103-106
                raise ValueError(
                    f"Polynomial degree {poly_degree(poly)} exceeds maximum allowed degree {max_degree}"
                )
            commitment = self.Z1                                  # :109
            for i, coeff in enumerate(poly):                      # :112
                if coeff == 0:                                    # :113
                    continue
                term = self.multiply(ck[i], int(coeff))           # :115
                commitment = self.add(commitment, term)           # :116
            commitments.append(commitment)
        return commitments

    def combine(self, polynomials, xi):
        """kzg.py:147-150: sum_i xi^(i+1) * poly_i  (exponent starts at 1)."""
        r = self.curve_order
        polys = [poly_strip(p, r) for p in polynomials]
        m = max((len(p) for p in polys), default=0)
        out = [0] * m
        for i, p in enumerate(polys):
            s = pow(xi, i + 1, r)
            for j, c in enumerate(p):
                out[j] = (out[j] + s * c) % r
        return poly_strip(out, r)

    def witness(self, polynomials, z, xi):
        """kzg.py:144-154: the quotient polynomial whose commitment is the proof."""
        r = self.curve_order
        z = int(z) % r
        xi = int(xi) % r
        return poly_div_linear(self.combine(polynomials, xi), z, r)

    def open(self, ck, polynomials, z, xi):
        """kzg.py:122-159."""
        return self.commit(ck, [self.witness(polynomials, z, xi)])[0]   # :157

    # ---- pairing-free acceptance (needs the trapdoor; tests only) -------
    def check_with_tau(self, tau, commitments, z, evaluations, proof, xi):
        """kzg.py:161-211 with e(A,G2)==e(B,(tau-z)G2) collapsed to A == (tau-z)*B."""
        r = self.curve_order
        z = int(z) % r
        xi = int(xi) % r
        C = self.Z1
        for i, comm in enumerate(commitments):                    # :184-188
            C = self.add(C, self.multiply(comm, pow(xi, i + 1, r)))
        v = 0
        for i, e in enumerate(evaluations):                       # :191-193
            v = (v + pow(xi, i + 1, r) * (int(e) % r)) % r
        lhs = self.add(C, self.neg(self.multiply(self.G1, v)))    # :200-201
        rhs = self.multiply(proof, (tau - z) % r)
        return self.eq(lhs, rhs) if not (self.cv.is_inf(lhs) or self.cv.is_inf(rhs)) \
            else (self.cv.is_inf(lhs) and self.cv.is_inf(rhs))
