"""Stand-in for `py_ecc.optimized_bn128` so that the reference's OWN kzg.py / plonk / marlin
modules can be imported and run in this image (py_ecc is not installable here).

Oracle / test infrastructure only (see oracle/__init__.py).  Used by oracle/refrun.py to
generate the golden call traces under tests/golden/ and by the CPU tests.

It exposes exactly the names kzg.py:27-30 imports -- G1, G2, multiply, add, curve_order,
pairing, neg, Z1, Z2, eq -- with py_ecc's value shapes: projective triples of `FQ` (G1) or
`FQ2` (G2), `FQ.n`, `repr(FQ) == repr(int)` (transcript.py:80-85 hashes str(point)).

  * G1 arithmetic delegates to oracle/curve.py (the restated optimized_curve.py formulas), so
    the projective representative -- which the Fiat-Shamir transcript hashes -- is the one
    those formulas produce.
  * G2 arithmetic runs the same formulas generically over FQ2 = Fp[i]/(i^2+1).
  * `pairing(Q, P)` is the ate pairing on BN254 restated from the published py_ecc bn128
    algorithm: twist G2 -> E(Fp12) with Fp12 = Fp[w]/(w^12 - 18 w^6 + 82), affine Miller loop
    over the loop count 29793968203157093288, two Frobenius line steps, final exponentiation
    (p^12-1)/r.  Only equality of pairing values is ever used (kzg.py:209,286), so any correct
    bilinear non-degenerate pairing decides accept/reject identically; tests/test_oracle.py
    checks bilinearity and non-degeneracy.

G2 generator: the standard BN254 (alt_bn128) generator; checked on the twist and of order r
in tests/test_oracle.py.
"""

from .curve import FQ as _FQ, get_curve

_PARAMS = {
    "bn254": {
        "fq12_mc": (82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0),              # w^12 - 18 w^6 + 82
        "g2_x": [10857046999023057135944570762232829481370756359578518086990519993285655852781,
                 11559732032986387107991004021392285783925812861821192530917403151452391805634],
        "g2_y": [8495653923123431417604973247489272438418190587263600148770280649306958101930,
                 4082367875863433681332203403145435568316851327593401208105741076214120093531],
        "b": 3, "b2": lambda FQ2: FQ2([3, 0]) / FQ2([9, 1]),             # twist: y^2 = x^3 + 3/(9+i)
        "twist_xi_real": 9, "twist_divides": False,
        "ate_loop_count": 29793968203157093288, "log_ate_loop_count": 63, "frobenius_steps": True,
    },
    # py_ecc.optimized_bls12_381 (kzg.py:32-35): Fp12 = Fp[w]/(w^12 - 2 w^6 + 2), twist y^2 = x^3 + 4(1+i),
    # ate loop over |x| = 0xd201000000010000 without Frobenius steps; the standard G2 generator
    "bls12_381": {
        "fq12_mc": (2, 0, 0, 0, 0, 0, -2, 0, 0, 0, 0, 0),
        "g2_x": [0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
                 0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e],
        "g2_y": [0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
                 0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be],
        "b": 4, "b2": lambda FQ2: FQ2([4, 4]),
        "twist_xi_real": 1, "twist_divides": True,
        "ate_loop_count": 15132376222941642752, "log_ate_loop_count": 62, "frobenius_steps": False,
    },
}


def make(curve_name):
    """All names of the stand-in for one curve, as a dict (installed as a module's globals below)."""
    prm = _PARAMS[curve_name]
    _cv = get_curve(curve_name)
    field_modulus = _cv.p
    curve_order = _cv.r
    P = field_modulus

    class FQ(_FQ):
        """py_ecc.fields.optimized_bn128_FQ look-alike (single modulus)."""
        __slots__ = ()

        def __init__(self, n, p=P):
            _FQ.__init__(self, n, P)


    # ------------------------------------------------------------------ extension fields
    class FQP:
        """Fp[x]/(x^deg + sum mc[i] x^i); coefficients are plain ints."""
        degree = 0
        mc = ()

        def __init__(self, coeffs):
            assert len(coeffs) == self.degree
            self.coeffs = [int(c) % P for c in coeffs]

        def __add__(self, o):
            return type(self)([a + b for a, b in zip(self.coeffs, o.coeffs)])

        def __sub__(self, o):
            return type(self)([a - b for a, b in zip(self.coeffs, o.coeffs)])

        def __neg__(self):
            return type(self)([-a for a in self.coeffs])

        def __mul__(self, o):
            if isinstance(o, (int, _FQ)):
                k = int(o)
                return type(self)([a * k for a in self.coeffs])
            d = self.degree
            b = [0] * (2 * d - 1)
            for i, x in enumerate(self.coeffs):
                if x:
                    for j, y in enumerate(o.coeffs):
                        b[i + j] += x * y
            for exp in range(2 * d - 2, d - 1, -1):
                top = b[exp] % P
                if top:
                    base = exp - d
                    for i, c in self.mc_items:
                        b[base + i] -= top * c
            return type(self)(b[:d])

        __rmul__ = __mul__

        def __pow__(self, e):
            e = int(e)
            out = type(self).one()
            base = self
            while e:
                if e & 1:
                    out = out * base
                base = base * base
                e >>= 1
            return out

        def inv(self):
            # extended Euclid on polynomials over Fp: self * inv = 1 mod modulus
            d = self.degree
            lm, hm = [1] + [0] * d, [0] * (d + 1)
            low, high = self.coeffs + [0], list(self.mc) + [1]
            while _deg(low):
                r = _poly_rounded_div(high, low)
                r += [0] * (d + 1 - len(r))
                nm, new = list(hm), list(high)
                for i in range(d + 1):
                    for j in range(d + 1 - i):
                        nm[i + j] -= lm[i] * r[j]
                        new[i + j] -= low[i] * r[j]
                nm = [x % P for x in nm]
                new = [x % P for x in new]
                lm, low, hm, high = nm, new, lm, low
            li = pow(low[0], -1, P)
            return type(self)([c * li for c in lm[:d]])

        def __truediv__(self, o):
            if isinstance(o, (int, _FQ)):
                return self * pow(int(o), -1, P)
            return self * o.inv()

        def __eq__(self, o):
            return isinstance(o, FQP) and self.coeffs == o.coeffs

        def __ne__(self, o):
            return not self == o

        def __hash__(self):
            return hash(tuple(self.coeffs))

        def __repr__(self):
            return repr(tuple(self.coeffs))

        def is_zero(self):
            return not any(self.coeffs)

        @classmethod
        def one(cls):
            return cls([1] + [0] * (cls.degree - 1))

        @classmethod
        def zero(cls):
            return cls([0] * cls.degree)


    def _deg(p):
        d = len(p) - 1
        while d and p[d] % P == 0:
            d -= 1
        return d


    def _poly_rounded_div(a, b):
        dega, degb = _deg(a), _deg(b)
        temp = [x % P for x in a]
        o = [0] * len(a)
        binv = pow(b[degb], -1, P)
        for i in range(dega - degb, -1, -1):
            q = temp[degb + i] * binv % P
            o[i] = (o[i] + q) % P
            for c in range(degb + 1):
                temp[c + i] = (temp[c + i] - b[c] * q) % P
        return o[:_deg(o) + 1]


    class FQ2(FQP):
        degree = 2
        mc = (1, 0)                                        # i^2 + 1
        mc_items = ((0, 1),)


    class FQ12(FQP):
        degree = 12
        mc = prm["fq12_mc"]
        mc_items = tuple((i, c) for i, c in enumerate(prm["fq12_mc"]) if c)


    # ------------------------------------------------------------------ curve points
    G1 = (FQ(_cv.G1[0]), FQ(_cv.G1[1]), FQ(1))
    Z1 = (FQ(1), FQ(1), FQ(0))
    G2 = (FQ2(prm["g2_x"]), FQ2(prm["g2_y"]), FQ2.one())
    Z2 = (FQ2.one(), FQ2.one(), FQ2.zero())
    b = FQ(prm["b"])
    b2 = prm["b2"](FQ2)


    def _is_g1(pt):
        return isinstance(pt[0], _FQ)


    def _ints(pt):
        return (pt[0].n, pt[1].n, pt[2].n)


    def _fq(t):
        return (FQ(t[0]), FQ(t[1]), FQ(t[2]))


    def is_inf(pt):
        z = pt[-1]
        return z.n == 0 if isinstance(z, _FQ) else z.is_zero()


    # generic projective formulas (py_ecc optimized_curve.py) for the FQ2 case
    def _double_g(pt):
        x, y, z = pt
        W = x * x * 3
        S = y * z
        B = x * y * S
        H = W * W - B * 8
        S2 = S * S
        return (H * S * 2, W * (B * 4 - H) - y * y * S2 * 8, S * S2 * 8)


    def _add_g(p1, p2):
        if is_inf(p1) or is_inf(p2):
            return p1 if is_inf(p2) else p2
        x1, y1, z1 = p1
        x2, y2, z2 = p2
        U1, U2, V1, V2 = y2 * z1, y1 * z2, x2 * z1, x1 * z2
        if V1 == V2 and U1 == U2:
            return _double_g(p1)
        if V1 == V2:
            one = type(x1).one()
            return (one, one, type(x1).zero())
        U, V = U1 - U2, V1 - V2
        Vsq = V * V
        VsqV2 = Vsq * V2
        Vcu = V * Vsq
        W = z1 * z2
        A = U * U * W - Vcu - VsqV2 * 2
        return (V * A, U * (VsqV2 - A) - Vcu * U2, Vcu * W)


    def double(pt):
        return _fq(_cv.double(_ints(pt))) if _is_g1(pt) else _double_g(pt)


    def add(p1, p2):
        if _is_g1(p1):
            return _fq(_cv.add(_ints(p1), _ints(p2)))
        return _add_g(p1, p2)


    def multiply(pt, n):
        n = int(n)
        if _is_g1(pt):
            return _fq(_cv.multiply(_ints(pt), n))
        if n == 0:
            one = type(pt[0]).one()
            return (one, one, type(pt[0]).zero())
        if n == 1:
            return pt
        if not n % 2:
            return multiply(_double_g(pt), n // 2)
        return _add_g(multiply(_double_g(pt), n // 2), pt)


    def neg(pt):
        x, y, z = pt
        return (x, -y, z)


    def eq(p1, p2):
        x1, y1, z1 = p1
        x2, y2, z2 = p2
        return x1 * z2 == x2 * z1 and y1 * z2 == y2 * z1


    def normalize(pt):
        x, y, z = pt
        return (x / z, y / z)


    def is_on_curve(pt, bb):
        if is_inf(pt):
            return True
        x, y, z = pt
        return y * y * z - x * x * x == bb * z * z * z


    # ------------------------------------------------------------------ pairing (affine, Fp12)
    ate_loop_count = prm["ate_loop_count"]
    log_ate_loop_count = prm["log_ate_loop_count"]
    _w = FQ12([0, 1] + [0] * 10)
    _w2, _w3 = _w * _w, _w * _w * _w


    def _twist(q_affine):
        x, y = q_affine
        k = prm["twist_xi_real"]                          # Fp2 = Fp[i] -> Fp[w^6] with w^6 = xi = k + i
        xc = [x.coeffs[0] - x.coeffs[1] * k, x.coeffs[1]]
        yc = [y.coeffs[0] - y.coeffs[1] * k, y.coeffs[1]]
        nx = FQ12([xc[0]] + [0] * 5 + [xc[1]] + [0] * 5)
        ny = FQ12([yc[0]] + [0] * 5 + [yc[1]] + [0] * 5)
        if prm["twist_divides"]:                          # M-type twist (BLS12-381): E' -> E is (x / w^2, y / w^3)
            return (nx / _w2, ny / _w3)
        return (nx * _w2, ny * _w3)


    def _embed(p_affine):
        x, y = p_affine
        return (FQ12([x.n] + [0] * 11), FQ12([y.n] + [0] * 11))


    def _aff_double(pt):
        x, y = pt
        m = (x * x * 3) / (y * 2)
        nx = m * m - x * 2
        return (nx, m * (x - nx) - y)


    def _aff_add(p1, p2):
        if p1 is None or p2 is None:
            return p1 if p2 is None else p2
        x1, y1 = p1
        x2, y2 = p2
        if x1 == x2:
            return _aff_double(p1) if y1 == y2 else None
        m = (y2 - y1) / (x2 - x1)
        nx = m * m - x1 - x2
        return (nx, m * (x1 - nx) - y1)


    def _linefunc(P1, P2, T):
        x1, y1 = P1
        x2, y2 = P2
        xt, yt = T
        if x1 != x2:
            m = (y2 - y1) / (x2 - x1)
            return m * (xt - x1) - (yt - y1)
        if y1 == y2:
            m = (x1 * x1 * 3) / (y1 * 2)
            return m * (xt - x1) - (yt - y1)
        return xt - x1


    def _miller_loop(Q, Pt):
        R, f = Q, FQ12.one()
        for i in range(log_ate_loop_count, -1, -1):
            f = f * f * _linefunc(R, R, Pt)
            R = _aff_double(R)
            if ate_loop_count & (1 << i):
                f = f * _linefunc(R, Q, Pt)
                R = _aff_add(R, Q)
        if prm["frobenius_steps"]:                        # BN curves: two more line steps with the Frobenius images of Q
            Q1 = (Q[0] ** P, Q[1] ** P)
            nQ2 = (Q1[0] ** P, -(Q1[1] ** P))
            f = f * _linefunc(R, Q1, Pt)
            R = _aff_add(R, Q1)
            f = f * _linefunc(R, nQ2, Pt)
        return f ** ((P ** 12 - 1) // curve_order)


    def pairing(Q, Pt):
        """e(P, Q) with Q in G2, P in G1 (argument order of py_ecc, kzg.py:205-206)."""
        assert is_on_curve(Q, b2), "Q not on the twist"
        assert is_on_curve(Pt, b), "P not on the curve"
        if is_inf(Pt) or is_inf(Q):
            return FQ12.one()
        return _miller_loop(_twist(normalize(Q)), _embed(normalize(Pt)))

    return {k: v for k, v in locals().items() if k not in ("prm",)}


globals().update(make("bn254"))
