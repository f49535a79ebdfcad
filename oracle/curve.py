"""Restated oracle of py_ecc's "optimized" G1 arithmetic (BN254, BLS12-381).

Oracle / test infrastructure only (see oracle/__init__.py).

Third-party dependency restated here: py_ecc (PyPI), modules
`py_ecc.optimized_bn128` and `py_ecc.optimized_bls12_381`, version NOT pinned by
the reference (README.md:15,24-25) and not present under /root/reference or in
this image.  The reference binds its callables at kzg.py:27-35,40-49 and calls
them at kzg.py:72,115-116.  Restated from the library's published algorithm
(optimized_curve.py): homogeneous projective points (x, y, z) over FQ,
Z1 = (1, 1, 0), G1 = (Gx, Gy, 1), `double`, `add`, recursive binary `multiply`,
`neg`, cross-multiplied `eq`, `normalize` = (x/z, y/z).  The library's internals
cannot be run here; these formulas are pinned by the group law (canonical affine
result), the public known-answer multiples in tests/golden/public_kat.json and by
the reference's own verifiers accepting proofs computed with them (pairing check,
tests/golden/ref_trace_*.json via oracle/refrun.py).

Points are tuples of plain ints (mod p) for speed; `FQ` is the presentation
type with py_ecc's `.n` attribute used at the drop-in boundary.
"""

from .params import curve as _curve


class FQ:
    """Shape of py_ecc.fields.optimized_*_FQ as far as kzg.py's callers see it:
    `.n`, int(), ==, + - * /, printable as the integer (transcript.py:80-85 hashes str())."""
    __slots__ = ("n", "p")

    def __init__(self, n, p):
        self.n = int(n) % p
        self.p = p

    def _c(self, o):
        return o.n if isinstance(o, FQ) else int(o) % self.p

    def __add__(self, o): return FQ(self.n + self._c(o), self.p)
    __radd__ = __add__
    def __sub__(self, o): return FQ(self.n - self._c(o), self.p)
    def __rsub__(self, o): return FQ(self._c(o) - self.n, self.p)
    def __mul__(self, o): return FQ(self.n * self._c(o), self.p)
    __rmul__ = __mul__
    def __neg__(self): return FQ(-self.n, self.p)
    def __truediv__(self, o): return FQ(self.n * pow(self._c(o), -1, self.p), self.p)
    def __pow__(self, e): return FQ(pow(self.n, int(e), self.p), self.p)
    def __eq__(self, o):
        try:
            return self.n == self._c(o)
        except (TypeError, ValueError):
            return NotImplemented
    def __hash__(self): return hash(self.n)
    def __int__(self): return self.n
    __index__ = __int__
    def __repr__(self): return str(self.n)
    @classmethod
    def one(cls, p): return cls(1, p)
    @classmethod
    def zero(cls, p): return cls(0, p)


class G1Curve:
    """py_ecc-style function set for one curve, on int triples."""

    def __init__(self, name):
        cv = _curve(name)
        self.name = name
        self.p = cv["p"]
        self.r = cv["r"]
        self.b = cv["b"]
        self.G1 = (cv["G1"][0], cv["G1"][1], 1)
        self.Z1 = (1, 1, 0)

    # py_ecc optimized_curve.double
    def double(self, pt):
        p = self.p
        x, y, z = pt
        W = 3 * x * x % p
        S = y * z % p
        B = x * y * S % p
        H = (W * W - 8 * B) % p
        S_squared = S * S % p
        newx = 2 * H * S % p
        newy = (W * (4 * B - H) - 8 * y * y * S_squared) % p
        newz = 8 * S * S_squared % p
        return (newx, newy, newz)

    # py_ecc optimized_curve.add
    def add(self, p1, p2):
        p = self.p
        if p1[2] == 0 or p2[2] == 0:
            return p1 if p2[2] == 0 else p2
        x1, y1, z1 = p1
        x2, y2, z2 = p2
        U1 = y2 * z1 % p
        U2 = y1 * z2 % p
        V1 = x2 * z1 % p
        V2 = x1 * z2 % p
        if V1 == V2 and U1 == U2:
            return self.double(p1)
        elif V1 == V2:
            return (1, 1, 0)
        U = (U1 - U2) % p
        V = (V1 - V2) % p
        V_squared = V * V % p
        V_squared_times_V2 = V_squared * V2 % p
        V_cubed = V * V_squared % p
        W = z1 * z2 % p
        A = (U * U * W - V_cubed - 2 * V_squared_times_V2) % p
        newx = V * A % p
        newy = (U * (V_squared_times_V2 - A) - V_cubed * U2) % p
        newz = V_cubed * W % p
        return (newx, newy, newz)

    # py_ecc optimized_curve.multiply (recursive binary double-and-add)
    def multiply(self, pt, n):
        if n == 0:
            return (1, 1, 0)
        elif n == 1:
            return pt
        elif not n % 2:
            return self.multiply(self.double(pt), n // 2)
        else:
            return self.add(self.multiply(self.double(pt), int(n // 2)), pt)

    def neg(self, pt):
        x, y, z = pt
        return (x, (-y) % self.p, z)

    def is_inf(self, pt):
        return pt[2] % self.p == 0

    def eq(self, p1, p2):
        p = self.p
        x1, y1, z1 = p1
        x2, y2, z2 = p2
        return (x1 * z2 - x2 * z1) % p == 0 and (y1 * z2 - y2 * z1) % p == 0

    def normalize(self, pt):
        """(x/z, y/z); py_ecc returns it as a 2-tuple.  Infinity -> None here."""
        x, y, z = pt
        if z % self.p == 0:
            return None
        zi = pow(z, -1, self.p)
        return (x * zi % self.p, y * zi % self.p)

    def is_on_curve(self, pt):
        if self.is_inf(pt):
            return True
        x, y, z = pt
        p = self.p
        return (y * y * z - x * x * x - self.b * z * z * z) % p == 0

    # presentation helpers ------------------------------------------------
    def to_fq(self, pt):
        return tuple(FQ(c, self.p) for c in pt)

    def from_any(self, pt):
        """Accept py_ecc triples of FQ, int triples, or affine pairs."""
        if len(pt) == 2:
            return (int(pt[0]) % self.p, int(pt[1]) % self.p, 1)
        return tuple(int(c) % self.p for c in pt)


_cache = {}


def get_curve(name):
    if name not in _cache:
        _cache[name] = G1Curve(name)
    return _cache[name]
