"""Sage-free stand-ins for the few SageMath types the reference's callers hand to kzg.py and
fft_ff.py: GF(q) elements and PolynomialRing(GF(q), 'X') polynomials.

Used ONLY when `sage.all` cannot be imported (it is not installed in this image, SURVEY.md
section 0); with Sage present the drop-in modules use the real Sage objects.  These classes are
caller-side value types (layer L2/L3 in SURVEY.md section 1), not part of the accelerated path:
commit / open / fft_ff / ifft_ff always convert to limb arrays and run on the GPU.

Semantics follow Sage where the reference relies on them:
  * residues print as their integer representative (transcript.py:80-85 hashes str(obj));
  * poly.list() is low->high with trailing zeros stripped, zero poly -> [], degree -1
    (kzg.py:103,110);
  * `/` between polynomials builds a fraction that R(...) converts back when the division is
    exact (plonk/prover.py:109,297-316).
"""

import random as _random

# GF(q).random_element() draws the KZG trapdoor (kzg.py:67), the batching challenge of batch_check (kzg.py:236) and the
# provers' blinding scalars: by default it is backed by the operating system's CSPRNG.  seed() switches to a seeded
# Mersenne Twister for reproducible tests and golden-trace generation only (Sage's set_random_seed).
_rng = _random.SystemRandom()


def seed(s):
    """Make GF(q).random_element() deterministic (tests / fixture generation): seeded Mersenne Twister from here on.
    seed(None) returns to the system CSPRNG."""
    global _rng
    _rng = _random.SystemRandom() if s is None else _random.Random(s)


class FieldElement:
    __slots__ = ("n", "F")

    def __init__(self, n, F):
        self.n = n % F.q
        self.F = F

    # -- coercion
    def _c(self, o):
        if isinstance(o, FieldElement):
            return o.n
        if isinstance(o, int):
            return o % self.F.q
        if hasattr(o, "__int__") and not isinstance(o, (Poly, Fraction)):
            return int(o) % self.F.q
        return None

    def __add__(self, o):
        c = self._c(o)
        return NotImplemented if c is None else FieldElement(self.n + c, self.F)

    __radd__ = __add__

    def __sub__(self, o):
        c = self._c(o)
        return NotImplemented if c is None else FieldElement(self.n - c, self.F)

    def __rsub__(self, o):
        c = self._c(o)
        return NotImplemented if c is None else FieldElement(c - self.n, self.F)

    def __mul__(self, o):
        c = self._c(o)
        return NotImplemented if c is None else FieldElement(self.n * c, self.F)

    __rmul__ = __mul__

    def __neg__(self):
        return FieldElement(-self.n, self.F)

    def __truediv__(self, o):
        c = self._c(o)
        if c is None:
            return NotImplemented
        if c == 0:
            raise ZeroDivisionError("inverse of Mod(0, q) does not exist")
        return FieldElement(self.n * pow(c, -1, self.F.q), self.F)

    def __rtruediv__(self, o):
        c = self._c(o)
        if c is None:
            return NotImplemented
        return FieldElement(c * pow(self.n, -1, self.F.q), self.F)

    def __pow__(self, e):
        return FieldElement(pow(self.n, int(e), self.F.q), self.F)

    def __eq__(self, o):
        c = self._c(o)
        return NotImplemented if c is None else self.n == c

    def __ne__(self, o):
        r = self.__eq__(o)
        return r if r is NotImplemented else not r

    def __hash__(self):
        return hash(self.n)

    def __int__(self):
        return self.n

    __index__ = __int__

    def __bool__(self):
        return self.n != 0

    def __repr__(self):
        return str(self.n)

    def parent(self):
        return self.F

    def is_zero(self):
        return self.n == 0

    def multiplicative_order(self):
        if self.n == 0:
            raise ArithmeticError("multiplicative order of 0 not defined")
        order = self.F.q - 1
        for p in self.F._factors():
            while order % p == 0 and pow(self.n, order // p, self.F.q) == 1:
                order //= p
        return order

    def nth_root(self, n):
        """An element of exact order n when self == 1 (how the callers use it,
        plonk/encoder.py:49): multiplicative_generator()^((q-1)/n)."""
        if self.n != 1:
            raise NotImplementedError("nth_root only for 1 (roots of unity)")
        q = self.F.q
        if (q - 1) % n:
            raise ValueError("no n-th root of unity in this field")
        return FieldElement(pow(self.F.multiplicative_generator().n, (q - 1) // n, q), self.F)


_KNOWN = {
    # q: (least primitive root, prime factors of q-1)   [SURVEY.md section 7 hard part 7, section 8d]
    21888242871839275222246405745257275088548364400416034343698204186575808495617: (
        5, [2, 3, 13, 29, 983, 11003, 237073, 405928799, 1670836401704629, 13818364434197438864469338081]),
    52435875175126190479447740508185965837690552500527637822603658699938581184513: (
        7, [2, 3, 11, 19, 10177, 125527, 859267, 906349, 2508409, 2529403, 52437899, 254760293]),
}


class GFShim:
    """Callable like Sage's GF(q): F(x) coerces, F.random_element(), F.order()."""

    _cache = {}

    def __new__(cls, q):
        q = int(q)
        if q not in cls._cache:
            obj = super().__new__(cls)
            obj.q = q
            cls._cache[q] = obj
        return cls._cache[q]

    def __call__(self, x=0):
        if isinstance(x, FieldElement):
            return FieldElement(x.n, self)
        if isinstance(x, Poly):
            if x.degree() > 0:
                raise TypeError("not a constant polynomial")
            return FieldElement(x.c[0] if x.c else 0, self)
        if isinstance(x, Fraction):
            return self(x.to_poly())
        return FieldElement(int(x), self)

    def from_canonical_ints(self, ints):
        """list of elements from residues already in [0, q) -- the results of a device call -- without the per-element
        coercion and reduction of __call__ (the Python-object boundary of fft_ff at 2^18 elements is otherwise mostly this)."""
        new, cls, out = object.__new__, FieldElement, []
        append = out.append
        for v in ints:
            e = new(cls)
            e.n = v
            e.F = self
            append(e)
        return out

    def order(self):
        return self.q

    cardinality = order
    characteristic = order

    def random_element(self):
        return FieldElement(_rng.randrange(self.q), self)

    def _factors(self):
        if self.q in _KNOWN:
            return _KNOWN[self.q][1]
        raise NotImplementedError("factorisation of q-1 unknown for this modulus")

    def multiplicative_generator(self):
        if self.q in _KNOWN:
            return FieldElement(_KNOWN[self.q][0], self)
        raise NotImplementedError

    def zero(self):
        return FieldElement(0, self)

    def one(self):
        return FieldElement(1, self)

    def __repr__(self):
        return f"Finite Field of size {self.q}"

    def __eq__(self, o):
        return isinstance(o, GFShim) and o.q == self.q

    def __hash__(self):
        return hash(("GFShim", self.q))


def GF(q):
    return GFShim(q)


# --------------------------------------------------------------------------- polynomials
def _strip(c):
    while c and c[-1] == 0:
        c.pop()
    return c


class Poly:
    """Dense univariate polynomial over GF(q); coefficients are plain ints in self.c."""
    __slots__ = ("c", "R")

    def __init__(self, c, R):
        self.c = c
        self.R = R

    # -- helpers
    @property
    def q(self):
        return self.R.F.q

    def _coerce(self, o):
        if isinstance(o, Poly):
            return o
        if isinstance(o, Fraction):
            return None
        if isinstance(o, FieldElement):
            return Poly([o.n] if o.n else [], self.R)
        if isinstance(o, int) or hasattr(o, "__int__"):
            v = int(o) % self.q
            return Poly([v] if v else [], self.R)
        return None

    # -- Sage surface used by the reference
    def list(self):
        F = self.R.F
        return [FieldElement(x, F) for x in self.c]

    coefficients_list = list

    def int_list(self):
        return self.c[:]

    def degree(self):
        return len(self.c) - 1

    def parent(self):
        return self.R

    def is_zero(self):
        return not self.c

    def leading_coefficient(self):
        return FieldElement(self.c[-1] if self.c else 0, self.R.F)

    def constant_coefficient(self):
        return FieldElement(self.c[0] if self.c else 0, self.R.F)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [FieldElement(x, self.R.F) for x in self.c[i]]
        return FieldElement(self.c[i] if 0 <= i < len(self.c) else 0, self.R.F)

    def __iter__(self):
        return iter(self.list())

    def __call__(self, x):
        if isinstance(x, Poly):                    # composition, e.g. z_poly(g * X)
            res = Poly([], self.R)
            for a in reversed(self.c):
                res = res * x + a
            return res
        q = self.q
        xv = int(x) % q
        acc = 0
        for a in reversed(self.c):
            acc = (acc * xv + a) % q
        return FieldElement(acc, self.R.F)

    def __add__(self, o):
        p = self._coerce(o)
        if p is None:
            return NotImplemented
        a, b = self.c, p.c
        if len(a) < len(b):
            a, b = b, a
        q = self.q
        out = a[:]
        for i, v in enumerate(b):
            out[i] = (out[i] + v) % q
        return Poly(_strip(out), self.R)

    __radd__ = __add__

    def __neg__(self):
        q = self.q
        return Poly([(-v) % q for v in self.c], self.R)

    def __sub__(self, o):
        p = self._coerce(o)
        if p is None:
            return NotImplemented
        return self + (-p)

    def __rsub__(self, o):
        p = self._coerce(o)
        if p is None:
            return NotImplemented
        return p + (-self)

    def __mul__(self, o):
        if isinstance(o, Fraction):
            return NotImplemented
        p = self._coerce(o)
        if p is None:
            return NotImplemented
        a, b = self.c, p.c
        if not a or not b:
            return Poly([], self.R)
        q = self.q
        if len(b) == 1:
            s = b[0]
            return Poly(_strip([v * s % q for v in a]), self.R)
        if len(a) == 1:
            s = a[0]
            return Poly(_strip([v * s % q for v in b]), self.R)
        out = [0] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            if x:
                for j, y in enumerate(b):
                    out[i + j] += x * y
        return Poly(_strip([v % q for v in out]), self.R)

    __rmul__ = __mul__

    def __pow__(self, e):
        e = int(e)
        if e < 0:
            raise ValueError("negative power of a polynomial")
        if len(self.c) == 2 and self.c[0] == 0 and self.c[1] == 1:      # X**e
            return Poly([0] * e + [1], self.R)
        res, base = Poly([1], self.R), self
        while e:
            if e & 1:
                res = res * base
            e >>= 1
            if e:
                base = base * base
        return res

    def quo_rem(self, o):
        p = self._coerce(o)
        if p is None or not p.c:
            raise ZeroDivisionError("polynomial division by zero")
        q = self.q
        a = [x for x in self.c]
        b = p.c
        db = len(b) - 1
        if len(a) - 1 < db:
            return Poly([], self.R), Poly(a, self.R)
        inv = pow(b[-1], -1, q)
        quo = [0] * (len(a) - db)
        for i in range(len(a) - 1, db - 1, -1):
            coef = a[i] * inv % q
            if coef:
                quo[i - db] = coef
                for j in range(db + 1):
                    a[i - db + j] = (a[i - db + j] - coef * b[j]) % q
        return Poly(_strip(quo), self.R), Poly(_strip(a[:db]), self.R)

    def __floordiv__(self, o):
        return self.quo_rem(o)[0]

    def __mod__(self, o):
        return self.quo_rem(o)[1]

    def __truediv__(self, o):
        if isinstance(o, Poly):
            return Fraction(self, o)
        if isinstance(o, Fraction):
            return Fraction(self * o.den, o.num)
        p = self._coerce(o)
        if p is None:
            return NotImplemented
        if not p.c:
            raise ZeroDivisionError
        inv = pow(p.c[0], -1, self.q)
        return self * inv

    def __rtruediv__(self, o):
        p = self._coerce(o)
        if p is None:
            return NotImplemented
        return Fraction(p, self)

    def __eq__(self, o):
        if isinstance(o, Fraction):
            return o == self
        p = self._coerce(o)
        return NotImplemented if p is None else self.c == p.c

    def __ne__(self, o):
        r = self.__eq__(o)
        return r if r is NotImplemented else not r

    def __hash__(self):
        return hash(tuple(self.c))

    def __bool__(self):
        return bool(self.c)

    def derivative(self):
        q = self.q
        return Poly(_strip([i * v % q for i, v in enumerate(self.c)][1:]), self.R)

    def __repr__(self):
        if not self.c:
            return "0"
        terms = []
        for i in range(len(self.c) - 1, -1, -1):
            v = self.c[i]
            if not v:
                continue
            if i == 0:
                terms.append(str(v))
            else:
                x = "X" if i == 1 else f"X^{i}"
                terms.append(x if v == 1 else f"{v}*{x}")
        return " + ".join(terms)


class Fraction:
    """num/den of polynomials (Sage's fraction-field element as far as the provers use it)."""
    __slots__ = ("num", "den")

    def __init__(self, num, den):
        if not den.c:
            raise ZeroDivisionError("fraction with zero denominator")
        self.num, self.den = num, den

    def _wrap(self, o):
        if isinstance(o, Fraction):
            return o
        p = self.num._coerce(o)
        if p is None:
            return None
        return Fraction(p, Poly([1], self.num.R))

    def to_poly(self):
        quo, rem = self.num.quo_rem(self.den)
        if rem.c:
            raise TypeError("denominator does not divide numerator: not a polynomial")
        return quo

    def __add__(self, o):
        f = self._wrap(o)
        if f is None:
            return NotImplemented
        if f.den == self.den:
            return Fraction(self.num + f.num, self.den)
        return Fraction(self.num * f.den + f.num * self.den, self.den * f.den)

    __radd__ = __add__

    def __neg__(self):
        return Fraction(-self.num, self.den)

    def __sub__(self, o):
        f = self._wrap(o)
        return NotImplemented if f is None else self + (-f)

    def __rsub__(self, o):
        f = self._wrap(o)
        return NotImplemented if f is None else f + (-self)

    def __mul__(self, o):
        f = self._wrap(o)
        if f is None:
            return NotImplemented
        return Fraction(self.num * f.num, self.den * f.den)

    __rmul__ = __mul__

    def __truediv__(self, o):
        f = self._wrap(o)
        if f is None:
            return NotImplemented
        return Fraction(self.num * f.den, self.den * f.num)

    def __rtruediv__(self, o):
        f = self._wrap(o)
        return NotImplemented if f is None else f / self

    def __call__(self, x):
        return self.num(x) / self.den(x)

    def __eq__(self, o):
        f = self._wrap(o)
        return NotImplemented if f is None else self.num * f.den == f.num * self.den

    def __hash__(self):
        return hash((self.num, self.den))

    def numerator(self):
        return self.num

    def denominator(self):
        return self.den


class PolyRingShim:
    _cache = {}

    def __new__(cls, F, name="X"):
        key = (F.q, name)
        if key not in cls._cache:
            obj = super().__new__(cls)
            obj.F = F
            obj.name = name
            cls._cache[key] = obj
        return cls._cache[key]

    def __call__(self, x=0):
        q = self.F.q
        if isinstance(x, Poly):
            return Poly(list(x.c), self)
        if isinstance(x, Fraction):
            return x.to_poly()
        if isinstance(x, (list, tuple)):
            return Poly(_strip([int(v) % q for v in x]), self)
        v = int(x) % q
        return Poly([v] if v else [], self)

    def gen(self):
        return Poly([0, 1], self)

    def base_ring(self):
        return self.F

    def zero(self):
        return Poly([], self)

    def one(self):
        return Poly([1], self)

    def lagrange_polynomial(self, points):
        """Interpolating polynomial through (x_i, y_i) (O(n^2); callers use it for <= n points)."""
        q = self.F.q
        xs = [int(x) % q for x, _ in points]
        ys = [int(y) % q for _, y in points]
        n = len(xs)
        # master polynomial prod (X - x_i)
        master = [1]
        for x in xs:
            nxt = [0] * (len(master) + 1)
            for i, v in enumerate(master):
                nxt[i + 1] = (nxt[i + 1] + v) % q
                nxt[i] = (nxt[i] - v * x) % q
            master = nxt
        res = [0] * n
        for i in range(n):
            # synthetic division of master by (X - x_i)
            quo = [0] * n
            quo[n - 1] = master[n]
            for j in range(n - 1, 0, -1):
                quo[j - 1] = (master[j] + xs[i] * quo[j]) % q
            denom = 0
            for v in reversed(quo):
                denom = (denom * xs[i] + v) % q
            s = ys[i] * pow(denom, -1, q) % q
            if s:
                for j in range(n):
                    res[j] = (res[j] + s * quo[j]) % q
        return Poly(_strip(res), self)

    def __repr__(self):
        return f"Univariate Polynomial Ring in {self.name} over {self.F!r}"

    def __eq__(self, o):
        return isinstance(o, PolyRingShim) and o.F == self.F and o.name == self.name

    def __hash__(self):
        return hash(("PolyRingShim", self.F.q, self.name))


def PolynomialRing(F, name="X"):
    return PolyRingShim(F, name)


# --------------------------------------------------------------------------- vector / matrix / prod
def prod(items, start=1):
    """sage.all.prod: product of an iterable (marlin/prover.py:63-64, marlin/encoder.py:158)."""
    out = start
    for it in items:
        out = out * it
    return out


class Vector:
    """vector(F, entries): only what marlin/encoder.py:202-207 needs (A * z, indexing, len)."""

    def __init__(self, F, entries):
        self.F = F
        self.v = [F(e) for e in entries]

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return self.v[i]

    def __iter__(self):
        return iter(self.v)

    def list(self):
        return list(self.v)

    def __repr__(self):
        return "(" + ", ".join(repr(e) for e in self.v) + ")"


def vector(F, entries=None):
    if entries is None:
        F, entries = entries, F
        F = entries[0].parent()
    return Vector(F, entries)


class Matrix:
    """Dense matrix over GF(q) with the methods the Marlin indexer/encoder call: nrows, ncols,
    [i, j] / [i][j] access, .T, * vector, rows(), columns() (marlin/indexer.py:48-52,
    marlin/encoder.py:37,98-125,205-207)."""

    def __init__(self, F, rows):
        self.F = F
        self.r = [[F(e) for e in row] for row in rows]

    def nrows(self):
        return len(self.r)

    def ncols(self):
        return len(self.r[0]) if self.r else 0

    def dimensions(self):
        return (self.nrows(), self.ncols())

    def nonzero_positions(self):
        """(i, j) of non-zero entries in row-major order, as Sage returns them (marlin/encoder.py:40,101)."""
        return [(i, j) for i, row in enumerate(self.r) for j, e in enumerate(row) if e.n]

    def base_ring(self):
        return self.F

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            i, j = idx
            if isinstance(i, slice) or isinstance(j, slice):
                rows = self.r[i] if isinstance(i, slice) else [self.r[i]]
                sub = [row[j] if isinstance(j, slice) else [row[j]] for row in rows]
                return Matrix(self.F, sub)
            return self.r[i][j]
        return self.r[idx]

    def __setitem__(self, idx, val):
        if isinstance(idx, tuple):
            i, j = idx
            if isinstance(i, slice) and not isinstance(j, slice):      # M[:, j] = column
                col = val.column(0) if isinstance(val, Matrix) else list(val)
                for k, row in enumerate(self.r[i]):
                    row[j] = self.F(col[k])
                return
            self.r[i][j] = self.F(val)
        else:
            self.r[idx] = [self.F(e) for e in val]

    @property
    def T(self):
        return self.transpose()

    def transpose(self):
        return Matrix(self.F, [list(c) for c in zip(*self.r)]) if self.r else Matrix(self.F, [])

    def rows(self):
        return [list(r) for r in self.r]

    def row(self, i):
        return list(self.r[i])

    def columns(self):
        return [list(c) for c in zip(*self.r)]

    def column(self, j):
        return [row[j] for row in self.r]

    def __copy__(self):
        return Matrix(self.F, self.r)

    def __mul__(self, o):
        if isinstance(o, Vector):
            q = self.F.q
            ov = [e.n for e in o.v]
            return Vector(self.F, [sum(a.n * b for a, b in zip(row, ov) if a.n) % q for row in self.r])
        if isinstance(o, Matrix):
            q = self.F.q
            cols = o.columns()
            return Matrix(self.F, [[sum(a.n * b.n for a, b in zip(row, col)) % q for col in cols] for row in self.r])
        return Matrix(self.F, [[e * o for e in row] for row in self.r])

    __rmul__ = __mul__

    def __eq__(self, o):
        return isinstance(o, Matrix) and self.r == o.r

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return "\n".join("[" + " ".join(repr(e) for e in row) + "]" for row in self.r)


def matrix(F, *args):
    """matrix(F, rows) / matrix(F, nrows, ncols, flat_or_rows)."""
    if len(args) == 1:
        return Matrix(F, args[0])
    nr, nc, data = args
    if data and not isinstance(data[0], (list, tuple)):
        data = [data[i * nc:(i + 1) * nc] for i in range(nr)]
    return Matrix(F, data)
