"""Minimal GF(q) element type on CPython ints.

Oracle / test infrastructure only (see oracle/__init__.py).

Stands in for Sage's `GF(curve_order)` elements (reference kzg.py:52) with just
the operations the hot path uses on them: `+ - * ** == int()`, `F(0)`, `F(1)`,
`F(n)` (fft_ff.py:29-30,57), `w**(-1)` (fft_ff.py:53) and
`g.multiplicative_order()` (fft_ff.py:77).
"""


class Fe:
    __slots__ = ("n", "F")

    def __init__(self, n, F):
        self.n = n % F.q
        self.F = F

    def _c(self, o):
        if isinstance(o, Fe):
            return o.n
        return int(o) % self.F.q

    def __add__(self, o):
        return Fe(self.n + self._c(o), self.F)

    __radd__ = __add__

    def __sub__(self, o):
        return Fe(self.n - self._c(o), self.F)

    def __rsub__(self, o):
        return Fe(self._c(o) - self.n, self.F)

    def __mul__(self, o):
        return Fe(self.n * self._c(o), self.F)

    __rmul__ = __mul__

    def __neg__(self):
        return Fe(-self.n, self.F)

    def __pow__(self, e):
        return Fe(pow(self.n, int(e), self.F.q), self.F)   # negative e -> modular inverse power

    def __truediv__(self, o):
        return Fe(self.n * pow(self._c(o), -1, self.F.q), self.F)

    def __eq__(self, o):
        try:
            return self.n == self._c(o)
        except (TypeError, ValueError):
            return NotImplemented

    def __hash__(self):
        return hash(self.n)

    def __int__(self):
        return self.n

    __index__ = __int__

    def __repr__(self):
        return str(self.n)        # Sage prints a residue as its integer representative

    def multiplicative_order(self):
        """Order of the element; factor-free for elements of 2-power order (all callers),
        falls back to a divisor search over the known factorisation otherwise."""
        assert self.n != 0
        q = self.F.q
        # 2-power order?
        x, k = self.n, 0
        while x != 1 and k <= 64:
            x = x * x % q
            k += 1
        if x == 1:
            return 1 << k
        return q - 1      # upper bound; only `order >= n` is ever asserted (fft_ff.py:78)


class GFp:
    """Callable coercion `F(x)` like Sage's GF(q)."""

    def __init__(self, q):
        self.q = q

    def __call__(self, x=0):
        if isinstance(x, Fe):
            return Fe(x.n, self)
        return Fe(int(x), self)

    def order(self):
        return self.q

    def random_element(self, rng=None):
        import random
        return Fe((rng or random).randrange(self.q), self)

    def __eq__(self, o):
        return isinstance(o, GFp) and o.q == self.q

    def __hash__(self):
        return hash(("GFp", self.q))
