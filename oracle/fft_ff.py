"""Restated oracle of the reference's fft_ff.py (radix-2 NTT / iNTT).

Oracle / test infrastructure only (see oracle/__init__.py).  Parity is PINNED: the reference
ships no test or vector for this module (fft_ff.py:1-85), so its own fft_ff.py is run
unmodified (oracle/refrun.py; BN254 and BLS12-381 scalar fields) and every call is recorded in
tests/golden/ref_trace_fft*.json; this restatement reproduces all of them bit for bit
(tests/test_reference_traces.py) and also satisfies the DFT definition
out[k] = sum_j c[j] * w^(j*k) (tests/test_oracle.py).

Two flavours of the same recursion:
  * fft_ff / ifft_ff / fft_ff_interpolation -- element-generic, statement for
    statement after fft_ff.py:3-37, 39-58, 60-85 (works on oracle.field.Fe or any
    type with + - * **);
  * fft_ff_int / ifft_ff_int -- identical recursion on bare ints mod q, ~10x
    faster; used as the CPU timing baseline and for larger parity sizes.
  * coset_fft_ff_int -- the north_star's coset variant, DEFINED by composition
    with fft_ff: the reference has no coset FFT.
"""


def fft_ff(coeffs, w, F):
    """fft_ff.py:3-37.  Natural order in / natural order out, caller-supplied root."""
    n = len(coeffs)
    if n == 1:
        return coeffs                       # fft_ff.py:16-17: the SAME list object
    even = coeffs[0::2]                     # fft_ff.py:20-21
    odd = coeffs[1::2]
    w_squared = w ** 2                      # fft_ff.py:24
    even_fft = fft_ff(even, w_squared, F)
    odd_fft = fft_ff(odd, w_squared, F)
    result = [F(0)] * n                     # fft_ff.py:29
    w_power = F(1)
    for i in range(n // 2):                 # fft_ff.py:32-35
        result[i] = even_fft[i] + w_power * odd_fft[i]
        result[i + n // 2] = even_fft[i] - w_power * odd_fft[i]
        w_power *= w
    return result


def ifft_ff(values, w, F):
    """fft_ff.py:39-58."""
    n = len(values)
    w_inv = w ** (-1)                       # fft_ff.py:53
    result = fft_ff(values, w_inv, F)
    n_inv = F(n) ** (-1)                    # fft_ff.py:57
    return [x * n_inv for x in result]


def fft_ff_interpolation(values, g, F):
    """fft_ff.py:60-85.  Returns the coefficient list low->high with trailing zeros
    stripped, i.e. what Sage's PolynomialRing(F,'X')(coeffs).list() would hold
    (the Sage polynomial wrapper itself is the drop-in module's business)."""
    n = len(values)
    assert (n & (n - 1)) == 0, "Length of values must be a power of 2"   # fft_ff.py:74
    order = g.multiplicative_order()                                     # fft_ff.py:77
    assert order >= n, f"Order of g ({order}) must be at least n ({n})"  # fft_ff.py:78
    coeffs = ifft_ff(values, g, F)
    coeffs = list(coeffs)
    while coeffs and coeffs[-1] == 0:
        coeffs.pop()
    return coeffs


# ---------------------------------------------------------------- int flavour
def fft_ff_int(coeffs, w, q):
    """Same recursion as fft_ff (fft_ff.py:16-37) on ints mod q."""
    n = len(coeffs)
    if n == 1:
        return coeffs
    even_fft = fft_ff_int(coeffs[0::2], w * w % q, q)
    odd_fft = fft_ff_int(coeffs[1::2], w * w % q, q)
    result = [0] * n
    w_power = 1
    h = n // 2
    for i in range(h):
        t = w_power * odd_fft[i] % q
        e = even_fft[i]
        result[i] = (e + t) % q
        result[i + h] = (e - t) % q
        w_power = w_power * w % q
    return result


def ifft_ff_int(values, w, q):
    """fft_ff.py:39-58 on ints."""
    n = len(values)
    w_inv = pow(w, -1, q)
    result = fft_ff_int(values, w_inv, q)
    n_inv = pow(n % q, -1, q)
    return [x * n_inv % q for x in result]


def coset_fft_ff_int(coeffs, w, shift, q):
    """coset_fft(c, w, s)[k] = sum_j c[j] * s^j * w^(jk)  ==  fft_ff([c[j]*s**j], w, F)
    (no reference counterpart)."""
    out, sp = [], 1
    for c in coeffs:
        out.append(c * sp % q)
        sp = sp * shift % q
    return fft_ff_int(out, w, q)


def coset_ifft_ff_int(values, w, shift, q):
    """Inverse of coset_fft_ff_int: ifft then multiply coefficient j by shift^-j."""
    c = ifft_ff_int(values, w, q)
    si = pow(shift, -1, q)
    out, sp = [], 1
    for x in c:
        out.append(x * sp % q)
        sp = sp * si % q
    return out


def dft_definition(coeffs, w, q, ks=None):
    """O(n) per output reference: out[k] = Horner(c, w^k).  Pins the recursion."""
    n = len(coeffs)
    ks = range(n) if ks is None else ks
    res = []
    for k in ks:
        x = pow(w, k, q)
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % q
        res.append(acc)
    return res
