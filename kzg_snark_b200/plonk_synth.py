"""Synthetic PLONK circuits of arbitrary size in the reference's arithmetisation format
(constraint-system/PLONK_ARITHMETIZATION_INSTANCE.pkl: selector value lists qM qL qR qO qC of
length n, a wire permutation `perm` of 0..3n-1 and the wire values w = a | b | c; main.py:68-79).

The bundled instance has 16 gates; BASELINE.json's PLONK config needs circuits large enough for
the GPU to matter.  Gates: the first `n_pub` are public-input gates (qL = 1, a_i = x_i,
plonk/encoder.py:218-223), the rest are random multiplication / addition gates whose inputs
are, with probability 1/2 each, copies of earlier outputs -- which is what makes the permutation
non-trivial."""
import random

import numpy as np


def synthetic_circuit(n, n_pub, r, seed=0):
    """Returns (qM, qL, qR, qO, qC, perm, w) with Python-int lists (values reduced mod r)."""
    assert n & (n - 1) == 0 and 0 < n_pub < n
    rng = random.Random(seed)
    qM, qL, qR, qO, qC = ([0] * n for _ in range(5))
    a, b, c = [0] * n, [0] * n, [0] * n
    var = np.arange(3 * n, dtype=np.int64)             # variable id of each wire position (own id = unconstrained)
    for i in range(n_pub):
        qL[i] = 1
        a[i] = rng.randrange(r)
    for i in range(n_pub, n):
        for side, vals in ((0, a), (1, b)):
            if i > n_pub and rng.random() < 0.5:       # copy an earlier gate's output
                j = rng.randrange(n_pub, i)
                vals[i] = c[j]
                var[side * n + i] = var[2 * n + j]
            elif rng.random() < 0.25:                  # or a public input
                j = rng.randrange(n_pub)
                vals[i] = a[j]
                var[side * n + i] = var[j]
            else:
                vals[i] = rng.randrange(r)
        kind = rng.randrange(3)
        if kind == 0:                                  # a * b - c = 0
            qM[i], qO[i] = 1, r - 1
            c[i] = a[i] * b[i] % r
        elif kind == 1:                                # a + b - c = 0
            qL[i], qR[i], qO[i] = 1, 1, r - 1
            c[i] = (a[i] + b[i]) % r
        else:                                          # 3 a b + 2 a - b + k - c = 0
            k = rng.randrange(r)
            qM[i], qL[i], qR[i], qO[i], qC[i] = 3, 2, r - 1, r - 1, k
            c[i] = (3 * a[i] * b[i] + 2 * a[i] - b[i] + k) % r
    # permutation: one cycle per variable (positions with equal id, rotated by one)
    order = np.argsort(var, kind="stable")
    sv = var[order]
    first = np.ones(3 * n, dtype=bool)
    first[1:] = sv[1:] != sv[:-1]
    start = np.maximum.accumulate(np.where(first, np.arange(3 * n), 0))
    nxt = np.empty(3 * n, dtype=np.int64)
    nxt[:-1] = order[1:]
    last = np.ones(3 * n, dtype=bool)
    last[:-1] = first[1:]
    nxt[last] = order[start[last]]
    perm = np.empty(3 * n, dtype=np.int64)
    perm[order] = nxt
    return qM, qL, qR, qO, qC, perm.tolist(), a + b + c
