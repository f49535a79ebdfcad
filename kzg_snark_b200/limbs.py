"""Python int / field element  <->  fixed-width little-endian limb arrays (numpy uint64).

The C ABI takes canonical residues as uint64 limbs (include/kzgpu.h).  Anything that
supports int() is accepted on the way in (Sage IntegerMod, py_ecc FQ, our shim elements,
plain ints) -- the duck-typing the reference's callers rely on.
"""

import operator

import numpy as np


_TO_BYTES = {}


def _packer(nbytes):
    f = _TO_BYTES.get(nbytes)
    if f is None:
        f = _TO_BYTES[nbytes] = operator.methodcaller("to_bytes", nbytes, "little")
    return f


def ints_to_limbs(values, modulus, nlimbs=4):
    """list of int-likes -> (len, nlimbs) uint64, each reduced mod `modulus`.

    This is the Python-object boundary of KZG.commit / open / fft_ff (kzg.py:110,115: `poly.list()` then `int(coeff)`), so it
    is written for throughput: plain ints and field elements that expose their canonical residue as `.n` (the Sage-free
    shim, py_ecc's FQ) are packed with one C-level `to_bytes` per element and a single join (~0.1 us per element instead
    of ~1 us for a per-element slice assignment); anything else goes through int().  Values outside [0, modulus) are
    found with one vectorised comparison on the packed limbs and reduced individually."""
    n = len(values)
    nbytes = nlimbs * 8
    if n == 0:
        return np.zeros((0, nlimbs), dtype=np.uint64)
    first = values[0]
    if type(first) is int:
        ints = values
    elif hasattr(first, "n") and not callable(first.n):
        try:
            ints = [v.n for v in values]
        except AttributeError:                              # mixed list
            ints = [int(v) for v in values]
    else:
        ints = [int(v) for v in values]
    try:
        raw = b"".join(map(_packer(nbytes), ints))
    except (OverflowError, AttributeError, TypeError):      # negative, wider than nbytes, or not all plain ints
        ints = [int(v) % modulus for v in ints]
        raw = b"".join(map(_packer(nbytes), ints))
    arr = np.frombuffer(raw, dtype="<u8").reshape(n, nlimbs).copy()
    # rows >= modulus (lexicographic compare, most significant limb first)
    ge = np.ones(n, dtype=bool)
    decided = np.zeros(n, dtype=bool)
    for i in range(nlimbs - 1, -1, -1):
        m = np.uint64((modulus >> (64 * i)) & 0xFFFFFFFFFFFFFFFF)
        lt = ~decided & (arr[:, i] < m)
        gt = ~decided & (arr[:, i] > m)
        ge[lt] = False
        decided |= lt | gt
    if ge.any():
        for k in np.nonzero(ge)[0]:
            arr[k] = np.frombuffer((int(ints[k]) % modulus).to_bytes(nbytes, "little"), dtype="<u8")
    return arr


def int_to_limbs(v, modulus, nlimbs=4):
    return np.frombuffer((int(v) % modulus).to_bytes(nlimbs * 8, "little"), dtype="<u8").copy()


def limbs_to_ints(arr):
    """(len, nlimbs) uint64 -> list of Python ints."""
    a = np.ascontiguousarray(arr, dtype="<u8")
    if a.ndim == 1:
        a = a.reshape(1, -1)
    nbytes = a.shape[1] * 8
    raw = memoryview(a.tobytes())
    fb = int.from_bytes
    return [fb(raw[i:i + nbytes], "little") for i in range(0, len(raw), nbytes)]


def limbs_to_int(arr):
    return int.from_bytes(np.ascontiguousarray(arr, dtype="<u8").tobytes(), "little")


def random_scalars(n, modulus, seed, nlimbs=4):
    """n residues uniform in [0, modulus) as a (n, nlimbs) uint64 array, generated with numpy
    (PCG64, rejection of values >= modulus) -- the synthetic inputs of SURVEY.md section 8d."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bits = modulus.bit_length()
    top_mask = (1 << (bits - 64 * (nlimbs - 1))) - 1
    mod_limbs = [(modulus >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(nlimbs)]
    out = np.empty((n, nlimbs), dtype=np.uint64)
    todo = np.arange(n)
    while todo.size:
        cand = rng.integers(0, 1 << 64, size=(todo.size, nlimbs), dtype=np.uint64, endpoint=False)
        cand[:, nlimbs - 1] &= np.uint64(top_mask)
        # lexicographic compare (most significant limb first): keep cand < modulus
        lt = np.zeros(todo.size, dtype=bool)
        eq = np.ones(todo.size, dtype=bool)
        for i in range(nlimbs - 1, -1, -1):
            m = np.uint64(mod_limbs[i])
            lt |= eq & (cand[:, i] < m)
            eq &= cand[:, i] == m
        out[todo[lt]] = cand[lt]
        todo = todo[~lt]
    return out
