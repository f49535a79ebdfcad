"""Python int / field element  <->  fixed-width little-endian limb arrays (numpy uint64).

The C ABI takes canonical residues as uint64 limbs (include/kzgpu.h).  Anything that
supports int() is accepted on the way in (Sage IntegerMod, py_ecc FQ, our shim elements,
plain ints) -- the duck-typing the reference's callers rely on.
"""

import numpy as np


def ints_to_limbs(values, modulus, nlimbs=4):
    """list of int-likes -> (len, nlimbs) uint64, each reduced mod `modulus`."""
    n = len(values)
    nbytes = nlimbs * 8
    buf = bytearray(n * nbytes)
    off = 0
    for v in values:
        buf[off:off + nbytes] = (int(v) % modulus).to_bytes(nbytes, "little")
        off += nbytes
    return np.frombuffer(bytes(buf), dtype="<u8").reshape(n, nlimbs).copy()


def int_to_limbs(v, modulus, nlimbs=4):
    return np.frombuffer((int(v) % modulus).to_bytes(nlimbs * 8, "little"), dtype="<u8").copy()


def limbs_to_ints(arr):
    """(len, nlimbs) uint64 -> list of Python ints."""
    a = np.ascontiguousarray(arr, dtype="<u8")
    if a.ndim == 1:
        a = a.reshape(1, -1)
    nbytes = a.shape[1] * 8
    raw = a.tobytes()
    return [int.from_bytes(raw[i:i + nbytes], "little") for i in range(0, len(raw), nbytes)]


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
