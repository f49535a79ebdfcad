"""Worker for tests/test_gpu_multi.py (one library context per process): runs a fixed set of hot-path calls from a PLAIN
python process -- no torch, no launcher -- and prints the results as JSON.  With KZGPU_DEVICES=all the library spreads
the work over every visible GPU (kzgpu_init_multi); without it everything runs on one.  Both must print the same thing."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                              # noqa: E402
from kzg_snark_b200 import _ffi, device                         # noqa: E402
from kzg_snark_b200.limbs import random_scalars, ints_to_limbs  # noqa: E402

R = device.FR[0]
TAU = 0x1D2C3B4A5F6E7D8C9BA % R


def hx(a):
    return np.ascontiguousarray(a).tobytes().hex()


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 14       # log2 of the base polynomial length
    _ffi.init()
    out = {"ndev": _ffi.device_count()}
    n = 1 << scale
    # the 11 commitments of one Marlin proof (marlin/prover.py:106,142,176), scaled: lengths n .. 12 n
    lens = [n + 2] * 4 + [n + 4, 2 * n + 1] + [n, n - 1, n + 2] + [2 * n - 1, 12 * n - 6]
    srs = device.Srs.generate("bn254", TAU, 12 * n)
    polys = [random_scalars(ln, R, seed=50 + j) for j, ln in enumerate(lens)]
    pts, infs = device.msm_batch(srs, polys)                    # KZG.commit(ck, [p_1 .. p_11]) at the buffer level
    out["commits"] = [hx(p) for p in pts]
    out["infs"] = [bool(f) for f in infs]
    # one long MSM from host scalars (point-sharded when several devices), also with an offset into the key
    big = random_scalars(8 * n, R, seed=7)
    o, f = device.msm(srs, big)
    out["msm"] = hx(o)
    o, f = device.msm(srs, big[: 5 * n + 3], first=n + 1)
    out["msm_offset"] = hx(o)
    # the same from device-resident scalars on the primary device (peers pull their slices)
    d = _ffi.DeviceBuffer(big.nbytes).upload(big)
    o, f = device.msm_dev(srs, d, 8 * n)
    out["msm_dev"] = hx(o)
    # batched open (kzg.py:122-159): quotient on the primary device, its MSM over all devices
    z, xi = ints_to_limbs([0x1234567], R)[0], ints_to_limbs([0x7654321], R)[0]
    o, f = device.open_proof(srs, [polys[10], polys[5], polys[0]], z, xi)
    out["open"] = hx(o)
    # batched NTTs: whole vectors per device
    m = 1 << (scale + 1)
    w = ints_to_limbs([pow(5, (R - 1) // m, R)], R)[0]
    vecs = random_scalars(9 * m, R, seed=3)
    y = device.ntt("bn254", vecs.copy(), w, batch=9)
    out["ntt"] = hx(y[::997])
    back = device.ntt("bn254", y.copy(), w, inverse=True, batch=9)
    out["ntt_roundtrip"] = bool((back == vecs).all())
    # the Python-object level: KZG.setup / commit / open exactly as the reference's callers use them (kzg.py:56-159), k polynomials per call
    if scale <= 12:
        from kzg_snark_b200.kzg import KZG
        kzg = KZG("bn254")
        ck, _ = kzg.setup(12 * n - 1, tau=TAU)
        F = kzg.Fq
        pys = [kzg.R([F(int(v)) for v in np.asarray(p_)[:, 0]]) for p_ in polys]        # 64-bit coefficients keep the conversion quick
        comm = kzg.commit(ck, pys)
        out["kzg_commit"] = [[int(c) for c in pt] for pt in comm]
        out["kzg_open"] = [int(c) for c in kzg.open(ck, pys[:4], 7, 42)]
    # the device Marlin prover (marlin/prover.py:25-245): every round's commitments and both openings use all devices
    from kzg_snark_b200 import marlin
    rows = 1 << max(scale - 2, 8)
    A, B, C, x, wit = marlin.synthetic_r1cs(rows, 8, R, seed=scale)
    mK = 1 << (2 * rows - 1).bit_length()
    idx = marlin.Indexer("bn254")
    ipk, _ = idx.preprocess(A, B, C, max_degree=6 * mK, tau=TAU)
    pr = marlin.Prover("bn254")
    proof = pr.prove(ipk, [idx.kzg.Fq(v) for v in x], ints_to_limbs(wit, R), draws=random_scalars(8 + 2 * rows + 1, R, seed=4))
    assert set(pr.checks.values()) == {0}
    out["marlin"] = {"commitments": {k: [[int(c) for c in p] for p in v] for k, v in proof["commitments"].items()},
                     "kzg_proofs": {k: [int(c) for c in v] for k, v in proof["kzg_proofs"].items()},
                     "evaluations": {k: [int(e) for e in v] for k, v in proof["evaluations"].items()}}
    out["launches"] = _ffi.launch_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
