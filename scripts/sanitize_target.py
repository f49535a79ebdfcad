"""Small, fast coverage of every kernel family for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python scripts/sanitize_target.py
MSM (fused short-key tables, per-table tables, plain key, batch pass, chunked host upload, skewed scalars), open, NTT
(generic and compile-time-shape passes, inverse, coset, batch), SRS generation, verifier-side combination; both curves."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars, ints_to_limbs

_ffi.init()
for cid in (0, 1):
    r = device.FR[cid]
    for n in (1, 37, 1 << 10, (1 << 14) + 5):
        srs = device.Srs.generate(cid, 0x1234567, n)
        sc = random_scalars(n, r, seed=n)
        sc[::7] = 0
        device.msm(srs, sc)
        device.msm_batch(srs, [sc, sc[: n // 2], sc[:1]])
        z, xi = ints_to_limbs([5], r)[0], ints_to_limbs([9], r)[0]
        device.open_proof(srs, [sc, sc[: n // 3 + 1]], z, xi)
        srs.destroy()
    for logn in (1, 5, 10, 13, 16):
        n = 1 << logn
        w = ints_to_limbs([pow(5 if cid == 0 else 7, (r - 1) // n, r)], r)[0]
        x = random_scalars(n, r, seed=logn)
        y = device.ntt(cid, x.copy(), w)
        back = device.ntt(cid, y.copy(), w, inverse=True)
        assert (back == x).all()
        device.ntt(cid, x.copy(), w, coset_limbs=ints_to_limbs([7], r)[0])
        if logn <= 10:
            device.ntt(cid, np.concatenate([x, x, x]), w, batch=3)
# long enough for the chunked / pipelined host paths and the 2^20-class kernels
srs = device.Srs.generate(0, 0x1234567, 1 << 20)
sc = random_scalars(1 << 20, device.FR[0], seed=3)
device.msm(srs, sc)
w = ints_to_limbs([pow(5, (device.FR[0] - 1) >> 20, device.FR[0])], device.FR[0])[0]
device.ntt(0, sc.copy(), w)
pts = srs.read(0, 4)
device.g1_lincomb(0, pts, sc[:4])
srs.destroy()
print("sanitize target ok, launches:", _ffi.launch_count())
