#!/usr/bin/env python3
"""A/B of the host-buffer NTT entry point (kzgpu_ntt: H2D + passes + D2H inside) on one box:
plain copy-compute-copy (KZGPU_NTT_NO_OVERLAP=1) vs slab-pipelined transfers, 2^24 BN254."""
import os
import subprocess
import sys
import time

if len(sys.argv) == 1:
    for mode in ("plain", "overlap", "strict", "plain", "overlap", "strict"):
        env = dict(os.environ)
        if mode == "plain":
            env["KZGPU_NTT_NO_OVERLAP"] = "1"
        if mode == "strict":                                  # overlap, canonical values between passes
            env["KZGPU_NTT_STRICT"] = "1"
        r = subprocess.run([sys.executable, __file__, mode], env=env, capture_output=True, text=True)
        print(r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
    sys.exit(0)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                            # noqa: E402
from kzg_snark_b200 import _ffi, device                      # noqa: E402
from kzg_snark_b200.limbs import random_scalars, ints_to_limbs  # noqa: E402
R = device.FR[0]
n = 1 << 24
_ffi.init()
pin = _ffi.PinnedArray((n, 4))
x = random_scalars(n, R, seed=5)
pin.array[:] = x
wl = ints_to_limbs([pow(5, (R - 1) // n, R)], R)[0]
d = _ffi.DeviceBuffer(n * 32).upload(x)
device.ntt_dev(0, d, n, wl)
ref = np.zeros((n, 4), dtype=np.uint64)
d.download(ref)
device.ntt(0, pin.array, wl)
ok = bool((pin.array == ref).all())
for _ in range(2):
    device.ntt(0, pin.array, wl)
t0 = time.perf_counter()
for _ in range(10):
    device.ntt(0, pin.array, wl)
e2e = (time.perf_counter() - t0) / 10
_ffi.timer_start()
for _ in range(10):
    device.ntt_dev(0, d, n, wl)
res = _ffi.timer_stop() / 10
print(f"{sys.argv[1]:8s} e2e {1e3 * e2e:.2f} ms | resident {res:.2f} ms | host result == device result: {ok}")
