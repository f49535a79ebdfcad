#!/usr/bin/env python3
"""A/B of the MSM sort/accumulate pipeline (sort of chunk k+1 on a second stream while chunk k is accumulated):
KZGPU_MSM_NO_PIPE=1 against the default and a few first-chunk fractions, device-resident and host-scalar entry
points, at 2^LOGN BN254 (argv[1], default 24).  Every mode must return the same point."""
import os
import subprocess
import sys
import time

if len(sys.argv) <= 2:
    logn = sys.argv[1] if len(sys.argv) == 2 else "24"
    for mode in ("nopipe", "8", "4", "6", "12", "16", "nopipe", "8"):
        env = dict(os.environ)
        if mode == "nopipe":
            env["KZGPU_MSM_NO_PIPE"] = "1"
        else:
            env["KZGPU_MSM_DEV_SPLIT"] = mode
        r = subprocess.run([sys.executable, __file__, logn, mode], env=env, capture_output=True, text=True)
        print(r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
    sys.exit(0)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi, device                      # noqa: E402
from kzg_snark_b200.limbs import random_scalars              # noqa: E402
n = 1 << int(sys.argv[1])
_ffi.init()
pin = _ffi.PinnedArray((n, 4))
pin.array[:] = random_scalars(n, device.FR[0], seed=3)
d = _ffi.DeviceBuffer(n * 32)
srs = device.Srs.generate(0, 0x123456789abcdef, n)
d.upload(pin.array)
for _ in range(3):
    out_h = device.msm(srs, pin.array)
t0 = time.perf_counter()
for _ in range(10):
    device.msm(srs, pin.array)
e2e = (time.perf_counter() - t0) / 10
for _ in range(2):
    out_d = device.msm_dev(srs, d, n)
_ffi.timer_start()
for _ in range(10):
    device.msm_dev(srs, d, n)
res = _ffi.timer_stop() / 10
print(f"{sys.argv[2]:8s} e2e {1e3 * e2e:.2f} ms | resident {res:.2f} ms | same point: {bool((out_h[0] == out_d[0]).all())} | x limb0 = {int(out_d[0].ravel()[0]):016x}")
