#!/usr/bin/env python3
"""A/B of the host-scalar MSM entry point (kzgpu_msm: H2D inside) on one box: uniform vs growing
upload chunks, plus the raw pinned H2D rate, at 2^24 BN254."""
import os
import subprocess
import sys
import time

if len(sys.argv) == 1:
    for mode in ("uniform", "geometric", "uniform", "geometric"):
        env = dict(os.environ, KZGPU_MSM_CHUNKS=mode)
        r = subprocess.run([sys.executable, __file__, mode], env=env, capture_output=True, text=True)
        print(r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
    sys.exit(0)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi, device                      # noqa: E402
from kzg_snark_b200.limbs import random_scalars              # noqa: E402
n = 1 << 24
_ffi.init()
pin = _ffi.PinnedArray((n, 4))
pin.array[:] = random_scalars(n, device.FR[0], seed=3)
d = _ffi.DeviceBuffer(n * 32)
srs = device.Srs.generate(0, 0x123456789abcdef, n)
d.upload(pin.array)
t0 = time.perf_counter()
for _ in range(5):
    d.upload(pin.array)
h2d = (time.perf_counter() - t0) / 5
for _ in range(3):
    device.msm(srs, pin.array)
t0 = time.perf_counter()
for _ in range(10):
    device.msm(srs, pin.array)
e2e = (time.perf_counter() - t0) / 10
for _ in range(2):
    device.msm_dev(srs, d, n)
_ffi.timer_start()
for _ in range(10):
    device.msm_dev(srs, d, n)
res = _ffi.timer_stop() / 10
print(f"{sys.argv[1]:10s} e2e {1e3 * e2e:.2f} ms | resident {res:.2f} ms | H2D 512 MiB {1e3 * h2d:.2f} ms ({n * 32 / h2d / 1e9:.1f} GB/s)")
