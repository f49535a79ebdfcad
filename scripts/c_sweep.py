#!/usr/bin/env python3
"""Calibrates the table-window cost model: MSM time at 2^logn for several forced window sizes
(KZGPU_SRS_TABLES=c=NN), one process per setting.  usage: c_sweep.py logn c1 c2 ..."""
import os
import subprocess
import sys

if len(sys.argv) > 2 and sys.argv[1] != "--one":
    logn = sys.argv[1]
    for c in sys.argv[2:]:
        env = dict(os.environ, KZGPU_SRS_TABLES=f"c={c}")
        r = subprocess.run([sys.executable, __file__, "--one", logn], env=env, capture_output=True, text=True)
        print(f"logn={logn} c={c}: {r.stdout.strip()} {r.stderr.strip()[-200:]}", flush=True)
    sys.exit(0)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi, device                      # noqa: E402
from kzg_snark_b200.limbs import random_scalars              # noqa: E402
logn = int(sys.argv[2])
n = 1 << logn
_ffi.init()
d = _ffi.DeviceBuffer(n * 32).upload(random_scalars(n, device.FR[0], seed=1))
srs = device.Srs.generate(0, 0x123456789abcdef, n)
for _ in range(3):
    device.msm_dev(srs, d, n)
_ffi.timer_start()
for _ in range(10):
    device.msm_dev(srs, d, n)
print(f"{_ffi.timer_stop() / 10:.3f} ms", srs.info())
