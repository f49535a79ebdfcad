"""Short single-GPU target for ncu: MSMs only, at 2^logn (default 24)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars


logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
curve = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # 0 = BN254, 1 = BLS12-381
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1 << logn
_ffi.init()
d = _ffi.DeviceBuffer(n * 32).upload(random_scalars(n, device.FR[curve], seed=1))
srs = device.Srs.generate(curve, 0x123456789abcdef, n)
for _ in range(reps):
    out = device.msm_dev(srs, d, n)
print("ok", _ffi.launch_count())
