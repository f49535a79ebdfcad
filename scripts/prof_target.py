"""Short single-GPU target for `ncu --set full`: two NTTs and two MSMs at the bench size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars, ints_to_limbs

R = device.FR[0]
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << logn
_ffi.init()
x = random_scalars(n, R, seed=1)
d = _ffi.DeviceBuffer(n * 32).upload(x)
wl = ints_to_limbs([pow(5, (R - 1) // n, R)], R)[0]
for _ in range(2):
    device.ntt_dev(0, d, n, wl)
srs = device.Srs.generate(0, 0x123456789abcdef, n)
for _ in range(2):
    out = device.msm_dev(srs, d, n)
print("ok", _ffi.launch_count())
