python -m pytest tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -2
python - <<'PY'
import time
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars
_ffi.init()
R = device.FR[0]
for n in (201, 1025, 2048, 4096, 8192):
    srs = device.Srs.generate(0, 12345, n); sc = random_scalars(n, R, seed=n); d = _ffi.DeviceBuffer(n * 32).upload(sc)
    for _ in range(3): device.msm_dev(srs, d, n)
    t0 = time.perf_counter()
    for _ in range(20): device.msm_dev(srs, d, n)
    print("msm n=%d: %.3f ms per call" % (n, (time.perf_counter() - t0) * 50), srs.info())
PY
