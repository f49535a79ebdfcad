set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY' 2>&1 | tail -8
import time, json, bench
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars
_ffi.init()
for nm in ("plonk", "marlin"):
    print(nm, "dropin replay", bench.dropin_hotpath(nm))
R = device.FR[0]
for n in (22, 201, 1024, 4096):
    srs = device.Srs.generate(0, 12345, n); sc = random_scalars(n, R, seed=n); d = _ffi.DeviceBuffer(n * 32).upload(sc)
    for _ in range(3): device.msm_dev(srs, d, n)
    t0 = time.perf_counter()
    for _ in range(20): device.msm_dev(srs, d, n)
    print("msm n=%d: %.3f ms per call, key" % (n, (time.perf_counter() - t0) * 50), srs.info())
import __graft_entry__ as g
g.smoke()
PY
