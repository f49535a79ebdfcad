"""Host-scalar (pinned) vs resident time of one rank's shard of a point-sharded MSM: `python scripts/e2e_shard.py 21 22 23`
(run once per KZGPU_MSM_NCHUNKS setting: the variable is read once per process)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars
_ffi.init()
R = device.FR[0]
for logn in [int(a) for a in sys.argv[1:]] or [21, 22, 23]:
    n = 1 << logn
    srs = device.Srs.generate(0, 0x1D2C3B4A5F6E7D8C9BA, n)
    pin = _ffi.PinnedArray((n, 4)); pin.array[:] = random_scalars(n, R, seed=logn)
    d = _ffi.DeviceBuffer(n * 32).upload(pin.array)
    part = _ffi.DeviceBuffer(256)
    def run(fn, reps=10):
        for _ in range(3): fn()
        _ffi.check(_ffi._lib.kzgpu_sync())
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        _ffi.check(_ffi._lib.kzgpu_sync())
        return (time.perf_counter() - t0) * 1e3 / reps
    t_res = run(lambda: device.msm_partial_dev(srs, d, n, part))
    t_host = run(lambda: device.msm_partial(srs, pin.array, part))
    t_full = run(lambda: device.msm(srs, pin.array))
    print(f"2^{logn} nchunks={os.environ.get('KZGPU_MSM_NCHUNKS','default')}: partial resident {t_res:.3f} ms, partial from host {t_host:.3f} ms, kzgpu_msm from host {t_full:.3f} ms", flush=True)
    srs.destroy(); d.free(); pin.free(); part.free()
