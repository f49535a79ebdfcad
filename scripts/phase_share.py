"""Per-phase device time of one MSM (sort+tasks / accumulate / merge+reduce) for a list of `curve:logn` cases, each against a
key of exactly that size -- the shard sizes of the strong-scaling run (2^24 / N) and the small-MSM fixed cost.
    python scripts/phase_share.py bn254:21 bn254:22 bls12_381:22"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars

_ffi.init()
print("device:", _ffi.device_info())
print("| curve | n | c | W | total ms | sort+tasks | accumulate | merge+reduce | launches |\n|---|---|---|---|---|---|---|---|---|")
for case in sys.argv[1:] or ["bn254:16", "bn254:18", "bn254:20", "bn254:21", "bn254:22", "bn254:23", "bn254:24", "bls12_381:22"]:
    curve, logn = case.split(":")
    cid = device.curve_id(curve)
    n = 1 << int(logn)
    srs = device.Srs.generate(cid, 0x1D2C3B4A5F6E7D8C9BA, n)
    d = _ffi.DeviceBuffer(n * 32).upload(random_scalars(n, device.FR[cid], seed=int(logn)))
    for _ in range(3):
        device.msm_dev(srs, d, n)
    reps = 7
    ts = []
    l0 = _ffi.launch_count()
    for _ in range(reps):
        _ffi.timer_start(); device.msm_dev(srs, d, n); ts.append(_ffi.timer_stop())
    launches = (_ffi.launch_count() - l0) // reps
    _ffi.profile_reset(); _ffi.profile_enable(True)
    for _ in range(reps):
        device.msm_dev(srs, d, n)
    _ffi.profile_enable(False)
    p = [_ffi.profile_get(i)["ms"] / reps for i in (2, 0, 3)]
    info = srs.info()
    print(f"| {curve} | 2^{logn} | {info['c']} | {info['tables']} | {sorted(ts)[reps // 2]:.3f} | {p[0]:.3f} | {p[1]:.3f} | {p[2]:.3f} | {launches} |", flush=True)
    d.free(); srs.destroy()
