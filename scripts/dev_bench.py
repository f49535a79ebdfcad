"""Developer timing script (not the contract bench): microbenchmarks + kernel timings."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars, ints_to_limbs

R_BN = device.FR[0]

def main():
    what = sys.argv[1:] or ["micro", "ntt", "msm"]
    _ffi.init()
    info = _ffi.device_info(); print(info, flush=True)
    sm = info["sm_count"]
    res = {}
    if "micro" in what:
        for kind, name in ((0, "imad_wide"), (5, "imad_wide+2alu"), (6, "imad_wide+4alu"), (7, "imad32"), (8, "dfma"), (9, "dfma+add64"), (10, "dfma+wide/2"), (11, "dfma+add64+wide/2"), (1, "fe_mul_bn254"), (2, "fe_mul_bls381"), (3, "madd_bn254"), (4, "madd_bls381")):
            for threads, bps in ((256, 2), (128, 4), (256, 4), (128, 2)):
                iters = 400 if kind in (3, 4) else 2000
                ms, ops = _ffi.microbench(kind, sm * bps, threads, iters)
                print(f"micro {name:14s} blocks/SM={bps} threads={threads}: {ms:8.3f} ms  {ops/ms/1e6:10.2f} Gop/s", flush=True)
                res.setdefault(name, []).append(ops / ms / 1e6)
    if "ntt" in what:
        for logn in (12, 16, 20, 22, 24, 26):
            n = 1 << logn
            x = random_scalars(n, R_BN, seed=logn)
            w = pow(5, (R_BN - 1) // n, R_BN)
            wl = ints_to_limbs([w], R_BN)[0]
            d = _ffi.DeviceBuffer(n * 32).upload(x)
            t0 = time.time(); device.ntt_dev(0, d, n, wl); _ffi.check(_ffi._lib.kzgpu_sync()); cold = time.time() - t0
            ts = []
            for _ in range(5):
                _ffi.timer_start(); device.ntt_dev(0, d, n, wl); ts.append(_ffi.timer_stop())
            ms = sorted(ts)[len(ts) // 2]
            print(f"ntt 2^{logn}: cold {cold*1e3:.2f} ms, warm {ms:.3f} ms, {n/ms/1e3:.1f} Melem/s, hbm-equiv {64*n/ms/1e6:.1f} GB/s", flush=True)
            d.free()
    if "msm" in what:
        tau = 0x1234567890abcdef1234567
        for logn in (16, 20, 22, 24):
            n = 1 << logn
            t0 = time.time(); srs = device.Srs.generate(0, tau, n); tg = time.time() - t0
            sc = random_scalars(n, R_BN, seed=logn)
            d = _ffi.DeviceBuffer(n * 32).upload(sc)
            cs = [None] if logn < 24 else [None, 19, 21, 22]
            for c in cs:
                if c is None: os.environ.pop("KZGPU_MSM_C", None)
                else: os.environ["KZGPU_MSM_C"] = str(c)
                device.msm_dev(srs, d, n)
                ts = []
                for _ in range(3):
                    _ffi.timer_start(); out = device.msm_dev(srs, d, n); ts.append(_ffi.timer_stop())
                ms = sorted(ts)[1]
                print(f"msm 2^{logn} c={c}: srs gen {tg:.2f}s, {ms:.2f} ms, {n/ms/1e3:.1f} Mpts/s", flush=True)
            os.environ.pop("KZGPU_MSM_C", None)
            d.free(); srs.destroy()
    print("launches", _ffi.launch_count())

if __name__ == "__main__":
    main()
