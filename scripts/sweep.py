"""Config 2 / config 3 sweeps: NTT 2^12..2^26 and MSM 2^16..2^26 on one B200,
device-resident timing with CUDA events (3 warm-ups, median of `reps`), correctness at every size by
a size-independent property (NTT: Horner spot checks + round trip; MSM: tau-identity).
Writes a markdown table to stdout."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kzg_snark_b200 import _ffi, device
from kzg_snark_b200.limbs import random_scalars, ints_to_limbs, limbs_to_ints

FR, FP = device.FR, device.FP
GEN = {0: 5, 1: 7}
G1 = {0: (1, 2)}

def med(ts):
    return sorted(ts)[len(ts) // 2]

def horner(c, x, r):
    acc = 0
    for v in reversed(c):
        acc = (acc * x + v) % r
    return acc

def ntt_row(cid, logn, reps=7):
    r = FR[cid]; n = 1 << logn
    x = random_scalars(n, r, seed=logn)
    w = pow(GEN[cid], (r - 1) // n, r)
    wl = ints_to_limbs([w], r)[0]; sl = ints_to_limbs([7], r)[0]
    d = _ffi.DeviceBuffer(n * 32).upload(x)
    out = {}
    for name, kw in (("fwd", {}), ("inv", {"inverse": True}), ("coset", {"coset_limbs": sl})):
        for _ in range(3):
            device.ntt_dev(cid, d, n, wl, **kw)
        ts = []
        for _ in range(reps):
            _ffi.timer_start(); device.ntt_dev(cid, d, n, wl, **kw); ts.append(_ffi.timer_stop())
        out[name] = med(ts)
    # correctness: forward transform of the original data, Horner at 3 sampled outputs, then round trip
    d.upload(x)
    device.ntt_dev(cid, d, n, wl)
    y = np.zeros_like(x); d.download(y)
    ok = True
    if logn <= 20:
        xi = limbs_to_ints(x)
        for k in (0, 1, n // 2 + 3 if n > 8 else 1):
            ok &= limbs_to_ints(y[k:k + 1])[0] == horner(xi, pow(w, k, r), r)
    device.ntt_dev(cid, d, n, wl, inverse=True)
    z = np.zeros_like(x); d.download(z)
    ok &= bool((z == x).all())
    d.free()
    return out, ok

def msm_row(cid, logn, kind="uniform", reps=5):
    r = FR[cid]; n = 1 << logn
    tau = 0x1D2C3B4A5F6E7D8C9BA % r
    t0 = time.time(); srs = device.Srs.generate(cid, tau, n); _ffi.check(_ffi._lib.kzgpu_sync()); tg = time.time() - t0
    sc = random_scalars(n, r, seed=logn)
    if kind == "skew":      # witness-like: 50 % zeros, 25 % below 2^16, 25 % uniform
        sc[::2] = 0; sc[1::4, 1:] = 0; sc[1::4, 0] &= np.uint64(0xFFFF)
    d = _ffi.DeviceBuffer(n * 32).upload(sc)
    for _ in range(2):
        out, inf = device.msm_dev(srs, d, n)
    ts = []
    for _ in range(reps):
        _ffi.timer_start(); out, inf = device.msm_dev(srs, d, n); ts.append(_ffi.timer_stop())
    info = srs.info()
    ok = None
    if logn <= 22 and cid == 0:       # tau-identity on the host: p(tau) * G1 via the oracle
        from oracle.curve import get_curve
        cv = get_curve("bn254")
        e = horner(limbs_to_ints(sc), tau, r)
        exp = cv.normalize(cv.multiply(cv.G1, e))
        got = None if inf else tuple(limbs_to_ints(out.reshape(2, 4)))
        ok = got == exp
    d.free(); srs.destroy()
    return med(ts), info, tg, ok

def main():
    _ffi.init()
    print(f"device: {_ffi.device_info()}")
    what = sys.argv[1:] or ["ntt", "msm"]
    if "ntt" in what:
        print("\n| NTT BN254 r | n | fwd ms | inv ms | coset ms | fwd elements/s | check |\n|---|---|---|---|---|---|---|")
        for logn in range(12, 27, 2):
            t, ok = ntt_row(0, logn)
            print(f"| bn254 | 2^{logn} | {t['fwd']:.3f} | {t['inv']:.3f} | {t['coset']:.3f} | {(1 << logn) / t['fwd'] * 1e3:.3e} | {'ok' if ok else 'FAIL'} |", flush=True)
        t, ok = ntt_row(1, 24)
        print(f"| bls12_381 | 2^24 | {t['fwd']:.3f} | {t['inv']:.3f} | {t['coset']:.3f} | {(1 << 24) / t['fwd'] * 1e3:.3e} | {'ok' if ok else 'FAIL'} |", flush=True)
    if "msm" in what:
        print("\n| MSM | n | scalars | ms | points/s | key layout | SRS build s | tau-identity |\n|---|---|---|---|---|---|---|---|")
        for logn in (16, 18, 20, 22, 24, 26):
            for kind in (("uniform", "skew") if logn == 24 else ("uniform",)):
                ms, info, tg, ok = msm_row(0, logn, kind)
                print(f"| bn254 | 2^{logn} | {kind} | {ms:.3f} | {(1 << logn) / ms * 1e3:.3e} | c={info['c']} W={info['tables']} {info['bytes'] / 2**30:.1f} GiB | {tg:.2f} | {'ok' if ok else ('FAIL' if ok is False else 'n/a (host Horner too slow)')} |", flush=True)
        ms, info, tg, ok = msm_row(1, 22)
        print(f"| bls12_381 | 2^22 | uniform | {ms:.3f} | {(1 << 22) / ms * 1e3:.3e} | c={info['c']} W={info['tables']} {info['bytes'] / 2**30:.1f} GiB | {tg:.2f} | n/a |", flush=True)

if __name__ == "__main__":
    main()
