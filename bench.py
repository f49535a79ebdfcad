#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for the KZG-MSM / NTT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload msm|ntt|sweep|plonk|marlin] [--curve bn254|bls12_381] [--logn 24]

Default workload (`msm`): one step = one KZG commit of a 2^24-coefficient polynomial, i.e. one BN254 G1 MSM of 2^24
points against the device-resident SRS (BASELINE.json metric "G1 MSM points/s ... at 2^24 (1/2/4/8 B200)"; configs[2]).
The same line also carries the NTT half of the metric (`"ntt"`: scalar-field NTT of 2^24 elements, configs[1]) with its
own roofline, and at N = 1 the PLONK / Marlin provers (configs[3], configs[4]).

N > 1 (torchrun, one process per GPU, NCCL): the metric's own configuration -- ONE 2^24-point MSM split over the N
ranks ("scaling": "strong"): SRS points and scalars are sharded by contiguous index range, every rank reduces its
2^24/N points to one XYZZ partial sum, the partials are all-gathered over NCCL (128 B per rank) and folded.  `value` is
2^24 points / step time.  The independent-shards variant (2^24 points per GPU) is reported beside it as `"weak"`.

`--workload sweep` prints configs[1] / configs[2] as tables (NTT 2^12..2^26, MSM 2^16..2^26: device-resident, host-buffer
and CPU-port columns, a correctness check per size).  `--impl reference` times the reference's own `KZG.commit` loop
(kzg.py:112-116, imported unmodified from baseline/_ref or /root/reference; SageMath / py_ecc replaced by the stand-ins
of oracle/refrun.py) on all host cores, on a bounded sample of the same workload.
"""

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R_BN254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R_BLS = 52435875175126190479447740508185965837690552500527637822603658699938581184513
FR = {"bn254": R_BN254, "bls12_381": R_BLS}
GEN = {"bn254": 5, "bls12_381": 7}                   # least primitive roots of the two scalar fields
FP_BITS = {"bn254": 254, "bls12_381": 381}
TAU = 0x2545F4914F6CDD1D9E3779B97F4A7C15F39CC0605CEDC834 % R_BN254      # fixed synthetic trapdoor

# SURVEY.md section 8(d) work model: 16 windows x (8M+2S = 10 modmul) x 272 32-bit IMADs
MODEL_IMAD_PER_POINT = 43520
IMAD32_PER_MODMUL = 272
# What the kernels EXECUTE, per N-limb Montgomery operation (field.cuh / mp_prims_gen.cuh; checked against the SASS of the
# accumulate loop with scripts/sass_loop.py and against ncu's smsp__inst_executed): a general product issues 2N^2 - N
# IMAD.WIDE (32x32+64 -> 64, one per 2 issue slots of the multiplier pipe) and 2N narrow IMAD / IMAD.HI (one slot: half
# the cost); the dedicated squaring N(N+1)/2 + N^2 - N wide and 2N narrow.
# The mixed addition's last coordinate, Y3 = R (Q - X3) - Y1 PPP, is a DUAL product with one reduction (fe_mul2_nofinal):
# 3N^2 - N wide and 2N narrow instead of two general products.
def wide_slots(nlimbs, products, squarings, duals=0):
    """multiplier-pipe work in IMAD.WIDE equivalents (narrow multiplies counted at half)."""
    n = nlimbs
    gen = (2 * n * n - n) + 0.5 * (2 * n)
    sqr = (n * (n + 1) // 2 + n * n - n) + 0.5 * (2 * n)
    dual = (3 * n * n - n) + 0.5 * (2 * n)
    return products * gen + squarings * sqr + duals * dual


MADD = (6, 2, 1)          # XYZZ mixed addition: 6 products + 2 squarings + 1 dual product (curve.cuh xyzz_madd_lz)
FULL_ADD = (12, 2)        # XYZZ + XYZZ (canonical arithmetic, bucket reduction)


def _timeit(f):
    t0 = time.perf_counter()
    f()
    return time.perf_counter() - t0


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` captures
    (profiles/r2_traffic.json, else r1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f).get(kernel, {}).get("dram_bytes_per_launch")
            if v:
                return v
        except Exception:
            pass
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            return None
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arms
def _cpu_commit_chunk(args):
    """Worker: the commit loop kzg.py:112-116 over `count` (coefficient, SRS point) pairs.
    kind "reference": the reference's OWN kzg.py (KZG.commit, imported unmodified) on the SageMath / py_ecc stand-ins;
    kind "port": the oracle's restatement of the same loop."""
    seed, count, tau, kind, curve = args
    import random
    from oracle.kzg import KZGOracle
    ko = KZGOracle(curve)
    rng = random.Random(seed)
    ck = ko.setup_fast(count - 1, tau)                   # input preparation (not timed): points tau^i * G1
    coeffs = [rng.randrange(ko.curve_order) for _ in range(count)]
    if kind == "reference":
        from oracle import refrun
        with refrun.ReferenceRun(seed=seed, record=False) as rr:
            kzg = rr.kzg.KZG(curve_type=curve)           # reference kzg.py:18
            from oracle import pyecc_standin, pyecc_standin_bls
            E = pyecc_standin if curve == "bn254" else pyecc_standin_bls
            rck = [tuple(E.FQ(int(c)) for c in pt) for pt in ck]
            poly = kzg.R([kzg.Fq(c) for c in coeffs])
            t0 = time.perf_counter()
            c = kzg.commit(rck, [poly])[0]               # reference kzg.py:80-120
            dt = time.perf_counter() - t0
        return dt, ko.cv.normalize(tuple(int(v) for v in c))
    t0 = time.perf_counter()
    c = ko.commit(ck, [coeffs])[0]
    dt = time.perf_counter() - t0
    return dt, ko.cv.normalize(c)


def reference_kind():
    """("reference", note) when the reference's own sources are reachable, else ("port", note)."""
    try:
        from oracle import refrun
        if refrun.available():
            return "reference", ("reference kzg.py / fft_ff.py imported unmodified from " + refrun.REFERENCE_ROOT +
                                 "; SageMath and py_ecc (not installable in this image) replaced by the stand-ins of oracle/refrun.py")
    except Exception:
        pass
    return "port", "restated reference algorithm on CPython ints (oracle/): no reference tree on this box"


def cpu_msm_rate(points_per_core, cores, kind="port", curve="bn254"):
    """points/s of the reference commit loop using `cores` processes."""
    jobs = [(1000 + i, points_per_core, TAU, kind, curve) for i in range(cores)]
    if cores == 1:
        res = [_cpu_commit_chunk(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_commit_chunk, jobs)
    wall = max(r[0] for r in res)              # commit loop time only (inputs prepared before)
    return points_per_core * cores / wall, wall


def _cpu_ntt_chunk(args):
    seed, logn, kind, curve = args
    import random
    r = FR[curve]
    rng = random.Random(seed)
    n = 1 << logn
    x = [rng.randrange(r) for _ in range(n)]
    w = pow(GEN[curve], (r - 1) // n, r)
    if kind == "reference":
        from oracle import refrun
        with refrun.ReferenceRun(seed=seed, record=False) as rr:
            F = rr.kzg.KZG(curve_type=curve).Fq
            xs, ws = [F(v) for v in x], F(w)
            t0 = time.perf_counter()
            rr.fft_ff.fft_ff(xs, ws, F)                  # reference fft_ff.py:3-37
            return time.perf_counter() - t0
    from oracle.fft_ff import fft_ff_int
    t0 = time.perf_counter()
    fft_ff_int(x, w, r)
    return time.perf_counter() - t0


def cpu_ntt_rate(logn, cores, kind="port", curve="bn254"):
    """elements/s of the recursive fft_ff; `cores` independent vectors in parallel
    (the reference itself is single-threaded: batched vectors are its only parallelism)."""
    jobs = [(2000 + i, logn, kind, curve) for i in range(cores)]
    if cores == 1:
        wall = _cpu_ntt_chunk(jobs[0])
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            wall = max(pool.map(_cpu_ntt_chunk, jobs))
    return (1 << logn) * cores / wall, wall


def _replay_prover_calls(name, make_key, commit, open_, fns, F):
    """Time the hot-path calls the reference's prover made on a bundled instance (the records of
    tests/golden/ref_trace_<name>.json after the indexer's) through one implementation of the boundary."""
    tr = json.load(open(os.path.join(ROOT, "tests", "golden", f"ref_trace_{name}.json")))
    H = lambda v: int(v, 16)                                          # noqa: E731
    keys = [make_key([(H(p[0]), H(p[1])) for p in k]) for k in tr["keys"]]
    calls = tr["calls"][tr["notes"]["index_calls"]:]
    t0 = time.perf_counter()
    for rec in calls:
        if rec["fn"] == "commit":
            commit(keys[rec["ck_id"]], [[H(c) for c in p] for p in rec["polys"]])
        elif rec["fn"] == "open":
            open_(keys[rec["ck_id"]], [[H(c) for c in p] for p in rec["polys"]], H(rec["z"]), H(rec["xi"]))
        else:
            fns[rec["fn"]]([F(H(v)) for v in rec["in"]], F(H(rec["w"])), F)
    return time.perf_counter() - t0, len(calls), tr["notes"]["prove_seconds"]


def cpu_hotpath(name):
    """The restated oracle on those calls, 1 core (the reference is single-threaded)."""
    from oracle import fft_ff as off
    from oracle.field import GFp
    from oracle.kzg import KZGOracle
    ko = KZGOracle("bn254")
    return _replay_prover_calls(name, lambda pts: [(x, y, 1) for x, y in pts], ko.commit, ko.open,
                                {"fft_ff": off.fft_ff, "ifft_ff": off.ifft_ff, "fft_ff_interpolation": off.fft_ff_interpolation},
                                GFp(R_BN254))[0]


def dropin_hotpath(name):
    """The same calls through the GPU drop-in modules (`kzg.KZG`, `fft_ff`): Python objects in, Python objects out,
    i.e. what the reference's unmodified prover pays when the two modules are shadowed."""
    from kzg_snark_b200.kzg import KZG
    from kzg_snark_b200 import fft_ff as gff
    kzg = KZG("bn254")
    fq = kzg._codec.fq
    fns = {"fft_ff": gff.fft_ff, "ifft_ff": gff.ifft_ff, "fft_ff_interpolation": gff.fft_ff_interpolation}
    mk = lambda pts: [(fq(x), fq(y), fq(1)) for x, y in pts]           # noqa: E731
    _replay_prover_calls(name, mk, kzg.commit, kzg.open, fns, kzg.Fq)                   # warm-up: key upload, NTT plans
    best = min(_replay_prover_calls(name, mk, kzg.commit, kzg.open, fns, kzg.Fq) for _ in range(3))
    return {"dropin_hotpath_s": best[0], "calls": best[1], "reference_prove_s": best[2]}


def reference_prover_on_gpu_dropin(name):
    """SURVEY.md 8(d) config 4 / 5 as defined there: the reference's UNMODIFIED plonk/ or marlin/ indexer, prover and verifier
    (imported from baseline/_ref or /root/reference) with `kzg` / `fft_ff` resolved to the GPU drop-in; wall seconds of
    `Prover.prove`, verifier must accept and the reference's tamper test must reject.  None without a reference tree."""
    import pickle
    from oracle import refrun
    if not refrun.available():
        return None
    with refrun.ReferenceRun(seed=11, record=False, gpu_dropin=True) as rr:
        Fq = rr.kzg.KZG("bn254").Fq
        cs = os.path.join(refrun.REFERENCE_ROOT, "constraint-system")
        if name == "plonk":
            inst = pickle.load(open(os.path.join(cs, "PLONK_ARITHMETIZATION_INSTANCE.pkl"), "rb"))       # main.py:68-69
            sel = [inst[k] for k in ("qM", "qL", "qR", "qO", "qC")]
            w = [Fq(v) for v in inst["w"]]
            x, wit = w[:5], w[5:]
            ipk, ivk = rr.load("plonk.indexer").Indexer(curve_type="bn254").preprocess(*sel, inst["perm"], max_degree=len(sel[0]) + 5)
            prover, V = rr.load("plonk.prover").Prover(curve_type="bn254"), rr.load("plonk.verifier").Verifier
            tamper = lambda p: {**p, "evaluations": {**p["evaluations"], "a": p["evaluations"]["a"] + 1}}       # noqa: E731
        else:
            inst = pickle.load(open(os.path.join(cs, "R1CS_INSTANCE.pkl"), "rb"))                            # main.py:43-44
            z = [Fq(v) for v in inst["z"]]
            x, wit = z[:5], z[5:]
            ipk, ivk = rr.load("marlin.indexer").Indexer(curve_type="bn254").preprocess(inst["A"], inst["B"], inst["C"], max_degree=200)
            prover, V = rr.load("marlin.prover").Prover(curve_type="bn254"), rr.load("marlin.verifier").Verifier
            tamper = lambda p: {**p, "evaluations": {**p["evaluations"], "beta1": [p["evaluations"]["beta1"][0] + 1] + list(p["evaluations"]["beta1"][1:])}}  # noqa: E731
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            proof = prover.prove(ipk, x, wit)
            ts.append(time.perf_counter() - t0)
        ok = bool(V(curve_type="bn254").verify(ivk, x, proof))
        rejected = not V(curve_type="bn254").verify(ivk, x, tamper(proof))
    return {"prove_s": min(ts), "verifier_accepts": ok, "tampered_rejected": bool(rejected),
            "note": f"reference {name}/prover.py unmodified (from {refrun.REFERENCE_ROOT}) on the GPU kzg / fft_ff drop-in; host side = "
                    "the reference's Python polynomial algebra on the Sage stand-in, so this is dominated by host work"}


def marlin_synthetic_prove(rows_logn, reps=3):
    """configs[4]: kzg_snark_b200.marlin.Prover.prove (the device-resident counterpart of marlin/prover.py:25-245) on a synthetic
    R1CS of 2^rows_logn rows, on whatever devices the library is initialised on (one, or several: its commitments and openings
    are then point-sharded / dealt over them inside the library).  The reference prover cannot run at this size (dense Sage
    matrices)."""
    from kzg_snark_b200 import _ffi, marlin
    from kzg_snark_b200.limbs import random_scalars, ints_to_limbs
    rows = 1 << rows_logn
    t0 = time.perf_counter()
    sA, sB, sC, sx, sw = marlin.synthetic_r1cs(rows, 8, R_BN254, seed=rows_logn)
    swl = ints_to_limbs(sw, R_BN254)
    sdraws = random_scalars(8 + 2 * rows + 1, R_BN254, seed=4)
    t_gen = time.perf_counter() - t0
    midx = marlin.Indexer("bn254")
    mK = 1 << (2 * rows - 1).bit_length()
    t0 = time.perf_counter()
    sipk, _ = midx.preprocess(sA, sB, sC, max_degree=6 * mK, tau=TAU)
    _ffi.check(_ffi._lib.kzgpu_sync())
    t_index = time.perf_counter() - t0
    sxs = [midx.kzg.Fq(v) for v in sx]
    mpr = marlin.Prover("bn254")
    mt = []
    for _ in range(2 + reps):
        t0 = time.perf_counter()
        mpr.prove(sipk, sxs, swl, draws=sdraws)
        mt.append(time.perf_counter() - t0)
    assert set(mpr.checks.values()) == {0}, "a Marlin linearisation identity failed"
    warm = sorted(mt[2:])
    return {"rows": rows, "H": sipk["subgroups"]["n"], "K": sipk["subgroups"]["m"], "gpus": _ffi.device_count(),
            "prove_s": warm[len(warm) // 2], "first_prove_s": mt[0], "index_s": t_index, "instance_generation_s": t_gen,
            "rounds_s": {k: round(v, 5) for k, v in mpr.timings.items()},
            "checks": "f_1(beta_1) = f_2(beta_1) = f_3(beta_2) = 0 asserted (the verifier's polynomial identities)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    kind, note = reference_kind()
    if args.force_port:
        kind, note = "port", "restated reference algorithm on CPython ints (oracle/), forced with --force-port"
    vals, t_all = [], 0.0
    curve = args.curve
    if args.workload in ("msm", "sweep", "plonk", "marlin"):
        per_core = 256
        unit, metric = "points/s", "g1_msm_points_per_s"
        for i in range(args.warmup + args.steps):
            v, wall = cpu_msm_rate(per_core, cores, kind, curve)
            if i >= args.warmup:
                vals.append(v); t_all += wall
        sample = (f"{per_core * cores} random scalars x SRS points per step ({per_core} per core, {cores} processes), "
                  f"KZG.commit loop kzg.py:112-116; the loop is strictly linear in the number of points")
        workload = f"{curve} G1 MSM (KZG commit) of 2^{args.logn} points"
    else:
        logn_s = 13
        unit, metric = "elements/s", "ntt_elements_per_s"
        for i in range(args.warmup + args.steps):
            v, wall = cpu_ntt_rate(logn_s, cores, kind, curve)
            if i >= args.warmup:
                vals.append(v); t_all += wall
        sample = f"{cores} vectors of 2^{logn_s} elements per step (one per core), recursive fft_ff.py:3-37"
        workload = f"{curve} scalar-field NTT of 2^{args.logn} elements"
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 (256-bit modular integers)" if curve == "bn254" else "u32x12 / u32x8 (384 / 256-bit modular integers)",
        "data": "synthetic",
        "config": {"workload": workload, "curve": curve, "sample": sample},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample, "note": note},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- GPU arm
class RawPtr:
    def __init__(self, p):
        import ctypes
        self.ptr = ctypes.c_void_p(p)


def on_curve(out, curve):
    from kzg_snark_b200 import device
    from kzg_snark_b200.limbs import limbs_to_ints
    cid = device.curve_id(curve)
    p, L = device.FP[cid], device.FP_LIMBS[cid]
    x, y = limbs_to_ints(out.reshape(2, L))
    return (y * y - x * x * x - (3 if cid == 0 else 4)) % p == 0


def horner_dev(curve, dbuf, n, x):
    """p(x) of the device-resident coefficient vector (kzgpu_poly_eval_dev): the scalar side of the tau-identity."""
    import ctypes
    import numpy as np
    from kzg_snark_b200 import _ffi, device
    from kzg_snark_b200.limbs import ints_to_limbs, limbs_to_int
    cid = device.curve_id(curve)
    out = np.zeros(4, dtype=np.uint64)
    xl = ints_to_limbs([x], device.FR[cid])[0]
    _ffi.check(_ffi._lib.kzgpu_poly_eval_dev(cid, dbuf.ptr, n, _ffi.ptr(xl), _ffi.ptr(out)))
    return limbs_to_int(out)


def tau_identity(curve, out, inf, e):
    """commit(ck, p) == p(tau) * G1 (kzg.py:108) with e = p(tau): one device-side scalar multiplication of the generator."""
    import numpy as np
    from kzg_snark_b200 import device
    from kzg_snark_b200.limbs import ints_to_limbs
    from kzg_snark_b200.kzg import _G1
    cid = device.curve_id(curve)
    L = device.FP_LIMBS[cid]
    g = np.concatenate([ints_to_limbs([c], device.FP[cid], L)[0] for c in _G1[curve]]).reshape(1, 2 * L)
    exp, einf = device.g1_lincomb(cid, g, ints_to_limbs([e], device.FR[cid]))
    return bool(inf) == bool(einf) and (bool(inf) or bool((exp == out).all()))


def run_gpu(args):
    import numpy as np
    from kzg_snark_b200 import _ffi, device
    from kzg_snark_b200.limbs import random_scalars, ints_to_limbs, limbs_to_ints

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch = None
    if world > 1:
        # NCCL_DEBUG is left as the caller set it (the harness reads the communicator lines); the result line is printed
        # last, on its own line, by rank 0
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _ffi.init(local)
    info = _ffi.device_info()
    peaks, peak_src = load_peaks()
    K, Wm = args.steps, args.warmup
    curve = args.curve
    cid = device.curve_id(curve)
    r_mod = FR[curve]
    nl = 8 if curve == "bn254" else 12                    # base-field limbs (32-bit)
    n_total = 1 << args.logn
    if world > 1:
        # one stream for torch (NCCL's wait lands on the current stream) and the library, so that the fold of the gathered
        # partials is ordered after the all-gather.  It must be a side stream: torch's default stream has handle 0, which
        # kzgpu_set_stream reads as "back to the library's own stream".
        from kzg_snark_b200.parallel import share_stream_with_torch
        side = share_stream_with_torch()                      # noqa: F841  (kept alive for the life of the run)

    def barrier():
        if dist is not None:
            dist.barrier()
        _ffi.check(_ffi._lib.kzgpu_sync())

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ integer roofline (live)
    sm = info["sm_count"]
    ms_i, ops_i = _ffi.microbench(0, sm * 4, 256, 2000, device=local)
    imad_wide_peak = ops_i / ms_i * 1e3                   # IMAD.WIDE.U32 / s, measured now on this GPU
    imad32_peak = 2.0 * imad_wide_peak                    # one wide = two 32-bit multiply-add slots (the SURVEY model's unit)

    # ------------------------------------------------------------------ MSM
    def msm_case(n, start, seed, want_e2e=True, want_profile=True, pageable=False):
        """One rank's share: `n` points [start, start + n) of the key; world ranks together form one MSM."""
        t0 = time.perf_counter()
        srs = device.Srs.generate(curve, TAU, n, start=start)
        _ffi.check(_ffi._lib.kzgpu_sync())
        build_s = time.perf_counter() - t0
        key = srs.info()
        pinned = _ffi.PinnedArray((n, 4))
        pinned.array[:] = random_scalars(n, r_mod, seed=seed)
        dsc = _ffi.DeviceBuffer(n * 32).upload(pinned.array)
        xyzz = 4 * nl * 4
        if world > 1:
            partial = torch.zeros(xyzz, dtype=torch.uint8, device="cuda")
            gathered = torch.zeros(xyzz * world, dtype=torch.uint8, device="cuda")

        def fold():
            dist.all_gather_into_tensor(gathered, partial)
            return device.g1_fold(curve, RawPtr(gathered.data_ptr()), world)

        def step_resident():
            if world == 1:
                return device.msm_dev(srs, dsc, n)
            device.msm_partial_dev(srs, dsc, n, RawPtr(partial.data_ptr()))
            return fold()

        def step_e2e(host):
            if world == 1:
                return device.msm(srs, host)                  # H2D of the scalars inside the call
            device.msm_partial(srs, host, RawPtr(partial.data_ptr()))       # this rank's H2D inside the call, overlapped
            return fold()

        for _ in range(Wm):
            if world > 1:
                gathered.zero_()                                 # a fold that ran ahead of the all-gather would see infinity
            out, inf = step_resident()
            assert not inf and on_curve(out, curve), "MSM result is not a curve point"
        # the tau-identity at full size (kzg.py:108): sum over the ranks of p_rank(tau) * tau^start
        e = horner_dev(curve, dsc, n, TAU % r_mod) * pow(TAU, start, r_mod) % r_mod
        if dist is not None:
            parts = [None] * world
            dist.all_gather_object(parts, e)
            e = sum(parts) % r_mod
        check_ok = tau_identity(curve, out, inf, e)
        assert check_ok, "MSM result violates the tau-identity commit == p(tau) * G1"
        # timed region: resident inputs, device clock
        sampler = ClockSampler(local)
        barrier()
        l0 = _ffi.launch_count()
        # CUDA events on the stream the library launches on (its own non-blocking stream: torch.cuda.Event on torch's
        # current stream would not see these kernels); at N > 1 every step ends with the host-synchronous fold of the
        # gathered partials, so the bracket covers the NCCL all-gather too
        _ffi.timer_start()
        for _ in range(K):
            step_resident()
        ms = _ffi.timer_stop()
        barrier()
        launches = _ffi.launch_count() - l0
        ms = max_over_ranks(ms)
        res = {"ms": ms, "launches": launches, "key": key, "build_s": build_s, "n": n, "check": bool(check_ok)}
        if want_profile:
            # per-kernel profile (second region, CUDA events around the kernels on the launching stream)
            _ffi.profile_reset(); _ffi.profile_enable(True)
            for _ in range(K):
                step_resident()
            _ffi.profile_enable(False)
            res["prof"] = {k: _ffi.profile_get(i) for k, i in (("accumulate", 0), ("sort", 2), ("reduce", 3))}
        if want_e2e:
            # e2e: host scalars in, affine point out, every step (warmed up like the resident region: the chunked entry
            # point allocates its staging buffer and second sort workspace on first use)
            for _ in range(Wm):
                step_e2e(pinned.array)
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                step_e2e(pinned.array)
            barrier()
            res["ms_e2e"] = max_over_ranks((time.perf_counter() - t0) * 1e3)
            if pageable:
                # the same call from an ordinary (pageable) numpy array -- what the Python drop-in hands over
                heap = np.array(pinned.array, copy=True)
                for _ in range(2):
                    step_e2e(heap)
                barrier()
                t0 = time.perf_counter()
                for _ in range(K):
                    step_e2e(heap)
                barrier()
                res["ms_e2e_pageable"] = max_over_ranks((time.perf_counter() - t0) * 1e3)
        res["clocks"] = sampler.stop()                           # sampled over the timed, per-kernel and e2e regions
        srs.destroy(); dsc.free(); pinned.free()
        return res

    def bench_msm():
        strong = world > 1
        n = n_total // world if strong else n_total
        r = msm_case(n, rank * n, seed=args.logn * 100 + rank, pageable=(world == 1))
        ms, prof = r["ms"], r["prof"]
        acc = prof["accumulate"]                                 # the bucket-accumulate kernel alone, one launch per MSM
        acc_ms = acc["ms"] / max(acc["launches"], 1)
        madds = acc["work"] / max(acc["launches"], 1)            # mixed additions per launch (n * windows upper bound)
        red = prof["reduce"]
        buckets = red["work"] / max(red["launches"], 1)          # buckets of the pass: 2 full additions each in the reduce
        acc_slots = madds * wide_slots(nl, *MADD)
        step_slots = acc_slots + 2.0 * buckets * wide_slots(nl, *FULL_ADD)
        total_pts = world * n
        pts_per_s = total_pts * K / (ms * 1e-3)
        frac = acc_slots / (acc_ms * 1e-3) / imad_wide_peak
        res = {
            "value": pts_per_s, "ms": ms, "launches": r["launches"], "clocks": r["clocks"], "n_per_rank": n, "key": r["key"],
            "build_s": r["build_s"], "check": r["check"],
            "e2e": {"value": total_pts * K / (r["ms_e2e"] * 1e-3), "unit": "points/s",
                    "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": (2 * nl * 4 + 4) * world,
                    "ms_per_step": r["ms_e2e"] / K, "host_buffers": "pinned (cudaHostAlloc)"},
            "roofline": {
                "bound": "imad", "kernel": "msm_accumulate_kernel",
                "achieved": acc_slots / (acc_ms * 1e-3) / 1e9, "peak": imad_wide_peak / 1e9, "unit": "G IMAD.WIDE/s",
                "frac": frac,
                "traffic": load_traffic("msm_accumulate_kernel") if (args.logn == 24 and world == 1 and curve == "bn254") else None,
                "model": f"executed multiplier work: {madds / n:.0f} mixed additions per point (fixed-base tables, c = {r['key']['c']}) x "
                         f"{wide_slots(nl, *MADD):.0f} IMAD.WIDE-equivalents (6 products + 2 squarings + 1 dual product on {nl} limbs; narrow IMAD / IMAD.HI "
                         "counted at half) / kernel time / IMAD.WIDE.U32 issue rate measured live (kzgpu_microbench 0); "
                         "compare ncu sm__pipe_fma_cycles_active / 50 % in profiles/",
                "step_frac": step_slots / (ms / K * 1e-3) / imad_wide_peak,
                "model_frac": (n / (acc_ms * 1e-3)) * MODEL_IMAD_PER_POINT / imad32_peak,
                "model_frac_note": "SURVEY 8(d) model figure (16 windows x 10 modmul x 272 IMAD32 per point) / (2 x IMAD.WIDE peak); "
                                   "exceeds 1 because the kernel executes fewer windows and cheaper squarings than the model charges",
                "kernel_ms": acc_ms, "kernel_share_of_step": acc["ms"] / max(sum(p["ms"] for p in prof.values()), 1e-9),
                "hbm_algorithmic_gbs": n * (2 * nl * 4 + 32) / (acc_ms * 1e-3) / 1e9,
                "hbm_peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src,
            },
            "profile_ms_per_step": {k: v["ms"] / K for k, v in prof.items()},
        }
        if "ms_e2e_pageable" in r:
            res["e2e"]["pageable"] = {"value": total_pts * K / (r["ms_e2e_pageable"] * 1e-3), "ms_per_step": r["ms_e2e_pageable"] / K,
                                      "host_buffers": "pageable numpy array (what the Python drop-in hands over): staged through page-locked double buffers by 4 host threads inside the call (kz_upload)"}
        if strong:
            # the independent-shards variant beside it: 2^logn points per GPU
            wr = msm_case(n_total, rank * n_total, seed=args.logn * 100 + 50 + rank, want_e2e=True, want_profile=False)
            res["weak"] = {"scaling": "weak", "points_per_gpu": n_total, "value": world * n_total * K / (wr["ms"] * 1e-3),
                           "ms_per_step": wr["ms"] / K, "key": wr["key"],
                           "e2e": {"value": world * n_total * K / (wr["ms_e2e"] * 1e-3), "ms_per_step": wr["ms_e2e"] / K,
                                   "h2d_bytes_per_step": n_total * 32 * world}}
        return res

    # ------------------------------------------------------------------ NTT
    def bench_ntt():
        w = pow(GEN[curve], (r_mod - 1) // n_total, r_mod)
        wl = ints_to_limbs([w], r_mod)[0]
        pinned = _ffi.PinnedArray((n_total, 4))
        pinned.array[:] = random_scalars(n_total, r_mod, seed=args.logn + 7 * rank)
        d = _ffi.DeviceBuffer(n_total * 32).upload(pinned.array)
        for _ in range(Wm):
            device.ntt_dev(curve, d, n_total, wl)
        sampler = ClockSampler(local)
        barrier()
        l0 = _ffi.launch_count()
        _ffi.timer_start()                                      # events on the library's launching stream, any world size
        for _ in range(K):
            device.ntt_dev(curve, d, n_total, wl)
        ms = _ffi.timer_stop()
        barrier()
        launches = _ffi.launch_count() - l0
        ms = max_over_ranks(ms)
        _ffi.profile_reset(); _ffi.profile_enable(True)
        for _ in range(K):
            device.ntt_dev(curve, d, n_total, wl)
        _ffi.profile_enable(False)
        pr = _ffi.profile_get(1)
        for _ in range(Wm):                                     # warm-up of the host-buffer entry point (staging buffer, events)
            device.ntt(curve, pinned.array, wl)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            device.ntt(curve, pinned.array, wl)              # H2D + kernels + D2H inside the call
        barrier()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
        # the transfers alone (same bytes, same buffers, all ranks at once): the host-side ceiling of the e2e number
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            d.upload(pinned.array)
            d.download(pinned.array)
        barrier()
        ms_copy = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {"value": world * n_total * K / (ms_e2e * 1e-3), "unit": "elements/s",
               "h2d_bytes_per_step": n_total * 32 * world, "d2h_bytes_per_step": n_total * 32 * world,
               "ms_per_step": ms_e2e / K, "host_buffers": "pinned (cudaHostAlloc)",
               "copy_only_ms_per_step": ms_copy / K,
               "copy_only_note": "H2D + D2H of the same buffers with no kernel, all ranks at once: every output depends on every "
                                 "input, so a single vector cannot overlap its two transfers; e2e - copy_only is what the NTT adds"}
        if world == 1:
            heap = np.array(pinned.array, copy=True)
            for _ in range(2):
                device.ntt(curve, heap, wl)
            t0 = time.perf_counter()
            for _ in range(K):
                device.ntt(curve, heap, wl)
            msp = (time.perf_counter() - t0) * 1e3
            e2e["pageable"] = {"value": n_total * K / (msp * 1e-3), "ms_per_step": msp / K, "host_buffers": "pageable numpy array, staged both ways (kz_upload / kz_download)"}
            # batched host vectors: vector k's download overlaps vector k+1's upload (full duplex)
            nb, bl = 4, args.logn - 2
            wb = ints_to_limbs([pow(GEN[curve], (r_mod - 1) >> bl, r_mod)], r_mod)[0]
            batch = pinned.array.reshape(nb, (1 << bl), 4)
            for _ in range(2):
                device.ntt(curve, batch.reshape(-1, 4), wb, batch=nb)
            t0 = time.perf_counter()
            for _ in range(K):
                device.ntt(curve, batch.reshape(-1, 4), wb, batch=nb)
            msb = (time.perf_counter() - t0) * 1e3
            e2e["batched"] = {"vectors": nb, "logn": bl, "value": n_total * K / (msb * 1e-3), "ms_per_step": msb / K,
                              "note": "kzgpu_ntt_batch from pinned host memory: per-vector pipeline, uploads and downloads on separate copy streams"}
        clocks = sampler.stop()
        ntt_ms = pr["ms"] / K                                   # all passes of one transform
        hbm = 64.0 * n_total / (ntt_ms * 1e-3) / 1e9
        modmuls = 0.5 * n_total * args.logn
        res = {
            "value": world * n_total * K / (ms * 1e-3), "ms": ms, "launches": launches, "clocks": clocks, "e2e": e2e,
            "roofline": {
                "bound": "hbm", "kernel": f"ntt_pass_kernel_c<Fr {curve}, 8> (all passes of one transform)",
                "achieved": hbm, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": hbm / peaks.get("hbm_gbs"),
                "traffic": load_traffic("ntt_pass_kernel") if (args.logn == 24 and curve == "bn254") else None, "peak_source": peak_src,
                "model": "SURVEY 8(d): algorithmic bytes = 2 x 32 B x n (twiddles not counted)",
                "passes": pr["launches"] // K, "kernel_ms": ntt_ms,
                "imad_model_frac": modmuls * IMAD32_PER_MODMUL / (ntt_ms * 1e-3) / imad32_peak,
                "imad_model": "SURVEY 8(d): (n/2) log2 n modmul x 272 IMAD32 -- the binding bound for 256-bit fields, but a MODEL: radix-8 "
                              "butterflies have trivial twiddles the model charges for; the executed figure is ncu's "
                              "sm__pipe_fma_cycles_active / 50 % (profiles/: 0.82 at 2^24)",
            },
        }
        d.free(); pinned.free()
        return res

    # ------------------------------------------------------------------ sweeps (configs[1], configs[2])
    def bench_sweep():
        """NTT 2^12..2^26 and MSM 2^16..2^26 on one GPU (per rank at N > 1: independent replicas are not what this workload is for,
        so it runs on rank 0's GPU only).  Per size: device-resident ms (CUDA events, median), host-buffer ms (pinned host arrays,
        copies inside the call), the CPU port (measured where it finishes in ~a second, else extrapolated and marked `~`), and a
        size-independent check: NTT -- Horner spot checks on the device + inverse round trip; MSM -- the tau-identity."""
        rows = {"ntt": [], "msm": []}
        med = lambda ts: sorted(ts)[len(ts) // 2]                        # noqa: E731
        reps = max(3, min(K, 7))
        cpu_ntt = {}
        if not args.no_cpu:
            for lg in (12, 14, 16):
                cpu_ntt[lg] = (1 << lg) / cpu_ntt_rate(lg, 1, "port", curve)[0]
            base_msm = 1.0 / cpu_msm_rate(1024, 1, "port", curve)[0]                  # seconds per point
        for lg in range(12, min(args.sweep_max, 26) + 1, 2):
            n = 1 << lg
            w = pow(GEN[curve], (r_mod - 1) // n, r_mod)
            wl = ints_to_limbs([w], r_mod)[0]
            pin = _ffi.PinnedArray((n, 4))
            x = random_scalars(n, r_mod, seed=lg)
            pin.array[:] = x
            d = _ffi.DeviceBuffer(n * 32).upload(x)
            keep = _ffi.DeviceBuffer(n * 32).upload(x)
            for _ in range(3):
                device.ntt_dev(curve, d, n, wl)
            ts = []
            for _ in range(reps):
                _ffi.timer_start(); device.ntt_dev(curve, d, n, wl); ts.append(_ffi.timer_stop())
            # check: forward transform of x, out[k] == p_x(w^k) at three k (device Horner), then the inverse returns x
            d.upload(x)
            device.ntt_dev(curve, d, n, wl)
            y = np.zeros_like(x); d.download(y)
            ok = all(limbs_to_ints(y[k:k + 1])[0] == horner_dev(curve, keep, n, pow(w, k, r_mod)) for k in (0, 1, n // 2 + 3))
            device.ntt_dev(curve, d, n, wl, inverse=True)
            z = np.zeros_like(x); d.download(z)
            ok = ok and bool((z == x).all())
            for _ in range(2):
                device.ntt(curve, pin.array, wl)
            th = []
            for _ in range(reps):
                t0 = time.perf_counter(); device.ntt(curve, pin.array, wl); th.append((time.perf_counter() - t0) * 1e3)
            cpu_s, cpu_mark = None, ""
            if cpu_ntt:
                if lg in cpu_ntt:
                    cpu_s = cpu_ntt[lg]
                else:
                    cpu_s, cpu_mark = cpu_ntt[16] * (n * lg) / ((1 << 16) * 16), "~"        # fft_ff is Theta(n log n)
            rows["ntt"].append({"logn": lg, "device_ms": med(ts), "host_buffer_ms": med(th), "elements_per_s": n / med(ts) * 1e3,
                                "cpu_port_s": cpu_s, "cpu_extrapolated": cpu_mark == "~", "check": bool(ok)})
            d.free(); keep.free(); pin.free()
        for lg in range(16, min(args.sweep_max, 26) + 1, 2):
            n = 1 << lg
            t0 = time.perf_counter()
            srs = device.Srs.generate(curve, TAU, n)
            _ffi.check(_ffi._lib.kzgpu_sync())
            tb = time.perf_counter() - t0
            pin = _ffi.PinnedArray((n, 4))
            pin.array[:] = random_scalars(n, r_mod, seed=100 + lg)
            d = _ffi.DeviceBuffer(n * 32).upload(pin.array)
            for _ in range(2):
                out, inf = device.msm_dev(srs, d, n)
            ts = []
            for _ in range(reps):
                _ffi.timer_start(); out, inf = device.msm_dev(srs, d, n); ts.append(_ffi.timer_stop())
            ok = tau_identity(curve, out, inf, horner_dev(curve, d, n, TAU % r_mod))
            for _ in range(2):
                device.msm(srs, pin.array)
            th = []
            for _ in range(reps):
                t0 = time.perf_counter(); device.msm(srs, pin.array); th.append((time.perf_counter() - t0) * 1e3)
            key = srs.info()
            rows["msm"].append({"logn": lg, "device_ms": med(ts), "host_buffer_ms": med(th), "points_per_s": n / med(ts) * 1e3,
                                "cpu_port_s": None if args.no_cpu else base_msm * n, "cpu_extrapolated": not args.no_cpu,
                                "key": key, "srs_build_s": tb, "check": bool(ok)})
            d.free(); pin.free(); srs.destroy()
        return rows

    # ------------------------------------------------------------------ the Python-object boundary (SURVEY.md section 7 hard part 3)
    def bench_python_boundary(logn=18):
        """What the reference's callers pay at the boundary they actually use: `KZG.commit(ck, [poly])` and `fft_ff(list, w, F)`
        with Python objects in and out (kzg.py:110,115 `poly.list()` / `int(coeff)`; fft_ff.py:3), next to the same work
        through the C ABI with a limb array, and the marshalling alone."""
        from kzg_snark_b200.kzg import KZG
        from kzg_snark_b200 import fft_ff as gff
        n = 1 << logn
        kzg = KZG("bn254")
        t0 = time.perf_counter()
        ck, _ = kzg.setup(n - 1, tau=TAU)                      # device SRS + the Python list of n points the reference API returns
        t_setup = time.perf_counter() - t0
        F = kzg.Fq
        vals = limbs_to_ints(random_scalars(n, R_BN254, seed=logn))
        poly = kzg.R([F(v) for v in vals])
        best = lambda f, reps=3: min(_timeit(f) for _ in range(reps))            # noqa: E731
        t_marshal = best(lambda: kzg._coeff_limbs(poly))
        limbs = kzg._coeff_limbs(poly)
        device.msm(ck.srs, limbs)
        t_cabi = best(lambda: device.msm(ck.srs, limbs))
        kzg.commit(ck, [poly])
        t_commit = best(lambda: kzg.commit(ck, [poly]))
        w = F(pow(5, (R_BN254 - 1) // n, R_BN254))
        xs = [F(v) for v in vals]
        gff.fft_ff(xs, w, F)
        t_fft = best(lambda: gff.fft_ff(xs, w, F))
        arr = ints_to_limbs(vals, R_BN254)
        wl = ints_to_limbs([int(w)], R_BN254)[0]
        t_fft_cabi = best(lambda: device.ntt("bn254", arr, wl))
        return {"coefficients": n, "setup_s": t_setup, "commit_python_objects_s": t_commit, "commit_marshalling_s": t_marshal,
                "commit_c_abi_s": t_cabi, "fft_ff_python_objects_s": t_fft, "fft_c_abi_s": t_fft_cabi,
                "note": "python objects = shim polynomial / list of field elements in, py_ecc-shaped point / list of elements out; "
                        "c_abi = the same call with a (n, 4) uint64 limb array (pageable numpy memory), PCIe included"}

    # ------------------------------------------------------------------ PLONK prove (configs[3])
    def bench_plonk():
        """End-to-end `Prover.prove` (kzg_snark_b200/plonk.py, the device-resident counterpart of
        plonk/prover.py:24) on (a) the reference's bundled 16-gate instance, checked bit for bit
        against the proof the reference's own prover produced in the build container, and (b) a
        synthetic circuit of 2^plonk_logn gates.  Wall-clock seconds, host lists / arrays in,
        proof (9 points + 6 scalars) out."""
        from kzg_snark_b200.plonk import Indexer, Prover
        from kzg_snark_b200.plonk_synth import synthetic_circuit
        gold = os.path.join(ROOT, "tests", "golden")
        H = lambda v: int(v, 16)                                      # noqa: E731
        d = json.load(open(os.path.join(gold, "ref_plonk_normalized.json")))
        inst = json.load(open(os.path.join(gold, "plonk_instance.json")))
        tr = json.load(open(os.path.join(gold, "ref_trace_plonk.json")))
        sel = [[H(v) for v in inst[k]] for k in ("qM", "qL", "qR", "qO", "qC")]
        nb = d["n"]
        idx = Indexer("bn254")
        ipk, _ = idx.preprocess(*sel, [H(v) for v in inst["perm"]], max_degree=nb + 5, tau=H(d["index_draws"][0]),
                                k1=H(d["k1"]), k2=H(d["k2"]))
        xs = [idx.kzg.Fq(H(v)) for v in d["x"]]
        ws = [H(v) for v in d["w"]]
        bl = [H(b) for b in d["prover_draws"][-11:]]
        prover = Prover("bn254")
        times = []
        for i in range(Wm + K):
            t0 = time.perf_counter()
            proof = prover.prove(ipk, xs, ws, blinders=bl)
            times.append(time.perf_counter() - t0)
        times = sorted(times[Wm:])
        same = all((int(proof[sec][k][0]), int(proof[sec][k][1])) == (H(v[0]), H(v[1])) if isinstance(v, list)
                   else int(proof[sec][k]) == H(v) for sec, body in d["proof"].items() for k, v in body.items())
        out = {"bundled": {"gates": nb, "prove_s": times[len(times) // 2], "proof_equals_reference_prover": bool(same),
                           "reference_prove_s": tr["notes"]["prove_seconds"],
                           "reference_note": "plonk/prover.py run unmodified in the build container on the Sage/py_ecc stand-ins "
                                             "(oracle/refrun.py), 1 core; recorded in tests/golden/ref_trace_plonk.json"}}
        n = 1 << args.plonk_logn
        t0 = time.perf_counter()
        qM, qL, qR, qO, qC, perm, w = synthetic_circuit(n, 16, R_BN254, seed=args.plonk_logn)
        wpin = _ffi.PinnedArray((3 * n - 16, 4))                      # the witness lives in page-locked host memory
        wpin.array[:] = ints_to_limbs(w[16:], R_BN254)
        wl = wpin.array
        t_gen_plonk = time.perf_counter() - t0
        t0 = time.perf_counter()
        ipk, _ = idx.preprocess(qM, qL, qR, qO, qC, perm, max_degree=n + 5, tau=TAU, k1=7, k2=13)
        _ffi.check(_ffi._lib.kzgpu_sync())
        t_index = time.perf_counter() - t0
        xs = [idx.kzg.Fq(v) for v in w[:16]]
        times, l0 = [], 0
        for i in range(3 + 5):
            if i == 3:
                l0 = _ffi.launch_count()
            t0 = time.perf_counter()
            prover.prove(ipk, xs, wl)
            times.append(time.perf_counter() - t0)
        launches = (_ffi.launch_count() - l0) // 5
        assert prover.last_r_zeta == 0 and not any(prover.last_t_top), "r(zeta) != 0: the proof would not verify"
        times = sorted(times[3:])
        out["bundled"].update({k: v for k, v in dropin_hotpath("plonk").items() if k != "reference_prove_s"})
        out["marlin_bundled"] = {**dropin_hotpath("marlin"),
                                 "note": "configs[4] bundled R1CS instance: the 19 commit / open / fft_ff / fft_ff_interpolation calls "
                                         "marlin/prover.py made (tests/golden/ref_trace_marlin.json), replayed through the drop-in"}
        if not args.no_cpu:
            for nm, key in (("plonk", "bundled"), ("marlin", "marlin_bundled")):
                try:
                    out[key]["reference_prover_on_gpu_dropin"] = reference_prover_on_gpu_dropin(nm)
                except Exception as exc:                                 # the bench line must survive a missing reference tree
                    out[key]["reference_prover_on_gpu_dropin"] = {"error": repr(exc)[:200]}
        # configs[4]: the device Marlin prover on the bundled R1CS instance (proof compared with the reference prover's) and on a
        # synthetic R1CS of 2^marlin_rows_logn rows
        from kzg_snark_b200 import marlin
        dm = json.load(open(os.path.join(gold, "ref_marlin_normalized.json")))
        r1 = json.load(open(os.path.join(gold, "r1cs_instance.json")))
        mats = [[[H(v) for v in row] for row in r1[k]] for k in "ABC"]
        midx = marlin.Indexer("bn254")
        mipk, _ = midx.preprocess(*mats, max_degree=200, tau=H(dm["index_draws"][0]))
        mx, mw, mdr = [midx.kzg.Fq(H(v)) for v in dm["x"]], [H(v) for v in dm["w"]], [H(v) for v in dm["prover_draws"]]
        mpr = marlin.Prover("bn254")
        mt = []
        for i in range(Wm + K):
            t0 = time.perf_counter()
            mproof = mpr.prove(mipk, mx, mw, draws=mdr)
            mt.append(time.perf_counter() - t0)
        mt = sorted(mt[Wm:])
        pt2 = lambda P: (int(P[0]), int(P[1]))                           # noqa: E731
        msame = (all([pt2(p) for p in mproof["commitments"][k]] == [(H(q[0]), H(q[1])) for q in v] for k, v in dm["proof"]["commitments"].items())
                 and all([int(e) for e in mproof["evaluations"][k]] == [H(q) for q in v] for k, v in dm["proof"]["evaluations"].items())
                 and all(pt2(mproof["kzg_proofs"][k]) == (H(v[0]), H(v[1])) for k, v in dm["proof"]["kzg_proofs"].items()))
        out["marlin_bundled"].update({"prove_s": mt[len(mt) // 2], "proof_equals_reference_prover": bool(msame)})
        out["marlin_synthetic"] = marlin_synthetic_prove(args.marlin_rows_logn)
        out["synthetic"] = {"gates": n, "prove_s": times[len(times) // 2], "index_s": t_index, "circuit_generation_s": t_gen_plonk,
                            "h2d_bytes": int(wl.nbytes), "host_buffers": "pinned (cudaHostAlloc)", "gpu_launches_per_prove": int(launches),
                            "rounds_s": {k: round(v, 5) for k, v in prover.timings.items()},
                            "checks": "r(zeta) == 0 and deg t <= 3n+5 asserted; verifier acceptance at this construction is "
                                      "covered by tests/test_gpu_plonk.py up to 2^14 gates"}
        wpin.free()
        return out

    # ------------------------------------------------------------------ Marlin kernel workload (configs[4])
    def bench_marlin():
        """The commit / open / NTT calls a Marlin proof over a 2^marlin_logn-constraint R1CS implies
        (call sites marlin/prover.py:106,142,176,226-227,439-449,469 and
        marlin/encoder.py:123-125 with |H| = n, |K| = m = 2n, b = 2).  The reference prover itself cannot
        run at this size (dense Sage matrices).  Items are independent: item j runs on rank j % world
        with the full SRS replicated on every GPU; no data-path collective."""
        import ctypes
        n_ = 1 << args.marlin_logn
        m_ = 2 * n_
        commits = [n_ + 2] * 4 + [n_ + 4, 2 * n_ + 1] + [n_, n_ - 1, n_ + 2] + [m_ - 1, 6 * m_ - 6]
        ntts = [(m_, True)] * 9 + [(n_, True)] * 4 + [(m_, False)] * 9 + [(m_, True)]
        opens = [[2 * n_ + 2, m_ - 1, n_ + 2, n_], [6 * m_ - 6] + [m_] * 6]
        srs_n = 6 * m_
        srs = device.Srs.generate("bn254", TAU, srs_n)
        lib = _ffi._lib
        from kzg_snark_b200.parallel import lpt_assign, shard_range
        L = lambda v: _ffi.ptr(ints_to_limbs([v], R_BN254)[0])           # noqa: E731
        SHARD_MIN = 1 << 22          # MSMs at least this long are point-sharded over all ranks instead of owned by one

        def rand_dev(cnt, seed):
            return _ffi.DeviceBuffer(cnt * 32).upload(random_scalars(cnt, R_BN254, seed=seed))

        # work items: (kind, index, cost); big MSMs are split over every rank (cost / world each), the rest are
        # assigned whole, longest first (parallel.lpt_assign)
        sharded_c = [j for j, ln in enumerate(commits) if world > 1 and ln >= SHARD_MIN]
        sharded_o = [j for j, ls in enumerate(opens) if world > 1 and max(ls) >= SHARD_MIN]
        whole = [("c", j, commits[j]) for j in range(len(commits)) if j not in sharded_c]
        whole += [("n", j, ntts[j][0] // 8) for j in range(len(ntts))]
        whole += [("o", j, max(opens[j]) + sum(opens[j]) // 4) for j in range(len(opens)) if j not in sharded_o]
        owner = lpt_assign([w[2] for w in whole], world)
        mine = [w for w, o in zip(whole, owner) if o == rank]
        my_c = [(j, rand_dev(commits[j], 500 + j)) for kind, j, _ in mine if kind == "c"]
        my_n = [(j, rand_dev(ntts[j][0], 600 + j)) for kind, j, _ in mine if kind == "n"]
        my_o = [(j, [rand_dev(c, 700 + 10 * j + i) for i, c in enumerate(opens[j])]) for kind, j, _ in mine if kind == "o"]
        sh_c = []                                                        # (start, count, this rank's slice of the scalars)
        for j in sharded_c:
            s0, cnt = shard_range(commits[j], world, rank)
            sh_c.append((s0, cnt, rand_dev(cnt, 800 + 16 * j + rank)))
        sh_o = [(j, [rand_dev(c, 700 + 10 * j + i) for i, c in enumerate(opens[j])], _ffi.DeviceBuffer(max(opens[j]) * 32))
                for j in sharded_o]                                      # polynomials replicated, quotient scratch
        if world > 1:
            partial = torch.zeros(128, dtype=torch.uint8, device="cuda")
            gathered = torch.zeros(128 * world, dtype=torch.uint8, device="cuda")
        w_of = {sz: ints_to_limbs([pow(5, (R_BN254 - 1) // sz, R_BN254)], R_BN254)[0] for sz in (n_, m_)}
        zl, xil = 0x1234567 % R_BN254, 0x7654321 % R_BN254

        def gather_fold():
            dist.all_gather_into_tensor(gathered, partial)
            return device.g1_fold("bn254", RawPtr(gathered.data_ptr()), world)

        def step():
            for j, d in my_c:
                device.msm_dev(srs, d, commits[j])
            for j, d in my_n:
                device.ntt_dev("bn254", d, ntts[j][0], w_of[ntts[j][0]], inverse=ntts[j][1])
            for j, ds in my_o:
                k = len(ds)
                out = np.zeros(8, dtype=np.uint64)
                fl = ctypes.c_int(0)
                ptrs = (ctypes.c_void_p * k)(*[d.ptr.value for d in ds])
                lens = (ctypes.c_size_t * k)(*opens[j])
                _ffi.check(lib.kzgpu_open_dev(srs.handle, ptrs, lens, k, L(zl), L(xil), _ffi.ptr(out), ctypes.byref(fl), None))
            for s0, cnt, d in sh_c:                                      # point-sharded commit: partial -> all-gather -> fold
                device.msm_partial_dev(srs, d, cnt, RawPtr(partial.data_ptr()), first=s0)
                gather_fold()
            for j, ds, dq in sh_o:                                       # point-sharded open: quotient everywhere, MSM by range
                k = len(ds)
                ptrs = (ctypes.c_void_p * k)(*[d.ptr.value for d in ds])
                lens = (ctypes.c_size_t * k)(*opens[j])
                ql = ctypes.c_size_t(0)
                _ffi.check(lib.kzgpu_open_quotient_dev(_ffi.BN254, ptrs, lens, k, L(zl), L(xil), dq.ptr, ctypes.byref(ql), None))
                s0, cnt = shard_range(ql.value, world, rank)
                device.msm_partial_dev(srs, RawPtr(dq.ptr.value + 32 * s0), cnt, RawPtr(partial.data_ptr()), first=s0)
                gather_fold()

        for _ in range(Wm):
            step()
        sampler = ClockSampler(local)
        barrier()
        l0 = _ffi.launch_count()
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        barrier()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        return {"ms": ms, "launches": _ffi.launch_count() - l0, "clocks": sampler.stop(),
                "items": {"commits": commits, "ntts": [[a, "inverse" if b else "forward"] for a, b in ntts], "opens": opens},
                "srs_points": srs_n, "srs": srs.info()}

    dtype = "u32x8 (256-bit modular integers, Montgomery)" if curve == "bn254" else \
        "u32x12 base field / u32x8 scalars (384 / 256-bit modular integers, Montgomery)"

    pending = []

    def finish():
        # the communicator is torn down FIRST, so that with NCCL_DEBUG=INFO its chatter precedes the result line, which is
        # the last thing this process prints (rank 0 only)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        for text in pending:
            sys.stdout.write("\n" + text + "\n")
        sys.stdout.flush()
        return 0

    def emit(line):
        pending.append(json.dumps(line))

    if args.workload == "marlin":
        res = bench_marlin()
        if rank == 0:
            emit({"metric": "marlin_kernel_workload_s", "value": res["ms"] / K / 1e3, "unit": "s", "n_gpus": world,
                  "steps": K, "warmup": Wm, "ms_per_step": res["ms"] / K, "higher_is_better": False,
                  "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                  "config": {"workload": f"commit/open/NTT calls of one Marlin proof, 2^{args.marlin_logn} constraints "
                                         f"(|H|=2^{args.marlin_logn}, |K|=2^{args.marlin_logn + 1}) on {world} GPUs: MSMs of >= 2^22 points point-sharded over all ranks, "
                                         "the rest assigned longest-first, SRS replicated", "items": res["items"], "srs_points": res["srs_points"],
                             "srs_layout": res["srs"]},
                  "clocks": res["clocks"], "gpu_launches": res["launches"], "device": info["name"]})
        return finish()

    if args.workload == "sweep":
        rows = bench_sweep() if rank == 0 else None
        if rank == 0:
            top = rows["msm"][-1] if rows["msm"] else {"points_per_s": 0.0, "device_ms": 0.0}
            emit({"metric": "g1_msm_points_per_s", "value": top["points_per_s"], "unit": "points/s", "n_gpus": 1, "steps": K, "warmup": Wm,
                  "ms_per_step": top["device_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype,
                  "data": "synthetic",
                  "config": {"workload": f"configs[1] + configs[2] sweeps on one GPU, {curve}: NTT 2^12..2^{min(args.sweep_max, 26)}, MSM 2^16..2^{min(args.sweep_max, 26)}; "
                                         "value = the largest MSM of the sweep", "curve": curve,
                             "columns": "device_ms: resident inputs, CUDA events, median; host_buffer_ms: pinned host arrays in / result out, wall clock; "
                                        "cpu_port_s: oracle port on 1 core (cpu_extrapolated marks n log n / linear extrapolation); "
                                        "check: NTT Horner spot checks + inverse round trip, MSM tau-identity commit == p(tau) * G1"},
                  "sweep": rows, "device": info["name"]})
            sys.stderr.write("\n| NTT n | device ms | host-buffer ms | elements/s | CPU port s | check |\n|---|---|---|---|---|---|\n")
            for r_ in rows["ntt"]:
                cpu = "-" if r_["cpu_port_s"] is None else f"{'~' if r_['cpu_extrapolated'] else ''}{r_['cpu_port_s']:.3g}"
                sys.stderr.write(f"| 2^{r_['logn']} | {r_['device_ms']:.3f} | {r_['host_buffer_ms']:.3f} | {r_['elements_per_s']:.3e} | {cpu} | {'ok' if r_['check'] else 'FAIL'} |\n")
            sys.stderr.write("\n| MSM n | device ms | host-buffer ms | points/s | CPU port s | key | build s | check |\n|---|---|---|---|---|---|---|---|\n")
            for r_ in rows["msm"]:
                cpu = "-" if r_["cpu_port_s"] is None else f"~{r_['cpu_port_s']:.3g}"
                k_ = r_["key"]
                sys.stderr.write(f"| 2^{r_['logn']} | {r_['device_ms']:.3f} | {r_['host_buffer_ms']:.3f} | {r_['points_per_s']:.3e} | {cpu} | "
                                 f"c={k_['c']} W={k_['tables']} {k_['bytes'] / 2**30:.2f} GiB | {r_['srs_build_s']:.2f} | {'ok' if r_['check'] else 'FAIL'} |\n")
        return finish()

    primary = bench_msm() if args.workload == "msm" else (bench_ntt() if args.workload == "ntt" else None)
    secondary = None
    if args.workload == "msm" and not args.no_secondary:
        secondary = bench_ntt()
    plonk = None
    if world == 1 and curve == "bn254" and (args.workload == "plonk" or not args.no_secondary):
        plonk = bench_plonk()
        plonk["python_boundary"] = bench_python_boundary()
    if args.workload == "plonk":
        if rank == 0:
            emit({"metric": "plonk_prove_s", "value": plonk["synthetic"]["prove_s"], "unit": "s", "n_gpus": 1,
                  "steps": 5, "warmup": 3, "ms_per_step": 1e3 * plonk["synthetic"]["prove_s"],
                  "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                  "config": {"workload": f"PLONK prove, synthetic circuit of 2^{args.plonk_logn} gates, BN254"},
                  "plonk": plonk, "device": info["name"]})
        return finish()

    # ------------------------------------------------------------------ configs[4] at N > 1: the Marlin prove on all the job's GPUs
    marlin_multi = None
    if world > 1 and args.workload == "msm" and not args.no_secondary:
        # The prover is one process (its polynomials live on one device); with several devices the LIBRARY spreads each round's
        # commitments and openings (kzgpu_init_multi).  Rank 0 therefore re-initialises the library on the job's N GPUs while the
        # other ranks -- idle, their buffers freed -- wait on a CPU-side (gloo) barrier.
        try:
            cpu_group = dist.new_group(backend="gloo")
        except Exception:                                                # no gloo: every rank skips this leg alike
            cpu_group = None
        barrier()
        if rank == 0 and cpu_group is not None:
            try:
                _ffi.shutdown()
                _ffi.init_multi(list(range(world)))
                marlin_multi = marlin_synthetic_prove(args.marlin_rows_logn)
            except Exception as exc:                                     # the primary line must survive
                marlin_multi = {"error": repr(exc)[:300]}
            try:
                _ffi.shutdown()
            except Exception:
                pass
        if cpu_group is not None:
            dist.barrier(group=cpu_group)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if args.workload == "msm":
            v, wall = cpu_msm_rate(4096 if curve == "bn254" else 1536, 1, "port", curve)
            cpu = {"value": v, "unit": "points/s", "cores": 1, "kind": "port",
                   "sample": f"{4096 if curve == 'bn254' else 1536}-point commit (kzg.py:112-116 loop on the restated py_ecc arithmetic), {wall:.1f} s",
                   "host_cpus": os.cpu_count()}
            if secondary is not None:
                v2, wall2 = cpu_ntt_rate(17, 1, "port", curve)
                secondary["cpu_baseline"] = {"value": v2, "unit": "elements/s", "cores": 1, "kind": "port",
                                             "sample": f"2^17-element recursive fft_ff (fft_ff.py:3-37 on CPython ints), {wall2:.1f} s"}
        else:
            v, wall = cpu_ntt_rate(18, 1, "port", curve)
            cpu = {"value": v, "unit": "elements/s", "cores": 1, "kind": "port",
                   "sample": f"2^18-element recursive fft_ff (fft_ff.py:3-37 on CPython ints), {wall:.1f} s",
                   "host_cpus": os.cpu_count()}

    cpu_plonk = None
    if rank == 0 and world == 1 and not args.no_cpu and plonk is not None:
        cpu_plonk = cpu_hotpath("plonk")

    if rank == 0:
        if args.workload == "msm":
            metric, unit = "g1_msm_points_per_s", "points/s"
            key = primary["key"]
            workload = (f"{curve} G1 MSM (KZG commit) of 2^{args.logn} points" +
                        (f", ONE MSM split over {world} GPUs: SRS / scalars sharded by contiguous index range ({primary['n_per_rank']} points per rank), "
                         "NCCL all-gather of the XYZZ partials + fold" if world > 1 else ""))
            pt = 2 * nl * 4
            footprint = (f"inputs per GPU: SRS {primary['n_per_rank'] * pt / 2**20:.0f} MiB (x {key['tables']} window tables = {key['bytes'] / 2**30:.2f} GiB) "
                         f"+ scalars {primary['n_per_rank'] * 32 / 2**20:.0f} MiB; the accumulate kernel gathers from the tables, far larger than the 126 MB L2")
            scaling = "strong"
        else:
            metric, unit = "ntt_elements_per_s", "elements/s"
            workload = f"{curve} scalar-field NTT of 2^{args.logn} elements" + (f", one vector per GPU ({world} replicas)" if world > 1 else "")
            footprint = "512 MiB vector at 2^24 exceeds the 126 MB L2"
            scaling = "weak"
        line = {
            "metric": metric, "value": primary["value"], "unit": unit, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": primary["ms"] / K, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload, "curve": curve, "logn": args.logn, "l2": footprint,
                       "srs": "tau^i*G1 generated on device from a fixed tau", "scalars": "uniform in [0,r), numpy PCG64"},
            "clocks": primary["clocks"], "e2e": primary["e2e"], "gpu_launches": primary["launches"],
            "roofline": primary["roofline"], "cpu_baseline": cpu,
            "imad_wide_peak_per_s": imad_wide_peak, "device": info["name"],
        }
        if args.workload == "msm":
            line["config"].update({"srs_window_bits": key["c"], "srs_window_tables": key["tables"], "srs_table_bytes_per_gpu": key["bytes"],
                                   "srs_build_s": round(primary["build_s"], 3), "tau_identity_check": primary["check"],
                                   "scaling_note": "N = 1 and N > 1 run the same total work (2^logn points): v_N / (N v_1) is strong-scaling efficiency"})
        if "profile_ms_per_step" in primary:
            line["profile_ms_per_step"] = primary["profile_ms_per_step"]
        if "weak" in primary:
            line["weak"] = primary["weak"]
        if marlin_multi is not None:
            line["marlin"] = {**marlin_multi, "note": f"configs[4]: device Marlin prover, synthetic R1CS, ONE process driving the job's {world} GPUs "
                                                      "through kzgpu_init_multi (commitments / openings point-sharded or dealt over the devices inside "
                                                      "the library); the N = 1 figure is plonk.marlin_synthetic of the 1-GPU line"}
        if plonk is not None:
            if cpu_plonk is not None:
                plonk["bundled"]["cpu_port_hotpath_s"] = cpu_plonk
                plonk["marlin_bundled"]["cpu_port_hotpath_s"] = cpu_hotpath("marlin")
            line["plonk"] = plonk
        if secondary is not None:
            line["ntt"] = {"metric": "ntt_elements_per_s", "value": secondary["value"], "unit": "elements/s",
                           "ms_per_step": secondary["ms"] / K, "e2e": secondary["e2e"], "roofline": secondary["roofline"],
                           "gpu_launches": secondary["launches"], "clocks": secondary["clocks"],
                           "cpu_baseline": secondary.get("cpu_baseline"), "scaling": "weak (one vector per GPU)" if world > 1 else None,
                           "config": {"workload": f"{curve} scalar-field NTT of 2^{args.logn} elements, natural order in/out" +
                                                  (f", one vector per GPU ({world} replicas)" if world > 1 else "")}}
        emit(line)
    return finish()


def run_inproc(args):
    """Multi-GPU INSIDE the library (kzgpu_init_multi): one plain python process, no torch, no launcher.
    One step = one 2^logn-point MSM through `kzgpu_msm` with pinned host scalars (the product entry point); beside it the same
    MSM from scalars resident on the primary device (peers pull their slices over NVLink), one batched commit of the 11
    Marlin-sized polynomials (placed longest-first / point-sharded) and one batched NTT of 8 host vectors."""
    import numpy as np
    from kzg_snark_b200 import _ffi, device
    from kzg_snark_b200.limbs import random_scalars, ints_to_limbs
    curve, r_mod = args.curve, FR[args.curve]
    K, Wm = args.steps, args.warmup
    n = 1 << args.logn
    _ffi.init_multi(None if args.devices <= 0 else list(range(args.devices)))
    nd = _ffi.device_count()
    info = _ffi.device_info()
    t0 = time.perf_counter()
    srs = device.Srs.generate(curve, TAU, n)
    _ffi.check(_ffi._lib.kzgpu_sync())
    build_s = time.perf_counter() - t0
    pinned = _ffi.PinnedArray((n, 4))
    pinned.array[:] = random_scalars(n, r_mod, seed=args.logn * 100)
    dsc = _ffi.DeviceBuffer(n * 32).upload(pinned.array)
    t0 = time.perf_counter()
    out, inf = device.msm(srs, pinned.array)                        # first call: builds the per-device shards (peer copy + tables)
    first_s = time.perf_counter() - t0
    ok = tau_identity(curve, out, inf, horner_dev(curve, dsc, n, TAU % r_mod))

    def timed(fn, reps):
        for _ in range(Wm):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) * 1e3 / reps

    sampler = ClockSampler(0)
    l0 = _ffi.launch_count()
    ms_host = timed(lambda: device.msm(srs, pinned.array), K)
    launches = (_ffi.launch_count() - l0) // (K + Wm)
    ms_dev = timed(lambda: device.msm_dev(srs, dsc, n), K)
    # batched commit of Marlin-sized polynomials (marlin/prover.py:106,142,176 at 2^20 constraints, scaled to the key)
    b = n // 16
    lens = [b + 2] * 4 + [b + 4, 2 * b + 1] + [b, b - 1, b + 2] + [2 * b - 1, 12 * b - 6]
    packed = _ffi.PinnedArray((sum(lens), 4))
    off = 0
    for j, ln in enumerate(lens):
        packed.array[off:off + ln] = random_scalars(ln, r_mod, seed=60 + j)
        off += ln
    ms_batch = timed(lambda: device.msm_batch_packed(srs, packed.array, lens), max(2, K // 2))
    # 8 host vectors of 2^(logn-3) elements each through kzgpu_ntt_batch
    m = n // 8
    wl = ints_to_limbs([pow(GEN[curve], (r_mod - 1) // m, r_mod)], r_mod)[0]
    vecs = pinned.array.reshape(8 * m, 4)
    ms_ntt = timed(lambda: device.ntt(curve, vecs, wl, batch=8), max(2, K // 2))
    pinned.free(); dsc.free(); srs.destroy(); packed.free()
    # configs[2] across devices: one MSM of 2^k points through kzgpu_msm_dev / kzgpu_msm on all the devices, k = 16 .. sweep_max
    # (below KZGPU_SHARD_MIN = 2^20 the library keeps the MSM on the primary device); tau-identity at every size
    sweep = []
    for lg in range(16, min(args.sweep_max, 26) + 1, 2):
        nn = 1 << lg
        t0 = time.perf_counter()
        s2 = device.Srs.generate(curve, TAU, nn)
        _ffi.check(_ffi._lib.kzgpu_sync())
        tb = time.perf_counter() - t0
        pin2 = _ffi.PinnedArray((nn, 4))
        pin2.array[:] = random_scalars(nn, r_mod, seed=300 + lg)
        d2 = _ffi.DeviceBuffer(nn * 32).upload(pin2.array)
        t0 = time.perf_counter()
        o2, f2 = device.msm_dev(s2, d2, nn)
        t_first = time.perf_counter() - t0
        ok2 = tau_identity(curve, o2, f2, horner_dev(curve, d2, nn, TAU % r_mod))
        reps = 5 if lg <= 24 else 3
        sweep.append({"logn": lg, "device_ms": timed(lambda: device.msm_dev(s2, d2, nn), reps),
                      "host_buffer_ms": timed(lambda: device.msm(s2, pin2.array), reps), "srs_build_s": tb,
                      "first_call_s": t_first, "check": bool(ok2)})
        sweep[-1]["points_per_s"] = nn / (sweep[-1]["device_ms"] * 1e-3)
        d2.free(); pin2.free(); s2.destroy()
    # configs[4]: the device Marlin prover on a synthetic R1CS of 2^marlin_rows_logn rows, its commitments and openings over all devices
    marlin_res = marlin_synthetic_prove(args.marlin_rows_logn)
    clocks = sampler.stop()
    line = {"metric": "g1_msm_points_per_s", "value": n / (ms_host * 1e-3), "unit": "points/s", "n_gpus": nd, "steps": K, "warmup": Wm,
            "ms_per_step": ms_host, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x8 (256-bit modular integers, Montgomery)" if curve == "bn254" else "u32x12 / u32x8", "data": "synthetic",
            "config": {"workload": f"{curve} G1 MSM of 2^{args.logn} points through kzgpu_msm from ONE process driving {nd} GPU(s) "
                                   "(kzgpu_init_multi): host scalars in pinned memory, each device uploads and reduces its shard, peer-to-peer "
                                   "partials, fold on the primary", "curve": curve, "logn": args.logn,
                       "srs_build_s": round(build_s, 3), "first_call_s_incl_shard_build": round(first_s, 3), "tau_identity_check": bool(ok)},
            "e2e": {"value": n / (ms_host * 1e-3), "unit": "points/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 68,
                    "ms_per_step": ms_host, "host_buffers": "pinned (cudaHostAlloc)"},
            "device_resident": {"ms_per_step": ms_dev, "value": n / (ms_dev * 1e-3),
                                "note": "scalars on the primary device; the other devices pull their slices over NVLink (cudaMemcpyPeerAsync)"},
            "batched_commit": {"polynomials": lens, "ms_per_call": ms_batch, "points_per_s": sum(lens) / (ms_batch * 1e-3),
                               "note": "kzgpu_msm_batch from one pinned host buffer: >= KZGPU_SHARD_MIN coefficients point-sharded over all devices "
                                       "(shard tables sized for the polynomial's length class), the rest placed longest-first (key replicated)"},
            "batched_ntt": {"vectors": 8, "n": m, "ms_per_call": ms_ntt, "elements_per_s": 8 * m / (ms_ntt * 1e-3),
                            "note": "kzgpu_ntt_batch from pinned host memory, whole vectors per device"},
            "msm_sweep": sweep,
            "marlin_synthetic": marlin_res,
            "gpu_launches": launches, "clocks": clocks, "device": info["name"], "cpu_baseline": None,
            "roofline": None}
    sys.stdout.write("\n" + json.dumps(line) + "\n")
    sys.stdout.flush()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="msm", choices=["msm", "ntt", "sweep", "plonk", "marlin", "inproc"])
    ap.add_argument("--devices", type=int, default=0, help="--workload inproc: GPUs driven from this ONE process (0 = all visible)")
    ap.add_argument("--curve", default="bn254", choices=["bn254", "bls12_381"])
    ap.add_argument("--marlin-logn", type=int, default=20, help="constraints of the Marlin kernel workload (log2)")
    ap.add_argument("--marlin-rows-logn", type=int, default=20, help="rows of the synthetic R1CS for the device Marlin prover (log2; configs[4] names 2^20)")
    ap.add_argument("--plonk-logn", type=int, default=20, help="gates of the synthetic PLONK circuit (log2)")
    ap.add_argument("--logn", type=int, default=24)
    ap.add_argument("--sweep-max", type=int, default=26, help="largest log2 size of --workload sweep")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the NTT half of the metric and the provers")
    ap.add_argument("--force-port", action="store_true", help="--impl reference: time the oracle port even when the reference tree is present")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "inproc":
        return run_inproc(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
