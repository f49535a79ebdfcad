#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for the KZG-MSM / NTT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload msm|ntt] [--logn 24] [--scaling weak|strong]

Default workload (`msm`): one step = one KZG commit of a 2^24-coefficient polynomial, i.e. one
BN254 G1 MSM of 2^24 points against the device-resident SRS (BASELINE.json metric "G1 MSM
points/s ... at 2^24"; configs[2]).  The same line also carries the NTT half of the metric
(`"ntt"`: BN254 scalar-field NTT of 2^24 elements, configs[1]) with its own roofline.
`--workload ntt` makes the NTT the primary metric instead.

N > 1 (torchrun, one process per GPU, NCCL): the SRS points and the scalars are sharded by
index range, every rank reduces its shard to one XYZZ partial sum, the partials are
all-gathered over NCCL (128 B per rank) and folded.  Default scaling is "weak" (2^logn points
per GPU); `--scaling strong` splits 2^logn points across the ranks.

`--impl reference` times the reference algorithm's CPU restatement (oracle/: py_ecc-style
double-and-add commit loop, kzg.py:112-116, recursive fft_ff, fft_ff.py:3-37) on all host
cores, on a bounded sample of the same workload.
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
TAU = 0x2545F4914F6CDD1D9E3779B97F4A7C15F39CC0605CEDC834 % R_BN254      # fixed synthetic trapdoor

# SURVEY.md section 8(d) work model: 16 windows x (8M+2S = 10 modmul) x 272 32-bit IMADs
MODEL_IMAD_PER_POINT = 43520
IMAD32_PER_MODMUL = 272
MODMUL_PER_MADD = 10
# what the accumulate kernel executes per mixed addition: 8 general products (64 + 64 + 8 wide multiplies = 272 IMAD32 each) and
# 2 dedicated squarings (36 + 64 + 8 wide = 216 IMAD32 each, field.cuh fe_sqr_nofinal)
IMAD32_EXECUTED_PER_MADD = 8 * 272 + 2 * 216


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/r1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            return json.load(f).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
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
    """Worker: reference commit loop (kzg.py:112-116) over a slice of (scalar, index) pairs."""
    seed, start, count, tau = args
    import random
    from oracle.kzg import KZGOracle
    ko = KZGOracle("bn254")
    rng = random.Random(seed)
    # SRS slice via the oracle's shared-doubling setup (not timed by the caller? it is part of
    # producing inputs) -- points tau^i * G1
    ck = ko.setup_fast(count - 1, tau)
    coeffs = [rng.randrange(ko.curve_order) for _ in range(count)]
    t0 = time.perf_counter()
    c = ko.commit(ck, [coeffs])[0]
    dt = time.perf_counter() - t0
    return dt, ko.cv.normalize(c)


def cpu_msm_rate(points_per_core, cores):
    """points/s of the restated reference commit loop using `cores` processes."""
    jobs = [(1000 + i, 0, points_per_core, TAU) for i in range(cores)]
    if cores == 1:
        res = [_cpu_commit_chunk(jobs[0])]
        wall = res[0][0]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_commit_chunk, jobs)
        wall = max(r[0] for r in res)          # commit loop time only (inputs prepared before)
    return points_per_core * cores / wall, wall


def _cpu_ntt_chunk(args):
    seed, logn = args
    import random
    from oracle.fft_ff import fft_ff_int
    rng = random.Random(seed)
    n = 1 << logn
    x = [rng.randrange(R_BN254) for _ in range(n)]
    w = pow(5, (R_BN254 - 1) // n, R_BN254)
    t0 = time.perf_counter()
    fft_ff_int(x, w, R_BN254)
    return time.perf_counter() - t0


def cpu_ntt_rate(logn, cores):
    """elements/s of the restated recursive fft_ff; `cores` independent vectors in parallel
    (the reference itself is single-threaded: batched vectors are its only parallelism)."""
    jobs = [(2000 + i, logn) for i in range(cores)]
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


def cpu_plonk_hotpath():
    return cpu_hotpath("plonk")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    vals, t_all = [], 0.0
    if args.workload == "msm":
        per_core = 256
        unit, metric = "points/s", "g1_msm_points_per_s"
        for i in range(args.warmup + args.steps):
            v, wall = cpu_msm_rate(per_core, cores)
            if i >= args.warmup:
                vals.append(v); t_all += wall
        sample = f"{per_core * cores} random BN254 scalars x SRS points per step ({per_core}/core), commit loop kzg.py:112-116"
        workload = f"BN254 G1 MSM (KZG commit) of 2^{args.logn} points"
    else:
        logn_s = 15
        unit, metric = "elements/s", "ntt_elements_per_s"
        for i in range(args.warmup + args.steps):
            v, wall = cpu_ntt_rate(logn_s, cores)
            if i >= args.warmup:
                vals.append(v); t_all += wall
        sample = f"{cores} vectors of 2^{logn_s} elements per step (one per core), recursive fft_ff.py:3-37"
        workload = f"BN254 scalar-field NTT of 2^{args.logn} elements"
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u32x8 (256-bit modular integers)", "data": "synthetic",
        "config": {"workload": workload, "curve": "bn254", "sample": sample},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                         "note": "restated reference algorithm on CPython ints (SageMath/py_ecc are not installable here)"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- GPU arm
class RawPtr:
    def __init__(self, p):
        import ctypes
        self.ptr = ctypes.c_void_p(p)


def on_curve_bn254(out):
    from kzg_snark_b200.limbs import limbs_to_ints
    p = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    x, y = limbs_to_ints(out.reshape(2, 4))
    return (y * y - x * x * x - 3) % p == 0


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
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _ffi.init(local)
    info = _ffi.device_info()
    peaks, peak_src = load_peaks()
    K, Wm = args.steps, args.warmup
    n_total = 1 << args.logn
    if world > 1 and args.scaling == "strong":
        n = n_total // world
    else:
        n = n_total                           # weak: per-GPU work fixed
    start = rank * n
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

    result = {}

    # ------------------------------------------------------------------ integer roofline (live)
    sm = info["sm_count"]
    ms_i, ops_i = _ffi.microbench(0, sm * 4, 256, 2000)
    imad_wide_peak = ops_i / ms_i * 1e3                   # IMAD.WIDE.U32 / s
    imad32_peak = 2.0 * imad_wide_peak                    # one wide = two 32-bit multiply-add slots

    # ------------------------------------------------------------------ MSM
    def bench_msm():
        srs = device.Srs.generate("bn254", TAU, n, start=start)
        pinned = _ffi.PinnedArray((n, 4))
        pinned.array[:] = random_scalars(n, R_BN254, seed=args.logn * 100 + rank)
        dsc = _ffi.DeviceBuffer(n * 32).upload(pinned.array)
        if world > 1:
            partial = torch.zeros(128, dtype=torch.uint8, device="cuda")
            gathered = torch.zeros(128 * world, dtype=torch.uint8, device="cuda")

        def step_resident():
            if world == 1:
                return device.msm_dev(srs, dsc, n)
            device.msm_partial_dev(srs, dsc, n, RawPtr(partial.data_ptr()))
            dist.all_gather_into_tensor(gathered, partial)
            return device.g1_fold("bn254", RawPtr(gathered.data_ptr()), world)

        def step_e2e():
            if world == 1:
                return device.msm(srs, pinned.array)       # H2D of the scalars inside the call
            device.msm_partial(srs, pinned.array, RawPtr(partial.data_ptr()))     # this rank's H2D inside the call, overlapped
            dist.all_gather_into_tensor(gathered, partial)
            return device.g1_fold("bn254", RawPtr(gathered.data_ptr()), world)

        for _ in range(Wm):
            if world > 1:
                gathered.zero_()                                 # a fold that ran ahead of the all-gather would see infinity
            out, inf = step_resident()
            assert not inf and on_curve_bn254(out), "MSM result is not a curve point"
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
        # per-kernel profile (second region, CUDA events around the kernels on the launching stream)
        _ffi.profile_reset(); _ffi.profile_enable(True)
        for _ in range(K):
            step_resident()
        _ffi.profile_enable(False)
        prof = {k: _ffi.profile_get(i) for k, i in (("accumulate", 0), ("sort", 2), ("reduce", 3))}
        # e2e: host (pinned) scalars in, affine point out, every step (warmed up like the resident region: the chunked
        # entry point allocates its staging buffer and second sort workspace on first use)
        for _ in range(Wm):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_e2e()
        barrier()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
        clocks = sampler.stop()                                  # sampled over the timed, per-kernel and e2e regions
        acc = prof["accumulate"]                                 # the bucket-accumulate kernel alone, one launch per MSM
        acc_ms = acc["ms"] / max(acc["launches"], 1)
        madds = acc["work"] / max(acc["launches"], 1)            # mixed additions per launch (n * windows upper bound)
        pts_per_s = world * n * K / (ms * 1e-3)
        res = {
            "value": pts_per_s, "ms": ms, "launches": launches, "clocks": clocks,
            "e2e": {"value": world * n * K / (ms_e2e * 1e-3), "unit": "points/s",
                    "h2d_bytes_per_step": n * 32 * world, "d2h_bytes_per_step": (64 + 4) * world,
                    "ms_per_step": ms_e2e / K, "host_buffers": "pinned (cudaHostAlloc)"},
            "roofline": {
                "bound": "imad", "kernel": "msm_accumulate_kernel",
                "achieved": (n / (acc_ms * 1e-3)) * MODEL_IMAD_PER_POINT / 1e9,
                "peak": imad32_peak / 1e9, "unit": "G IMAD32/s",
                "frac": (n / (acc_ms * 1e-3)) * MODEL_IMAD_PER_POINT / imad32_peak,
                "traffic": load_traffic("msm_accumulate_kernel") if args.logn == 24 else None,
                "model": "SURVEY 8(d): 43,520 32-bit IMAD per point (16 windows x 10 modmul x 272); "
                         "peak = live IMAD.WIDE.U32 microbenchmark x 2",
                "executed_frac": madds * IMAD32_EXECUTED_PER_MADD / (acc_ms * 1e-3) / imad32_peak,
                "kernel_ms": acc_ms, "kernel_share_of_step": acc["ms"] / max(sum(p["ms"] for p in prof.values()), 1e-9),
                "hbm_algorithmic_gbs": n * 96 / (acc_ms * 1e-3) / 1e9,
                "hbm_peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src,
            },
            "profile_ms_per_step": {k: v["ms"] / K for k, v in prof.items()},
        }
        srs.destroy(); dsc.free(); pinned.free()
        return res

    # ------------------------------------------------------------------ NTT
    def bench_ntt():
        w = pow(5, (R_BN254 - 1) // n_total, R_BN254)
        wl = ints_to_limbs([w], R_BN254)[0]
        pinned = _ffi.PinnedArray((n_total, 4))
        pinned.array[:] = random_scalars(n_total, R_BN254, seed=args.logn + 7 * rank)
        d = _ffi.DeviceBuffer(n_total * 32).upload(pinned.array)
        for _ in range(Wm):
            device.ntt_dev("bn254", d, n_total, wl)
        sampler = ClockSampler(local)
        barrier()
        l0 = _ffi.launch_count()
        _ffi.timer_start()                                      # events on the library's launching stream, any world size
        for _ in range(K):
            device.ntt_dev("bn254", d, n_total, wl)
        ms = _ffi.timer_stop()
        barrier()
        launches = _ffi.launch_count() - l0
        ms = max_over_ranks(ms)
        _ffi.profile_reset(); _ffi.profile_enable(True)
        for _ in range(K):
            device.ntt_dev("bn254", d, n_total, wl)
        _ffi.profile_enable(False)
        pr = _ffi.profile_get(1)
        for _ in range(Wm):                                     # warm-up of the host-buffer entry point (staging buffer, events)
            device.ntt("bn254", pinned.array, wl)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            device.ntt("bn254", pinned.array, wl)            # H2D + kernels + D2H inside the call
        barrier()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
        clocks = sampler.stop()
        ntt_ms = pr["ms"] / K                                   # all passes of one transform
        hbm = 64.0 * n_total / (ntt_ms * 1e-3) / 1e9
        modmuls = 0.5 * n_total * args.logn
        res = {
            "value": world * n_total * K / (ms * 1e-3), "ms": ms, "launches": launches, "clocks": clocks,
            "e2e": {"value": world * n_total * K / (ms_e2e * 1e-3), "unit": "elements/s",
                    "h2d_bytes_per_step": n_total * 32 * world, "d2h_bytes_per_step": n_total * 32 * world,
                    "ms_per_step": ms_e2e / K, "host_buffers": "pinned (cudaHostAlloc)"},
            "roofline": {
                "bound": "hbm", "kernel": "ntt_pass_kernel_c<FrBN254, 8, lazy> (all passes of one transform)",
                "achieved": hbm, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": hbm / peaks.get("hbm_gbs"),
                "traffic": load_traffic("ntt_pass_kernel") if args.logn == 24 else None, "peak_source": peak_src,
                "model": "SURVEY 8(d): algorithmic bytes = 2 x 32 B x n (twiddles not counted)",
                "passes": pr["launches"] // K, "kernel_ms": ntt_ms,
                "imad_frac": modmuls * IMAD32_PER_MODMUL / (ntt_ms * 1e-3) / imad32_peak,
                "imad_model": "SURVEY 8(d): (n/2) log2 n modmul x 272 IMAD32; the binding bound for 256-bit fields",
            },
        }
        d.free(); pinned.free()
        return res

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
        t_gen = time.perf_counter() - t0
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
        rows = 1 << args.marlin_rows_logn
        t0 = time.perf_counter()
        sA, sB, sC, sx, sw = marlin.synthetic_r1cs(rows, 8, R_BN254, seed=args.marlin_rows_logn)
        swl = ints_to_limbs(sw, R_BN254)
        sdraws = random_scalars(8 + 2 * rows + 1, R_BN254, seed=4)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        mK = 1 << (2 * rows - 1).bit_length()
        sipk, _ = midx.preprocess(sA, sB, sC, max_degree=6 * mK, tau=TAU)
        _ffi.check(_ffi._lib.kzgpu_sync())
        t_mindex = time.perf_counter() - t0
        sxs = [midx.kzg.Fq(v) for v in sx]
        mt = []
        for i in range(3 + 5):
            t0 = time.perf_counter()
            mpr.prove(sipk, sxs, swl, draws=sdraws)
            mt.append(time.perf_counter() - t0)
        assert set(mpr.checks.values()) == {0}, "a Marlin linearisation identity failed"
        mt = sorted(mt[3:])
        out["marlin_synthetic"] = {"rows": rows, "H": sipk["subgroups"]["n"], "K": sipk["subgroups"]["m"], "prove_s": mt[len(mt) // 2],
                                   "index_s": t_mindex, "instance_generation_s": t_gen,
                                   "rounds_s": {k: round(v, 5) for k, v in mpr.timings.items()},
                                   "checks": "f_1(beta_1) = f_2(beta_1) = f_3(beta_2) = 0 asserted (the verifier's polynomial identities)"}
        out["synthetic"] = {"gates": n, "prove_s": times[len(times) // 2], "index_s": t_index, "circuit_generation_s": t_gen,
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

    if args.workload == "marlin":
        res = bench_marlin()
        if rank == 0:
            print(json.dumps({"metric": "marlin_kernel_workload_s", "value": res["ms"] / K / 1e3, "unit": "s", "n_gpus": world,
                              "steps": K, "warmup": Wm, "ms_per_step": res["ms"] / K, "higher_is_better": False,
                              "scaling": "strong", "vs_baseline": None,
                              "dtype": "u32x8 (256-bit modular integers, Montgomery)", "data": "synthetic",
                              "config": {"workload": f"commit/open/NTT calls of one Marlin proof, 2^{args.marlin_logn} constraints "
                                                     f"(|H|=2^{args.marlin_logn}, |K|=2^{args.marlin_logn + 1}) on {world} GPUs: MSMs of >= 2^22 points point-sharded over all ranks, "
                                                     "the rest assigned longest-first, SRS replicated", "items": res["items"], "srs_points": res["srs_points"],
                                         "srs_layout": res["srs"]},
                              "clocks": res["clocks"], "gpu_launches": res["launches"], "device": info["name"]}))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    primary = bench_msm() if args.workload == "msm" else (bench_ntt() if args.workload == "ntt" else None)
    secondary = None
    if args.workload == "msm" and not args.no_secondary:
        secondary = bench_ntt()
    plonk = None
    if world == 1 and (args.workload == "plonk" or not args.no_secondary):
        plonk = bench_plonk()
    if args.workload == "plonk":
        if rank == 0:
            print(json.dumps({"metric": "plonk_prove_s", "value": plonk["synthetic"]["prove_s"], "unit": "s", "n_gpus": 1,
                              "steps": 5, "warmup": 3, "ms_per_step": 1e3 * plonk["synthetic"]["prove_s"],
                              "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
                              "dtype": "u32x8 (256-bit modular integers, Montgomery)", "data": "synthetic",
                              "config": {"workload": f"PLONK prove, synthetic circuit of 2^{args.plonk_logn} gates, BN254"},
                              "plonk": plonk, "device": info["name"]}))
        return 0

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if args.workload == "msm":
            v, wall = cpu_msm_rate(4096, 1)
            cpu = {"value": v, "unit": "points/s", "cores": 1, "kind": "port",
                   "sample": f"4096-point commit (kzg.py:112-116 loop on the restated py_ecc arithmetic), {wall:.1f} s",
                   "host_cpus": os.cpu_count()}
            if secondary is not None:
                v2, wall2 = cpu_ntt_rate(17, 1)
                secondary["cpu_baseline"] = {"value": v2, "unit": "elements/s", "cores": 1, "kind": "port",
                                             "sample": f"2^17-element recursive fft_ff (fft_ff.py:3-37 on CPython ints), {wall2:.1f} s"}
        else:
            v, wall = cpu_ntt_rate(18, 1)
            cpu = {"value": v, "unit": "elements/s", "cores": 1, "kind": "port",
                   "sample": f"2^18-element recursive fft_ff (fft_ff.py:3-37 on CPython ints), {wall:.1f} s",
                   "host_cpus": os.cpu_count()}

    cpu_plonk = None
    if rank == 0 and world == 1 and not args.no_cpu and plonk is not None:
        cpu_plonk = cpu_plonk_hotpath()

    if rank == 0:
        if args.workload == "msm":
            metric, unit = "g1_msm_points_per_s", "points/s"
            workload = (f"BN254 G1 MSM (KZG commit) of 2^{args.logn} points" +
                        (f" per GPU, SRS/scalars sharded by index range over {world} GPUs, NCCL all-gather of XYZZ partials"
                         if world > 1 and args.scaling == "weak" else
                         (f" split over {world} GPUs" if world > 1 else "")))
            footprint = "inputs 1.5 GiB/GPU (SRS 1 GiB + scalars 512 MiB at 2^24) exceed the 126 MB L2"
        else:
            metric, unit = "ntt_elements_per_s", "elements/s"
            workload = f"BN254 scalar-field NTT of 2^{args.logn} elements" + (f", one vector per GPU ({world} replicas)" if world > 1 else "")
            footprint = "512 MiB vector at 2^24 exceeds the 126 MB L2"
        line = {
            "metric": metric, "value": primary["value"], "unit": unit, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": primary["ms"] / K, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": "u32x8 (256-bit modular integers, Montgomery)", "data": "synthetic",
            "config": {"workload": workload, "curve": "bn254", "logn": args.logn, "l2": footprint,
                       "srs": "tau^i*G1 generated on device from a fixed tau", "scalars": "uniform in [0,r), numpy PCG64"},
            "clocks": primary["clocks"], "e2e": primary["e2e"], "gpu_launches": primary["launches"],
            "roofline": primary["roofline"], "cpu_baseline": cpu,
            "imad_wide_peak_per_s": imad_wide_peak, "device": info["name"],
        }
        if "profile_ms_per_step" in primary:
            line["profile_ms_per_step"] = primary["profile_ms_per_step"]
        if plonk is not None:
            if cpu_plonk is not None:
                plonk["bundled"]["cpu_port_hotpath_s"] = cpu_plonk
                plonk["marlin_bundled"]["cpu_port_hotpath_s"] = cpu_hotpath("marlin")
            line["plonk"] = plonk
        if secondary is not None:
            line["ntt"] = {"metric": "ntt_elements_per_s", "value": secondary["value"], "unit": "elements/s",
                           "ms_per_step": secondary["ms"] / K, "e2e": secondary["e2e"], "roofline": secondary["roofline"],
                           "gpu_launches": secondary["launches"], "clocks": secondary["clocks"],
                           "cpu_baseline": secondary.get("cpu_baseline"),
                           "config": {"workload": f"BN254 scalar-field NTT of 2^{args.logn} elements, natural order in/out"}}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="msm", choices=["msm", "ntt", "plonk", "marlin"])
    ap.add_argument("--marlin-logn", type=int, default=20, help="constraints of the Marlin kernel workload (log2)")
    ap.add_argument("--marlin-rows-logn", type=int, default=16, help="rows of the synthetic R1CS for the device Marlin prover (log2)")
    ap.add_argument("--plonk-logn", type=int, default=20, help="gates of the synthetic PLONK circuit (log2)")
    ap.add_argument("--logn", type=int, default=24)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the NTT half of the metric")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
