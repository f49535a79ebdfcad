"""ctypes binding of libkzgpu.so (include/kzgpu.h) -- numpy + ctypes only, no torch.

This is the only way the package reaches the GPU.  There is no CPU fallback: if the
library has not been built, or no sm_100 device answers, every entry point raises
`KzgpuError` (a RuntimeError).
"""

import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkzgpu.so")

BN254 = 0
BLS12_381 = 1
CURVE_IDS = {"bn254": BN254, "bls12_381": BLS12_381}

E_INVAL, E_CUDA, E_NOTINIT, E_RANGE, E_HANDLE = -1, -2, -3, -4, -5


class KzgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"kzgpu error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None
_inited = False

_u64p = ctypes.POINTER(ctypes.c_uint64)
_szp = ctypes.POINTER(ctypes.c_size_t)
_intp = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); must list every symbol include/kzgpu.h declares
# (tests/test_abi.py cross-checks this table against the header).
SIGNATURES = {
    "kzgpu_init": (ctypes.c_int, [ctypes.c_int]),
    "kzgpu_init_multi": (ctypes.c_int, [ctypes.c_int, _intp]),
    "kzgpu_device_count": (ctypes.c_int, [_intp]),
    "kzgpu_shutdown": (ctypes.c_int, []),
    "kzgpu_last_error": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t]),
    "kzgpu_device_info": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, _intp, _szp]),
    "kzgpu_fp_limbs64": (ctypes.c_int, [ctypes.c_int]),
    "kzgpu_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]),
    "kzgpu_free": (ctypes.c_int, [ctypes.c_void_p]),
    "kzgpu_h2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "kzgpu_d2h": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "kzgpu_d2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "kzgpu_memset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]),
    "kzgpu_sync": (ctypes.c_int, []),
    "kzgpu_host_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]),
    "kzgpu_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "kzgpu_set_stream": (ctypes.c_int, [ctypes.c_void_p]),
    "kzgpu_timer_start": (ctypes.c_int, []),
    "kzgpu_timer_stop": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float)]),
    "kzgpu_srs_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, _u64p]),
    "kzgpu_srs_generate": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, _u64p]),
    "kzgpu_srs_generate_range": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, _u64p]),
    "kzgpu_srs_destroy": (ctypes.c_int, [ctypes.c_uint64]),
    "kzgpu_srs_size": (ctypes.c_int, [ctypes.c_uint64, _szp]),
    "kzgpu_srs_info": (ctypes.c_int, [ctypes.c_uint64, _intp, _intp, _szp]),
    "kzgpu_srs_read": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "kzgpu_msm": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, _intp]),
    "kzgpu_msm_dev": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, _intp]),
    "kzgpu_msm_batch": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_void_p, _szp, ctypes.c_size_t, ctypes.c_void_p, _intp]),
    "kzgpu_msm_batch_dev": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, _intp]),
    "kzgpu_msm_partial_dev": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "kzgpu_msm_partial": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "kzgpu_g1_fold": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, _intp]),
    "kzgpu_g1_lincomb": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, _intp]),
    "kzgpu_ntt": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kzgpu_ntt_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kzgpu_ntt_batch": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kzgpu_ntt_batch_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kzgpu_open": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_void_p, _szp, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _intp, ctypes.c_void_p]),
    "kzgpu_open_quotient": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, _szp, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _szp, ctypes.c_void_p]),
    "kzgpu_open_dev": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_void_p, _szp, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _intp, ctypes.c_void_p]),
    "kzgpu_open_quotient_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, _szp, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _szp, ctypes.c_void_p]),
    "kzgpu_poly_eval_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "kzgpu_poly_lincomb_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, _szp, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "kzgpu_powers_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "kzgpu_plonk_permutation_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t] + [ctypes.c_void_p] * 10 + [_intp]),
    "kzgpu_plonk_quotient_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "kzgpu_poly_mul_pointwise_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "kzgpu_spmv_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t] + [ctypes.c_void_p] * 5),
    "kzgpu_marlin_h2_evals_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t] + [ctypes.c_void_p] * 6),
    "kzgpu_marlin_f2_evals_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t] + [ctypes.c_void_p] * 8),
    "kzgpu_marlin_t_evals_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t] + [ctypes.c_void_p] * 8),
    "kzgpu_field_op": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "kzgpu_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "kzgpu_profile_reset": (ctypes.c_int, []),
    "kzgpu_profile_get": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), _u64p, ctypes.POINTER(ctypes.c_double)]),
    "kzgpu_launch_count": (ctypes.c_int, [_u64p]),
}


def load_library():
    """dlopen the in-tree library and bind every declared symbol.  No device is touched."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KzgpuError(E_NOTINIT, f"{LIB_PATH} not built; run `python -m kzg_snark_b200.build` "
                                    "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    buf = ctypes.create_string_buffer(512)
    load_library().kzgpu_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc):
    if rc != 0:
        raise KzgpuError(rc, last_error())


def init(device=None):
    """Initialise the process-wide GPU context (a singleton, not per-KZG).

    One device by default (`device`, else $KZGPU_DEVICE, else $LOCAL_RANK, else 0).  $KZGPU_DEVICES = "all" or a comma
    separated list ("0,1,2,3") initialises the library on several devices of this ONE process (kzgpu_init_multi): large
    MSMs are then point-sharded and batched commits / NTTs spread over them inside the library, no launcher involved."""
    global _inited
    lib = load_library()
    if _inited:
        return lib
    multi = os.environ.get("KZGPU_DEVICES") if device is None else None
    if multi and "LOCAL_RANK" not in os.environ:
        return init_multi(None if multi.strip().lower() == "all" else [int(x) for x in multi.split(",")])
    if device is None:
        device = int(os.environ.get("KZGPU_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    check(lib.kzgpu_init(device))
    _inited = True
    return lib


def init_multi(devices=None):
    """Initialise on several devices of this process: `devices` = list of CUDA ordinals (first = primary), None = all."""
    global _inited
    lib = load_library()
    if _inited:
        return lib
    if devices is None:
        check(lib.kzgpu_init_multi(0, None))
    else:
        arr = (ctypes.c_int * len(devices))(*devices)
        check(lib.kzgpu_init_multi(len(devices), arr))
    _inited = True
    return lib


def device_count():
    n = ctypes.c_int(0)
    load_library().kzgpu_device_count(ctypes.byref(n))
    return n.value


def shutdown():
    global _inited
    if _lib is not None and _inited:
        _lib.kzgpu_shutdown()
    _inited = False


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def device_info():
    lib = init()
    name = ctypes.create_string_buffer(128)
    sm = ctypes.c_int(0)
    mem = ctypes.c_size_t(0)
    check(lib.kzgpu_device_info(name, 128, ctypes.byref(sm), ctypes.byref(mem)))
    return {"name": name.value.decode(), "sm_count": sm.value, "total_mem": mem.value}


def launch_count():
    c = ctypes.c_uint64(0)
    load_library().kzgpu_launch_count(ctypes.byref(c))
    return c.value


class DeviceBuffer:
    """RAII wrapper over kzgpu_alloc / kzgpu_free."""

    def __init__(self, nbytes):
        lib = init()
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p(0)
        check(lib.kzgpu_alloc(ctypes.byref(p), self.nbytes))
        self.ptr = p

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(_lib.kzgpu_h2d(self.ptr, ptr(arr), arr.nbytes))
        return self

    def download(self, arr):
        assert arr.flags["C_CONTIGUOUS"] and arr.nbytes <= self.nbytes
        check(_lib.kzgpu_d2h(ptr(arr), self.ptr, arr.nbytes))
        return arr

    def free(self):
        if self.ptr is not None and _lib is not None and _inited:
            _lib.kzgpu_free(self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def timer_start():
    check(_lib.kzgpu_timer_start())


def timer_stop():
    ms = ctypes.c_float(0)
    check(_lib.kzgpu_timer_stop(ctypes.byref(ms)))
    return ms.value


BENCH_LIB_PATH = os.path.join(_HERE, "libkzgpu_bench.so")
_bench_lib = None


def microbench(kind, blocks, threads, iters, device=None):
    """Throughput microbenchmark from the separate measurement library (include/kzgpu_bench.h; not part of the product ABI)."""
    global _bench_lib
    if _bench_lib is None:
        if not os.path.exists(BENCH_LIB_PATH):
            raise KzgpuError(E_NOTINIT, f"{BENCH_LIB_PATH} not built; run `python -m kzg_snark_b200.build`")
        _bench_lib = ctypes.CDLL(BENCH_LIB_PATH)
        _bench_lib.kzgpu_microbench.restype = ctypes.c_int
        _bench_lib.kzgpu_microbench.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)]
    if device is None:
        device = int(os.environ.get("KZGPU_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ms = ctypes.c_float(0)
    ops = ctypes.c_double(0)
    rc = _bench_lib.kzgpu_microbench(device, kind, blocks, threads, iters, ctypes.byref(ms), ctypes.byref(ops))
    if rc:
        raise KzgpuError(rc, "kzgpu_microbench failed")
    return ms.value, ops.value


def set_stream(cuda_stream):
    """Run the library on the caller's CUDA stream (int handle, e.g. torch's cuda_stream); 0/None = own."""
    check(init().kzgpu_set_stream(ctypes.c_void_p(cuda_stream or 0)))


def profile_enable(on=True):
    check(init().kzgpu_profile_enable(1 if on else 0))


def profile_reset():
    check(init().kzgpu_profile_reset())


def profile_get(which):
    ms = ctypes.c_double(0)
    n = ctypes.c_uint64(0)
    w = ctypes.c_double(0)
    check(init().kzgpu_profile_get(which, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(w)))
    return {"ms": ms.value, "launches": n.value, "work": w.value}


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (for the e2e leg of bench.py)."""

    def __init__(self, shape, dtype=np.uint64):
        lib = init()
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = ctypes.c_void_p(0)
        check(lib.kzgpu_host_alloc(ctypes.byref(p), self.nbytes))
        self.ptr = p
        buf = (ctypes.c_uint8 * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr is not None and _lib is not None and _inited:
            self.array = None
            _lib.kzgpu_host_free(self.ptr)
        self.ptr = None
