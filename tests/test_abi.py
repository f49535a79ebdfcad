"""CPU: the C-ABI library loads and exports every symbol include/kzgpu.h declares; the ctypes
table matches the header; the product fails loudly without a GPU (no compute calls here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kzgpu.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(kzgpu_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from kzg_snark_b200 import build
    build.build()
    from kzg_snark_b200 import _ffi
    return _ffi.load_library()


def test_header_lists_the_expected_surface():
    fns = header_functions()
    for need in ("kzgpu_init", "kzgpu_srs_create", "kzgpu_srs_generate", "kzgpu_msm", "kzgpu_msm_batch",
                 "kzgpu_ntt", "kzgpu_ntt_batch", "kzgpu_open", "kzgpu_last_error"):
        assert need in fns


def test_library_exports_every_declared_symbol(lib):
    from kzg_snark_b200 import _ffi
    fns = header_functions()
    assert sorted(_ffi.SIGNATURES) == fns, "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(kzgpu_\w+)", out))
    assert set(fns) <= exported
    for f in fns:
        assert hasattr(lib, f)


def test_microbenchmarks_live_in_their_own_library(lib):
    """Measurement tooling is not part of the product ABI: kzgpu_microbench is exported by libkzgpu_bench.so
    (include/kzgpu_bench.h) and by nothing else."""
    from kzg_snark_b200 import _ffi
    assert os.path.exists(_ffi.BENCH_LIB_PATH)
    nm = lambda p: set(re.findall(r"\sT\s+(kzgpu_\w+)", subprocess.run(["nm", "-D", "--defined-only", p], capture_output=True, text=True).stdout))   # noqa: E731
    assert nm(_ffi.BENCH_LIB_PATH) == {"kzgpu_microbench"}
    assert "kzgpu_microbench" not in nm(_ffi.LIB_PATH)
    assert "kzgpu_microbench" in open(os.path.join(ROOT, "include", "kzgpu_bench.h")).read()


def test_library_is_sm100a_only():
    from kzg_snark_b200 import _ffi
    out = subprocess.run(["cuobjdump", "-lelf", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_calls_fail_loudly_without_init_or_device(lib):
    import ctypes
    from kzg_snark_b200 import _ffi
    if _ffi._inited:
        pytest.skip("a GPU context is live in this process")
    # not initialised -> ENOTINIT from every entry point that needs the device
    h = ctypes.c_uint64(0)
    assert lib.kzgpu_srs_create(0, None, 0, ctypes.byref(h)) == _ffi.E_NOTINIT
    assert lib.kzgpu_ntt(0, None, 4, None, 0, None) == _ffi.E_NOTINIT
    assert "kzgpu_init" in _ffi.last_error()
    assert lib.kzgpu_fp_limbs64(0) == 4 and lib.kzgpu_fp_limbs64(1) == 6 and lib.kzgpu_fp_limbs64(7) < 0
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(_ffi.KzgpuError):
            _ffi.init(0)                      # no device: error, never a CPU fallback
        from kzg_snark_b200.kzg import KZG
        k = KZG("bn254")
        with pytest.raises(_ffi.KzgpuError):
            k.setup(4, tau=5)
        from kzg_snark_b200.fft_ff import fft_ff
        with pytest.raises(_ffi.KzgpuError):
            fft_ff([k.Fq(1), k.Fq(2)], k.Fq(k.curve_order - 1), k.Fq)


def test_product_never_imports_the_oracle():
    """The package must not route through oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, "kzg_snark_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
