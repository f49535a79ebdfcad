#!/usr/bin/env python3
"""Batched-affine feasibility probe: additions/s of kzgpu_microbench kinds 12 (consecutive
operands) and 13 (operands gathered from an 8 GiB table) next to the XYZZ mixed addition (kind 3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kzg_snark_b200 import _ffi                               # noqa: E402

sm = _ffi.device_info()["sm_count"]
ms, ops = _ffi.microbench(3, sm * 4, 128, 2000)
print(f"XYZZ mixed addition (registers only)      {ops / ms / 1e6:8.2f} G adds/s")
for per_sm in (4, 6, 8):
    for iters in (165, 330, 660):
        for kind, name in ((12, "consecutive"), (13, "gathered   ")):
            ms, ops = _ffi.microbench(kind, sm * per_sm, 128, iters)
            print(f"batched affine, {name}, {per_sm} blocks/SM x 128 thr, {iters:4d} pairs/thread ({ops / 1e6:6.1f} M pairs): "
                  f"{ms:8.3f} ms  {ops / ms / 1e6:8.2f} G adds/s", flush=True)
