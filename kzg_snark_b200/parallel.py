"""Multi-GPU layout, harness flavour: one process per GPU, torch.distributed for the plumbing only.

(The PRODUCT path needs none of this: `KZGPU_DEVICES=all` / `_ffi.init_multi([...])` initialises libkzgpu.so on several devices
of one plain process and the library shards MSMs, places batched commits and deals batched NTTs itself -- `kzgpu_init_multi`,
DESIGN.md section 6.  This module serves `bench.py --gpus N` under torchrun, which is how the driver launches the benchmark.)

The reference is single-process and has no parallelism (SURVEY.md sections 0, 8e); this layout
is new.  The path shards in exactly two ways:

  * one large MSM (KZG.commit of a long polynomial, kzg.py:108-116): the SRS points and the
    scalars are split into contiguous index ranges, one per rank; each rank reduces its range to
    ONE un-normalised XYZZ partial sum (`kzgpu_msm_partial_dev`), the partials (128 B for BN254,
    192 B for BLS12-381) are all-gathered over NCCL and every rank folds them
    (`kzgpu_g1_fold`).  That all-gather is the only exchange step of the path; it is
    latency-bound, NVLink bandwidth is irrelevant at this payload.
  * batched work (the k polynomials of one commit() call, kzg.py:102; independent NTT
    vectors): item j goes to rank j % world, no data-path collective at all; results are
    collected with an all-gather of the 64-byte points.

torch is imported lazily and only here (and in bench.py for N > 1): the single-GPU product path
is ctypes + numpy.
"""

import ctypes

import numpy as np

XYZZ_BYTES = {0: 128, 1: 192}          # 4 * fp_limbs32 * 4 bytes


def shard_range(n_total, world, rank):
    """Contiguous balanced split of [0, n_total): returns (start, count) for `rank`."""
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def round_robin(n_items, world, rank):
    """Indices of the batch items (polynomials / vectors) that `rank` owns."""
    return list(range(rank, n_items, world))


def lpt_assign(costs, world):
    """Longest-processing-time-first assignment of independent batch items to ranks: items sorted by
    cost (descending, ties by index) go to the currently least-loaded rank.  Returns owner[i].
    Used where the items of a batch differ widely in size (the commits of one Marlin proof range
    from 2^20 to 12 * 2^20 coefficients); plain round-robin is the special case of equal costs."""
    owner = [0] * len(costs)
    load = [0.0] * world
    for i in sorted(range(len(costs)), key=lambda j: (-costs[j], j)):
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += costs[i]
    return owner


class _Raw:
    """Adapter: a raw device address as the `.ptr` the device.* helpers expect."""

    def __init__(self, address):
        self.ptr = ctypes.c_void_p(address)


def all_gather_bytes(local, world, dist, device="cpu"):
    """All-gather a fixed-size uint8 vector from every rank; returns a (world, len) tensor.
    Works on gloo (CPU tensors) and NCCL (CUDA tensors)."""
    import torch
    out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out.view(world, local.numel())


def share_stream_with_torch():
    """Put torch (and with it NCCL's completion wait) and the library on ONE side stream, so that library kernels that read
    a collective's output are ordered after it.  torch's default stream will not do: its handle is 0, which
    kzgpu_set_stream reads as "back to the library's own non-blocking stream" -- and that stream is not ordered against
    the legacy default stream.  Call once per process, after torch.cuda.set_device."""
    import torch
    from . import _ffi
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    _ffi.set_stream(side.cuda_stream)
    return side


def sharded_msm_step(srs_shard, d_scalars, n_local, curve_id, dist, partial, gathered):
    """One point-sharded MSM on the GPU path: partial (device) -> NCCL all-gather -> fold.
    `partial` / `gathered` are pre-allocated CUDA uint8 tensors (XYZZ_BYTES, world*XYZZ_BYTES).
    Returns (affine limbs, is_inf) on every rank.  Requires share_stream_with_torch() (the fold must not run ahead of
    the all-gather)."""
    from . import device
    device.msm_partial_dev(srs_shard, d_scalars, n_local, _Raw(partial.data_ptr()))
    dist.all_gather_into_tensor(gathered, partial)
    world = gathered.numel() // partial.numel()
    return device.g1_fold(curve_id, _Raw(gathered.data_ptr()), world)


def sharded_reduce(local_value, world, rank, dist, encode, decode, fold, nbytes):
    """Backend-agnostic skeleton of the same exchange, used by the CPU (gloo) tests:
    `encode(local_value) -> bytes[nbytes]`, all-gather, `fold([decode(b) for b in rows])`."""
    import torch
    buf = torch.frombuffer(bytearray(encode(local_value)), dtype=torch.uint8).clone()
    assert buf.numel() == nbytes
    rows = all_gather_bytes(buf, world, dist)
    return fold([decode(bytes(rows[r].tolist())) for r in range(world)])


def gather_batch_results(local_results, n_items, world, rank, dist, item_bytes):
    """Collect the per-item results of a round-robin batch (item j computed on rank j % world)
    into item order on every rank.  local_results: {item index: bytes[item_bytes]}."""
    import torch
    per_rank = (n_items + world - 1) // world
    buf = torch.zeros(per_rank * item_bytes, dtype=torch.uint8)
    for slot, j in enumerate(round_robin(n_items, world, rank)):
        buf[slot * item_bytes:(slot + 1) * item_bytes] = torch.frombuffer(bytearray(local_results[j]), dtype=torch.uint8)
    rows = all_gather_bytes(buf, world, dist)
    out = [None] * n_items
    for r in range(world):
        for slot, j in enumerate(round_robin(n_items, world, r)):
            out[j] = bytes(rows[r][slot * item_bytes:(slot + 1) * item_bytes].tolist())
    return out
