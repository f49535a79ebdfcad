"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (index-range sharding,
the all-gather of fixed-size partials, the fold, round-robin batches).  The per-rank compute is
supplied by the oracle here -- on the GPU box the same skeleton carries the CUDA partial sums
(bench.py --gpus N, kzg_snark_b200/parallel.py)."""
import os
import random
import socket

import pytest

from oracle.curve import get_curve
from oracle.kzg import KZGOracle, poly_eval

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp          # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    from kzg_snark_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cv = get_curve("bn254")
        ko = KZGOracle("bn254")
        rng = random.Random(99)                      # same seed on both ranks: same global inputs
        tau = rng.randrange(1, cv.r)
        scalars = [rng.randrange(cv.r) for _ in range(n_total)]
        start, count = parallel.shard_range(n_total, world, rank)
        # this rank's SRS shard [tau^(start+i) G] and scalar slice
        shard_pts = [cv.multiply(cv.G1, pow(tau, start + i, cv.r)) for i in range(count)]
        partial = ko.commit(shard_pts, [scalars[start:start + count]])[0] if count else cv.Z1

        def encode(pt):                              # projective ints, 3 x 32 bytes (the CUDA path sends XYZZ limbs)
            return b"".join(int(c).to_bytes(32, "little") for c in pt)

        def decode(b):
            return tuple(int.from_bytes(b[i:i + 32], "little") for i in (0, 32, 64))

        def fold(parts):
            acc = cv.Z1
            for p in parts:
                acc = cv.add(acc, p)
            return acc

        total = parallel.sharded_reduce(partial, world, rank, dist, encode, decode, fold, 96)
        exp = cv.multiply(cv.G1, poly_eval(scalars, tau, cv.r))
        ok_msm = cv.normalize(total) == cv.normalize(exp)
        # round-robin batch: 5 "polynomials", item j on rank j % world
        k = 5
        mine = {j: (j * 1000 + rank).to_bytes(8, "little") for j in parallel.round_robin(k, world, rank)}
        got = parallel.gather_batch_results(mine, k, world, rank, dist, 8)
        ok_batch = [int.from_bytes(b, "little") for b in got] == [j * 1000 + (j % world) for j in range(k)]
        q.put((rank, ok_msm, ok_batch))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from kzg_snark_b200.parallel import shard_range, round_robin
    for n in (0, 1, 7, 16, 1 << 24, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert sorted(sum((round_robin(11, 4, r) for r in range(4)), [])) == list(range(11))


def test_lpt_assignment_balances_unequal_items():
    from kzg_snark_b200.parallel import lpt_assign
    costs = [1, 1, 1, 1, 1, 2, 1, 1, 1, 2, 12]
    owner = lpt_assign(costs, 4)
    load = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(4)]
    assert sorted(set(owner)) == [0, 1, 2, 3] and max(load) == 12 and sum(load) == sum(costs)
    assert lpt_assign([3, 3, 3, 3], 4) == [0, 1, 2, 3] and lpt_assign([5], 8) == [0] and lpt_assign([], 2) == []
    assert lpt_assign(costs, 1) == [0] * len(costs)


@pytest.mark.parametrize("n_total", [9, 32])
def test_point_sharded_commit_over_gloo(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_strong_scaling_check_formula():
    """bench.py's correctness check of a point-sharded MSM at N ranks: rank k evaluates its slice p_k at tau on the device and the
    ranks' values are combined as sum_k p_k(tau) * tau^(start_k) = p(tau), the scalar of the tau-identity commit == p(tau) * G1
    (kzg.py:108).  Checked here on plain integers for every split bench.py uses."""
    from kzg_snark_b200.parallel import shard_range
    cv = get_curve("bn254")
    rng = random.Random(5)
    n = 1 << 10
    tau = rng.randrange(1, cv.r)
    p = [rng.randrange(cv.r) for _ in range(n)]
    whole = poly_eval(p, tau, cv.r)
    for world in (1, 2, 4, 8):
        parts = []
        for rank in range(world):
            start, cnt = shard_range(n, world, rank)
            assert (start, cnt) == (rank * (n // world), n // world)          # bench.py: start = rank * n, n = n_total // world
            parts.append(poly_eval(p[start:start + cnt], tau, cv.r) * pow(tau, start, cv.r) % cv.r)
        assert sum(parts) % cv.r == whole
