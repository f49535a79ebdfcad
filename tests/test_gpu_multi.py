"""GPU (>= 2 devices): multi-GPU INSIDE the library, from a plain python process -- kzgpu_init_multi, point-sharded MSMs,
batched commits placed longest-first, batched NTTs per device.  The same calls
on one device must give the same points and vectors bit for bit."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    try:
        return int(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU "))
    except Exception:
        return 0


def _run(env_extra, scale):
    env = {k: v for k, v in os.environ.items() if k not in ("KZGPU_DEVICES", "LOCAL_RANK", "RANK", "WORLD_SIZE")}
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(HERE, "multi_gpu_worker.py"), str(scale)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("scale,shard_min", [(12, 1 << 14), (17, 1 << 20)])
def test_marlin_sized_commits_on_all_gpus_match_one_gpu(scale, shard_min):
    """11 Marlin-sized polynomials through the batched commit (kzg.py:102), a long MSM, a batched open and 9 batched NTTs: all
    visible GPUs against one GPU.  The small case lowers KZGPU_SHARD_MIN so that sharding, placement and replicas are all
    exercised at 2^12; the second is the default policy at 2^17 .. 12 * 2^17 coefficients."""
    ng = _gpus()
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    one = _run({"KZGPU_SHARD_MIN": str(shard_min)}, scale)
    many = _run({"KZGPU_DEVICES": "all", "KZGPU_SHARD_MIN": str(shard_min)}, scale)
    assert one["ndev"] == 1 and many["ndev"] == ng
    for k in ("commits", "infs", "msm", "msm_offset", "msm_dev", "open", "ntt", "ntt_roundtrip", "marlin"):
        assert one[k] == many[k], k
    if scale <= 12:                                   # KZG.commit / KZG.open with Python objects, all GPUs against one
        assert one["kzg_commit"] == many["kzg_commit"] and len(many["kzg_commit"]) == 11 and one["kzg_open"] == many["kzg_open"]
    assert many["ntt_roundtrip"] is True and many["launches"] > one["launches"] // 2
