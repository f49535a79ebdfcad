"""The measurement contract of bench.py, checked without a GPU: the reference arm (the oracle port on the host cores, the
one leg of bench.py that may execute oracle/) prints a well-formed line, and the committed GPU line of the round
(profiles/*_bench.json, written by `python bench.py` on a B200) carries every key the contract names."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and BASE_KEYS <= set(line)
    assert line["metric"] == "g1_msm_points_per_s" and line["unit"] == "points/s" and line["value"] > 0
    # "reference" = the reference's own kzg.py (from /root/reference or its staged copy baseline/_ref), "port" = the oracle
    from oracle import refrun
    assert line["cpu_baseline"]["kind"] == ("reference" if refrun.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_gpu_line_carries_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2*_bench.json"))) or sorted(glob.glob(os.path.join(ROOT, "profiles", "r1*_bench.json")))
    assert files, "no committed bench line under profiles/"
    line = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    r2 = os.path.basename(files[-1]).startswith("r2")
    assert BASE_KEYS | {"roofline", "clocks"} <= set(line)
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["gpu_launches"] > 0 and line["data"] == "synthetic" and "workload" in line["config"] and "l2" in line["config"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 32 << 24 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] < line["value"]
    rf = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["traffic"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    if r2:          # round 2: frac is the EXECUTED multiplier work against the live IMAD.WIDE peak, so it is a real fraction
        assert 0.5 < rf["frac"] <= 1.02 and 0.5 < rf["step_frac"] <= rf["frac"] and rf["model_frac"] > 0
        assert line["scaling"] == "strong" and line["config"]["tau_identity_check"] is True and line["config"]["srs_table_bytes_per_gpu"] > 0
        assert line["e2e"]["pageable"]["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0 and cb["sample"]
    ck = line["clocks"]
    assert ck["sm_mhz"] > 0.9 * ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ntt = line["ntt"]
    assert ntt["roofline"]["bound"] == "hbm" and 0 < ntt["roofline"]["frac"] < 1
    assert 0 < ntt["roofline"]["imad_model_frac" if r2 else "imad_frac"] < 1.05
    assert ntt["e2e"]["h2d_bytes_per_step"] == ntt["e2e"]["d2h_bytes_per_step"] == 32 << 24
    assert line["plonk"]["bundled"]["proof_equals_reference_prover"] is True
    assert line["plonk"]["marlin_bundled"]["proof_equals_reference_prover"] is True
