"""bench.py's output contract, checked without a GPU: the reference arm runs here (it is the oracle on host cores),
and the committed bench lines of the CUDA arm (profiles/) are checked for the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--config", "test_small", "--ref-cores", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("name", ["r01_bench_n1.json", "r01_bench_n2.json", "r01_bench_n8.json", "r02_bench_n1.json", "r02_bench_n2.json"])
def test_committed_bench_lines_have_the_contract_keys(name):
    path = os.path.join(ROOT, "profiles", name)
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d), (BASE_KEYS | {"roofline", "clocks"}) - set(d)
    assert d["metric"].startswith("voxelized frames/sec") and d["unit"] == "frames/s" and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["gpu_launches"] > 0 and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]


def test_round2_lines_carry_the_parity_check_and_the_extra_records():
    n1 = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
    pc = n1["parity_check"]
    assert pc["oracle"] == "ref" and pc["mismatches"] == 0 and pc["files"] == 4 * pc["frames"] >= 16
    assert n1["single_step"]["value"] > 0 and n1["single_step"]["value"] <= n1["value"] * 1.05
    assert set(n1["extra_configs"]) == {"dense", "stress_512", "past1000"}
    for rec in n1["extra_configs"].values():
        assert rec["value"] > 0 and rec["frames"] > 0 and "reference" in rec
    assert n1["cpu_baseline"]["single_thread"]["cores"] == 1
    assert "profiles/r02_fill_traffic.json" in n1["roofline"]["traffic_source"]
    n2 = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n2.json")).read().strip().splitlines()[-1])
    for rec in n2["strong"].values():
        assert rec["files_equal"] is True and rec["speedup"] > 1.0 and rec["n_gpus"] == 2
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] == "reference" and len(ref["frame_hashes"]) == ref["config"]["frames_per_step"]
