"""The bench line contract (driver-facing): the records committed under profiles/ are real outputs of bench.py on B200; this
checks that they carry every key the contract names, with consistent values, so that a change of bench.py that drops one is
caught on CPU.  (The numbers themselves are measurements, not assertions.)"""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    txt = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()
    return json.loads([l for l in txt if l.startswith("{")][-1])


def test_single_gpu_record_has_the_contract_keys():
    d = _line("r1s3_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and d["warmup"] >= 3
    # value = units / time: 128 chains x 50 loci per step
    assert abs(d["value"] - 128 * 50 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] > 0
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms_per_launch"][r["kernel"]] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] and k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] == d["steps"] * 6       # propose, accept, split_t, accept_t, changeu, swap


def test_reference_arm_record():
    d = _line("r1s3_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == _line("r1s3_bench_n1.json")["metric"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["workload"] == _line("r1s3_bench_n1.json")["config"]["workload"]


@pytest.mark.parametrize("name,n", [("r1s2_bench_n2.json", 2), ("r1s2_bench_n4.json", 4), ("r1s3_bench_n8.json", 8), ("r1s3_bench_n8_sim300x256.json", 8)])
def test_multi_gpu_records(name, n):
    d = _line(name)
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["config"]["chains_total"] % n == 0
    per_step = d["config"]["chains_total"] * d["config"]["loci"]
    assert abs(d["value"] - per_step / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["e2e"]["value"] > 0 and d["clocks"]["sm_mhz"]
