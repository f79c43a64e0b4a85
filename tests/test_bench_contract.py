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


# ---- round 2 records (profiles/r2s*): real input by default, nine launches a step, the target configuration at N >= 2 ----------
def test_round2_single_gpu_record():
    d = _line("r2s5_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "models", "dropped_for_capacity"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["data"] == "real" and d["config"]["input"] == "real" and d["dtype"] == "f64" and d["warmup"] >= 3
    assert "L2" in d["config"]["l2"] and d["dropped_for_capacity"] == 0
    assert abs(d["value"] - 128 * 50 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    # k_move, k_weigh, k_propose_redo, k_accept, k_split_t_fast, k_split_t_redo, k_accept_t, k_changeu, k_swap
    assert d["gpu_launches"] == d["steps"] * 9
    e = d["e2e"]
    assert e["statistic"].startswith("median") and len(e["repetitions_s"]) >= 3 and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["peak_kind"] in ("measured", "fallback") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms_per_launch"][r["kernel"]] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert set(r["per_kernel"]) >= {"k_move", "k_weigh", "k_accept", "k_split_t", "k_accept_t", "k_changeu", "k_swap"} and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and "real input" in c["sample"]
    for m in ("hky", "stepwise", "joint_is_sw"):
        assert d["models"][m]["value"] > 0 and d["models"][m]["dropped_for_capacity"] == 0
    assert d["lmode"]["fp64_fma_per_sec_measured"] > 0 and d["lmode"]["fp64_exp_per_sec_measured"] > 0


@pytest.mark.parametrize("name,n", [("r2s4_bench_n2.json", 2), ("r2s4_bench_n8.json", 8)])
def test_round2_multi_gpu_records_carry_the_target_configuration(name, n):
    d = _line(name)
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["data"] == "real" and d["config"]["chains_total"] == 128 * n
    assert abs(d["value"] - d["config"]["chains_total"] * d["config"]["loci"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert "peer memory" in d["multi_gpu_step"] and d["gpu_launches"] == d["steps"] * 9
    c3 = d["config3"]
    assert c3["config"]["loci"] == 300 and c3["config"]["chains_total"] == 256 * n and "larger than" in c3["config"]["l2"]
    assert abs(c3["value"] - c3["config"]["chains_total"] * 300 / (c3["ms_per_step"] * 1e-3)) <= 1e-6 * c3["value"]
    assert c3["cpu_baseline"]["kind"] == "reference" and "300 loci" in c3["cpu_baseline"]["sample"] and c3["dropped_for_capacity"] == 0
    assert abs(c3["ratio_to_cpu_baseline"] - c3["value"] / c3["cpu_baseline"]["value"]) < 1e-9 * c3["ratio_to_cpu_baseline"]


def test_round2_reference_arm_record():
    d = _line("r2s3_bench_ref.json")
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and "qupdate" in d["cpu_baseline"]["sample"]


# ---- the round's final records (profiles/r2s6_*, r2s7_*): launches counted from the engine's settings, uploads overlapping the steps ----
def test_final_single_gpu_record():
    d = _line("r2s7_bench_n1.json")
    assert d["n_gpus"] == 1 and d["data"] == "real" and d["dtype"] == "f64" and d["warmup"] >= 3 and d["dropped_for_capacity"] == 0
    assert abs(d["value"] - 128 * 50 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["gpu_launches"] == d["steps"] * 17          # two chain groups x eight kernels + k_swap (ima2p_engine_launches_per_step)
    e = d["e2e"]
    assert "upload_block" in e["upload"] and e["statistic"].startswith("median") and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms_per_launch"][r["kernel"]] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] > 0 and r["peak_kind"] in ("measured", "fallback")
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and 0.2 < c["accept_rate"] < 0.5 and abs(c["accept_rate"] - d["accept_rate"]) < 0.05
    ref = _line("r2s7_bench_ref.json")
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["config"]["workload"] == d["config"]["workload"]
    assert ref["e2e"]["value"] == ref["value"] and ref["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.parametrize("name,n", [("r2s6_bench_n2.json", 2), ("r2s6_bench_n8.json", 8)])
def test_final_multi_gpu_records(name, n):
    d = _line(name)
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["config"]["chains_total"] == 128 * n and d["gpu_launches"] == d["steps"] * 17
    assert abs(d["value"] - d["config"]["chains_total"] * d["config"]["loci"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    c3 = d["config3"]
    assert c3["config"]["chains_total"] == 256 * n and c3["gpu_launches"] == c3["steps"] * 33      # four groups where the kernels run many waves
    assert abs(c3["ratio_to_cpu_baseline"] - c3["value"] / c3["cpu_baseline"]["value"]) < 1e-9 * c3["ratio_to_cpu_baseline"]


# ---- session 8 (profiles/r2s8_*): results read through the two-slot report, jointp in one pass of 512 vectors ----
def test_session8_records():
    d, old = _line("r2s8_bench_n1.json"), _line("r2s7_bench_n1.json")
    assert d["n_gpus"] == 1 and d["data"] == "real" and d["config"] == old["config"] and d["gpu_launches"] == d["steps"] * 17
    assert abs(d["value"] - 128 * 50 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert "step_report_begin" in e["read_back"] and e["h2d_bytes_per_step"] > 5e6 and e["d2h_bytes_per_step"] == 4280
    assert old["e2e"]["value"] < e["value"] < d["value"]
    lm = d["lmode"]
    assert lm["jointp_vectors"] == 512 and lm["jointp_geneval_per_sec"] > 3 * old["lmode"]["jointp_geneval_per_sec"]
    assert abs(lm["margincalc_geneval_per_sec"] / old["lmode"]["margincalc_geneval_per_sec"] - 1) < 0.05      # untouched kernel, same figure
    d2 = _line("r2s8_bench_n2.json")
    assert d2["n_gpus"] == 2 and d2["config"]["chains_total"] == 256 and d2["value"] > 1.9 * d["value"]
    assert d2["lmode"]["jointp_geneval_per_sec"] > 1.8 * lm["jointp_geneval_per_sec"]                       # rows sharded over the two GPUs
    d8 = _line("r2s8_bench_n8.json")
    assert d8["n_gpus"] == 8 and d8["config"]["chains_total"] == 1024 and d8["value"] > 7.5 * d["value"] and d8["e2e"]["value"] > 1e8
    # the VERDICT's "jointp flat at 2.4e10 for every N": now 4.6 x N = 1 on eight GPUs, 7.5 x the figure the round began with
    assert d8["lmode"]["jointp_geneval_per_sec"] > 4.5 * lm["jointp_geneval_per_sec"] > 4.5 * 3 * old["lmode"]["jointp_geneval_per_sec"]
    assert d8["lmode"]["margincalc_geneval_per_sec"] > 5.5 * lm["margincalc_geneval_per_sec"]
    # the session's last build (the scan queues the rows that enter the sum): same step, jointp past 1e11 on one GPU
    f = _line("r2s8_bench_n1_final.json")
    assert f["config"] == d["config"] and abs(f["value"] / d["value"] - 1) < 0.01 and f["e2e"]["value"] > 15.5e6
    assert f["lmode"]["jointp_geneval_per_sec"] > 1e11 > lm["jointp_geneval_per_sec"]
