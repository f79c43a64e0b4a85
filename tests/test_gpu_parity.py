"""GPU parity tests (run with -m gpu on the B200): the product library libima2p_b200.so, called through
the C ABI, against the CPU oracle and the reference-generated golden fixtures.

Bars (BASELINE.json north_star): integer coalescent/migration counts bit-exact; likelihoods and priors
within 1e-9 relative in fp64."""
import numpy as np
import pytest

import engine_checks as ec

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def lib():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ima2p_b200 import capi
    return capi.lib()


@pytest.mark.parametrize("name", ec.STATIC_FIXTURES)
def test_static_eval_matches_reference(lib, name):
    ec.static_eval_matches_reference(lib, name, rtol=RTOL)


@pytest.mark.parametrize("name,nsteps", [("state_sim5_hn4", 10), ("state_sim3_hn3", 20), ("state_sim5_3pop_hn2", 12),
                                         ("state_sim2_hn2", 20), ("state_sim5_hky_hn2", 12), ("state_sim5_4popA_hn2", 10),
                                         ("state_sim5_4popB_hn2", 10)])
def test_device_proposals_match_oracle(lib, name, nsteps):
    # Sim2 is one locus of 100 genes: root moves are rare in 40 proposals, so they are not demanded there
    ec.proposals_match_oracle(lib, name, nsteps, rtol=RTOL, need_root_moves=(name != "state_sim2_hn2"))


@pytest.mark.parametrize("name,nsteps", [("state_sim5_hn4", 2000), ("state_sim5_3pop_hn2", 500), ("state_sim50_hn3", 300),
                                         ("state_sim300_hn1", 100), ("state_sim5_hky_hn2", 300), ("state_sim3_sw_hn2", 400)])
def test_incremental_sums_match_fresh_evaluation(lib, name, nsteps):
    cnt = ec.incremental_sums_match_fresh_evaluation(lib, name, nsteps, rtol=RTOL)
    assert cnt["steps"] == nsteps and cnt["dropped"] == 0


def test_lmode_joint_models_of_three_populations(lib):
    ec.lmode_joint_models_match_reference(lib)


@pytest.mark.parametrize("name", ["lmode_sim5_hn2", "lmode_sim5_expo_hn2"])
def test_lmode_matches_reference(lib, name):
    ec.lmode_matches_reference(lib, name, rtol=RTOL)


@pytest.mark.parametrize("name", ["state_sim3_sw_hn2", "state_sim3_joint_hn2"])
def test_stepwise_updates_match_oracle(lib, name):
    ec.stepwise_updates_match_oracle(lib, name, 40, rtol=RTOL)


@pytest.mark.parametrize("name,nchains,burn,sweeps", [("trace_sim3", 256, 6000, 4000), ("trace_sim5", 256, 6000, 4000)])
def test_long_run_statistics_match_reference_sampler(lib, name, nchains, burn, sweeps):
    """north_star: posterior summaries from long runs agree with the reference statistically."""
    z, m_e, m_r, eng_acc, ref_acc = ec.long_run_summaries_match_reference(lib, name, nchains, burn, sweeps, nsigma=5.0)
    assert abs(z).max() < 5.0


@pytest.mark.parametrize("name", ["lmode_extra_sim5_hn2", "lmode_extra_sim5_expo_hn2", "lmode_extra_sim5_3pop_hn2"])
def test_lmode_moments_and_popmig(lib, name):
    assert ec.lmode_moments_and_popmig_match_reference(lib, name) >= 6


def test_device_incomplete_gamma_matches_reference_tables(lib):
    assert ec.gamma_tables_match_reference(lib, rtol=1e-10) > 400


def test_runs_are_reproducible_and_seed_dependent(lib):
    from support import engine_from_fixture, load_golden
    d = load_golden("state_sim5_hn4")
    outs = []
    for seed in (7, 7, 8):
        eng, _ = engine_from_fixture(d, lib=lib, seed=seed)
        eng.eval()
        eng.run(50)
        eng.sync()
        outs.append(np.array([eng.chain(c)["probg"] for c in range(eng.nchains)]))
        eng.close()
    assert np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("nloci,nchains,nsteps", [(50, 128, 400), (300, 256, 60)])
def test_full_size_workload_properties(lib, nloci, nchains, nsteps):
    # BASELINE configs[1] and the per-GPU shard of configs[2], through size-independent properties
    ec.full_size_workload_properties(lib, nloci, nchains, nsteps)


def test_packed_upload_equals_plain_upload(lib):
    ec.packed_upload_equals_plain_upload(lib)


def test_lmode_f3_matches_oracle_on_bootstrapped_rows(lib):
    ec.lmode_f3_matches_oracle_on_bootstrapped_rows(lib, 200000)


def test_speculation_depth_does_not_change_the_run(lib):
    ec.speculation_depth_does_not_change_the_run(lib)


@pytest.mark.parametrize("name", ["state_sim50_hn3", "state_sim5_3pop_hn2", "state_sim5_hky_hn2", "state_sim5_4popA_hn2"])
def test_fast_path_equals_general_path(lib, name):
    ec.fast_path_equals_general_path(lib, name, ppws=(4, 8, 16))


def test_capacity_grows_without_changing_the_chain(lib):
    ec.capacity_grows_without_changing_the_chain(lib)


def test_swap_walk_by_levels_equals_the_walk_in_order(lib):
    ec.swap_walk_by_levels_equals_the_walk_in_order(lib)


def test_hky_partials_stay_consistent(lib):
    ec.hky_partials_stay_consistent(lib)


def test_scalar_walk_by_levels_equals_the_walk_in_order(lib):
    ec.scalar_walk_by_levels_equals_the_walk_in_order(lib)


def test_programmatic_launches_do_not_change_the_run(lib):
    """Inside the step graph kernels are launched with programmatic stream serialisation and open with griddepcontrol.wait
    (IMA2P_PDL, on by default): the chains must be those of plainly launched kernels."""
    import os
    got = []
    for pdl in ("1", "0"):
        os.environ["IMA2P_PDL"] = pdl
        try:
            got.append(ec.run_summary(lib, nsteps=120))
        finally:
            os.environ.pop("IMA2P_PDL", None)
    assert got[0] is not None and np.array_equal(got[0], got[1])


def test_pipeline_does_not_change_the_run(lib):
    ec.pipeline_does_not_change_the_run(lib)


# ---- section 8 (f1): the rest of the step on the device ------------------------------------------------------------
@pytest.mark.parametrize("name", ["tupdates_sim5_hn2", "tupdates_sim5_3pop_hn2", "tupdates_sim3_sw_hn2", "tupdates_sim3_joint_hn2"])
def test_split_time_update_matches_reference(lib, name):
    ec.split_time_update_matches_reference(lib, name)


@pytest.mark.parametrize("name", ["uupdates_sim5_hn2", "uupdates_sim5_hky_hn2", "uupdates_sim3_sw_hn2", "uupdates_sim3_joint_hn2"])
def test_mutation_scalar_update_matches_reference(lib, name):
    ec.mutation_scalar_update_matches_reference(lib, name)


@pytest.mark.parametrize("name,nsteps", [("state_sim5_hn4", 2000), ("state_sim5_3pop_hn2", 500), ("state_sim50_hn3", 300),
                                         ("state_sim3_sw_hn2", 300), ("state_sim5_hky_hn2", 100), ("state_sim3_joint_hn2", 300),
                                         ("state_sim5_4popA_hn2", 300), ("state_sim5_4popB_hn2", 300)])
def test_incremental_sums_with_full_schedule(lib, name, nsteps):
    ec.incremental_sums_match_fresh_evaluation(lib, name, nsteps, full_schedule=True)


@pytest.mark.parametrize("name,nchains,burn,sweeps", [("trace_full_sim3", 64, 15000, 20000), ("trace_full_sim5", 64, 15000, 20000)])
def test_full_schedule_posterior_matches_reference_sampler(lib, name, nchains, burn, sweeps):
    z, zt, zu, tm, tr, uc = ec.long_run_summaries_match_reference(lib, name, nchains, burn, sweeps, full_schedule=True)
    assert abs(zt).max() < 5.0 and abs(zu).max() < 5.0


def test_recent_split_time_statistics(lib):
    z, _, _, _, _ = ec.long_run_summaries_match_reference(lib, "trace_sim3_recent", 256, 4000, 4000)
    assert abs(z).max() < 5.0


def test_thermodynamic_integration(lib):
    ec.thermodynamic_integration_matches_reference(lib)
    # accumulators: after k recorded steps every temperature slot holds k chain likelihood sums
    d = ec.load_golden("state_sim5_hn4")
    eng, fm = ec.engine_from_fixture(d, lib=lib)
    eng.eval()
    tot = 0.0
    for _ in range(7):
        eng.run(3)
        eng.thermo_accumulate()
        tot += sum(eng.chain(c)["pdg"] for c in range(eng.nchains))
    s = eng.thermo_sums(reset=True)
    assert ec.rel_close(s.sum(), tot, 1e-12) and np.all(s < 0) and np.all(eng.thermo_sums() == 0)
    eng.close()


@pytest.mark.parametrize("name", ["state_sim5_nomig_hn2", "state_sim5_3pop_nomig_hn2"])
def test_no_migration_model(lib, name):
    ec.proposals_match_oracle(lib, name, 15, need_root_moves=False)
    cnt = ec.incremental_sums_match_fresh_evaluation(lib, name, 500)
    assert cnt["accepted"] > 0 and cnt["topology"] > 0


@pytest.mark.parametrize("name", ["trace_sim3_nomig", "trace_sim5_3pop_nomig"])
def test_no_migration_statistics(lib, name):
    z, _, _, _, _ = ec.long_run_summaries_match_reference(lib, name, 256, 4000, 4000)
    assert abs(z).max() < 5.0


@pytest.mark.parametrize("name", ["nwupdates_sim5_hn2", "nwupdates_sim3_hn2", "nwupdates_sim5_3pop_hn2"])
def test_nielsen_wakeley_update_matches_oracle(lib, name):
    ec.nielsen_wakeley_update_matches_oracle(lib, name)


@pytest.mark.parametrize("name", ["mcf_sim5_hn2", "mcf_sim3_sw_hn2", "mcf_sim5_hky_hn2", "mcf_sim3_joint_hn2"])
def test_reference_state_file_evaluates_to_reference_values(lib, name, tmp_path):
    from test_mcf import mcf_read_matches_reference
    mcf_read_matches_reference(lib, name, tmp_path, rtol=RTOL)


def test_step_report(lib):
    ec.step_report_matches_separate_reads(lib)


def test_cold_chain_counters(lib):
    ec.cold_chain_counters_are_consistent(lib)


def test_front_end_on_the_device(lib, tmp_path):
    """ima2p_b200/IMa2p_b200 (the product executable) from a .u file: posterior summaries against the reference's."""
    import os
    from test_frontend import frontend_posterior_matches_reference
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ima2p_b200", "IMa2p_b200")
    assert os.path.exists(exe), "build the front end with __graft_entry__.build()"
    frontend_posterior_matches_reference(exe, lib, tmp_path)


def test_front_end_report_head_on_the_device(lib, tmp_path):
    """The opening sections of the product executable's report against the reference's own (run information, the cold
    chain's update-rate tables per locus and update type, swaps between adjacent temperatures)."""
    import os
    from test_frontend import frontend_report_head_matches_reference
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ima2p_b200", "IMa2p_b200")
    assert os.path.exists(exe), "build the front end with __graft_entry__.build()"
    frontend_report_head_matches_reference(exe, tmp_path, lib=lib)


@pytest.mark.parametrize("name", ["lmode_report_sim3", "lmode_report_3pop"])
def test_front_end_l_mode_report_on_the_device(lib, tmp_path, name):
    """The product executable's L mode on the device: the report sections of the reference's own L-mode run on the same .ti file
    character for character (greater-than tables, moments, the marginal peak table found by the lock-step searches, histogram
    groups), the joint-posterior peaks (-c2; three populations: the size-only and migration-only models) and the nested models
    of -w, all against the reference's tables."""
    import os
    import test_frontend as tf
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ima2p_b200", "IMa2p_b200")
    assert os.path.exists(exe), "build the front end with __graft_entry__.build()"
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    tf.test_l_mode_report_sections_equal_the_reference_text(exe, tmp_path / "a", name)
    tf.test_nested_models_of_the_joint_search(exe, tmp_path / "b", name)


def test_two_gpus_front_end_writes_the_single_gpu_ti_file(lib, tmp_path):
    """The product executable started once per GPU (RANK / WORLD_SIZE / LOCAL_RANK): the ranks exchange swap sums and the cold
    chain's record through peer memory (cudaIpc handles over the rendezvous files) and rank 0 writes the .ti file the one-GPU
    run of the same seed writes.  Needs two GPUs."""
    import os
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from ima2p_b200 import synth
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ima2p_b200", "IMa2p_b200")
    u = tmp_path / "two.u"
    synth.write_u(str(u), synth.make_dataset(6, 10, 10, seed=5))
    common = ["-i", str(u), "-q10", "-m1", "-t3", "-b2000", "-l100", "-d10", "-hfg", "-ha0.96", "-hb0.9", "-s11"]
    r1 = subprocess.run([exe] + common + ["-hn16", "-o", str(tmp_path / "one.out")], capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stderr
    env = dict(os.environ, WORLD_SIZE="2", MASTER_PORT="29998", IMA2P_RENDEZVOUS_DIR=str(tmp_path))
    procs = [subprocess.Popen([exe] + common + ["-hn8", "-o", str(tmp_path / "two.out")], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-400:] for o in outs]
    body = lambda path: open(path).read().split("VALUESSTART", 1)[1]
    assert body(tmp_path / "one.out.ti") == body(tmp_path / "two.out.ti")
