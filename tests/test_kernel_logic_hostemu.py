"""CPU tests of the KERNEL SOURCES through the tests-only host-emulation build (tests/hostemu/build.sh):
the same ima2p_b200/csrc/*.h|*.cu files compiled by g++ with a "warp" of one lane.  They check the kernel
logic against the oracle and the golden fixtures where there is no GPU.  The product never loads this
library -- the GPU parity tests proper are tests/test_gpu_parity.py (-m gpu) through libima2p_b200.so."""
import os
import subprocess

import pytest

import engine_checks as ec
from ima2p_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "hostemu", "libima2p_hostemu.so")


@pytest.fixture(scope="module")
def emu():
    subprocess.run([os.path.join(HERE, "hostemu", "build.sh")], check=True)
    return capi.bind(EMU)


@pytest.mark.parametrize("name", ec.STATIC_FIXTURES)
def test_static_eval(emu, name):
    ec.static_eval_matches_reference(emu, name, rtol=1e-12)


@pytest.mark.parametrize("name,nsteps", [("state_sim5_hn4", 12), ("state_sim3_hn3", 25), ("state_sim5_3pop_hn2", 15), ("state_sim5_hky_hn2", 10), ("state_sim5_4popA_hn2", 10), ("state_sim5_4popB_hn2", 10)])
def test_proposals(emu, name, nsteps):
    ec.proposals_match_oracle(emu, name, nsteps)


@pytest.mark.parametrize("name,nsteps", [("state_sim5_hn4", 300), ("state_sim5_3pop_hn2", 150), ("state_sim50_hn3", 20)])
def test_incremental_sums(emu, name, nsteps):
    cnt = ec.incremental_sums_match_fresh_evaluation(emu, name, nsteps)
    assert cnt["steps"] == nsteps and cnt["accepted"] > 0


@pytest.mark.parametrize("name", ["lmode_sim5_hn2", "lmode_sim5_expo_hn2"])
def test_lmode(emu, name):
    ec.lmode_matches_reference(emu, name, rtol=1e-12)


def test_workload_properties_small(emu):
    # the GPU suite runs this at BASELINE sizes (50 x 128 and 300 x 256); here a size the one-lane emulation finishes in seconds
    ec.full_size_workload_properties(emu, 12, 6, 40, noracle=16)


def test_packed_upload_equals_plain_upload(emu):
    ec.packed_upload_equals_plain_upload(emu)


def test_lmode_f3_matches_oracle_on_bootstrapped_rows(emu):
    ec.lmode_f3_matches_oracle_on_bootstrapped_rows(emu, 2500)


def test_speculation_depth_does_not_change_the_run(emu):
    ec.speculation_depth_does_not_change_the_run(emu, nsteps=25)


@pytest.mark.parametrize("name", ["state_sim50_hn3", "state_sim5_3pop_hn2", "state_sim5_hky_hn2", "state_sim5_4popA_hn2"])
def test_fast_path_equals_general_path(emu, name):
    ec.fast_path_equals_general_path(emu, name, nsteps=25)


def test_capacity_grows_without_changing_the_chain(emu):
    ec.capacity_grows_without_changing_the_chain(emu)


def test_hky_partials_stay_consistent(emu):
    ec.hky_partials_stay_consistent(emu, nsteps=200)


def test_scalar_walk_by_levels_equals_the_walk_in_order(emu):
    ec.scalar_walk_by_levels_equals_the_walk_in_order(emu, nsteps=30)


def test_pipeline_setting_is_accepted_and_does_not_change_the_run(emu):
    ec.pipeline_does_not_change_the_run(emu, nsteps=9)


def test_lmode_joint_models_of_three_populations(emu):
    ec.lmode_joint_models_match_reference(emu, rtol=1e-12)


@pytest.mark.parametrize("name", ["lmode_extra_sim5_hn2", "lmode_extra_sim5_expo_hn2", "lmode_extra_sim5_3pop_hn2"])
def test_lmode_moments_and_popmig(emu, name):
    assert ec.lmode_moments_and_popmig_match_reference(emu, name) >= 6


def test_gamma_tables(emu):
    assert ec.gamma_tables_match_reference(emu, rtol=1e-12) > 400


@pytest.mark.parametrize("name", ["state_sim3_sw_hn2", "state_sim3_joint_hn2"])
def test_stepwise_updates(emu, name):
    ec.stepwise_updates_match_oracle(emu, name, 40, rtol=1e-10)


def test_long_run_statistics_match_reference_sampler(emu):
    # statistical parity of the whole sampler (proposal distributions, Hastings terms, prior, likelihood)
    z, _, _, eng_acc, ref_acc = ec.long_run_summaries_match_reference(emu, "trace_sim3", 64, 5000, 3000)
    assert abs(z).max() < 5.0


@pytest.mark.parametrize("name", ["tupdates_sim5_hn2", "tupdates_sim5_3pop_hn2", "tupdates_sim3_sw_hn2", "tupdates_sim3_joint_hn2"])
def test_split_time_update(emu, name):
    ec.split_time_update_matches_reference(emu, name)


@pytest.mark.parametrize("name", ["uupdates_sim5_hn2", "uupdates_sim5_hky_hn2", "uupdates_sim3_sw_hn2", "uupdates_sim3_joint_hn2"])
def test_mutation_scalar_update(emu, name):
    ec.mutation_scalar_update_matches_reference(emu, name)


@pytest.mark.parametrize("name,nsteps", [("state_sim5_hn4", 200), ("state_sim5_3pop_hn2", 100), ("state_sim3_sw_hn2", 60), ("state_sim5_hky_hn2", 30), ("state_sim3_joint_hn2", 60), ("state_sim5_4popA_hn2", 80),
                                         ("state_sim5_4popB_hn2", 80)])
def test_incremental_sums_full_schedule(emu, name, nsteps):
    ec.incremental_sums_match_fresh_evaluation(emu, name, nsteps, full_schedule=True)


def test_thermodynamic_integration(emu):
    ec.thermodynamic_integration_matches_reference(emu)


def test_full_schedule_posterior_matches_reference_sampler(emu):
    # genealogies + split time + mutation scalars: the posterior of t and of the scalars against whole qupdate() runs
    # of the reference (the split time mixes slowly: a long burn-in is part of the test)
    z, zt, zu, _, _, uc = ec.long_run_summaries_match_reference(emu, "trace_full_sim3", 8, 8000, 30000, full_schedule=True)
    assert abs(zt).max() < 5.0 and abs(zu).max() < 5.0 and uc["t_accepts"] > 0 and uc["u_accepts"] > 0


def test_recent_split_time_statistics(emu):
    # the genealogy sampler where most of every genealogy lies in the ancestral population
    z, _, _, _, _ = ec.long_run_summaries_match_reference(emu, "trace_sim3_recent", 16, 3000, 6000)
    assert abs(z).max() < 5.0


@pytest.mark.parametrize("name", ["state_sim5_nomig_hn2", "state_sim5_3pop_nomig_hn2"])
def test_no_migration_model(emu, name):
    # -m 0: slider_nomigration (update_gtree.cpp:78-253), no migration path, no migration terms in the prior
    ec.proposals_match_oracle(emu, name, 15, need_root_moves=False)
    cnt = ec.incremental_sums_match_fresh_evaluation(emu, name, 200)
    assert cnt["accepted"] > 0 and cnt["topology"] > 0


def test_no_migration_statistics(emu):
    z, _, _, _, _ = ec.long_run_summaries_match_reference(emu, "trace_sim5_3pop_nomig", 16, 2000, 6000)
    assert abs(z).max() < 5.0


@pytest.mark.parametrize("name", ["nwupdates_sim5_hn2", "nwupdates_sim3_hn2", "nwupdates_sim5_3pop_hn2"])
def test_nielsen_wakeley_update(emu, name):
    ec.nielsen_wakeley_update_matches_oracle(emu, name)


def test_step_report(emu):
    ec.step_report_matches_separate_reads(emu)


def test_cold_chain_counters(emu):
    ec.cold_chain_counters_are_consistent(emu)
