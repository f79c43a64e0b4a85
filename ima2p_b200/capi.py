"""ctypes binding of the C ABI declared in include/ima2p_b200.h.

The product library is ima2p_b200/libima2p_b200.so (CUDA, sm_100a), built in-tree by
``__graft_entry__.build()``.  There is no CPU implementation behind this binding: loading fails loudly
when the library is missing, and every entry point returns IMA2P_E_CUDA when no device is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# IMA2P_B200_LIB selects another BUILD of the same CUDA library (e.g. a tuning variant); there is no non-CUDA build
LIB_PATH = os.environ.get("IMA2P_B200_LIB", os.path.join(HERE, "libima2p_b200.so"))

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)
c_u64_p = C.POINTER(C.c_uint64)
c_u32_p = C.POINTER(C.c_uint)

# every symbol include/ima2p_b200.h declares: name -> (restype, argtypes)
_i, _d, _v, _ll = C.c_int, C.c_double, C.c_void_p, C.c_longlong
SIGNATURES = {
    "ima2p_version": (C.c_char_p, []),
    "ima2p_last_error": (C.c_char_p, []),
    "ima2p_engine_create": (_i, [C.POINTER(_v), _i, _i, _i, _i, _i, _i, C.c_uint64]),
    "ima2p_engine_destroy": (None, [_v]),
    "ima2p_engine_set_model": (_i, [_v, _i, _i, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, _i, _i, c_int_p, c_int_p,
                                    c_int_p, c_dbl_p, c_dbl_p, _i, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p,
                                    c_dbl_p, _i, c_int_p, c_int_p, c_int_p, _i, _i, _i, _d]),
    "ima2p_engine_set_locus": (_i, [_v, _i, _i, _i, _i, _i, _d, c_int_p, c_int_p, c_int_p, _i, c_int_p, c_int_p, _d]),
    "ima2p_engine_finalize": (_i, [_v]),
    "ima2p_engine_set_heating": (_i, [_v, _i, _d, _d]),
    "ima2p_engine_set_betas": (_i, [_v, c_dbl_p]),
    "ima2p_engine_set_chain": (_i, [_v, _i, c_dbl_p]),
    "ima2p_engine_set_genealogy": (_i, [_v, _i, _i, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p, c_dbl_p, c_int_p,
                                        _i, _d, c_dbl_p, _d, c_dbl_p, c_int_p]),
    "ima2p_engine_get_genealogy": (_i, [_v, _i, _i, _i, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p, c_dbl_p,
                                        c_int_p, _i, c_int_p, c_dbl_p]),
    "ima2p_engine_get_alleles": (_i, [_v, _i, _i, _i, c_int_p, c_dbl_p, c_dbl_p]),
    "ima2p_engine_upload": (_i, [_v]),
    "ima2p_engine_eval": (_i, [_v]),
    "ima2p_engine_get_pair": (_i, [_v, _i, _i, c_int_p, c_dbl_p, c_dbl_p, c_int_p]),
    "ima2p_engine_get_chain": (_i, [_v, _i, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p]),
    "ima2p_engine_dims": (_i, [_v, c_int_p]),
    "ima2p_engine_run": (_i, [_v, _i, _i, _v]),
    "ima2p_engine_set_pipeline": (_i, [_v, _i, _i, _i]),
    "ima2p_engine_launches_per_step": (_i, [_v, _i]),
    "ima2p_engine_set_proposal_path": (_i, [_v, _i, _i]),
    "ima2p_engine_set_debug_records": (_i, [_v, _i]),
    "ima2p_engine_grow_capacity": (_i, [_v, _i]),
    "ima2p_engine_exchange_create": (_i, [_v, C.POINTER(_v), c_u64_p]),
    "ima2p_engine_exchange_attach": (_i, [_v, C.POINTER(_v)]),
    "ima2p_ipc_export": (_i, [_v, C.c_char_p]),
    "ima2p_ipc_import": (_i, [_i, C.c_char_p, C.POINTER(_v)]),
    "ima2p_engine_run_sharded": (_i, [_v, _i, _i, _v]),
    "ima2p_engine_cold_message": (_i, [_v, c_dbl_p, _v]),
    "ima2p_engine_sharded_update": (_i, [_v, _v]),
    "ima2p_engine_sharded_swap": (_i, [_v, _i, _v]),
    "ima2p_engine_set_speculation": (_i, [_v, _i]),
    "ima2p_engine_run_timed": (_i, [_v, _i, _i, _v, c_flt_p]),
    "ima2p_engine_update_genealogies": (_i, [_v, _v, _v]),
    "ima2p_engine_swap_replay": (_i, [_v, _v, _i, _v]),
    "ima2p_engine_step_propose": (_i, [_v, _v]),
    "ima2p_engine_step_decide": (_i, [_v, _v, _v]),
    "ima2p_engine_swap_replay_late": (_i, [_v, _v, _i, _v]),
    "ima2p_engine_get_proposal": (_i, [_v, _i, _i, c_dbl_p, c_u32_p, c_int_p]),
    "ima2p_debug_gamma": (_i, [_i, c_int_p, c_dbl_p, _i, c_dbl_p]),
    "ima2p_engine_counters": (_i, [_v, c_u64_p]),
    "ima2p_engine_set_update_schedule": (_i, [_v, _i, _i]),
    "ima2p_engine_set_update_priors": (_i, [_v, c_dbl_p, c_dbl_p, _d, _d, _d, _d]),
    "ima2p_engine_update_counters": (_i, [_v, c_u64_p]),
    "ima2p_engine_fetch_chain_pdg": (_i, [_v, _i, c_dbl_p]),
    "ima2p_engine_cold_counters": (_i, [_v, c_u64_p, c_u64_p, c_u64_p, c_u64_p]),
    "ima2p_engine_get_split_times": (_i, [_v, _i, c_dbl_p]),
    "ima2p_engine_fetch_parameters": (_i, [_v, c_dbl_p, c_dbl_p, c_dbl_p]),
    "ima2p_engine_get_scalars": (_i, [_v, _i, _i, c_dbl_p, c_dbl_p]),
    "ima2p_engine_debug_split_time": (_i, [_v, _i, _i, c_dbl_p, _i, c_dbl_p]),
    "ima2p_engine_debug_changeu": (_i, [_v, _i, _i, _i, _d, _d, _d, c_dbl_p]),
    "ima2p_engine_thermo_accumulate": (_i, [_v, _v]),
    "ima2p_engine_thermo_sums": (_i, [_v, c_dbl_p, _i]),
    "ima2p_thermo_marginlike": (_i, [c_dbl_p, _i, _i, c_dbl_p]),
    "ima2p_engine_get_betas": (_i, [_v, c_dbl_p]),
    "ima2p_engine_cold_row": (_i, [_v, c_flt_p, c_int_p]),
    "ima2p_engine_sync": (_i, [_v]),
    "ima2p_engine_state_bytes": (_i, [_v, c_u64_p]),
    "ima2p_engine_put_state": (_i, [_v, _v, _v, _v, _v, _v, _v, _v, _v, c_dbl_p, _v]),
    "ima2p_engine_state_block_layout": (_i, [_v, _ll, c_u64_p]),
    "ima2p_engine_put_state_block": (_i, [_v, _v, _ll, _v]),
    "ima2p_engine_upload_block": (_i, [_v, _v, _ll, _v]),
    "ima2p_engine_adopt_block": (_i, [_v, _v]),
    "ima2p_engine_put_state_packed": (_i, [_v, _v, _v, _v, _v, _v, _v, _v, _v, c_dbl_p, _v]),
    "ima2p_engine_fetch_state": (_i, [_v, _v, _v, _v, _v, _v, _v, _v, _v]),
    "ima2p_engine_fetch_pair_summaries": (_i, [_v, c_dbl_p, c_int_p, c_int_p, _v]),
    "ima2p_engine_fetch_chain_summary": (_i, [_v, c_dbl_p, _v]),
    "ima2p_modelspec_create": (_i, [C.POINTER(_v), _i, C.c_char_p, _d, _d, _i, _d, _i, _d]),
    "ima2p_modelspec_free": (None, [_v]),
    "ima2p_modelspec_dims": (_i, [_v, c_int_p]),
    "ima2p_modelspec_tables": (_i, [_v] + [c_int_p] * 13),
    "ima2p_engine_set_model_spec": (_i, [_v, _v]),
    "ima2p_dataset_read": (_i, [C.c_char_p, C.POINTER(_v)]),
    "ima2p_dataset_free": (None, [_v]),
    "ima2p_dataset_dims": (_i, [_v, c_int_p, c_int_p, C.c_char_p, _i]),
    "ima2p_dataset_text": (_i, [_v, _i, _i, C.c_char_p, _i]),
    "ima2p_dataset_locus": (_i, [_v, _i, c_int_p, c_dbl_p, c_int_p, C.c_char_p, _i]),
    "ima2p_dataset_locus_data": (_i, [_v, _i, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p]),
    "ima2p_engine_step_report": (_i, [_v, c_dbl_p, c_flt_p, c_int_p, _v]),
    "ima2p_engine_step_report_begin": (_i, [_v, _i, _v]),
    "ima2p_engine_step_report_end": (_i, [_v, _i, c_dbl_p, c_flt_p, c_int_p]),
    "ima2p_engine_write_mcf": (_i, [_v, C.c_char_p]),
    "ima2p_engine_read_mcf": (_i, [_v, C.c_char_p]),
    "ima2p_ti_create": (_i, [C.c_char_p, C.c_char_p]),
    "ima2p_ti_append": (_i, [C.c_char_p, c_flt_p, _ll, _i]),
    "ima2p_ti_load": (_i, [C.c_char_p, _i, c_flt_p, _ll, C.POINTER(_ll)]),
    "ima2p_lmode_create": (_i, [C.POINTER(_v), _i, _i, _i, _i, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, _i]),
    "ima2p_lmode_destroy": (None, [_v]),
    "ima2p_lmode_load": (_i, [_v, c_flt_p, _i, _i, _ll]),
    "ima2p_lmode_marginal_sums": (_i, [_v, _i, c_dbl_p, _i, _i, _i, _i, c_dbl_p, _v, _v]),
    "ima2p_lmode_margincalc": (_i, [_v, _i, c_dbl_p, _i, _d, _i, c_dbl_p]),
    "ima2p_lmode_marginp": (_i, [_v, _i, _i, _i, c_dbl_p, _i, c_dbl_p]),
    "ima2p_lmode_set_joint_model": (_i, [_v, _i]),
    "ima2p_lmode_marginal_many": (_i, [_v, _i, c_int_p, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p]),
    "ima2p_lmode_jointp": (_i, [_v, c_dbl_p, _i, _i, c_dbl_p, c_dbl_p]),
    "ima2p_lmode_moments": (_i, [_v, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p]),
    "ima2p_lmode_moments_finish": (None, [_i, c_dbl_p, _ll, c_dbl_p, c_dbl_p, c_dbl_p]),
    "ima2p_lmode_popmig_sums": (_i, [_v, _i, _i, c_dbl_p, _i, _i, _i, c_dbl_p]),
    "ima2p_lmode_popmig": (_i, [_v, _i, _i, c_dbl_p, _i, _i, c_dbl_p]),
    "ima2p_lmode_marginpopmig": (_i, [_v, _i, _i, _i, _i, c_dbl_p, _i, c_dbl_p]),
    "ima2p_lmode_greater_than": (_i, [_v, _i, _i, _i, c_dbl_p]),
    "ima2p_lmode_joint_phase1": (_i, [_v, c_dbl_p, _i, c_dbl_p, c_dbl_p]),
    "ima2p_lmode_joint_reseed": (_i, [_v, _i, c_dbl_p, c_dbl_p]),
    "ima2p_lmode_joint_phase2": (_i, [_v, _i, c_dbl_p, _ll, c_dbl_p]),
    "ima2p_lmode_joint_finish": (None, [c_dbl_p, _d, _ll, _i, c_dbl_p, c_dbl_p]),
    "ima2p_lmode_joint_begin": (_i, [_v, c_dbl_p, _i, _v, _v]),
    "ima2p_lmode_joint_finish_gathered": (None, [c_dbl_p, _i, _i, _ll, _i, c_dbl_p, c_dbl_p]),
    "ima2p_debug_fp64_peaks": (_i, [_i, c_dbl_p]),
    "ima2p_lmode_joint_middle": (_i, [_v, _i, _v, _i, _i, _ll, _v, _v]),
}

MAX_LINKED = 4          # IMA2P_MAX_LINKED
E_ARG, E_CUDA, E_UNSUPPORTED, E_DEVICE, E_CAPACITY = -1, -2, -3, -4, -5


class Ima2pError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ima2p_b200 error %d: %s" % (code, msg))
        self.code = code


def bind(path=None):
    """Load a build of the C ABI and set the prototypes of every declared symbol."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(ima2p_b200 has no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib():
    """The product library.  Only a CUDA build is accepted here: the tests-only host emulation of the kernels
    (tests/hostemu) identifies itself in ima2p_version() and is refused, whatever IMA2P_B200_LIB says."""
    global _LIB
    if _LIB is None:
        l = bind()
        v = l.ima2p_version().decode()
        if "sm_100a" not in v:
            raise ImportError("%s is not a CUDA build of ima2p_b200 (%s): there is no CPU path" % (LIB_PATH, v))
        _LIB = l
    return _LIB


def check(l, rc):
    if rc != 0:
        raise Ima2pError(rc, l.ima2p_last_error().decode())
