"""ctypes binding of libfedem_b200.so (the C ABI declared in include/fedem_b200.h).

The library is the product; nothing here computes.  Loading fails loudly when the CUDA
extension has not been built -- there is deliberately no CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class FsrError(RuntimeError):
    pass


def library_path():
    return os.path.join(_HERE, "lib", "libfedem_b200.so")


class FsrSam(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nnod", "nel", "ndof", "ndof1", "ndof2", "ngen", "neq", "nceq",
                                        "nmmnpc", "nmmceq")] + [
        ("madof", C.POINTER(C.c_int)), ("msc", C.POINTER(C.c_int)), ("mpmnpc", C.POINTER(C.c_int)),
        ("mmnpc", C.POINTER(C.c_int)), ("melcon", C.POINTER(C.c_int)), ("mpmceq", C.POINTER(C.c_int)),
        ("mmceq", C.POINTER(C.c_int)), ("ttcc", C.POINTER(C.c_double)), ("meqn", C.POINTER(C.c_int)),
        ("meqn1", C.POINTER(C.c_int)), ("meqn2", C.POINTER(C.c_int))]


class FsrElmData(C.Structure):
    _fields_ = [("xyz", C.POINTER(C.c_double)), ("emod", C.POINTER(C.c_double)),
                ("rny", C.POINTER(C.c_double)), ("thk", C.POINTER(C.c_double)),
                ("elmid", C.POINTER(C.c_int)), ("beam", C.POINTER(C.c_double))]


class FsrRosette(C.Structure):
    _fields_ = [("id", C.c_int), ("numnod", C.c_int), ("ngage", C.c_int), ("zero_init", C.c_int),
                ("nodes", C.c_int * 4), ("rpos", C.c_double * 12), ("zpos", C.c_double), ("emod", C.c_double),
                ("nu", C.c_double), ("alpha_gages", C.c_double), ("gate", C.c_double), ("sncurve", C.c_double * 4)]


class FsrRdbOptions(C.Structure):
    _fields_ = [("out_mask", C.c_uint), ("double_precision", C.c_int), ("rdbinc", C.c_int), ("part_base_id", C.c_int),
                ("part_user_id", C.c_int), ("part_descr", C.c_char_p), ("model_file", C.c_char_p),
                ("link_file", C.c_char_p), ("elmid", C.POINTER(C.c_int)), ("module_name", C.c_char_p),
                ("minex", C.POINTER(C.c_int)), ("sup_tr_init", C.POINTER(C.c_double))]


class FsrStrainCoat(C.Structure):
    _fields_ = [("id", C.c_int), ("nnod", C.c_int), ("npts", C.c_int), ("elm_id", C.c_int), ("nodes", C.c_int * 8),
                ("mat_id", C.c_int * 3), ("res_set", C.c_int * 3), ("sn_curve", (C.c_int * 2) * 3),
                ("emod", C.c_double * 3), ("nu", C.c_double * 3), ("zpos", C.c_double * 3), ("scf", C.c_double * 3)]


class FsrOptions(C.Structure):
    _fields_ = [("device", C.c_int), ("stressForm", C.c_int), ("step_tile", C.c_int),
                ("reserved", C.c_int * 5)]


# every symbol include/fedem_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
SYMBOLS = [
    ("fsr_part_create", C.c_int, [C.POINTER(_P), C.POINTER(FsrSam), C.POINTER(FsrElmData), C.POINTER(FsrOptions)]),
    ("fsr_set_recovery", C.c_int, [_P, _D, C.c_int, _D, C.c_int]),
    ("fsr_part_destroy", None, [_P]),
    ("fsr_set_stream", C.c_int, [_P, _P]),
    ("fsr_num_result_points", C.c_int, [_P]),
    ("fsr_result_point_offsets", C.c_int, [_P, _I]),
    ("fsr_ndim", C.c_int, [_P]),
    ("fsr_recover", C.c_int, [_P, _D, C.c_int, C.c_int, _D]),
    ("fsr_recover_dev", C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    ("fsr_split_elements", C.c_int, [C.POINTER(FsrSam), C.POINTER(FsrElmData), C.c_int, _I]),
    ("fsr_blockdef_create", C.c_int, [C.POINTER(_P), C.POINTER(FsrSam), C.POINTER(FsrElmData), C.POINTER(FsrOptions), C.c_int, C.c_int]),
    ("fsr_blockdef_sam", C.POINTER(FsrSam), [_P]),
    ("fsr_blockdef_elm", C.POINTER(FsrElmData), [_P]),
    ("fsr_blockdef_info", C.c_int, [_P, _I, _I, _I]),
    ("fsr_blockdef_destroy", None, [_P]),
    ("fsr_part_create_block", C.c_int, [C.POINTER(_P), C.POINTER(FsrSam), C.POINTER(FsrElmData), C.POINTER(FsrOptions), C.c_int, C.c_int]),
    ("fsr_block_info", C.c_int, [_P, _I]),
    ("fsr_block_rows", C.c_int, [_P, _I, _I]),
    ("fsr_set_recovery_parent", C.c_int, [_P, _D, C.c_int, _D, C.c_int]),
    ("fsr_group_create", C.c_int, [C.POINTER(_P), C.POINTER(FsrSam), C.POINTER(FsrElmData), C.POINTER(FsrOptions), _I, C.c_int]),
    ("fsr_group_set_recovery", C.c_int, [_P, _D, C.c_int, _D, C.c_int]),
    ("fsr_group_num_blocks", C.c_int, [_P]),
    ("fsr_group_num_result_points", C.c_int, [_P]),
    ("fsr_group_ndim", C.c_int, [_P]),
    ("fsr_group_block", _P, [_P, C.c_int]),
    ("fsr_group_recover", C.c_int, [_P, _D, C.c_int, C.c_int, _D]),
    ("fsr_group_synchronize", C.c_int, [_P]),
    ("fsr_group_reset_envelope", C.c_int, [_P]),
    ("fsr_group_get_envelope", C.c_int, [_P, _D, _D]),
    ("fsr_group_last_timing", C.c_int, [_P, _D, C.c_int]),
    ("fsr_group_timing_reset", C.c_int, [_P]),
    ("fsr_group_destroy", None, [_P]),
    ("fsr_comm_unique_id", C.c_int, [C.c_char_p, C.c_int]),
    ("fsr_comm_init_rank", C.c_int, [C.POINTER(_P), C.c_char_p, C.c_int, C.c_int, C.c_int]),
    ("fsr_comm_broadcast", C.c_int, [_P, _P, C.c_longlong, C.c_int, _P]),
    ("fsr_comm_gather_envelope", C.c_int, [_P, _P, _I, _I, _P, _P, C.c_int, _P]),
    ("fsr_comm_destroy", None, [_P]),
    ("fsr_nccl_version", C.c_int, []),
    ("fsr_recover_async", C.c_int, [_P, _D, C.c_int, C.c_int]),
    ("fsr_get_envelope_async", C.c_int, [_P, _D, _D]),
    ("fsr_envelope_wait", C.c_int, [_P]),
    ("fsr_synchronize", C.c_int, [_P]),
    ("fsr_recovery_update_parts", C.c_int, [C.c_int, _I, C.c_int, C.c_double, C.c_double, C.POINTER(_D)]),
    ("fsr_family_counts", C.c_int, [_P, _I, C.c_int]),
    ("fsr_vm_path_info", C.c_int, [_P, C.POINTER(C.c_longlong), C.c_int]),
    ("fsr_recover_displacements", C.c_int, [_P, _D, C.c_int, _D]),
    ("fsr_rdb_write_steps_displacements", C.c_int, [_P, _D, C.c_int, _I, _D, _D]),
    ("fsr_reset_envelope", C.c_int, [_P]),
    ("fsr_get_envelope", C.c_int, [_P, _D, _D]),
    ("fsr_envelope_dev", C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    ("fsr_copy_envelope_dev", C.c_int, [_P, _P, _P, _P]),
    ("fsr_recover_step_full", C.c_int, [_P, _D, _D, _D, _D, _D, _D]),
    ("fsr_expand", C.c_int, [_P, _D, C.c_int, C.c_int, _D]),
    ("fsr_expand_rows", C.c_int, [_P, _D, C.c_int, C.c_int, _I, C.c_int, _D]),
    ("fsr_fatigue", C.c_int, [C.c_int, _D, C.c_int, C.c_int, C.c_double, _D, C.c_double, C.c_int, _D, _I, _I]),
    ("fsr_fatigue_dev", C.c_int, [C.c_int, _P, C.c_size_t, C.c_int, C.c_int, C.c_double, _D, C.c_double,
                                  C.c_int, _P, _P, _P, _P]),
    ("fsr_fatigue_create", C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_double, _D, C.c_double, C.c_int, C.c_int]),
    ("fsr_fatigue_set_gage_params", C.c_int, [_P, _D, _D]),
    ("fsr_fatigue_reset", C.c_int, [_P]),
    ("fsr_fatigue_locate_dev", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int, C.c_int, _I, _P]),
    ("fsr_fatigue_feed_dev", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int, C.c_int, _P]),
    ("fsr_fatigue_finish", C.c_int, [_P, _D, _I, _I, _I]),
    ("fsr_fatigue_finish_dev", C.c_int, [_P, _P]),
    ("fsr_fatigue_results_dev", C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    ("fsr_fatigue_destroy", None, [_P]),
    ("fsr_vms_size", C.c_int, [_P]),
    ("fsr_get_vms", C.c_int, [_P, _D, _D, C.c_int]),
    ("fsr_gage_create", C.c_int, [C.POINTER(_P), _P, C.POINTER(FsrRosette), C.c_int]),
    ("fsr_gage_num_series", C.c_int, [_P]),
    ("fsr_gage_get_bcart", C.c_int, [_P, _D]),
    ("fsr_gage_recover", C.c_int, [_P, _D, C.c_int, C.c_int, _D]),
    ("fsr_gage_recover_dev", C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    ("fsr_gage_fatigue", C.c_int, [_P, _D, C.c_int, C.c_int, C.c_double, C.c_double, _D, C.c_double, C.c_int,
                                   _D, _I, _I, _I]),
    ("fsr_gage_fatigue_begin", C.c_int, [_P, C.c_double, C.c_double, _D, C.c_double, C.c_int, C.c_int]),
    ("fsr_gage_fatigue_feed_dev", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _I, _P]),
    ("fsr_gage_fatigue_end", C.c_int, [_P, _D, _I, _I, _I]),
    ("fsr_gage_destroy", None, [_P]),
    ("fsr_coat_begin", C.c_int, [_P, C.c_int, C.c_double]),
    ("fsr_coat_feed", C.c_int, [_P, _D, C.c_int, C.c_int]),
    ("fsr_coat_feed_dev", C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    ("fsr_coat_end", C.c_int, [_P, _D, _D, _I]),
    ("fsr_fmx_write", C.c_int, [C.c_char_p, C.c_char_p, C.c_int, _D, C.c_longlong, C.c_int]),
    ("fsr_fmx_read", C.c_int, [C.c_char_p, C.c_char_p, C.c_int, _I, _I, _D, C.c_longlong]),
    ("fsr_fsm_read_mpar", C.c_int, [C.c_char_p, _I, _I, C.c_int]),
    ("fsr_fsm_read", C.c_int, [C.c_char_p, _I, _I, _I, _I, _I, _I, _I, _I, _I, _D, _I, _I, _I]),
    ("fsr_fsm_write", C.c_int, [C.c_char_p, C.c_int, C.c_int, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _D, _I, _I, _I]),
    ("fsr_build_finit", C.c_int, [C.c_int, C.c_int, _D, _D, _D, _I, _I, C.c_int, _D, C.c_int, _D, C.c_int]),
    ("fsr_build_mode_finit", C.c_int, [C.c_int, _D, _I, _I, _D, C.c_int, C.c_int, _D, C.c_int, _D, C.c_int]),
    ("fsr_frs_open", C.c_int, [C.POINTER(_P), C.POINTER(C.c_char_p), C.c_int]),
    ("fsr_frs_close", None, [_P]),
    ("fsr_frs_num_steps", C.c_int, [_P]),
    ("fsr_frs_get_steps", C.c_int, [_P, _I, _D, C.c_int]),
    ("fsr_frs_find", C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_int]),
    ("fsr_frs_var_size", C.c_int, [_P, C.c_int]),
    ("fsr_frs_read", C.c_int, [_P, C.c_int, C.c_int, C.c_int, _D, C.c_int, C.c_int]),
    ("fsr_frs_reduced_history", C.c_int, [_P, C.c_int, C.c_int, _I, _I, _I, _D, C.c_int, C.c_int, C.c_int, C.c_int,
                                          _D, C.c_int]),
    ("fsr_frs_create", C.c_int, [C.POINTER(_P), C.c_char_p, C.c_int, C.c_char_p, C.c_longlong]),
    ("fsr_frs_create_tagged", C.c_int, [C.POINTER(_P), C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_longlong]),
    ("fsr_frs_write_step", C.c_int, [_P, C.c_int, C.c_double, _P]),
    ("fsr_frs_finish", C.c_int, [_P]),
    ("fsr_ftl_open", C.c_int, [C.POINTER(_P), C.c_char_p]),
    ("fsr_ftl_close", None, [_P]),
    ("fsr_ftl_version", C.c_int, [_P]),
    ("fsr_ftl_activate_groups", C.c_int, [_P, C.c_char_p]),
    ("fsr_ftl_sizes", C.c_int, [_P, _I]),
    ("fsr_ftl_get_nodes", C.c_int, [_P, _I, _I, _I, _I, _D]),
    ("fsr_ftl_get_topology", C.c_int, [_P, C.c_int, _I, _I, _I]),
    ("fsr_ftl_get_elmdata", C.c_int, [_P, _D, _D, _D, _D, _I, _D, _I]),
    ("fsr_ftl_ext2int", C.c_int, [_P, C.c_int, C.c_int]),
    ("fsr_ftl_num_strain_coats", C.c_int, [_P]),
    ("fsr_ftl_get_strain_coats", C.c_int, [_P, C.POINTER(FsrStrainCoat), C.c_int]),
    ("fsr_gage_set_coat_fatigue", C.c_int, [_P, _D]),
    ("fsr_sn_read", C.c_int, [C.POINTER(_P), C.c_char_p]),
    ("fsr_sn_free", None, [_P]),
    ("fsr_sn_num_standards", C.c_int, [_P]),
    ("fsr_sn_num_curves", C.c_int, [_P, C.c_int]),
    ("fsr_sn_get", C.c_int, [_P, C.c_int, C.c_int, _I, _D, _D, _D, C.c_int]),
    ("fsr_sn_value", C.c_double, [_P, C.c_int, C.c_int, C.c_double]),
    ("fsr_fsi_open", C.c_int, [C.POINTER(_P), C.c_char_p, C.c_int]),
    ("fsr_fsi_close", None, [_P]),
    ("fsr_fsi_part", C.c_int, [_P, _I, C.c_char_p, C.c_int, _I, _I, _D, _D, C.c_char_p, C.c_int]),
    ("fsr_fsi_triads", C.c_int, [_P, _I, _I, _I, _I, _D, _D]),
    ("fsr_fsi_read_rosettes", C.c_int, [C.c_char_p, C.c_int, C.POINTER(FsrRosette), _I, C.c_char_p, C.c_int, C.c_int]),
    ("fsr_rdb_create", C.c_int, [C.POINTER(_P), _P, C.c_char_p, C.POINTER(FsrRdbOptions)]),
    ("fsr_rdb_create_group", C.c_int, [C.POINTER(_P), _P, C.c_char_p, C.POINTER(FsrRdbOptions)]),
    ("fsr_rdb_build_header", C.c_int, [C.c_int, _I, C.c_int, _I, C.POINTER(FsrRdbOptions), C.c_char_p, C.c_int, C.POINTER(C.c_longlong)]),
    ("fsr_rdb_step_bytes", C.c_longlong, [_P]),
    ("fsr_rdb_header", C.c_int, [_P, C.c_char_p, C.c_int]),
    ("fsr_rdb_path", C.c_int, [_P, C.c_char_p, C.c_int]),
    ("fsr_rdb_write_steps", C.c_int, [_P, _D, C.c_int, C.c_int, _I, _D, _D]),
    ("fsr_total_nodal_displacement", None, [_D, _D, C.c_int, _D, _D, _D]),
    ("fsr_rdb_flush", C.c_int, [_P, _D, C.c_int]),
    ("fsr_rdb_close", C.c_int, [_P]),
    ("initSolverArgs", None, [C.c_int, C.POINTER(C.c_char_p)]),
    ("solveStress", C.c_int, []),
    ("solveGage", C.c_int, []),
    ("solveModes", C.c_int, []),
    ("solveFpp", C.c_int, []),
    ("fsr_fpp_define_options", None, []),
    ("fsr_modes_define_options", None, []),
    ("fsr_gage_define_options", None, []),
    ("fsr_select_steps", C.c_int, [_D, C.c_int, C.c_double, C.c_double, C.c_double, _I, C.c_int]),
    ("fsr_cmdline_reset", None, []),
    ("fsr_stress_define_options", None, []),
    ("fsr_cmdline_add_bool", None, [C.c_char_p, C.c_int]),
    ("fsr_cmdline_add_int", None, [C.c_char_p, C.c_int]),
    ("fsr_cmdline_add_double", None, [C.c_char_p, C.c_double]),
    ("fsr_cmdline_add_string", None, [C.c_char_p, C.c_char_p]),
    ("fsr_cmdline_init", None, [C.c_int, C.POINTER(C.c_char_p)]),
    ("fsr_cmdline_read_file", C.c_int, [C.c_char_p]),
    ("fsr_cmdline_get_bool", C.c_int, [C.c_char_p]),
    ("fsr_cmdline_get_int", C.c_int, [C.c_char_p]),
    ("fsr_cmdline_get_double", C.c_double, [C.c_char_p]),
    ("fsr_cmdline_get_string", C.c_int, [C.c_char_p, C.c_char_p, C.c_int]),
    ("fsr_cmdline_is_set", C.c_int, [C.c_char_p]),
    ("fsr_recovery_register", C.c_int, [C.c_int, _P, _I]),
    ("fsr_recovery_unregister", C.c_int, [C.c_int]),
    ("fsr_recovery_options", C.c_int, [C.c_char_p]),
    ("fsr_recovery_register_part", C.c_int, [C.c_int, C.c_int, C.c_char_p, _P, _I, _D]),
    ("fsr_recovery_update_parts_save", C.c_int, [C.c_int, _I, C.c_int, C.c_double, C.c_double, C.POINTER(_D), C.POINTER(_D), C.c_int]),
    ("fsr_recovery_close", C.c_int, []),
    ("fsr_recovery_file", C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    ("fsr_recovery_update", C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, _D]),
    ("getPartDeformationStateSize", C.c_int, [C.c_int]),
    ("getPartStressStateSize", C.c_int, [C.c_int]),
    ("savePartDeformationState", C.c_bool, [C.c_int, _D, C.c_int]),
    ("savePartStressState", C.c_bool, [C.c_int, _D, C.c_int]),
    ("fsr_last_error", C.c_char_p, []),
    ("fsr_kernel_launches", C.c_longlong, [C.c_int]),
    ("fsr_last_timing", C.c_int, [_P, _D, C.c_int]),
    ("fsr_timing_reset", C.c_int, [_P]),
]


def load_library():
    """Loads libfedem_b200.so and declares all prototypes.  Raises FsrError if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise FsrError(f"{path} not found: build the CUDA extension first (./build.sh or "
                       "__graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc, what):
    """Raises on fatal (<0) return codes; returns the warning count otherwise."""
    if rc < 0:
        msg = load_library().fsr_last_error().decode(errors="replace")
        raise FsrError(f"{what} failed (code {rc}): {msg}")
    return rc
