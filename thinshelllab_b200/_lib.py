"""ctypes binding of libtsl.so -- signatures transcribed from include/tsl.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtsl.so")


class LibraryMissing(RuntimeError):
    pass


class TslError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtsl error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_int), ("n_verts", C.c_int), ("dt", C.c_double), ("k_contact", C.c_double),
                ("eps_contact", C.c_double), ("eps_v", C.c_double), ("damping", C.c_double), ("gravity", C.c_double * 3),
                ("max_n_constraints", C.c_int), ("grid_h", C.c_double), ("grid_n", C.c_int)]


class StepStatsC(C.Structure):
    _fields_ = [("newton_iters", C.c_int), ("linear_iters", C.c_int), ("linesearch_evals", C.c_int), ("n_contacts", C.c_int),
                ("converged", C.c_int), ("flags", C.c_int), ("delta", C.c_double), ("energy", C.c_double),
                ("ms_contact", C.c_double), ("ms_assembly", C.c_double), ("ms_solve", C.c_double), ("ms_linesearch", C.c_double)]


class SolveStatsC(C.Structure):
    _fields_ = [("iters", C.c_int), ("flags", C.c_int), ("rel_residual", C.c_double)]


class SizesC(C.Structure):
    _fields_ = [("n_verts", C.c_int), ("n_tris", C.c_int), ("n_hinges", C.c_int), ("nnzb", C.c_int), ("nnzb_padded", C.c_int),
                ("n_contacts", C.c_int), ("bytes_matrix_f32", C.c_longlong), ("bytes_matrix_f64", C.c_longlong), ("n_solve", C.c_int), ("nnzb_solve", C.c_int)]


_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/tsl.h declares
SIGNATURES = {
    "tsl_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "tsl_destroy": (_i, [_vp]),
    "tsl_last_error": (C.c_char_p, [_vp]),
    "tsl_version": (C.c_char_p, []),
    "tsl_set_stream": (_i, [_vp, _vp]),
    "tsl_add_cloth": (_i, [_vp, _i, _i, _i, _d, _d, _d, _d, _d, _d, _vp]),
    "tsl_set_cloth_params": (_i, [_vp, _i, _d, _d, _d, _d]),
    "tsl_cloth_update_ref_angle": (_i, [_vp, _i]),
    "tsl_get_cloth_topology": (_i, [_vp, _i, _vp, _vp, _vp]),
    "tsl_set_side_test_override": (_i, [_vp, _i, _vp]),
    "tsl_add_tets": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _d, _d, _d, _vp]),
    "tsl_set_tet_params": (_i, [_vp, _i, _d, _d]),
    "tsl_set_surfaces": (_i, [_vp, _vp, _i, _vp, _i]),
    "tsl_add_contact_pair": (_i, [_vp, _i, _i, _i, _d]),
    "tsl_set_contact_mu": (_i, [_vp, _i, _d]),
    "tsl_bind_state": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsl_finalize": (_i, [_vp]),
    "tsl_reset_contact_state": (_i, [_vp]),
    "tsl_contact_detect": (_i, [_vp, _ip]),
    "tsl_energy": (_i, [_vp, _dp]),
    "tsl_assemble": (_i, [_vp, _i]),
    "tsl_solve": (_i, [_vp, _vp, _vp, _d, _i, C.POINTER(SolveStatsC)]),
    "tsl_step_forward": (_i, [_vp, _i, _d, C.POINTER(StepStatsC)]),
    "tsl_step_forward_host": (_i, [_vp, _vp, _vp, _i, _d, C.POINTER(StepStatsC)]),
    "tsl_step_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, C.POINTER(SolveStatsC)]),
    "tsl_step_backward_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _i, C.POINTER(SolveStatsC)]),
    "tsl_elastic_param_grad": (_i, [_vp, _vp, _vp, _vp, _dp]),
    "tsl_cloth_param_deri": (_i, [_vp, _i, _vp, _vp, _vp]),
    "tsl_friction_coef_grad": (_i, [_vp, _vp, _i, _i, _dp]),
    "tsl_elastic_force": (_i, [_vp, _i, _vp]),
    "tsl_gripper_apply": (_i, [_vp, _i, _i, _vp, _vp, _dp, C.POINTER(C.c_float)]),
    "tsl_gripper_gather": (_i, [_vp, _vp, _i, _i, _vp, _vp, C.POINTER(C.c_float), _d, _d, _dp]),
    "tsl_get_contact_blocks": (_i, [_vp, _ip, _vp, _vp, _vp]),
    "tsl_dist_unique_id": (_i, [_vp]),
    "tsl_dist_init": (_i, [_vp, _vp, _i, _i, _i, _i, _i]),
    "tsl_dist_stats": (_i, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "tsl_get_residual": (_i, [_vp, _vp]),
    "tsl_get_matrix_nnzb": (_i, [_vp, _ip]),
    "tsl_get_matrix": (_i, [_vp, _vp, _vp, _vp]),
    "tsl_get_projection": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "tsl_get_constraints": (_i, [_vp, _ip, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsl_get_sizes": (_i, [_vp, C.POINTER(SizesC)]),
    "tsl_bench_kernel": (_i, [_vp, _i, _i, C.POINTER(C.c_float)]),
    "tsl_set_option": (_i, [_vp, _i, _d]),
    "tsl_mg_get_level": (_i, [_vp, _i, _vp, _vp, _vp]),
    "tsl_precond_apply": (_i, [_vp, _vp, _vp]),
    "tsl_dense_solve_host": (_i, [_vp, _i, _vp, _vp, _vp]),
    "tsl_launch_count": (C.c_longlong, [_vp]),
}

OPT_PRECOND, OPT_MG_DEGREE, OPT_MG_COARSE_DEGREE, OPT_MG_RATIO, OPT_MG_SAFETY, OPT_GRAPHS, OPT_NEWTON_MODE = 0, 1, 2, 3, 4, 5, 6
OPT_ADJOINT_SOLVER, OPT_DIRECT_MAX_DOF, OPT_GMRES_M, OPT_FAST_ASSEMBLY = 7, 8, 9, 10
ADJ_AUTO, ADJ_DIRECT, ADJ_FGMRES, ADJ_BICGSTAB = 0, 1, 2, 3
ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_NUMERIC = -1, -2, -3, -4, -5
ASM_RESIDUAL, ASM_HESSIAN, ASM_SPD, ASM_SYM, ASM_F64, ASM_NEWTON = 1, 2, 4, 8, 16, 32

_LIB = None


def lib():
    """Loads libtsl.so (built in-tree by __graft_entry__.build() / csrc/build.sh).  Raises if it is missing:
    the product has no other execution path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(f"{LIB_PATH} not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(nvcc, sm_100a).  thinshelllab_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB
