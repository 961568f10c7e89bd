"""ctypes binding of include/icem_b200.h.  No fallback: a missing library or device raises."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.environ.get("ICEM_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                           "libicem_b200.so")

ICEM_ABI_VERSION = 9
INTEGRATOR = {"euler": 0, "rk4": 1}
DYN = {"dense_tanh": 0, "halfcheetah": 1, "humanoid_standup": 2, "mlp": 3, "articulated": 4}
COST = {"halfcheetah": 0, "humanoid_standup": 1, "locomotion": 2, "reacher": 3, "goal_distance": 4}
REDUCE = {"sum": 0, "best": 1, "final": 2}
PLANNER = {"icem": 0, "cem_std": 1, "random": 2}
UNIQUE_ID_BYTES = 128


class IcemConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("horizon", C.c_int32), ("act_dim", C.c_int32),
        ("num_simulated_trajectories", C.c_int32), ("opt_iterations", C.c_int32), ("elites_size", C.c_int32),
        ("use_mean_actions", C.c_int32), ("keep_previous_elites", C.c_int32),
        ("shift_elites_over_time", C.c_int32), ("cost_along_trajectory", C.c_int32), ("dynamics", C.c_int32),
        ("cost", C.c_int32), ("cost_penalise_flipping", C.c_int32), ("obs_dim", C.c_int32),
        ("colorednoise_v2", C.c_int32), ("keep_iteration_actions", C.c_int32), ("world_size", C.c_int32),
        ("rank", C.c_int32), ("planner", C.c_int32), ("execute_best_elite", C.c_int32), ("shift_means", C.c_int32),
        ("bounds_like_levine", C.c_int32), ("action_change_frequency", C.c_int32),
        ("num_problems", C.c_int32), ("cost_z_index", C.c_int32), ("cost_z_strict", C.c_int32),
        ("cost_velocity_index1", C.c_int32), ("cost_reserved", C.c_int32),
        ("cost_goal_index", C.c_int32), ("cost_achieved_index", C.c_int32), ("cost_goal_sparse", C.c_int32),
        ("cost_goal_shaped", C.c_int32),
        ("factor_decrease_num", C.c_double), ("alpha", C.c_double), ("init_std", C.c_double),
        ("fraction_elites_reused", C.c_double), ("noise_beta", C.c_double),
        ("cost_dt", C.c_double), ("cost_ctrl_weight", C.c_double), ("cost_unhealthy_weight", C.c_double),
        ("cost_z_lo", C.c_double), ("cost_z_hi", C.c_double), ("cost_state_bound", C.c_double),
        ("cost_forward_weight", C.c_double), ("cost_goal_threshold", C.c_double), ("cost_reach", C.c_double * 4),
        ("seed", C.c_uint64),
        ("action_low", C.POINTER(C.c_float)), ("action_high", C.POINTER(C.c_float)),
    ]


class IcemArticulatedModel(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("nb", "nq", "nv", "nu", "nc", "nsub", "obs_offset")]
        + [(n, C.c_float) for n in ("dt", "gravity", "ctrl_limit", "contact_stiffness", "contact_damping",
                                    "contact_damping_max", "friction_viscous", "friction")]
        + [(n, C.POINTER(C.c_int32)) for n in ("body_parent", "body_dof_start", "body_dof_count")]
        + [(n, C.POINTER(C.c_float)) for n in ("body_pos", "body_mass", "body_com", "body_inertia")]
        + [(n, C.POINTER(C.c_int32)) for n in ("dof_body", "dof_type", "dof_qadr", "dof_parent", "dof_limited",
                                               "dof_act")]
        + [(n, C.POINTER(C.c_float)) for n in ("dof_axis", "dof_anchor", "dof_stiffness", "dof_damping",
                                               "dof_armature", "dof_lo", "dof_hi", "dof_klim", "dof_blim", "dof_gear")]
        + [("con_body", C.POINTER(C.c_int32)), ("con_pos", C.POINTER(C.c_float)), ("con_radius", C.POINTER(C.c_float))]
        + [("integrator", C.c_int32)]
    )


class IcemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_F = C.POINTER(C.c_float)
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int32)
_H = C.c_void_p

# every symbol include/icem_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "icem_last_error": (C.c_char_p, []),
    "icem_abi_version": (C.c_int, []),
    "icem_config_sizeof": (C.c_int, []),
    "icem_articulated_model_sizeof": (C.c_int, []),
    "icem_kernel_launch_count": (C.c_uint64, []),
    "icem_create": (C.c_int, [C.POINTER(IcemConfig), C.POINTER(_H)]),
    "icem_destroy": (C.c_int, [_H]),
    "icem_set_dense_model": (C.c_int, [_H, C.c_int32, _F, _F, _F]),
    "icem_set_mlp_model": (C.c_int, [_H, C.c_int32, _I, C.POINTER(_F), C.POINTER(_F)]),
    "icem_set_articulated_model": (C.c_int, [_H, C.POINTER(IcemArticulatedModel)]),
    "icem_begin_rollout": (C.c_int, [_H]),
    "icem_plan": (C.c_int, [_H, _D, C.c_int32, _D]),
    "icem_plan_batch": (C.c_int, [_H, _D, C.c_int32, C.c_int32, _D]),
    "icem_num_problems": (C.c_int, [_H]),
    "icem_set_active_problem": (C.c_int, [_H, C.c_int32]),
    "icem_plan_async": (C.c_int, [_H, _D, C.c_int32]),
    "icem_plan_finish": (C.c_int, [_H, _D]),
    "icem_plan_device": (C.c_int, [_H]),
    "icem_advance_state_device": (C.c_int, [_H]),
    "icem_sync": (C.c_int, [_H]),
    "icem_last_plan_ms": (C.c_int, [_H, _F, _F]),
    "icem_inject_noise": (C.c_int, [_H, C.c_int32, C.c_int32, _F, _F]),
    "icem_get_mean": (C.c_int, [_H, _F]),
    "icem_get_std": (C.c_int, [_H, _F]),
    "icem_num_elites": (C.c_int, [_H]),
    "icem_get_elites": (C.c_int, [_H, _F, _F, _I]),
    "icem_population_size": (C.c_int, [_H, C.c_int32, C.c_int32, _I, _I]),
    "icem_get_iteration": (C.c_int, [_H, C.c_int32, _F, _F, _F, _I]),
    "icem_get_costs": (C.c_int, [_H, C.c_int32, _F, C.c_int32]),
    "icem_get_actions": (C.c_int, [_H, C.c_int32, _F, C.c_int32]),
    "icem_sim_step": (C.c_int, [_H, _D, C.c_int32, _D, _D, _D, C.c_int32, _D]),
    "icem_sim_step_batch": (C.c_int, [_H, C.c_int32, _D, C.c_int32, _D, _D]),
    "icem_state_dim": (C.c_int, [_H]),
    "icem_observe": (C.c_int, [_H, _D, C.c_int32, _D, C.c_int32]),
    "icem_op_sample": (C.c_int, [_H, C.c_int32, _F, _F, _F, _F, _F]),
    "icem_op_rollout_cost": (C.c_int, [_H, C.c_int32, _D, C.c_int32, _F, _F]),
    "icem_op_rollout_observations": (C.c_int, [_H, C.c_int32, _D, C.c_int32, _F, C.c_int32, _D]),
    "icem_op_topk": (C.c_int, [_H, C.c_int32, _F, C.c_int32, _I, _F]),
    "icem_comm_get_unique_id": (C.c_int, [C.c_char_p]),
    "icem_comm_init": (C.c_int, [_H, C.c_char_p]),
    "icem_mlp_trainer_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_H)]),
    "icem_mlp_trainer_destroy": (C.c_int, [_H]),
    "icem_mlp_trainer_set_weights": (C.c_int, [_H, C.POINTER(_F), C.POINTER(_F), C.c_int32]),
    "icem_mlp_trainer_get_weights": (C.c_int, [_H, C.POINTER(_F), C.POINTER(_F)]),
    "icem_mlp_trainer_set_data": (C.c_int, [_H, C.c_int64, _F, _F]),
    "icem_mlp_trainer_fit": (C.c_int, [_H, C.c_int32, C.c_int32, _I, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_float, _F]),
    "icem_mlp_trainer_predict": (C.c_int, [_H, C.c_int32, _F, _F]),
    "icem_bench_device": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, _F, _F, _I]),
    "icem_bench_op": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _F]),
}

_lib = None


def load():
    """dlopen the in-tree library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `python -m icem_b200.build` (needs nvcc). "
                              "icem_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.icem_abi_version() != ICEM_ABI_VERSION:
            raise ImportError("libicem_b200.so ABI version mismatch; rebuild")
        if (lib.icem_config_sizeof() != C.sizeof(IcemConfig)
                or lib.icem_articulated_model_sizeof() != C.sizeof(IcemArticulatedModel)):
            raise ImportError("libicem_b200.so was built from another include/icem_b200.h (struct sizes differ); "
                              "run `python -m icem_b200.build`")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().icem_last_error().decode("utf-8", "replace")
        raise IcemError(rc, msg)


def fptr(a):
    return a.ctypes.data_as(_F)


def dptr(a):
    return a.ctypes.data_as(_D)


def iptr(a):
    return a.ctypes.data_as(_I)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
