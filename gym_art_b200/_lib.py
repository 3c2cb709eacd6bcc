"""ctypes binding of libquadsim.so (include/quadsim.h).

This is the stub a maintainer of the reference would add to call the B200 path from Python: the reference has
no FFI of its own (its seam is duck typing, quadrotor.py:966-970 / :261), so the binding mirrors the header
one-to-one.  There is NO CPU fallback: if the shared object is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QUADSIM_LIB") or os.path.join(HERE, "libquadsim.so")   # override: A/B builds only

QS_ABI_VERSION = 8
QS_MODELS_PLAIN = 1
QS_OK, QS_EINVAL, QS_ECUDA, QS_ENONFINITE, QS_ESTATE, QS_EUNSUPPORTED, QS_EABI = 0, -1, -2, -3, -4, -5, -6
QS_F32, QS_F64 = 0, 1
(QS_CTRL_RAW, QS_CTRL_VERTICAL, QS_CTRL_VERTPLANE, QS_CTRL_SHIFTED, QS_CTRL_OMEGA_THRUST, QS_CTRL_VELOCITY_YAW,
 QS_CTRL_MELLINGER) = range(7)
QS_NOISE_PHILOX, QS_NOISE_INJECTED = 0, 1
QS_OBS_Z, QS_RESET_Z, QS_INFO_DIM, QS_MODEL_ROWS, QS_GOAL_Z = 33, 51, 48, 14, 4
QS_INIT_DEFAULT, QS_INIT_RANDOM, QS_INIT_PITCH_ROLL_RESTRICTED = 0, 1, 2
(QS_INFO_ACC, QS_INFO_ACCELEROMETER, QS_INFO_TORQUE, QS_INFO_OMEGA_DOT, QS_INFO_CMDS, QS_INFO_CRASHED, QS_INFO_COLLIDED,
 QS_INFO_COSTS, QS_INFO_EP_RETURN, QS_INFO_TICK, QS_INFO_REWARD, QS_INFO_DONE, QS_INFO_COSTS_RAW) = (
    0, 3, 6, 9, 12, 16, 17, 18, 28, 29, 30, 31, 32)

OBS_REPRS = ["xyz_vxyz_R_omega", "xyz_vxyz_R_omega_h", "xyzr_vxyzr_R_omega", "xyzr_vxyzr_R_omega_h",
             "xyz_vxyz_R_omega_acc_act", "xyz_vxyz_R_omega_act",
             # get_state.py:276-382: run in the reference once the module is given the names it forgot to import
             "xyz_vxyz_quat_omega", "xyzr_vxyzr_quat_omega", "xyzr_vxyzr_quat_omega_h", "xyzr_vxyzr_R_omega_t2w",
             "xyz_vxyz_R_omega_t2w", "xyz_vxyz_R_omega_t2w_t2t"]

STATUS_NAMES = {QS_EINVAL: "QS_EINVAL", QS_ECUDA: "QS_ECUDA", QS_ENONFINITE: "QS_ENONFINITE", QS_ESTATE: "QS_ESTATE",
                QS_EUNSUPPORTED: "QS_EUNSUPPORTED", QS_EABI: "QS_EABI"}

# every symbol include/quadsim.h declares (tests/test_abi.py checks the library exports each one)
EXPORTS = ["qs_abi_version", "qs_state_layout", "qs_create", "qs_destroy", "qs_last_error", "qs_bind_state", "qs_bind_models", "qs_bind_terminal_obs", "qs_bind_info",
           "qs_set_noise_injection", "qs_set_goal_injection", "qs_reset", "qs_observe", "qs_step", "qs_step_k", "qs_dynamics_step", "qs_affine_fg", "qs_step_host", "qs_get_stats",
           "qs_stats_device_ptr", "qs_check", "qs_get_clock", "qs_set_clock", "qs_launch_count", "qs_last_variant", "qs_last_host_path"]


class QsModel(C.Structure):
    _fields_ = [("mass", C.c_double), ("inertia", C.c_double * 3), ("thrust_max", C.c_double * 4),
                ("torque_max", C.c_double * 4), ("prop_pos", (C.c_double * 3) * 4), ("prop_ccw", C.c_double * 4),
                ("linearity", C.c_double), ("C_drag", C.c_double), ("C_roll", C.c_double),
                ("damp_time_up", C.c_double), ("damp_time_down", C.c_double),
                ("vel_damp", C.c_double), ("damp_omega_quadratic", C.c_double),
                ("ou_theta", C.c_double), ("ou_sigma", C.c_double), ("arm", C.c_double), ("omega_max", C.c_double),
                ("gravity", C.c_double), ("thrust_to_weight", C.c_double), ("torque_to_thrust", C.c_double), ("Jinv", (C.c_double * 4) * 4)]


class QsReward(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("pos", "effort", "action_change", "crash", "orient", "yaw", "rot", "attitude",
                                           "spin", "vel", "pos_offset", "pos_log_weight", "pos_linear_weight",
                                           "collision")] + [("log_pos", C.c_int32), ("_pad", C.c_int32)]


class QsSensorNoise(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("_pad", C.c_int32)] + [
        (k, C.c_double) for k in ("pos_norm_std", "pos_unif_range", "vel_norm_std", "vel_unif_range", "quat_norm_std",
                                  "quat_unif_range", "gyro_norm_std", "gyro_noise_density", "gyro_random_walk",
                                  "gyro_bias_correlation_time", "acc_static_noise_std", "acc_dynamic_noise_ratio")]


class QsConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("struct_size", C.c_int32), ("n_envs", C.c_int64),
                ("env_id_offset", C.c_int64), ("seed", C.c_uint64), ("n_agents", C.c_int32), ("device", C.c_int32),
                ("dtype", C.c_int32), ("controller", C.c_int32), ("zero_middle", C.c_int32), ("obs_repr", C.c_int32),
                ("noise_mode", C.c_int32), ("init_random_state", C.c_int32), ("auto_reset", C.c_int32),
                ("dim_mode", C.c_int32), ("sim_steps", C.c_int32), ("ep_len", C.c_int32), ("per_env_goal", C.c_int32),
                ("resample_goal", C.c_int32), ("excite", C.c_int32), ("_pad", C.c_int32), ("dt", C.c_double), ("svd_period", C.c_double), ("goal", C.c_double * 3),
                ("spawn_box", C.c_double), ("spawn_z_min", C.c_double), ("room_box", (C.c_double * 3) * 2), ("max_init_vel", C.c_double),
                ("max_init_omega", C.c_double), ("collision_k", C.c_double), ("pitch_max", C.c_double), ("roll_max", C.c_double),
                ("yaw_max", C.c_double), ("t2w_std", C.c_double), ("t2t_std", C.c_double), ("use_formation", C.c_int32), ("per_env_model", C.c_int32),
                ("formation", (C.c_double * 3) * 32), ("model", QsModel), ("reward", QsReward),
                ("sense", QsSensorNoise)]


class QsStateLayout(C.Structure):
    _fields_ = [("rows", C.c_int32), ("row_ou", C.c_int32), ("row_prev_action", C.c_int32),
                ("row_gyro_bias", C.c_int32), ("row_goal", C.c_int32), ("obs_dim", C.c_int32), ("act_dim", C.c_int32),
                ("_pad", C.c_int32)]


class QuadsimError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "%s (%d): %s" % (STATUS_NAMES.get(code, "QS_?"), code, msg))
        self.code = code


_lib = None


def load():
    """dlopen the in-tree library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libquadsim.so is not built: run `python -m gym_art_b200._build` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    P, H = C.c_void_p, C.c_void_p
    sig = {
        "qs_abi_version": (C.c_int, []),
        "qs_state_layout": (C.c_int, [C.POINTER(QsConfig), C.POINTER(QsStateLayout)]),
        "qs_create": (C.c_int, [C.POINTER(QsConfig), C.POINTER(H)]),
        "qs_destroy": (C.c_int, [H]),
        "qs_last_error": (C.c_char_p, [H]),
        "qs_bind_state": (C.c_int, [H, P]),
        "qs_bind_info": (C.c_int, [H, P]),
        "qs_bind_terminal_obs": (C.c_int, [H, P]),
        "qs_bind_models": (C.c_int, [H, P, C.c_int32]),
        "qs_set_noise_injection": (C.c_int, [H, P, P, P]),
        "qs_set_goal_injection": (C.c_int, [H, P]),
        "qs_reset": (C.c_int, [H, P, P, P]),
        "qs_observe": (C.c_int, [H, P, P, P]),
        "qs_step": (C.c_int, [H, P, P, P, P, P]),
        "qs_step_k": (C.c_int, [H, C.c_int, P, P, P, P, P]),
        "qs_dynamics_step": (C.c_int, [H, C.c_int, P, C.c_int, P]),
        "qs_affine_fg": (C.c_int, [H, P, C.c_double, P, P, C.c_int64, P]),
        "qs_step_host": (C.c_int, [H, P, P, P, P, P]),
        "qs_get_stats": (C.c_int, [H, C.POINTER(C.c_double * 3), C.c_int, P]),
        "qs_stats_device_ptr": (C.c_int, [H, C.POINTER(C.c_void_p)]),
        "qs_check": (C.c_int, [H, P]),
        "qs_get_clock": (C.c_int, [H, C.POINTER(C.c_double), C.POINTER(C.c_uint64), P]),
        "qs_set_clock": (C.c_int, [H, C.c_double, C.c_uint64, P]),
        "qs_launch_count": (C.c_int, [H, C.POINTER(C.c_uint64)]),
        "qs_last_variant": (C.c_char_p, [H]),
        "qs_last_host_path": (C.c_char_p, [H]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.qs_abi_version() != QS_ABI_VERSION:
        raise ImportError("libquadsim.so ABI %d != binding ABI %d: rebuild" % (lib.qs_abi_version(), QS_ABI_VERSION))
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != QS_OK:
        msg = load().qs_last_error(handle)
        raise QuadsimError(rc, msg.decode() if msg else "")
    return rc


def state_layout(cfg):
    lay = QsStateLayout()
    check(load().qs_state_layout(C.byref(cfg), C.byref(lay)))
    return lay
