"""ctypes loader for libpmg.so, the CUDA kernels + C-ABI declared in include/pmg.h.

There is no CPU fallback: if the shared library is missing or cannot be loaded the import
fails loudly.  Build it with `python -m pybullet_multigoal_gym_b200.build` (or
`__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PMG_LIBRARY") or os.path.join(_HERE, "libpmg.so")  # PMG_LIBRARY: instrumented development builds

ABI_VERSION = 6

# every symbol include/pmg.h declares
SYMBOLS = [
    "pmg_abi_version", "pmg_last_error", "pmg_create", "pmg_destroy", "pmg_dims", "pmg_seed",
    "pmg_reset", "pmg_set_device_rng", "pmg_reset_device", "pmg_set_auto_reset", "pmg_spawn_width", "pmg_last_spawn", "pmg_set_curriculum_update", "pmg_get_curriculum", "pmg_set_sub_goal", "pmg_step", "pmg_gather_create", "pmg_gather_connect", "pmg_gather_layout", "pmg_step_gather", "pmg_step_host", "pmg_step_host_blocks",
    "pmg_compute_reward", "pmg_her_sample", "pmg_her_relabel", "pmg_state_width", "pmg_get_state", "pmg_set_state",
    "pmg_kernel_timing", "pmg_kernel_time_ms", "pmg_launch_count", "pmg_overflow_count", "pmg_debug_box_box", "pmg_action_error",
]


class PmgConfig(C.Structure):
    _fields_ = [("task", C.c_int32), ("num_block", C.c_int32), ("batch", C.c_int32),
                ("binary_reward", C.c_int32), ("distance_threshold", C.c_float),
                ("max_episode_steps", C.c_int32), ("device", C.c_int32),
                ("grip_informed_goal", C.c_int32), ("joint_control", C.c_int32), ("task_decomposition", C.c_int32),
                ("use_curriculum", C.c_int32), ("num_goals_to_generate", C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "pybullet_multigoal_gym_b200: %s is missing. This package has no CPU fallback; build the "
            "CUDA extension with `python -m pybullet_multigoal_gym_b200.build`." % SO_PATH)
    L = C.CDLL(SO_PATH)
    missing = [s for s in SYMBOLS if not hasattr(L, s)]
    if missing:
        raise ImportError("libpmg.so does not export: %s" % ", ".join(missing))
    vp, fp, u8p = C.c_void_p, C.c_void_p, C.c_void_p
    L.pmg_abi_version.restype = C.c_int
    L.pmg_last_error.restype = C.c_char_p
    L.pmg_create.argtypes = [C.POINTER(PmgConfig), C.POINTER(vp)]
    L.pmg_destroy.argtypes = [vp]
    L.pmg_dims.argtypes = [vp, C.POINTER(C.c_int32)]
    L.pmg_seed.argtypes = [vp, vp, vp, C.c_int32]
    L.pmg_reset.argtypes = [vp, u8p, fp, fp, vp]
    L.pmg_set_device_rng.argtypes = [vp, C.c_uint64, C.c_int64]
    L.pmg_reset_device.argtypes = [vp, u8p, fp, vp]
    L.pmg_set_auto_reset.argtypes = [vp, C.c_int32, fp]
    L.pmg_spawn_width.argtypes = [vp]
    L.pmg_last_spawn.argtypes = [vp, fp]
    L.pmg_set_curriculum_update.argtypes = [vp, C.c_int32]
    L.pmg_get_curriculum.argtypes = [vp, fp, vp]
    L.pmg_set_sub_goal.argtypes = [vp, vp, vp]
    L.pmg_step.argtypes = [vp, fp, fp, fp, u8p, u8p, vp]
    L.pmg_gather_create.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.pmg_gather_connect.argtypes = [vp, vp]
    L.pmg_gather_layout.argtypes = [vp, C.POINTER(C.c_int64)]
    L.pmg_step_gather.argtypes = [vp, fp, C.POINTER(vp), vp]
    L.pmg_step_host.argtypes = [vp, fp, fp, fp, u8p, u8p, vp]
    L.pmg_debug_box_box.argtypes = [fp, C.c_int64, C.c_int32, fp, C.c_int32]
    L.pmg_action_error.argtypes = [vp, C.c_int32]
    L.pmg_step_host_blocks.argtypes = [vp, fp, fp, fp, u8p, u8p, vp]
    L.pmg_compute_reward.argtypes = [fp, fp, C.c_int64, C.c_int32, C.c_float, C.c_int32, fp, u8p, vp]
    L.pmg_her_sample.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_uint64, vp, vp, vp, vp]
    L.pmg_her_relabel.argtypes = [fp, fp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, C.c_int64, C.c_float, C.c_int32,
                                  fp, fp, u8p, vp]
    L.pmg_state_width.argtypes = [vp]
    L.pmg_get_state.argtypes = [vp, fp]
    L.pmg_set_state.argtypes = [vp, fp]
    L.pmg_kernel_timing.argtypes = [vp, C.c_int32]
    L.pmg_kernel_time_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.pmg_launch_count.argtypes = [vp]
    L.pmg_launch_count.restype = C.c_int64
    L.pmg_overflow_count.argtypes = [vp]
    L.pmg_overflow_count.restype = C.c_int64
    if L.pmg_abi_version() != ABI_VERSION:
        raise ImportError("libpmg.so ABI version %d, expected %d" % (L.pmg_abi_version(), ABI_VERSION))
    _lib = L
    return L


class PmgError(RuntimeError):
    pass


def check(rc):
    """Maps the C status convention onto Python exceptions (PMG_ERR_INVALID -> ValueError, the
    condition the reference asserts / raises ValueError on)."""
    if rc == 0:
        return
    msg = load().pmg_last_error().decode("utf8", "replace")
    if rc == -1:
        raise ValueError(msg)
    raise PmgError("pmg error %d: %s" % (rc, msg))
