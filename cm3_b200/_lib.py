"""ctypes binding of include/cm3env.h (the C ABI of libcm3env.so)."""
import ctypes as C
import os

from .build import library_path

MAX_AGENTS = 8
ABI_VERSION = 2
MAX_DST = 8
REAL_F32, REAL_F64 = 0, 1
TILE_REAL, TILE_I8, TILE_U2 = 0, 1, 2

STATUS_NAMES = {0: "CM3_OK", -1: "CM3_ERR_BAD_ARG", -2: "CM3_ERR_BAD_SHAPE", -3: "CM3_ERR_CUDA",
                -4: "CM3_ERR_UNSUPPORTED", -5: "CM3_ERR_NO_DEVICE", -6: "CM3_ERR_NCCL"}


class Cm3Error(RuntimeError):
    def __init__(self, status, message):
        self.status = status
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))


class CheckersConfig(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("n_columns", C.c_int32), ("n_obs", C.c_int32),
                ("n_agents", C.c_int32), ("max_steps", C.c_int32),
                ("agents_r", C.c_int32 * MAX_AGENTS), ("agents_c", C.c_int32 * MAX_AGENTS),
                ("num_envs", C.c_int32), ("real", C.c_int32), ("device", C.c_int32),
                ("tile", C.c_int32), ("env_id_offset", C.c_int64),
                ("random_goal", C.c_int32), ("reserved", C.c_int32)]


class CheckersState(C.Structure):
    _fields_ = [("remaining", C.c_void_p), ("agents", C.c_void_p), ("meta", C.c_void_p),
                ("sync", C.c_void_p)]


class CheckersOutputs(C.Structure):
    FIELDS = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "reward", "local_rewards",
              "done", "goal_idx")
    _fields_ = [(f, C.c_void_p) for f in FIELDS]


class ParticleConfig(C.Structure):
    _fields_ = [("n_agents", C.c_int32), ("max_steps", C.c_int32), ("num_envs", C.c_int32),
                ("real", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32),
                ("env_id_offset", C.c_int64),
                ("dt", C.c_double), ("damping", C.c_double), ("contact_force", C.c_double),
                ("contact_margin", C.c_double), ("agent_size", C.c_double), ("mass", C.c_double),
                ("sensitivity", C.c_double), ("reach_thresh", C.c_double),
                ("agents_x", C.c_double * MAX_AGENTS), ("agents_y", C.c_double * MAX_AGENTS),
                ("landmarks_x", C.c_double * MAX_AGENTS), ("landmarks_y", C.c_double * MAX_AGENTS),
                ("initial_std", C.c_double), ("prob_random", C.c_double), ("contact_cutoff", C.c_double)]


class ParticleState(C.Structure):
    _fields_ = [("sv", C.c_void_p), ("landmarks", C.c_void_p), ("steps", C.c_void_p),
                ("collisions", C.c_void_p), ("reached", C.c_void_p), ("sync", C.c_void_p)]


class ParticleOutputs(C.Structure):
    FIELDS = ("global_state", "obs_others", "obs_self", "reward", "reward_n", "done", "collisions", "reached")
    _fields_ = [(f, C.c_void_p) for f in FIELDS]


# every symbol include/cm3env.h declares: (restype, argtypes)
_vp, _i32, _i64, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
SYMBOLS = {
    "cm3_abi_version": (C.c_int, []),
    "cm3_last_error": (C.c_char_p, []),
    "cm3_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "cm3_stream_synchronize": (C.c_int, [_vp]),
    "cm3_checkers_create": (C.c_int, [C.POINTER(CheckersConfig), C.POINTER(_vp)]),
    "cm3_checkers_destroy": (C.c_int, [_vp]),
    "cm3_checkers_tiles": (C.c_int, [_vp, C.POINTER(_i32)]),
    "cm3_checkers_step_chained": (C.c_int, [_vp, C.POINTER(CheckersState), _vp, _u64, _i64, _i32,
                                            C.POINTER(CheckersOutputs), _vp]),
    "cm3_checkers_reset": (C.c_int, [_vp, C.POINTER(CheckersState), _vp, _vp,
                                     C.POINTER(CheckersOutputs), _vp]),
    "cm3_checkers_step": (C.c_int, [_vp, C.POINTER(CheckersState), _vp,
                                    C.POINTER(CheckersOutputs), _vp]),
    "cm3_checkers_rollout": (C.c_int, [_vp, C.POINTER(CheckersState), _vp, _u64, _i64, _i32, _i32,
                                       _vp, C.POINTER(CheckersOutputs), _vp]),
    "cm3_checkers_rollout_gather": (C.c_int, [_vp, C.POINTER(CheckersState), _vp, _u64, _i64, _i32,
                                              _i32, _vp, _i32, C.POINTER(CheckersOutputs), _i64,
                                              _i64, _vp]),
    "cm3_checkers_step_host": (C.c_int, [_vp, C.POINTER(CheckersState), _vp, _vp,
                                         C.POINTER(CheckersOutputs), C.POINTER(CheckersOutputs),
                                         _vp]),
    "cm3_checkers_get_state": (C.c_int, [_vp, C.POINTER(CheckersState), C.POINTER(CheckersState), _vp]),
    "cm3_checkers_set_state": (C.c_int, [_vp, C.POINTER(CheckersState), C.POINTER(CheckersState), _vp]),
    "cm3_checkers_step_host_packed": (C.c_int, [_vp, C.POINTER(CheckersState), _vp, _vp,
                                                C.POINTER(CheckersOutputs), _vp, _vp, C.c_size_t, _vp]),
    "cm3_checkers_rollout_host": (C.c_int, [_vp, C.POINTER(CheckersState), _vp, _vp, _i32, _u64, _i64, _i32,
                                            C.POINTER(CheckersOutputs), C.POINTER(_vp), _vp, C.c_size_t,
                                            C.c_size_t, _vp]),
    "cm3_particle_rollout_host": (C.c_int, [_vp, C.POINTER(ParticleState), _vp, _vp, _i32, _u64, _i64, _i32,
                                            C.POINTER(ParticleOutputs), C.POINTER(_vp), _vp, C.c_size_t,
                                            C.c_size_t, _vp]),
    "cm3_particle_default_config": (None, [C.POINTER(ParticleConfig), _i32, _i32]),
    "cm3_particle_create": (C.c_int, [C.POINTER(ParticleConfig), C.POINTER(_vp)]),
    "cm3_particle_destroy": (C.c_int, [_vp]),
    "cm3_particle_tiles": (C.c_int, [_vp, C.POINTER(_i32)]),
    "cm3_particle_step_chained": (C.c_int, [_vp, C.POINTER(ParticleState), _vp, _u64, _i64, _i32,
                                            C.POINTER(ParticleOutputs), _vp]),
    "cm3_particle_reset": (C.c_int, [_vp, C.POINTER(ParticleState), _vp, _vp, _vp, _u64, _i64,
                                     C.POINTER(ParticleOutputs), _vp]),
    "cm3_particle_step": (C.c_int, [_vp, C.POINTER(ParticleState), _vp,
                                    C.POINTER(ParticleOutputs), _vp]),
    "cm3_particle_rollout": (C.c_int, [_vp, C.POINTER(ParticleState), _vp, _u64, _i64, _i32, _i32,
                                       _vp, C.POINTER(ParticleOutputs), _vp]),
    "cm3_particle_rollout_gather": (C.c_int, [_vp, C.POINTER(ParticleState), _vp, _u64, _i64, _i32,
                                              _i32, _vp, _i32, C.POINTER(ParticleOutputs), _i64,
                                              _i64, _vp]),
    "cm3_particle_step_host": (C.c_int, [_vp, C.POINTER(ParticleState), _vp, _vp,
                                         C.POINTER(ParticleOutputs), C.POINTER(ParticleOutputs),
                                         _vp]),
    "cm3_particle_get_state": (C.c_int, [_vp, C.POINTER(ParticleState), C.POINTER(ParticleState), _vp]),
    "cm3_particle_set_state": (C.c_int, [_vp, C.POINTER(ParticleState), C.POINTER(ParticleState), _vp]),
    "cm3_particle_step_host_packed": (C.c_int, [_vp, C.POINTER(ParticleState), _vp, _vp,
                                                C.POINTER(ParticleOutputs), _vp, _vp, C.c_size_t, _vp]),
    "cm3_comm_unique_id": (C.c_int, [_vp]),
    "cm3_comm_init": (C.c_int, [_vp, _i32, _i32, _i32, C.POINTER(_vp)]),
    "cm3_comm_allgather": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "cm3_comm_destroy": (C.c_int, [_vp]),
}

_lib = None


def load_library():
    """Loads libcm3env.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path):
        raise Cm3Error(-5, "%s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(cm3_b200 has no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cm3_abi_version() != ABI_VERSION:
        raise Cm3Error(-1, "ABI version mismatch: library %d, binding %d" % (lib.cm3_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise Cm3Error(status, load_library().cm3_last_error().decode("utf-8", "replace"))


def abi_version():
    return load_library().cm3_abi_version()


def device_count():
    n = C.c_int(0)
    check(load_library().cm3_device_count(C.byref(n)))
    return n.value
