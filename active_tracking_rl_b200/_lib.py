"""ctypes binding of libtrack2d.so (include/track2d.h).  There is NO CPU fallback: if the CUDA library
is missing or fails to load, importing the env raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtrack2d.so")

ABI_VERSION = 1
MAP = {"Block": 0, "Maze": 1, "Empty": 2}
OBS = {"Partial": 0, "Full": 1}
TARGET = {"Adv": 0, "PZR": 1, "Far": 2, "Nav": 3, "Ram": 4, "RPF": 5}
RNG = {"philox": 0, "numpy": 1}
FLAG_AUTO_RESET = 1
FLAG_KEEP_F64 = 2
FLAG_PLAN_AHEAD = 4
STATUS_BAD_ACTION, STATUS_PLAN_OVERFLOW, STATUS_ASTAR_REPLACE, STATUS_HEAP_OVERFLOW = 1, 2, 4, 8
NAV_MAXPLAN = 1024
RAM_MAXPLAN = 9


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("num_envs", C.c_int32), ("map_type", C.c_int32), ("obs_type", C.c_int32),
        ("target_mode", C.c_int32), ("level", C.c_int32), ("rng_mode", C.c_int32), ("device", C.c_int32),
        ("max_episode_steps", C.c_int32), ("flags", C.c_int32), ("seed", C.c_uint64),
    ]


class Track2DError(RuntimeError):
    pass


# every symbol include/track2d.h declares: name -> (restype, argtypes)
_vp, _i32, _u32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_uint32, C.c_int64, C.c_double
SYMBOLS = {
    "track2d_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "track2d_destroy": (C.c_int, [_vp]),
    "track2d_last_error": (C.c_char_p, []),
    "track2d_abi_version": (C.c_int, []),
    "track2d_launch_count": (C.c_uint64, []),
    "track2d_obs_cells": (C.c_int, [_vp]),
    "track2d_num_envs": (C.c_int, [_vp]),
    "track2d_map_height": (C.c_int, [_vp]),
    "track2d_map_width": (C.c_int, [_vp]),
    "track2d_reset": (C.c_int, [_vp, _vp, _vp, _vp]),
    "track2d_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "track2d_step_u8": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "track2d_reset_u8": (C.c_int, [_vp, _vp, _vp, _vp]),
    "track2d_reset_host": (C.c_int, [_vp, _vp, _vp]),
    "track2d_step_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "track2d_reset_host_u8": (C.c_int, [_vp, _vp, _vp]),
    "track2d_step_host_u8": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "track2d_join": (C.c_int, [_vp, _vp]),
    "track2d_step_host_begin": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _i32]),
    "track2d_host_chunk_wait": (C.c_int, [_vp, _i32]),
    "track2d_host_chunk_wait_stream": (C.c_int, [_vp, _i32, _vp]),
    "track2d_get_maps": (C.c_int, [_vp, _i32, _i32, _vp]),
    "track2d_set_maps": (C.c_int, [_vp, _i32, _i32, _vp]),
    "track2d_get_agents": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "track2d_set_agents": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "track2d_get_goals": (C.c_int, [_vp, _i32, _i32, _vp]),
    "track2d_get_ram": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "track2d_set_ram": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "track2d_get_nav": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "track2d_set_nav": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "track2d_astar_solve": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "track2d_get_rewards_f64": (C.c_int, [_vp, _i32, _i32, _vp]),
    "track2d_get_target_actions": (C.c_int, [_vp, _i32, _i32, _vp]),
    "track2d_seed_env": (C.c_int, [_vp, _i32, _u32]),
    "track2d_get_rng_numpy": (C.c_int, [_vp, _i32, _vp, _vp]),
    "track2d_init_maze": (C.c_int, [_vp, _vp, _vp]),
    "track2d_get_status": (C.c_int, [_vp, _vp, _vp]),
    "track2d_get_counters": (C.c_int, [_vp, _vp, _vp]),
    "track2d_maze_conv_forward": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "track2d_maze_conv_backward": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "track2d_maze_conv_forward_ex": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "track2d_maze_conv_backward_ex": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "track2d_lstm_heads_forward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                             C.c_uint64, _u32, _i32, _i64, _i64, _vp]),
    "track2d_policy_post_step": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "track2d_embed_add": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i64, _vp]),
    "track2d_a3c_loss_grad": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _dbl, _dbl, _dbl, _dbl, _dbl, _i32, _i32, _i32, _vp]),
    "track2d_peer_create": (C.c_int, [_i32, _i32, _i64, _i32, C.POINTER(C.c_void_p)]),
    "track2d_peer_handle": (C.c_int, [_vp, _vp]),
    "track2d_peer_connect": (C.c_int, [_vp, _vp]),
    "track2d_peer_segment": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "track2d_peer_connect_local": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "track2d_peer_allreduce": (C.c_int, [_vp, _vp, _vp]),
    "track2d_peer_sharedadam_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
    "track2d_peer_status": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "track2d_peer_destroy": (None, [_vp]),
    "track2d_lstm_heads_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "track2d_relu_backward_workspace_floats": (C.c_int64, [_i64, _i32]),
    "track2d_relu_backward_groupsum": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "track2d_gae_returns": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _dbl, _dbl, _vp]),
    "track2d_gemm_workspace_floats": (C.c_int64, [_i64, _i64, _i64]),
    "track2d_gemm_tf32x3": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _i32, _vp, _i64, _vp]),
    "track2d_lstm_bias_workspace_floats": (C.c_int64, [_i64, _i32]),
    "track2d_lstm_cell_forward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp]),
    "track2d_lstm_cell_backward": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp]),
    "track2d_colsum_workspace_floats": (C.c_int64, [_i64, _i32]),
    "track2d_colsum": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _i64, _vp]),
    "track2d_sharedadam_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _dbl, _dbl, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
}

_lib = None


def load():
    """dlopen libtrack2d.so and bind every exported symbol.  Raises Track2DError if the library has not
    been built (python -m active_tracking_rl_b200.build) -- the product has no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Track2DError("libtrack2d.so is not built: run `python -m active_tracking_rl_b200.build` "
                           "(needs nvcc; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.track2d_abi_version() != ABI_VERSION:
        raise Track2DError("libtrack2d.so ABI %d != binding ABI %d" % (lib.track2d_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


class on_device(object):
    """The learner kernels (GEMM, conv stack, LSTM, A3C loss, SharedAdam) launch on the CURRENT CUDA device with a stream of the
    operands' device: make that device current for the duration of the call (no-op when it already is)."""

    def __init__(self, device):
        self.idx = device.index if device.index is not None else 0
        self.prev = None

    def __enter__(self):
        import torch
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            import torch
            torch.cuda.set_device(self.prev)
        return False


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        raise Track2DError("libtrack2d error %d: %s" % (rc, lib.track2d_last_error().decode("utf-8", "replace")))


def parse_env_id(env_id):
    """'Track2D-BlockPartialPZR-v0' -> ('Block', 'Partial', 'PZR', 0)  (gym_track2d/__init__.py:3-18)"""
    if not env_id.startswith("Track2D-") or "-v" not in env_id:
        raise KeyError("unknown env id %r" % (env_id,))
    body, ver = env_id[len("Track2D-"):].rsplit("-v", 1)
    for m in MAP:
        if body.startswith(m):
            rest = body[len(m):]
            for o in OBS:
                if rest.startswith(o) and rest[len(o):] in TARGET and ver in ("0", "1"):
                    return m, o, rest[len(o):], int(ver)
    raise KeyError("unknown env id %r" % (env_id,))


def all_env_ids():
    """the 72 ids the reference registers"""
    return ["Track2D-%s%s%s-v%d" % (m, o, t, lv) for m in ("Maze", "Block", "Empty") for o in ("Full", "Partial")
            for t in ("Adv", "PZR", "Far", "Nav", "Ram", "RPF") for lv in range(2)]
