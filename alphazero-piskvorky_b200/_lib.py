"""ctypes binding of libazp_b200.so (the C ABI declared in include/azp_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded
every entry point raises immediately."""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# AZP_B200_LIB selects another build of the same library (the bring-up build with clock stamps: make BRINGUP=1 OUT=...)
LIB_PATH = os.environ.get("AZP_B200_LIB") or os.path.join(HERE, "libazp_b200.so")

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_INVALID_MOVE, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_NO_WEIGHTS = range(7)
MORE = 100
NET_STEM, NET_TOWER = 0, 1
DTYPE_F32, DTYPE_BF16 = 0, 1
AUGMENT_REFERENCE, AUGMENT_D4, AUGMENT_NONE = 0, 1, 2
AUGMENT_MULT = {AUGMENT_REFERENCE: 4, AUGMENT_D4: 8, AUGMENT_NONE: 1}
PROF_KINDS = ("dominant", "advance", "conv1", "conv2", "fc", "root_expand", "end_move", "finish", "augment")


class Config(C.Structure):
    _fields_ = [
        ("board", C.c_int32), ("win_length", C.c_int32), ("games", C.c_int32), ("sims", C.c_int32),
        ("c_puct", C.c_double), ("dirichlet_alpha", C.c_double), ("dirichlet_weight", C.c_double),
        ("net_kind", C.c_int32), ("net_dtype", C.c_int32), ("tower_blocks", C.c_int32),
        ("tower_channels", C.c_int32), ("augment_mode", C.c_int32), ("max_plies", C.c_int32),
        ("seed", C.c_uint64), ("rank", C.c_int32), ("world", C.c_int32), ("device", C.c_int32),
        ("leaves_per_step", C.c_int32),
    ]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "simulations", "leaf_evals", "terminal_leaves", "children_scanned", "path_length", "expansions",
        "games_finished", "plies", "positions", "examples", "kernel_launches", "x_wins", "o_wins", "draws")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


P = C.c_void_p
# name -> (restype, argtypes); must list EVERY symbol include/azp_b200.h declares
SIGNATURES = {
    "azp_version": (C.c_int, []),
    "azp_create": (C.c_int, [C.POINTER(Config), C.POINTER(P)]),
    "azp_destroy": (None, [P]),
    "azp_last_error": (C.c_char_p, [P]),
    "azp_pack_boards": (C.c_int, [C.c_int, C.c_int, P, P, P, P]),
    "azp_unpack_boards": (C.c_int, [C.c_int, C.c_int, P, P, P, P]),
    "azp_legal_mask": (C.c_int, [C.c_int, C.c_int, P, P, P, P]),
    "azp_apply_moves": (C.c_int, [C.c_int, C.c_int, P, P, P, P, P, P, P]),
    "azp_terminal_status": (C.c_int, [C.c_int, C.c_int, C.c_int, P, P, P, P]),
    "azp_encode": (C.c_int, [C.c_int, C.c_int, P, P, P, P, P, P]),
    "azp_extract_policy": (C.c_int, [C.c_int, C.c_int, P, P, P, P, P]),
    "azp_augment_write": (C.c_int, [C.c_int, C.c_int, C.c_int, P, P, P, P, P, P, P]),
    "azp_augment_write_ring": (C.c_int, [C.c_int, C.c_int, C.c_int, P, P, P, P, P, P, P, P, C.c_int64, P]),
    "azp_load_weights": (C.c_int, [P, C.c_int, P, C.c_size_t]),
    "azp_load_weights_device": (C.c_int, [P, C.c_int, P, C.c_size_t, P]),
    "azp_weight_count": (C.c_size_t, [P]),
    "azp_net_forward": (C.c_int, [P, C.c_int, C.c_int, P, P, P, P, P]),
    "azp_debug_activations": (C.c_int, [P, C.c_int, C.c_int, P, P]),
    "azp_search_begin": (C.c_int, [P, C.c_int, P, P, P, P, P, P, P]),
    "azp_root_expand": (C.c_int, [P, P, C.c_int, P]),
    "azp_select": (C.c_int, [P, P]),
    "azp_leaf_states": (C.c_int, [P, P, P, P, P, P, P]),
    "azp_expand_backup": (C.c_int, [P, P, P, P]),
    "azp_expand_backup_f64": (C.c_int, [P, P, P, P]),
    "azp_root_visits": (C.c_int, [P, P, P]),
    "azp_root_priors": (C.c_int, [P, P, P]),
    "azp_search": (C.c_int, [P, C.c_int, C.c_int, P, P, P, P, P, P, C.c_int, P, P]),
    "azp_set_temperature_table": (C.c_int, [P, P, C.c_int]),
    "azp_set_augment_mode": (C.c_int, [P, C.c_int]),
    "azp_set_slab_capacity": (C.c_int, [P, C.c_int64]),
    "azp_selfplay_run": (C.c_int, [P, C.c_int64, C.c_int64, C.c_int, C.c_int, P, P, P, C.c_int64,
                                   C.POINTER(C.c_int64), P]),
    "azp_arena_run": (C.c_int, [P, C.c_int64, C.c_int64, C.POINTER(C.c_int64), P]),
    "azp_arena_run_steps": (C.c_int, [P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), P]),
    "azp_episode_snapshot": (C.c_int, [P, P, P, P, P, P, P]),
    "azp_get_counters": (C.c_int, [P, C.POINTER(Counters)]),
    "azp_reset_counters": (C.c_int, [P]),
    "azp_set_profiling": (C.c_int, [P, C.c_int]),
    "azp_get_profile": (C.c_int, [P, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "azp_get_profile_kernel": (C.c_int, [P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_double)]),
    "azp_trainer_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(P)]),
    "azp_trainer_destroy": (None, [P]),
    "azp_trainer_last_error": (C.c_char_p, [P]),
    "azp_trainer_weight_count": (C.c_size_t, [P]),
    "azp_train_loss_grad": (C.c_int, [P, C.c_int, P, P, P, P, P, P, P]),
    "azp_adamw_step": (C.c_int, [C.c_size_t, P, P, P, P, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.c_double, P]),
}

_lib = None


class EngineError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"azp_b200 status {status}: {message}")
        self.status = status


def load() -> C.CDLL:
    """Load the CUDA library or fail loudly (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The B200 self-play engine has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, engine=None):
    if status in (OK, MORE):
        return status
    msg = load().azp_last_error(engine)
    text = msg.decode() if msg else ""
    if status == ERR_INVALID_MOVE:
        raise ValueError("Invalid move")           # games.py:76-77
    raise EngineError(status, text)


def stream_ptr(device=None):
    """torch's current stream ON ``device`` (default: the current device) as a cudaStream_t."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device (or host) address of a tensor / None."""
    return None if t is None else C.c_void_p(t.data_ptr())
