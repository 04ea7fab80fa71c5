"""ctypes loader of ``libbgs_b200.so`` (C ABI: ``include/bgs_b200.h``).

There is no CPU fallback anywhere in this package: if the CUDA library has not been built, or no
CUDA device is visible, every compute call raises ``RuntimeError``.  Host-only logic (constructors,
JSON, equality) works without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libbgs_b200.so")

STATS_LEN = 256
STAT_GAMES, STAT_WIN0, STAT_WIN1, STAT_DRAWS, STAT_STEPS, STAT_TRUNCATED = 0, 1, 2, 3, 4, 5
STAT_HIST0 = 16
WINNER_DRAW, WINNER_TRUNCATED = -1, -2

_lib = None

_u64, _i32, _vp = C.c_uint64, C.c_int, C.c_void_p

_SIGNATURES = {
    "bgs_version": (C.c_int, []),
    "bgs_last_error": (C.c_char_p, []),
    "bgs_device_count": (C.c_int, []),
    "bgs_connect_supported": (C.c_int, [_i32, _i32, _i32]),
    "bgs_connect_packed_words": (C.c_int, [_i32, _i32]),
    "bgs_connect_rollout": (C.c_int, [_i32, _i32, _i32, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bgs_connect_rollout_export": (C.c_int, [_i32, _i32, _i32, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bgs_connect_start_words": (C.c_int, [_i32, _i32]),
    "bgs_connect_rollout_from": (C.c_int, [_i32, _i32, _i32, _u64, _u64, _u64] + [_vp] * 10),
    "bgs_connect_pack": (C.c_int, [_i32, _i32, _u64] + [_vp] * 6),
    "bgs_connect_rollout_from_packed": (C.c_int, [_i32, _i32, _i32, _u64, _u64, _u64] + [_vp] * 9),
    "bgs_connect_sample_step": (C.c_int, [_i32, _i32, _i32, _u64, _vp, _vp, _vp, _vp, _u64, _u64] + [_vp] * 11),
    "bgs_connect_keys": (C.c_int, [_i32, _i32, _u64] + [_vp] * 5),
    "bgs_bounce_sample_step": (C.c_int, [_i32, _i32, _i32, _u64, _vp, _vp, _vp, _vp, _vp, _u64, _u64] + [_vp] * 10),
    "bgs_bounce_keys": (C.c_int, [_i32, _i32, _u64] + [_vp] * 5),
    "bgs_connect_export": (C.c_int, [_i32, _i32, _u64, _vp, _vp, _vp, _vp, _vp]),
    "bgs_connect_trajectory_grids": (C.c_int, [_i32, _i32, _u64, _vp, _vp, _vp, _vp]),
    "bgs_connect_pack_results": (C.c_int, [_u64, _vp, _vp, _vp, _vp]),
    "bgs_connect_pack_results_wide": (C.c_int, [_u64, _vp, _vp, _vp, _vp]),
    "bgs_bounce_pack_results": (C.c_int, [_u64, _vp, _vp, _vp, _vp]),
    "bgs_connect_dense_results": (C.c_int, [_i32, _i32, _i32, _vp, _vp]),
    "bgs_connect_pack_results_dense": (C.c_int, [_i32, _i32, _i32, _u64, _vp, _vp, _vp, _vp]),
    "bgs_connect_step": (C.c_int, [_i32, _i32, _i32, _u64] + [_vp] * 12),
    "bgs_connect_query": (C.c_int, [_i32, _i32, _u64] + [_vp] * 6),
    "bgs_connect_rollout_host": (C.c_int, [_i32, _i32, _i32, _i32, _u64, _u64, _u64] + [_vp] * 6),
    "bgs_bounce_supported": (C.c_int, [_i32, _i32, _i32]),
    "bgs_bounce_moves": (C.c_int, [_i32, _i32, _i32, _u64] + [_vp] * 7),
    "bgs_bounce_step": (C.c_int, [_i32, _i32, _i32, _u64] + [_vp] * 12),
    "bgs_bounce_rollout": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _u64, _u64, _u64] + [_vp] * 7),
    "bgs_bounce_rollout_from": (C.c_int, [_i32, _i32, _i32, _i32, _u64, _u64, _u64] + [_vp] * 11),
    "bgs_bounce_rollout_host": (C.c_int, [_i32, _vp, _i32, _i32, _i32, _i32, _u64, _u64, _u64] + [_vp] * 6),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded library. Raises RuntimeError (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C board-game-simulator-python_b200/csrc`. This package has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().bgs_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libbgs_b200: {last_error()} (code {rc})")


def device_count() -> int:
    return int(lib().bgs_device_count())


def require_cuda():
    """torch on a CUDA device, or a loud failure."""
    import torch

    lib()
    if not torch.cuda.is_available() or device_count() < 1:
        raise RuntimeError(
            "simulator (b200): no CUDA device is available and this package has no CPU fallback"
        )
    return torch


def ptr(t) -> int | None:
    """Device/host address of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(torch) -> int:
    return torch.cuda.current_stream().cuda_stream
