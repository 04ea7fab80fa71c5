"""ctypes binding of the CPU oracle (``oracle/bgs_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package never does (tests/test_no_oracle_in_product.py
checks that).  The library is built by ``oracle/Makefile`` (``__graft_entry__.build()`` runs it).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbgs_oracle.so")

STATS_LEN = 256
STAT_GAMES, STAT_WIN0, STAT_WIN1, STAT_DRAWS, STAT_STEPS, STAT_TRUNCATED = 0, 1, 2, 3, 4, 5
STAT_HIST0 = 16
DOMAIN_CONNECT, DOMAIN_BOUNCE = 0, 1

SOURCE_EMPTY, SOURCE_BLOCKED, SOURCE_PIECE, ALLOW_NULL_MOVE = 0, 1, 2, 4

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc if the shared library is missing or stale."""
    srcs = [os.path.join(_HERE, f) for f in ("bgs_oracle.c", "fast_connect.c", "bgs_oracle.h", "Makefile")]
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs)
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libbgs_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _p(arr, ctype):
    if arr is None:
        return None
    return arr.ctypes.data_as(C.POINTER(ctype))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.bgso_draw.restype = C.c_uint32
        L.bgso_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.bgso_connect_replay.restype = C.c_int64
        L.bgso_bounce_replay.restype = C.c_int64
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------- RNG
def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().bgso_philox4x32_10(c, k, out)
    return [int(x) for x in out]


def draw(seed: int, gid: int, t: int, domain: int) -> int:
    return int(lib().bgso_draw(seed, gid, t, domain))


def reward(winner: int) -> np.ndarray:
    out = np.zeros(2, dtype=np.float32)
    lib().bgso_reward(C.c_int(winner), _p(out, C.c_float))
    return out


# ------------------------------------------------------------------------------------------ Connect
def _grid8(grid):
    g = np.ascontiguousarray(grid, dtype=np.int8)
    assert g.ndim == 2
    return g


def connect_ended(grid, winner: int) -> bool:
    g = _grid8(grid)
    return bool(lib().bgso_connect_ended(_p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(winner)))


def connect_actions(grid, winner: int) -> list[int]:
    g = _grid8(grid)
    cols = np.zeros(g.shape[1], dtype=np.int32)
    n = lib().bgso_connect_actions(_p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(winner), _p(cols, C.c_int32))
    return [int(c) for c in cols[:n]]


def connect_next(grid, count: int, player: int, winner: int, col: int):
    """Returns (grid, player, winner) or None if the move is illegal."""
    g = _grid8(grid)
    out = np.empty_like(g)
    p, w = C.c_int(0), C.c_int(0)
    rc = lib().bgso_connect_next(
        _p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(count), C.c_int(player), C.c_int(winner), C.c_int(col),
        _p(out, C.c_int8), C.byref(p), C.byref(w),
    )
    if rc != 0:
        return None
    return out, p.value, w.value


def connect_rollout(H, W, K, n, gid0=0, seed=0, want_actions=True, want_grid=True):
    HW = H * W
    res = {
        "actions": np.empty((n, HW), dtype=np.uint8) if want_actions else None,
        "length": np.empty(n, dtype=np.uint8),
        "winner": np.empty(n, dtype=np.int8),
        "final_grid": np.empty((n, H, W), dtype=np.int8) if want_grid else None,
        "reward": np.empty((n, 2), dtype=np.float32),
        "stats": np.zeros(STATS_LEN, dtype=np.int64),
    }
    rc = lib().bgso_connect_rollout(
        C.c_int(H), C.c_int(W), C.c_int(K), C.c_uint64(n), C.c_uint64(gid0), C.c_uint64(seed),
        _p(res["actions"], C.c_uint8), _p(res["length"], C.c_uint8), _p(res["winner"], C.c_int8),
        _p(res["final_grid"], C.c_int8), _p(res["reward"], C.c_float), _p(res["stats"], C.c_int64),
    )
    if rc != 0:
        raise ValueError("oracle: unsupported Connect configuration")
    return res


def fast_connect_rollout(H, W, K, n, gid0=0, seed=0, per_game=True):
    """fast_connect.c: the bitboard loop (same draws, same results as connect_rollout without trajectories)."""
    res = {
        "length": np.empty(n, dtype=np.uint8) if per_game else None,
        "winner": np.empty(n, dtype=np.int8) if per_game else None,
        "stats": np.zeros(STATS_LEN, dtype=np.int64),
    }
    rc = lib().bgso_fast_connect_rollout(
        C.c_int(H), C.c_int(W), C.c_int(K), C.c_uint64(n), C.c_uint64(gid0), C.c_uint64(seed),
        _p(res["length"], C.c_uint8), _p(res["winner"], C.c_int8), _p(res["stats"], C.c_int64),
    )
    if rc != 0:
        raise ValueError("fast_connect: unsupported board")
    return res


def connect_rollout_from(K, grid, player, winner_in, gid0=0, seed=0):
    """Rollouts from supplied positions: grid int8[n,H,W], player int8[n], winner_in int8[n]."""
    grid = np.ascontiguousarray(grid, dtype=np.int8)
    player = np.ascontiguousarray(player, dtype=np.int8)
    winner_in = np.ascontiguousarray(winner_in, dtype=np.int8)
    n, H, W = grid.shape
    res = {
        "actions": np.empty((n, H * W), dtype=np.uint8),
        "length": np.empty(n, dtype=np.uint8),
        "winner": np.empty(n, dtype=np.int8),
        "final_grid": np.empty((n, H, W), dtype=np.int8),
        "reward": np.empty((n, 2), dtype=np.float32),
        "stats": np.zeros(STATS_LEN, dtype=np.int64),
    }
    rc = lib().bgso_connect_rollout_from(
        C.c_int(H), C.c_int(W), C.c_int(K), C.c_uint64(n), C.c_uint64(gid0), C.c_uint64(seed),
        _p(grid, C.c_int8), _p(player, C.c_int8), _p(winner_in, C.c_int8), _p(res["actions"], C.c_uint8),
        _p(res["length"], C.c_uint8), _p(res["winner"], C.c_int8), _p(res["final_grid"], C.c_int8),
        _p(res["reward"], C.c_float), _p(res["stats"], C.c_int64),
    )
    if rc != 0:
        raise ValueError("oracle: unsupported Connect configuration")
    return res


def connect_replay(H, W, K, actions, length, winner=None, final_grid=None, reward=None):
    """Replays trajectories through the oracle's transition; returns (n_bad, first_bad)."""
    actions = np.ascontiguousarray(actions, dtype=np.uint8)
    length = np.ascontiguousarray(length, dtype=np.uint8)
    n = length.shape[0]
    assert actions.shape == (n, H * W)
    winner = None if winner is None else np.ascontiguousarray(winner, dtype=np.int8)
    final_grid = None if final_grid is None else np.ascontiguousarray(final_grid, dtype=np.int8)
    reward = None if reward is None else np.ascontiguousarray(reward, dtype=np.float32)
    first = C.c_int64(-1)
    bad = lib().bgso_connect_replay(
        C.c_int(H), C.c_int(W), C.c_int(K), C.c_uint64(n), _p(actions, C.c_uint8), _p(length, C.c_uint8),
        _p(winner, C.c_int8), _p(final_grid, C.c_int8), _p(reward, C.c_float), C.byref(first),
    )
    return int(bad), int(first.value)


# ------------------------------------------------------------------------------------------- Bounce
def bounce_source_row(grid, player: int) -> int:
    g = _grid8(grid)
    return int(lib().bgso_bounce_source_row(_p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(player)))


def bounce_targets(grid, player: int, sx: int, sy: int, rules: int = 0) -> np.ndarray:
    g = _grid8(grid)
    out = np.zeros(g.shape, dtype=np.uint8)
    lib().bgso_bounce_targets(
        _p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(player), C.c_int(sx), C.c_int(sy), C.c_int(rules),
        _p(out, C.c_uint8),
    )
    return out


def bounce_actions(grid, player: int, ended: bool, rules: int = 0) -> np.ndarray:
    """int32[count, 4] rows (sx, sy, tx, ty) in ascending (sy, sx, ty, tx) order."""
    g = _grid8(grid)
    cap = g.shape[1] * g.size
    moves = np.zeros((cap, 4), dtype=np.int32)
    n = lib().bgso_bounce_actions(
        _p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(player), C.c_int(int(ended)), C.c_int(rules),
        _p(moves, C.c_int32), C.c_int(cap),
    )
    return moves[:n].copy()


def bounce_next(grid, player: int, ended: bool, sx, sy, tx, ty, rules: int = 0):
    """Returns (grid, player, winner, ended) or None if the move is illegal."""
    g = _grid8(grid)
    out = np.empty_like(g)
    p, w, e = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib().bgso_bounce_next(
        _p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(player), C.c_int(int(ended)), C.c_int(rules),
        C.c_int(sx), C.c_int(sy), C.c_int(tx), C.c_int(ty), _p(out, C.c_int8), C.byref(p), C.byref(w), C.byref(e),
    )
    if rc != 0:
        return None
    return out, p.value, w.value, bool(e.value)


def bounce_rollout(grid0, n, max_plies=512, gid0=0, seed=0, rules=0, want_moves=True, want_grid=True):
    g = _grid8(grid0)
    H, W = g.shape
    res = {
        "moves": np.empty((n, max_plies, 2), dtype=np.uint8) if want_moves else None,
        "length": np.empty(n, dtype=np.uint16),
        "winner": np.empty(n, dtype=np.int8),
        "final_grid": np.empty((n, H, W), dtype=np.int8) if want_grid else None,
        "reward": np.empty((n, 2), dtype=np.float32),
        "stats": np.zeros(STATS_LEN, dtype=np.int64),
    }
    rc = lib().bgso_bounce_rollout(
        _p(g, C.c_int8), C.c_int(H), C.c_int(W), C.c_int(rules), C.c_int(max_plies), C.c_uint64(n),
        C.c_uint64(gid0), C.c_uint64(seed), _p(res["moves"], C.c_uint8), _p(res["length"], C.c_uint16),
        _p(res["winner"], C.c_int8), _p(res["final_grid"], C.c_int8), _p(res["reward"], C.c_float),
        _p(res["stats"], C.c_int64),
    )
    if rc != 0:
        raise ValueError("oracle: unsupported Bounce configuration")
    return res


def bounce_rollout_from(grids, player, winner_in, ended_in, max_plies=512, gid0=0, seed=0, rules=0):
    grids = np.ascontiguousarray(grids, dtype=np.int8)
    player = np.ascontiguousarray(player, dtype=np.int8)
    winner_in = np.ascontiguousarray(winner_in, dtype=np.int8)
    ended_in = np.ascontiguousarray(ended_in, dtype=np.uint8)
    n, H, W = grids.shape
    res = {
        "moves": np.empty((n, max_plies, 2), dtype=np.uint8),
        "length": np.empty(n, dtype=np.uint16),
        "winner": np.empty(n, dtype=np.int8),
        "final_grid": np.empty((n, H, W), dtype=np.int8),
        "reward": np.empty((n, 2), dtype=np.float32),
        "stats": np.zeros(STATS_LEN, dtype=np.int64),
    }
    rc = lib().bgso_bounce_rollout_from(
        _p(grids, C.c_int8), _p(player, C.c_int8), _p(winner_in, C.c_int8), _p(ended_in, C.c_uint8),
        C.c_int(H), C.c_int(W), C.c_int(rules), C.c_int(max_plies), C.c_uint64(n), C.c_uint64(gid0), C.c_uint64(seed),
        _p(res["moves"], C.c_uint8), _p(res["length"], C.c_uint16), _p(res["winner"], C.c_int8),
        _p(res["final_grid"], C.c_int8), _p(res["reward"], C.c_float), _p(res["stats"], C.c_int64),
    )
    if rc != 0:
        raise ValueError("oracle: unsupported Bounce configuration")
    return res


def bounce_replay(grid0, moves, length, winner=None, final_grid=None, reward=None, rules=0):
    g = _grid8(grid0)
    H, W = g.shape
    moves = np.ascontiguousarray(moves, dtype=np.uint8)
    length = np.ascontiguousarray(length, dtype=np.uint16)
    n, max_plies = moves.shape[0], moves.shape[1]
    winner = None if winner is None else np.ascontiguousarray(winner, dtype=np.int8)
    final_grid = None if final_grid is None else np.ascontiguousarray(final_grid, dtype=np.int8)
    reward = None if reward is None else np.ascontiguousarray(reward, dtype=np.float32)
    first = C.c_int64(-1)
    bad = lib().bgso_bounce_replay(
        _p(g, C.c_int8), C.c_int(H), C.c_int(W), C.c_int(rules), C.c_int(max_plies), C.c_uint64(n),
        _p(moves, C.c_uint8), _p(length, C.c_uint16), _p(winner, C.c_int8), _p(final_grid, C.c_int8),
        _p(reward, C.c_float), C.byref(first),
    )
    return int(bad), int(first.value)


# ------------------------------------------------------- weighted choice / keys (B200-build conventions)
def connect_sample(grid, winner: int, probs, seed: int, gid: int, t: int) -> int:
    g = _grid8(grid)
    pr = np.ascontiguousarray(probs, dtype=np.float32)
    return int(lib().bgso_connect_sample(_p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(winner), _p(pr, C.c_float),
                                         C.c_uint64(seed), C.c_uint64(gid), C.c_uint32(t)))


def bounce_sample(grid, player: int, ended: bool, probs, seed: int, gid: int, t: int, rules: int = 0):
    """(sx, sy, tx, ty) or None."""
    g = _grid8(grid)
    pr = np.ascontiguousarray(probs, dtype=np.float32)
    mv = np.zeros(4, dtype=np.int32)
    rc = lib().bgso_bounce_sample(_p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(player), C.c_int(int(ended)),
                                  C.c_int(rules), _p(pr, C.c_float), C.c_uint64(seed), C.c_uint64(gid), C.c_uint32(t),
                                  _p(mv, C.c_int32))
    return None if rc != 0 else tuple(int(x) for x in mv)


def state_key(game: int, grid, player: int, winner: int) -> tuple[int, int]:
    g = _grid8(grid)
    key = np.zeros(2, dtype=np.uint64)
    lib().bgso_state_key(C.c_int(game), _p(g, C.c_int8), g.shape[0], g.shape[1], C.c_int(player), C.c_int(winner),
                         _p(key, C.c_uint64))
    return int(key[0]), int(key[1])
