"""The Bounce lane state machine (csrc/bounce_lane.cuh: bit-parallel move generation in mover-relative
orientation, k-th action selection, transition, blocked / draw test) compiled for the HOST and compared
game by game with the oracle.  This is the same source the CUDA rollout kernel inlines; the GPU parity
tests (tests/test_gpu_bounce.py) check the kernel itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import DEFAULT_BOUNCE_GRID, ROOT

SRC = os.path.join(ROOT, "tests", "native", "bounce_lane_host.cpp")
HDR = os.path.join(ROOT, "board-game-simulator-python_b200", "csrc", "bounce_lane.cuh")
OUT = os.path.join(ROOT, "tests", "native", "_build", "libbounce_lane_host.so")

GRID = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
SMALL = np.array([[0, 0, 0], [1, 2, 3], [0, 0, 0], [0, 0, 0], [1, 2, 3], [0, 0, 0]], dtype=np.int8)
BIG_VALUES = np.array(
    [[0] * 6, [1, 0, 4, 0, 7, 2], [0] * 6, [0, 5, 0, 0, 0, 0], [0] * 6, [0] * 6, [0, 0, 6, 0, 0, 0], [3, 0, 7, 1, 0, 2], [0] * 6],
    dtype=np.int8,
)


@pytest.fixture(scope="module")
def host():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", OUT, SRC], check=True)
    return C.CDLL(OUT)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def lane_rollout(host, grid0, n, max_plies, gid0, seed, rules, mode=0, start=None):
    g = np.ascontiguousarray(grid0, dtype=np.int8) if grid0 is not None else None
    if start is not None:
        grids, player, winner_in, ended_in = start
        H, W = grids.shape[1:]
    else:
        grids = player = winner_in = ended_in = None
        H, W = g.shape
    res = {
        "moves": np.empty((n, max_plies, 2), dtype=np.uint8),
        "length": np.empty(n, dtype=np.uint16),
        "winner": np.empty(n, dtype=np.int8),
        "final_grid": np.empty((n, H, W), dtype=np.int8),
        "reward": np.empty((n, 2), dtype=np.float32),
        "stats": np.zeros(256, dtype=np.int64),
    }
    rc = host.bgs_lane_host_bounce_rollout(
        C.c_int(mode), _p(g, C.c_int8), _p(grids, C.c_int8), _p(player, C.c_int8), _p(winner_in, C.c_int8),
        _p(ended_in, C.c_uint8), C.c_int(H), C.c_int(W), C.c_int(rules), C.c_int(max_plies), C.c_uint64(n),
        C.c_uint64(gid0), C.c_uint64(seed), _p(res["moves"], C.c_uint8), _p(res["length"], C.c_uint16),
        _p(res["winner"], C.c_int8), _p(res["final_grid"], C.c_int8), _p(res["reward"], C.c_float),
        _p(res["stats"], C.c_int64))
    assert rc == 0
    return res


def assert_same(got, ref, msg=""):
    for key in ("length", "winner", "moves", "final_grid", "reward", "stats"):
        np.testing.assert_array_equal(got[key], ref[key], err_msg=f"{key} {msg}")


@pytest.mark.parametrize("rules", [0, 1, 2, 4, 5, 6])
@pytest.mark.parametrize("grid0", [GRID, SMALL, BIG_VALUES], ids=["default", "small", "big_values"])
def test_lane_rollouts_equal_oracle(host, oracle, grid0, rules):
    n, cap = 1500, 96
    got = lane_rollout(host, grid0, n, cap, 17, 3, rules)
    ref = oracle.bounce_rollout(grid0, n, max_plies=cap, gid0=17, seed=3, rules=rules)
    assert_same(got, ref)


def test_compile_time_default_board_equals_oracle(host, oracle):
    n, cap = 20000, 512
    got = lane_rollout(host, GRID, n, cap, 5, 11, 0, mode=1)
    ref = oracle.bounce_rollout(GRID, n, max_plies=cap, gid0=5, seed=11, rules=0)
    assert_same(got, ref)
    assert abs(got["stats"][4] / n - 28.9) < 1.0


def test_lane_random_start_grids_equal_oracle(host, oracle):
    rng = np.random.default_rng(3)
    for trial in range(60):
        W = int(rng.integers(1, 9))
        H = int(rng.integers(3, min(64 // W, 12) + 1))
        grid0 = np.zeros((H, W), dtype=np.int8)
        maxv = int(rng.choice([2, 3, 3, 5, 7, 15]))
        cells = rng.random((H - 2, W)) < rng.uniform(0.1, 0.5)
        grid0[1:-1][cells] = rng.integers(1, maxv + 1, size=int(cells.sum()))
        rules = int(rng.choice([0, 0, 1, 2, 4, 6]))
        n, cap = 300, 48
        got = lane_rollout(host, grid0, n, cap, 9 * trial, trial, rules)
        ref = oracle.bounce_rollout(grid0, n, max_plies=cap, gid0=9 * trial, seed=trial, rules=rules)
        assert_same(got, ref, f"trial {trial} {H}x{W} rules {rules}")


def test_lane_rollouts_from_positions_equal_oracle(host, oracle):
    """Per-game start positions (either side to move, some already ended) as bgs_bounce_rollout_from."""
    rng = np.random.default_rng(5)
    n, cap = 600, 64
    base = oracle.bounce_rollout(GRID, n, max_plies=12, gid0=0, seed=2)
    grids = base["final_grid"].copy()
    player = (base["length"] % 2).astype(np.int8)
    winner_in = np.where(base["winner"] >= 0, base["winner"], -1).astype(np.int8)
    ended_in = (base["winner"] != -2).astype(np.uint8)
    grids[::7] = GRID
    player[::7] = rng.integers(0, 2, size=len(player[::7]))
    winner_in[::7] = -1
    ended_in[::7] = 0
    for rules in (0, 6):
        got = lane_rollout(host, None, n, cap, 40, 8, rules, start=(grids, player, winner_in, ended_in))
        ref = oracle.bounce_rollout_from(grids, player, winner_in, ended_in, max_plies=cap, gid0=40, seed=8, rules=rules)
        assert_same(got, ref, f"rules {rules}")


def test_lane_large_boards_equal_oracle(host, oracle):
    """Boards of more than 64 cells or more than 8 columns (up to 128 cells, 16 columns): the same state
    machine on unsigned __int128 board words."""
    rng = np.random.default_rng(11)
    shapes = [(8, 9), (10, 10), (9, 12), (16, 8), (8, 16), (12, 10), (11, 11), (5, 16), (14, 9), (3, 13)]
    for trial, (H, W) in enumerate(shapes):
        assert H * W <= 128 and W <= 16 and (H * W > 64 or W > 8)
        grid0 = np.zeros((H, W), dtype=np.int8)
        maxv = int(rng.choice([3, 3, 5, 9, 15]))
        cells = rng.random((H - 2, W)) < rng.uniform(0.1, 0.4)
        grid0[1:-1][cells] = rng.integers(1, maxv + 1, size=int(cells.sum()))
        rules = int(rng.choice([0, 0, 1, 2, 4, 6]))
        n, cap = 200, 64
        got = lane_rollout(host, grid0, n, cap, 5 * trial, trial, rules)
        ref = oracle.bounce_rollout(grid0, n, max_plies=cap, gid0=5 * trial, seed=trial, rules=rules)
        assert_same(got, ref, f"trial {trial} {H}x{W} rules {rules}")


def test_segment_table_hash_of_the_default_board_is_a_bijection(host):
    """GeoCT<9,6> indexes its segment table with (window & mask) * kHashMul >> 24 instead of 8 gathered bits; the
    kernel builds the table THROUGH that hash (seg_lut_slot), so a collision would silently drop an entry."""
    host.bgs_lane_host_seg_hash_is_perfect.restype = C.c_int
    assert host.bgs_lane_host_seg_hash_is_perfect() == 1

