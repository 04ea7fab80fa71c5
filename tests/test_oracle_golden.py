"""The oracle against every golden vector the reference holds for this path, plus self-consistency
properties (hypothesis) that do not depend on the reference: CPU only."""
import importlib
import os
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import golden_replay
from conftest import DEFAULT_BOUNCE_GRID, ROOT


@pytest.fixture(scope="module")
def oracle_api():
    """The oracle-backed stand-in package, imported under a private name so that it can never be
    confused with the product's `simulator` package."""
    path = os.path.join(ROOT, "oracle", "pyapi")
    saved = {k: v for k, v in sys.modules.items() if k == "simulator" or k.startswith("simulator.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, path)
    try:
        connect = importlib.import_module("simulator.game.connect")
        bounce = importlib.import_module("simulator.game.bounce")
    finally:
        sys.path.remove(path)
        for k in [k for k in sys.modules if k == "simulator" or k.startswith("simulator.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return connect, bounce


def test_fixture_is_complete(golden):
    n = sum(len(t["steps"]) for g in ("connect", "bounce") for t in golden[g].values())
    assert n == 20  # 4 Connect + 16 Bounce pictured positions
    assert set(golden["connect"]) == {"test_small", "test_json"}
    assert len(golden["bounce"]) == 7


def test_oracle_connect_golden(oracle_api, golden):
    assert golden_replay.replay_connect(oracle_api[0], golden) == 4
    golden_replay.replay_connect_json(oracle_api[0], golden)


def test_oracle_bounce_golden(oracle_api, golden):
    assert golden_replay.replay_bounce(oracle_api[1], golden) == 16
    golden_replay.replay_bounce_json(oracle_api[1], golden)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox4x32_10(
        [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]
    ) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    # the draw of (seed, game, ply) is word ply%4 of block ply//4
    seed, gid = 0x1234_5678_9ABC_DEF0, 0x0FED_CBA9_8765_4321
    blk = oracle.philox4x32_10([gid & 0xFFFFFFFF, gid >> 32, 2, 0], [seed & 0xFFFFFFFF, seed >> 32])
    assert [oracle.draw(seed, gid, 8 + j, 0) for j in range(4)] == blk


# ------------------------------------------------------------------------------ Connect properties
def _py_has_run(grid, K, who):
    H, W = grid.shape
    for r in range(H):
        for c in range(W):
            for dr, dc in ((0, 1), (1, 0), (1, 1), (1, -1)):
                ok = True
                for j in range(K):
                    rr, cc = r + j * dr, c + j * dc
                    if not (0 <= rr < H and 0 <= cc < W and grid[rr, cc] == who):
                        ok = False
                        break
                if ok:
                    return True
    return False


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 7), st.integers(1, 8), st.integers(1, 5), st.integers(0, 2**32 - 1))
def test_connect_transition_properties(oracle, H, W, K, seed):
    rng = np.random.default_rng(seed)
    grid = np.full((H, W), -1, dtype=np.int8)
    player, winner = 0, -1
    for ply in range(H * W + 1):
        legal = oracle.connect_actions(grid, winner)
        ended = oracle.connect_ended(grid, winner)
        assert ended == (len(legal) == 0)
        assert legal == ([] if ended else [c for c in range(W) if grid[H - 1, c] < 0])
        if ended:
            break
        col = int(rng.choice(legal))
        height = int((grid[:, col] >= 0).sum())
        nxt = oracle.connect_next(grid, K, player, winner, col)
        assert nxt is not None
        g2, p2, w2 = nxt
        assert g2[height, col] == player and (g2 != grid).sum() == 1  # lowest empty cell
        assert p2 == 1 - player
        assert (w2 == player) == _py_has_run(g2, K, player)
        assert w2 in (-1, player)
        grid, player, winner = g2, p2, w2
    # illegal moves are rejected
    assert oracle.connect_next(grid, K, player, winner, 0) is None
    assert oracle.connect_next(np.full((H, W), -1, np.int8), K, 0, -1, W) is None
    assert oracle.connect_next(np.full((H, W), -1, np.int8), K, 0, -1, -1) is None


def test_connect_rollout_statistics(oracle):
    r = oracle.connect_rollout(6, 7, 4, 20000, gid0=0, seed=0)
    s = r["stats"]
    assert s[oracle.STAT_GAMES] == 20000
    assert s[oracle.STAT_WIN0] + s[oracle.STAT_WIN1] + s[oracle.STAT_DRAWS] == 20000
    assert s[oracle.STAT_STEPS] == int(r["length"].astype(np.int64).sum())
    assert 20.9 < s[oracle.STAT_STEPS] / 20000 < 21.9  # 21.45 plies per random game
    assert 0.53 < s[oracle.STAT_WIN0] / 20000 < 0.59
    hist = s[oracle.STAT_HIST0:]
    assert hist.sum() == 20000 and (hist * np.arange(len(hist))).sum() == s[oracle.STAT_STEPS]
    assert r["length"].min() >= 7 and r["length"].max() <= 42
    assert oracle.connect_replay(6, 7, 4, r["actions"], r["length"], r["winner"], r["final_grid"], r["reward"]) == (0, -1)
    # replay detects corruption
    bad = r["actions"].copy()
    bad[5, 0] = (bad[5, 0] + 1) % 7
    assert oracle.connect_replay(6, 7, 4, bad, r["length"], r["winner"], r["final_grid"], r["reward"])[0] >= 1
    # game ids, not positions in the batch, key the RNG
    r2 = oracle.connect_rollout(6, 7, 4, 100, gid0=500, seed=0)
    np.testing.assert_array_equal(r2["actions"], r["actions"][500:600])
    r3 = oracle.connect_rollout(6, 7, 4, 100, gid0=500, seed=1)
    assert (r3["actions"] != r2["actions"]).any()


# ------------------------------------------------------------------------------- Bounce properties
def _py_targets(grid, player, sx, sy, rules=0):
    """Independent breadth-first restatement of SURVEY.md 4.4 rule 3 (sets instead of recursion)."""
    H, W = grid.shape
    g = grid.copy()
    variant = rules & 3
    wall = None
    if variant in (0, 1):
        g[sy, sx] = 0
    if variant == 1:
        wall = (sx, sy)
    fwd = 1 if player == 0 else -1
    far = H - 1 if player == 0 else 0
    start = (sx, sy, int(grid[sy, sx]), None)
    seen, todo, targets = {start}, [start], set()
    while todo:
        x, y, rem, last = todo.pop()
        for d, (dx, dy) in (("f", (0, fwd)), ("l", (-1, 0)), ("r", (1, 0))):
            if (last, d) in (("l", "r"), ("r", "l")):
                continue
            nx, ny = x + dx, y + dy
            if not (0 <= nx < W and 0 <= ny < H) or (nx, ny) == wall:
                continue
            v = int(g[ny, nx])
            if rem > 1:
                if v != 0 or ny == far:
                    continue
                nxt = (nx, ny, rem - 1, d)
            elif v > 0:
                nxt = (nx, ny, v, None)
            else:
                targets.add((nx, ny))
                continue
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
    if not rules & 4:
        targets.discard((sx, sy))
    return targets


@pytest.mark.parametrize("rules", [0, 1, 2, 4, 5, 6])
def test_bounce_targets_match_independent_search(oracle, rules):
    grid0 = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
    res = oracle.bounce_rollout(grid0, 40, max_plies=60, seed=7, rules=rules)
    checked = 0
    for i in range(40):
        g = grid0.copy()
        player = 0
        for t in range(int(res["length"][i])):
            row = oracle.bounce_source_row(g, player)
            expect = []
            for sx in range(g.shape[1]):
                if g[row, sx] > 0:
                    tm = oracle.bounce_targets(g, player, sx, row, rules)
                    got = {(x, y) for y, x in zip(*np.nonzero(tm))}
                    assert got == _py_targets(g, player, sx, row, rules)
                    expect += [(sx, row, x, y) for (y, x) in sorted((y, x) for (x, y) in got)]
            acts = oracle.bounce_actions(g, player, False, rules)
            assert [tuple(a) for a in acts] == expect  # ascending (sy, sx, ty, tx)
            checked += 1
            s, tcell = res["moves"][i, t]
            W = g.shape[1]
            nxt = oracle.bounce_next(g, player, False, s % W, s // W, tcell % W, tcell // W, rules)
            assert nxt is not None
            g, player = nxt[0], nxt[1]
        np.testing.assert_array_equal(g, res["final_grid"][i])
    assert checked > 300


def test_bounce_rollout_statistics(oracle):
    grid0 = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
    r = oracle.bounce_rollout(grid0, 1500, max_plies=512, seed=0)
    s = r["stats"]
    assert s[oracle.STAT_GAMES] == 1500
    assert s[oracle.STAT_WIN0] + s[oracle.STAT_WIN1] + s[oracle.STAT_DRAWS] + s[oracle.STAT_TRUNCATED] == 1500
    assert 24 < s[oracle.STAT_STEPS] / 1500 < 34  # 28.9 plies per random game
    assert oracle.bounce_replay(grid0, r["moves"], r["length"], r["winner"], r["final_grid"], r["reward"]) == (0, -1)
    # truncation
    rt = oracle.bounce_rollout(grid0, 200, max_plies=6, seed=0)
    assert rt["stats"][oracle.STAT_TRUNCATED] > 0 and (rt["winner"][rt["length"] == 6] <= -1).all()
    assert set(np.unique(rt["winner"])) <= {-2, -1, 0, 1}
    assert oracle.bounce_replay(grid0, rt["moves"], rt["length"], rt["winner"], rt["final_grid"], rt["reward"]) == (0, -1)
    np.testing.assert_array_equal(rt["moves"][:, :5], r["moves"][:200, :5])


def test_bitboard_cpu_baseline_equals_the_oracle(oracle):
    """oracle/fast_connect.c (the "best CPU" loop timed by bench.py) plays exactly the oracle's games."""
    for (H, W, K), n in (((6, 7, 4), 20000), ((4, 5, 3), 5000), ((7, 8, 5), 3000), ((2, 3, 2), 500), ((6, 7, 1), 100), ((3, 3, 4), 300)):
        a = oracle.connect_rollout(H, W, K, n, gid0=77, seed=9, want_actions=False, want_grid=False)
        b = oracle.fast_connect_rollout(H, W, K, n, gid0=77, seed=9)
        for k in ("length", "winner", "stats"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"{k} {(H, W, K)}")
