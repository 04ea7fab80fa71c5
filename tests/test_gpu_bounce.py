"""Parity of the CUDA Bounce path against the oracle and the reference's pictured positions."""
import numpy as np
import pytest
import torch

import golden_replay
from conftest import DEFAULT_BOUNCE_GRID

pytestmark = pytest.mark.gpu

GRID = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
SMALL = np.array([[0, 0, 0], [1, 2, 3], [0, 0, 0], [0, 0, 0], [1, 2, 3], [0, 0, 0]], dtype=np.int8)
BIG_VALUES = np.array(
    [[0] * 6, [1, 0, 4, 0, 7, 2], [0] * 6, [0, 5, 0, 0, 0, 0], [0] * 6, [0] * 6, [0, 0, 6, 0, 0, 0], [3, 0, 7, 1, 0, 2], [0] * 6],
    dtype=np.int8,
)


def test_golden_positions_through_the_object_api(golden):
    """All 16 pictured Bounce positions of the reference (exhaustive target sets, wins, blocked
    victory, draw, JSON) through simulator.game.bounce on the GPU."""
    from simulator.game import bounce

    assert golden_replay.replay_bounce(bounce, golden) == 16
    golden_replay.replay_bounce_json(bounce, golden)


@pytest.mark.parametrize("rules", [0, 1, 2, 4, 5, 6])
@pytest.mark.parametrize("grid0", [GRID, SMALL, BIG_VALUES], ids=["default", "small", "big_values"])
def test_rollouts_equal_oracle(oracle, grid0, rules):
    from simulator import batch

    n, cap = 1500, 96
    res = batch.bounce_rollout(grid0, n, seed=3, game_id0=17, max_plies=cap, rules=rules,
                               moves=True, final_grid=True, reward=True)
    ref = oracle.bounce_rollout(grid0, n, max_plies=cap, gid0=17, seed=3, rules=rules)
    np.testing.assert_array_equal(res.length.cpu().numpy().astype(np.uint16), ref["length"])
    np.testing.assert_array_equal(res.winner.cpu().numpy(), ref["winner"])
    np.testing.assert_array_equal(res.actions.cpu().numpy(), ref["moves"])
    np.testing.assert_array_equal(res.final_grid.cpu().numpy(), ref["final_grid"])
    np.testing.assert_array_equal(res.reward.cpu().numpy(), ref["reward"])
    np.testing.assert_array_equal(res.stats.cpu().numpy(), ref["stats"])


def test_rollouts_from_random_start_grids_equal_oracle(oracle):
    """Random start positions of many shapes (H*W <= 64, W <= 8, values up to 15, pieces anywhere but
    the goal rows): GPU rollouts equal the oracle's, under the default and one alternative rule set."""
    from simulator import batch

    rng = np.random.default_rng(3)
    done = 0
    for trial in range(40):
        W = int(rng.integers(1, 9))
        H = int(rng.integers(3, min(64 // W, 12) + 1))
        grid0 = np.zeros((H, W), dtype=np.int8)
        maxv = int(rng.choice([2, 3, 3, 5, 7, 15]))
        cells = rng.random((H - 2, W)) < rng.uniform(0.1, 0.5)
        grid0[1:-1][cells] = rng.integers(1, maxv + 1, size=int(cells.sum()))
        rules = int(rng.choice([0, 0, 1, 2, 4, 6]))
        n, cap = 400, 48
        res = batch.bounce_rollout(grid0, n, seed=trial, game_id0=9 * trial, max_plies=cap, rules=rules,
                                   moves=True, final_grid=True, reward=True)
        ref = oracle.bounce_rollout(grid0, n, max_plies=cap, gid0=9 * trial, seed=trial, rules=rules)
        np.testing.assert_array_equal(res.actions.cpu().numpy(), ref["moves"], err_msg=f"trial {trial}")
        np.testing.assert_array_equal(res.winner.cpu().numpy(), ref["winner"], err_msg=f"trial {trial}")
        np.testing.assert_array_equal(res.final_grid.cpu().numpy(), ref["final_grid"], err_msg=f"trial {trial}")
        np.testing.assert_array_equal(res.stats.cpu().numpy(), ref["stats"], err_msg=f"trial {trial}")
        done += 1
    assert done == 40


def test_object_api_random_games_match_oracle(oracle):
    """The reference's loop (textual/examples/arena.py:60-69) through simulator.game.bounce on the GPU,
    checked state by state against the oracle."""
    import random

    from simulator.game.bounce import Config

    random.seed(1)
    config = Config(GRID)
    for _ in range(2):
        state = config.sample_initial_state()
        grid, player, ended, winner = GRID.copy(), 0, False, -1
        plies = 0
        while not state.has_ended and plies < 60:
            assert state.player == player
            np.testing.assert_array_equal(state.grid, grid)
            ref = [tuple(a) for a in oracle.bounce_actions(grid, player, ended)]
            actions = state.actions
            assert [(*a.source.tolist(), *a.target.tolist()) for a in actions] == ref
            action = random.choice(actions)
            grid, player, winner, ended = oracle.bounce_next(grid, player - 0, ended, *action.source.tolist(), *action.target.tolist())
            state = action.sample_next_state()
            plies += 1
        assert state.has_ended == ended
        if ended:
            assert state.actions == [] and state.reward.tolist() == oracle.reward(winner).tolist()
    s0 = config.sample_initial_state()
    with pytest.raises(RuntimeError):
        s0.action_at(np.array([0, 1]), np.array([5, 5]))
    assert s0.actions_at(np.array([0, 0])) == []  # not a movable piece
    with pytest.raises(RuntimeError):
        s0.actions_at(np.array([9, 9]))  # outside the board


def test_replay_default_games_through_the_oracle(oracle):
    from simulator import batch

    n, cap = 50000, 512
    res = batch.bounce_rollout(GRID, n, seed=1, max_plies=cap, moves=True, final_grid=True, reward=True)
    bad, first = oracle.bounce_replay(
        GRID, res.actions.cpu().numpy(), res.length.cpu().numpy().astype(np.uint16), res.winner.cpu().numpy(),
        res.final_grid.cpu().numpy(), res.reward.cpu().numpy())
    assert (bad, first) == (0, -1)
    s = res.stats_dict()
    assert s["games"] == n and abs(s["steps"] / n - 28.9) < 1.0 and s["truncated"] <= n // 1000


def test_moves_and_step_kernels_equal_oracle(oracle):
    """BounceBatch.moves / step on states sampled from random play, with illegal moves mixed in."""
    from simulator import batch

    rng = np.random.default_rng(0)
    H, W = GRID.shape
    n = 256
    b = batch.BounceBatch.initial(GRID, n)
    grids = np.repeat(GRID[None], n, 0)
    players = np.zeros(n, np.int64)
    ended = np.zeros(n, bool)
    winners = np.full(n, -1, np.int64)
    for _ in range(40):
        row, targets, count = (t.cpu().numpy() for t in b.moves())
        moves = np.zeros((n, 4), np.int32)
        for i in range(n):
            acts = oracle.bounce_actions(grids[i], players[i], ended[i])
            assert count[i] == len(acts)
            got = []
            for sx in range(W):
                m = int(targets[i, sx]) & (2**64 - 1)
                got += [(sx, int(row[i]), c % W, c // W) for c in range(H * W) if (m >> c) & 1]
            assert got == [tuple(a) for a in acts]
            if len(acts) and rng.random() < 0.9:
                moves[i] = acts[rng.integers(len(acts))]
            else:
                moves[i] = rng.integers(-1, 10, size=4)
        nb, status = b.step(torch.from_numpy(moves))
        status = status.cpu().numpy()
        for i in range(n):
            nxt = oracle.bounce_next(grids[i], players[i], ended[i], *[int(v) for v in moves[i]])
            assert (nxt is None) == (status[i] == 1), (i, moves[i])
            if nxt is not None:
                grids[i], players[i], winners[i], ended[i] = nxt
        np.testing.assert_array_equal(nb.grid.cpu().numpy(), grids)
        np.testing.assert_array_equal(nb.player.cpu().numpy(), players)
        np.testing.assert_array_equal(nb.has_ended.cpu().numpy().astype(bool), ended)
        np.testing.assert_array_equal(nb.winner.cpu().numpy(), winners)
        b = nb
    assert ended.mean() > 0.3


def test_outputs_stay_inside_their_buffers():
    """Hand-made bounds check (compute-sanitizer is closed on this pool): canaries around every output."""
    from simulator import _native as N

    H, W = GRID.shape
    n, cap, pad = 1003, 40, 4096
    L = N.lib()
    sizes = {"moves": n * cap * 2, "length": n * 2, "winner": n, "grid": n * H * W, "reward": n * 8}
    off, total = {}, pad
    for k, v in sizes.items():
        off[k] = total
        total += (v + pad + 255) // 256 * 256
    buf = torch.full((total,), 0x5A, dtype=torch.uint8, device="cuda")
    base = buf.data_ptr()
    stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device="cuda")
    N.check(L.bgs_bounce_rollout(GRID.ctypes.data, H, W, 0, cap, n, 3, 7, base + off["moves"], base + off["length"],
                                 base + off["winner"], base + off["grid"], base + off["reward"], N.ptr(stats),
                                 N.stream_ptr(torch)))
    torch.cuda.synchronize()
    used = torch.zeros(total, dtype=torch.bool, device="cuda")
    for k, v in sizes.items():
        used[off[k]: off[k] + v] = True
    assert bool((buf[~used] == 0x5A).all()), "a kernel wrote outside its output buffer"
    assert int(stats[0]) == n


def test_rollouts_from_supplied_positions_equal_oracle(oracle):
    """Leaf-evaluation mode for Bounce: rollouts continued from positions reached by random play
    (either side to move, some already ended) equal the oracle's."""
    from simulator import batch

    n = 1500
    pre = oracle.bounce_rollout(GRID, n, max_plies=40, seed=4)  # positions after <= 40 random plies
    rng = np.random.default_rng(0)
    cut = np.minimum(rng.integers(0, 30, size=n), pre["length"]).astype(np.int64)
    grids = np.repeat(GRID[None], n, 0)
    player = np.zeros(n, np.int8)
    winner = np.full(n, -1, np.int8)
    ended = np.zeros(n, np.uint8)
    H, W = GRID.shape
    for i in range(n):
        g, pl, w, e = GRID.copy(), 0, -1, False
        for t in range(cut[i]):
            s, tc = pre["moves"][i, t]
            g, pl, w, e = oracle.bounce_next(g, pl, e, s % W, s // W, tc % W, tc // W)
        grids[i], player[i], winner[i], ended[i] = g, pl, w, e
    assert ended.sum() > 0 and (player == 1).sum() > 0
    b = batch.BounceBatch(torch.from_numpy(grids).cuda(), torch.from_numpy(player).cuda(), torch.from_numpy(winner).cuda(),
                          torch.from_numpy(ended).cuda())
    res = batch.bounce_rollout(None, n, seed=9, game_id0=70, max_plies=64, moves=True, final_grid=True, reward=True, start=b)
    ref = oracle.bounce_rollout_from(grids, player, winner, ended, max_plies=64, gid0=70, seed=9)
    np.testing.assert_array_equal(res.actions.cpu().numpy(), ref["moves"])
    np.testing.assert_array_equal(res.length.cpu().numpy().astype(np.uint16), ref["length"])
    np.testing.assert_array_equal(res.winner.cpu().numpy(), ref["winner"])
    np.testing.assert_array_equal(res.final_grid.cpu().numpy(), ref["final_grid"])
    np.testing.assert_array_equal(res.reward.cpu().numpy(), ref["reward"])
    np.testing.assert_array_equal(res.stats.cpu().numpy(), ref["stats"])


def test_truncation_and_unsupported_boards():
    from simulator import batch

    res = batch.bounce_rollout(GRID, 2000, seed=0, max_plies=6)
    s = res.stats_dict()
    assert s["truncated"] > 0 and s["truncated"] == int((res.winner == -2).sum())
    assert int(res.length.max()) == 6
    with pytest.raises(RuntimeError):
        batch.bounce_rollout(np.ones((12, 11), np.int8), 10)  # 132 cells
    with pytest.raises(RuntimeError):
        batch.bounce_rollout(np.ones((4, 17), np.int8), 10)  # 17 columns
    with pytest.raises(RuntimeError):
        batch.bounce_rollout(np.full((4, 4), 16, np.int8), 10)


LARGE_SHAPES = [(8, 9), (10, 10), (9, 12), (16, 8), (8, 16), (12, 10), (11, 11), (5, 16)]


def _large_grid(rng, H, W):
    grid0 = np.zeros((H, W), dtype=np.int8)
    maxv = int(rng.choice([3, 3, 5, 9, 15]))
    cells = rng.random((H - 2, W)) < rng.uniform(0.1, 0.4)
    grid0[1:-1][cells] = rng.integers(1, maxv + 1, size=int(cells.sum()))
    return grid0


def test_large_board_rollouts_equal_oracle(oracle):
    """Boards of more than 64 cells / more than 8 columns (128-bit board words): rollouts and rollouts from
    positions equal the oracle's."""
    from simulator import batch

    rng = np.random.default_rng(11)
    for trial, (H, W) in enumerate(LARGE_SHAPES):
        grid0 = _large_grid(rng, H, W)
        rules = int(rng.choice([0, 0, 1, 2, 4, 6]))
        n, cap = 300, 64
        res = batch.bounce_rollout(grid0, n, seed=trial, game_id0=5 * trial, max_plies=cap, rules=rules,
                                   moves=True, final_grid=True, reward=True)
        ref = oracle.bounce_rollout(grid0, n, max_plies=cap, gid0=5 * trial, seed=trial, rules=rules)
        for got, key in ((res.actions, "moves"), (res.winner, "winner"), (res.final_grid, "final_grid"),
                         (res.reward, "reward"), (res.stats, "stats")):
            np.testing.assert_array_equal(got.cpu().numpy(), ref[key], err_msg=f"{key} trial {trial} {H}x{W}")
        np.testing.assert_array_equal(res.length.cpu().numpy().astype(np.uint16), ref["length"])


def test_large_board_moves_and_step_equal_oracle(oracle):
    """BounceBatch.moves / step and the object API on a 10x10 and an 8x16 board, against the oracle."""
    from simulator import batch
    from simulator.game.bounce import Config

    rng = np.random.default_rng(5)
    for H, W in ((10, 10), (8, 16), (9, 12)):
        grid0 = _large_grid(rng, H, W)
        n = 64
        b = batch.BounceBatch.initial(grid0, n)
        grids = [grid0.copy() for _ in range(n)]
        players, ended = [0] * n, [False] * n
        for ply in range(6):
            row, targets, count = (t.cpu().numpy() for t in b.moves())
            assert targets.shape == (n, W, 2)
            mv = np.zeros((n, 4), dtype=np.int32)
            for i in range(n):
                ref = [tuple(a) for a in oracle.bounce_actions(grids[i], players[i], ended[i])]
                got = []
                for sx in range(W):
                    m = (int(targets[i, sx, 0]) & (2**64 - 1)) | ((int(targets[i, sx, 1]) & (2**64 - 1)) << 64)
                    for cell in range(H * W):
                        if (m >> cell) & 1:
                            got.append((sx, int(row[i]), cell % W, cell // W))
                assert sorted(got) == sorted(ref), (H, W, ply, i)
                assert int(count[i]) == len(ref)
                if ref:
                    mv[i] = ref[int(rng.integers(len(ref)))]
                else:
                    mv[i] = (0, 0, 0, 0)
            nb, status = b.step(torch.from_numpy(mv))
            st = status.cpu().numpy()
            for i in range(n):
                nxt = oracle.bounce_next(grids[i], players[i], ended[i], *[int(v) for v in mv[i]])
                if nxt is None:
                    assert st[i] == 1
                    continue
                assert st[i] == 0
                grids[i], players[i], _, ended[i] = nxt
            np.testing.assert_array_equal(nb.grid.cpu().numpy(), np.stack(grids))
            np.testing.assert_array_equal(nb.player.cpu().numpy(), np.array(players, dtype=np.int8))
            np.testing.assert_array_equal(nb.has_ended.cpu().numpy().astype(bool), np.array(ended))
            b = nb
        # the reference's object loop on the same board
        state = Config(grid0).sample_initial_state()
        ref = [tuple(a) for a in oracle.bounce_actions(grid0, 0, False)]
        assert [(*a.source.tolist(), *a.target.tolist()) for a in state.actions] == ref


@pytest.mark.parametrize("n", [1, 2, 33, 65, 129, 4097])
def test_small_and_ragged_batches_equal_oracle(oracle, n):
    """Fewer games than a warp has slots / lanes, and counts that leave the last warp partly empty."""
    from simulator import batch

    for grid0, cap in ((GRID, 80), (BIG_VALUES, 40), (SMALL, 1)):
        res = batch.bounce_rollout(grid0, n, seed=4, game_id0=1000, max_plies=cap, moves=True, final_grid=True, reward=True)
        ref = oracle.bounce_rollout(grid0, n, max_plies=cap, gid0=1000, seed=4)
        np.testing.assert_array_equal(res.actions.cpu().numpy(), ref["moves"])
        np.testing.assert_array_equal(res.length.cpu().numpy().astype(np.uint16), ref["length"])
        np.testing.assert_array_equal(res.winner.cpu().numpy(), ref["winner"])
        np.testing.assert_array_equal(res.final_grid.cpu().numpy(), ref["final_grid"])
        np.testing.assert_array_equal(res.stats.cpu().numpy(), ref["stats"])


def test_replay_one_million_default_games_at_512_plies(oracle):
    """BASELINE.json configs[2] (default 9x6 board, max_plies 512): 1 Mi games of the rollout kernel, every
    recorded move replayed through the oracle's transition on all host threads; the truncated games (cycles)
    must be exactly the ones the oracle still finds running at 512 plies."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    from simulator import batch

    n, cap = 2**20, 512
    res = batch.bounce_rollout(GRID, n, 20261018, 0, max_plies=cap, moves=True, final_grid=True, reward=True)
    torch.cuda.synchronize()
    moves, length, winner = res.actions.cpu().numpy(), res.length.cpu().numpy().astype(np.uint16), res.winner.cpu().numpy()
    fgrid, reward = res.final_grid.cpu().numpy(), res.reward.cpu().numpy()
    oracle.lib()
    workers = os.cpu_count() or 1
    chunks = np.array_split(np.arange(n), workers * 8)

    def check(ix):
        lo, hi = int(ix[0]), int(ix[-1]) + 1
        return oracle.bounce_replay(GRID, moves[lo:hi], length[lo:hi], winner[lo:hi], fgrid[lo:hi], reward[lo:hi])

    with ThreadPoolExecutor(workers) as ex:
        results = list(ex.map(check, chunks))
    assert sum(r[0] for r in results) == 0, results
    s = res.stats_dict()
    assert s["games"] == n and s["wins0"] + s["wins1"] + s["draws"] + s["truncated"] == n
    assert s["steps"] == int(length.astype(np.int64).sum()) and s["truncated"] == int((winner == -2).sum())
    assert 27.0 < s["steps"] / n < 30.0


@pytest.mark.parametrize("packed", [False, True])
def test_pipelined_host_stream_equals_oracle(oracle, packed):
    """HostRollout(game="bounce"): BASELINE.json configs[2] end to end -- every batch's per-game results and
    statistics in pinned host memory (one copy per batch; 2 bytes per game when packed) equal the oracle's."""
    from simulator import batch

    n, k, cap = 3001, 4, 96
    host = batch.HostRollout(GRID, n, depth=2, packed=packed, game="bounce", max_plies=cap)
    assert host.d2h_bytes == (2 if packed else 3) * ((n + 15) // 16 * 16) + 2048
    for i, out in enumerate(host.stream(17, 40, k)):
        ref = oracle.bounce_rollout(GRID, n, max_plies=cap, gid0=40 + i * n, seed=17, want_moves=False, want_grid=False)
        if packed:
            length, winner = host.unpack_results(out[1])
        else:
            length, winner = out[1], out[2]
        np.testing.assert_array_equal(length.numpy().astype(np.uint16), ref["length"])
        np.testing.assert_array_equal(winner.numpy(), ref["winner"])
        np.testing.assert_array_equal(out[0].numpy(), ref["stats"])
    assert (ref["winner"] == -2).sum() > 0  # some games were cut at the cap
