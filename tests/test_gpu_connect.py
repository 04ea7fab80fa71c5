"""Parity of the CUDA Connect-k path against the oracle (bit-exact: all integer / byte work).

Everything here calls the product through its public Python API, which calls the C ABI of
libbgs_b200.so; the oracle is only the checker."""
import numpy as np
import pytest
import torch

import golden_replay

pytestmark = pytest.mark.gpu

BOARDS = [(6, 7, 4), (8, 9, 5), (10, 12, 6)]
ODD_BOARDS = [(2, 3, 2), (1, 1, 1), (1, 5, 2), (5, 1, 3), (4, 4, 5), (7, 8, 4), (3, 16, 3), (15, 8, 4), (8, 16, 7), (6, 7, 1)]


def _run(cfg, n, seed=0, gid0=0, **kw):
    from simulator import batch

    res = batch.connect_rollout(cfg, n, seed, gid0, per_game=True, actions=True, final_grid=True, reward=True, **kw)
    torch.cuda.synchronize()
    return res


def _assert_equal_to_oracle(oracle, cfg, res, n, seed, gid0):
    ref = oracle.connect_rollout(*cfg, n, gid0=gid0, seed=seed)
    np.testing.assert_array_equal(res.length.cpu().numpy(), ref["length"])
    np.testing.assert_array_equal(res.winner.cpu().numpy(), ref["winner"])
    np.testing.assert_array_equal(res.actions.cpu().numpy(), ref["actions"])
    np.testing.assert_array_equal(res.final_grid.cpu().numpy(), ref["final_grid"])
    np.testing.assert_array_equal(res.reward.cpu().numpy(), ref["reward"])
    np.testing.assert_array_equal(res.stats.cpu().numpy(), ref["stats"])


def test_native_library_is_the_one_running():
    from simulator import _native as N

    assert N.lib().bgs_version() == 100 and N.device_count() >= 1
    import os

    maps = open("/proc/self/maps").read()
    assert os.path.realpath(N.LIB_PATH) in maps


@pytest.mark.parametrize("cfg", BOARDS)
def test_rollout_trajectories_equal_oracle(oracle, cfg):
    n = 30000 if cfg == (6, 7, 4) else 6000
    for seed, gid0 in ((0, 0), (0xDEADBEEFCAFEF00D, 2**40 + 12345)):
        _assert_equal_to_oracle(oracle, cfg, _run(cfg, n, seed, gid0), n, seed, gid0)


@pytest.mark.parametrize("cfg", ODD_BOARDS)
def test_rollout_generic_boards_equal_oracle(oracle, cfg):
    _assert_equal_to_oracle(oracle, cfg, _run(cfg, 3000, 5, 77), 3000, 5, 77)


def test_rollout_board_sweep(oracle):
    """Every board shape from 1x1 to 8x8 (plus wide / tall extremes) with several K, through the
    run-time-geometry kernel (with and without the opening phase, packed and per-byte trajectories,
    one- and two-word boards): trajectories and outcomes equal to the oracle's."""
    from simulator import batch

    shapes = [(h, w) for h in range(1, 9) for w in range(1, 9)] + [(2, 16), (15, 2), (8, 16), (15, 8), (11, 11), (9, 14)]
    checked = 0
    for i, (h, w) in enumerate(shapes):
        for k in sorted({1 + (i % 3), 4, max(h, w)}):
            n = 300
            res = batch.connect_rollout((h, w, k), n, 7 + i, 50 * i, per_game=True, actions=True, final_grid=True)
            ref = oracle.connect_rollout(h, w, k, n, gid0=50 * i, seed=7 + i)
            np.testing.assert_array_equal(res.actions.cpu().numpy(), ref["actions"], err_msg=str((h, w, k)))
            np.testing.assert_array_equal(res.winner.cpu().numpy(), ref["winner"], err_msg=str((h, w, k)))
            np.testing.assert_array_equal(res.final_grid.cpu().numpy(), ref["final_grid"], err_msg=str((h, w, k)))
            np.testing.assert_array_equal(res.stats.cpu().numpy(), ref["stats"], err_msg=str((h, w, k)))
            checked += 1
    assert checked >= 190


def test_ragged_and_tiny_batches(oracle):
    for n in (1, 2, 31, 33, 255, 257):
        _assert_equal_to_oracle(oracle, (6, 7, 4), _run((6, 7, 4), n, 9, 3), n, 9, 3)
    from simulator import batch

    res = batch.connect_rollout((6, 7, 4), 0)
    assert res.length.numel() == 0 and int(res.stats.sum()) == 0


def test_replay_one_million_games_through_the_oracle(oracle):
    """BASELINE.json: 100% bit-exact replay agreement on >= 1e6 sampled games (6x7x4)."""
    n = 1_048_576
    res = _run((6, 7, 4), n, seed=20261018, gid0=0)
    bad, first = oracle.connect_replay(
        6, 7, 4, res.actions.cpu().numpy(), res.length.cpu().numpy(), res.winner.cpu().numpy(),
        res.final_grid.cpu().numpy(), res.reward.cpu().numpy(),
    )
    assert (bad, first) == (0, -1)
    s = res.stats_dict()
    assert s["games"] == n and s["wins0"] + s["wins1"] + s["draws"] == n
    assert s["steps"] == int(res.length.sum(dtype=torch.int64))
    assert abs(s["steps"] / n - 21.35) < 0.15 and abs(s["wins0"] / n - 0.556) < 0.01


@pytest.mark.parametrize("cfg", [(8, 9, 5), (10, 12, 6)])
def test_replay_larger_boards_through_the_oracle(oracle, cfg):
    n = 60000
    res = _run(cfg, n, seed=11)
    bad, first = oracle.connect_replay(
        *cfg, res.actions.cpu().numpy(), res.length.cpu().numpy(), res.winner.cpu().numpy(),
        res.final_grid.cpu().numpy(), res.reward.cpu().numpy(),
    )
    assert (bad, first) == (0, -1)


def test_sharding_is_invisible(oracle):
    """Per-game streams are keyed by the GLOBAL game id: 1, 2, 4 or 8 shards give the same games,
    and the summed statistics equal the single-shard statistics (test T4)."""
    from simulator import batch

    n = 40001
    full = _run((6, 7, 4), n, seed=4)
    for world in (2, 8):
        stats = torch.zeros_like(full.stats)
        acts = []
        for r in range(world):
            start, count = batch.shard_range(n, r, world)
            part = batch.connect_rollout((6, 7, 4), count, 4, start, actions=True)
            stats += part.stats
            acts.append(part.actions)
        assert torch.equal(torch.cat(acts), full.actions)
        assert torch.equal(stats, full.stats)


def test_full_size_invariants():
    """16 Mi concurrent games (BASELINE.json configs[1]): size-independent properties."""
    from simulator import batch

    n = 16 * 2**20
    res = batch.connect_rollout((6, 7, 4), n, seed=1, per_game=True)
    torch.cuda.synchronize()
    s = res.stats.cpu().numpy()
    assert s[0] == n and s[1] + s[2] + s[3] == n
    hist = s[16:]
    assert hist.sum() == n and (hist * np.arange(len(hist))).sum() == s[4]
    assert hist[:7].sum() == 0 and hist[43:].sum() == 0
    L = res.length.to(torch.int64)
    assert int(L.sum()) == s[4]
    assert torch.equal(torch.bincount(L, minlength=len(hist)).cpu(), torch.from_numpy(hist))
    w = res.winner
    assert int((w == 0).sum()) == s[1] and int((w == 1).sum()) == s[2] and int((w == -1).sum()) == s[3]
    # a win by player 0 takes an odd number of plies, by player 1 an even number; draws fill the board
    assert bool(((L[w == 0] % 2) == 1).all()) and bool(((L[w == 1] % 2) == 0).all()) and bool((L[w == -1] == 42).all())
    # determinism + idempotence: the same call gives the same bytes, another seed does not
    again = batch.connect_rollout((6, 7, 4), n, seed=1, per_game=True)
    assert torch.equal(again.length, res.length) and torch.equal(again.winner, res.winner)
    other = batch.connect_rollout((6, 7, 4), 2**20, seed=2, per_game=True)
    assert not torch.equal(other.length, res.length[: 2**20])


def test_stats_accumulate_and_buffers_are_reused():
    from simulator import batch

    a = batch.connect_rollout((6, 7, 4), 5000, seed=1)
    b = batch.connect_rollout((6, 7, 4), 5000, seed=1, game_id0=5000, stats=a.stats, out=a)
    assert b.length.data_ptr() == a.length.data_ptr()
    assert int(b.stats[0]) == 10000
    c = batch.connect_rollout((6, 7, 4), 10000, seed=1)
    assert torch.equal(c.stats, b.stats)


@pytest.mark.parametrize("cfg", BOARDS + [(2, 3, 2), (4, 5, 3)])
def test_batched_step_equals_oracle(oracle, cfg):
    """ConnectBatch.step / legal / reward vs the oracle's transition on random play, with illegal
    and post-terminal actions mixed in."""
    from simulator import batch

    H, W, K = cfg
    n = 512
    rng = np.random.default_rng(1)
    b = batch.ConnectBatch.initial(cfg, n)
    grids = np.full((n, H, W), -1, dtype=np.int8)
    players = np.zeros(n, dtype=np.int64)
    winners = np.full(n, -1, dtype=np.int64)
    for _ in range(H * W + 2):
        np.testing.assert_array_equal(b.grid.cpu().numpy(), grids)
        legal_ref = np.zeros((n, W), dtype=bool)
        ended_ref = np.zeros(n, dtype=bool)
        for i in range(n):
            legal_ref[i, oracle.connect_actions(grids[i], winners[i])] = True
            ended_ref[i] = oracle.connect_ended(grids[i], winners[i])
        np.testing.assert_array_equal(b.legal_mask().cpu().numpy(), legal_ref)
        np.testing.assert_array_equal(b.has_ended.cpu().numpy().astype(bool), ended_ref)
        np.testing.assert_array_equal(b.reward.cpu().numpy(), np.stack([oracle.reward(w) for w in winners]))
        acts = rng.integers(-1, W + 1, size=n)
        nb, status = b.step(torch.from_numpy(acts))
        status = status.cpu().numpy()
        for i in range(n):
            nxt = oracle.connect_next(grids[i], K, players[i], winners[i], int(acts[i]))
            assert (nxt is None) == (status[i] == 1)
            if nxt is not None:
                grids[i], players[i], winners[i] = nxt
        np.testing.assert_array_equal(nb.player.cpu().numpy(), players)
        np.testing.assert_array_equal(nb.winner.cpu().numpy(), winners)
        b = nb
    assert ended_ref.mean() > 0.3


def test_golden_positions_through_the_object_api(golden):
    """The reference's own pictured positions (tests/test_connect.py) through simulator.game.connect."""
    from simulator.game import connect

    assert golden_replay.replay_connect(connect, golden) == 4
    golden_replay.replay_connect_json(connect, golden)


def test_object_api_random_games_match_oracle(oracle):
    import random

    from simulator.game.connect import Config

    random.seed(0)
    config = Config(6, 7, 4)
    for _ in range(3):
        state = config.sample_initial_state()
        grid, player, winner = np.full((6, 7), -1, np.int8), 0, -1
        while not state.has_ended:  # the loop of reference README.md:52-69
            assert state.player == player
            np.testing.assert_array_equal(state.grid, grid)
            actions = state.actions
            assert [a.column for a in actions] == oracle.connect_actions(grid, winner)
            action = random.choice(actions)
            grid, player, winner = oracle.connect_next(grid, 4, player, winner, action.column)
            state = action.sample_next_state()
        assert oracle.connect_ended(grid, winner) and state.actions == []
        np.testing.assert_array_equal(state.reward, oracle.reward(winner))
        with pytest.raises(RuntimeError):
            state.action_at(0)
    s = config.sample_initial_state()
    for bad in (-1, 7, 100):
        with pytest.raises(RuntimeError):
            s.action_at(bad)


def test_host_buffer_c_abi_entry_point(oracle):
    """bgs_connect_rollout_host: the call a non-Python FFI user makes (host pointers in, host out)."""
    import ctypes as C

    from simulator import _native as N

    n, H, W, K = 5000, 6, 7, 4
    actions = np.zeros((n, H * W), np.uint8)
    length = np.zeros(n, np.uint8)
    winner = np.zeros(n, np.int8)
    grid = np.zeros((n, H, W), np.int8)
    reward = np.zeros((n, 2), np.float32)
    stats = np.zeros(256, np.int64)
    N.check(N.lib().bgs_connect_rollout_host(
        0, H, W, K, n, 10, 3, actions.ctypes.data, length.ctypes.data, winner.ctypes.data, grid.ctypes.data,
        reward.ctypes.data, stats.ctypes.data))
    ref = oracle.connect_rollout(H, W, K, n, gid0=10, seed=3)
    for got, key in ((actions, "actions"), (length, "length"), (winner, "winner"), (grid, "final_grid"),
                     (reward, "reward"), (stats, "stats")):
        np.testing.assert_array_equal(got, ref[key])


def test_pipelined_host_stream_equals_oracle(oracle):
    """HostRollout.stream (the end-to-end path bench.py times): every batch's host results are the
    oracle's, although copies and kernels of consecutive batches overlap."""
    from simulator import batch

    n, k = 20000, 5
    host = batch.HostRollout((6, 7, 4), n)
    seen = 0
    for i, (st, length, winner) in enumerate(host.stream(11, 1000, k)):
        ref = oracle.connect_rollout(6, 7, 4, n, gid0=1000 + i * n, seed=11, want_actions=False, want_grid=False)
        np.testing.assert_array_equal(length.numpy(), ref["length"])
        np.testing.assert_array_equal(winner.numpy(), ref["winner"])
        np.testing.assert_array_equal(st.numpy(), ref["stats"])
        seen += 1
    assert seen == k
    st, length, winner = host.run(11, 1000)
    ref = oracle.connect_rollout(6, 7, 4, n, gid0=1000, seed=11, want_actions=False, want_grid=False)
    np.testing.assert_array_equal(length.numpy(), ref["length"])
    np.testing.assert_array_equal(st.numpy(), ref["stats"])


@pytest.mark.parametrize("cfg", [(6, 7, 4), (8, 9, 5), (10, 12, 6), (4, 5, 3), (5, 5, 4)])
def test_rollouts_from_supplied_positions_equal_oracle(oracle, cfg):
    """Leaf-evaluation mode: rollouts that continue from arbitrary positions (either side to move,
    some already won, some full) are bit-identical to the oracle's."""
    from simulator import batch

    H, W, K = cfg
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(5)
    b = batch.ConnectBatch.initial(cfg, n)
    plies = torch.randint(0, H * W + 1, (n,), device="cuda", generator=g)
    for t in range(H * W):  # random play for a per-state number of plies; illegal / post-end moves are no-ops
        acts = torch.randint(0, W, (n,), device="cuda", generator=g)
        acts = torch.where(plies > t, acts, torch.full_like(acts, -1))
        b, _ = b.step(acts)
    assert int(b.has_ended.sum()) > 0 and int((b.player == 1).sum()) > 0
    res = batch.connect_rollout(cfg, n, 21, 1000, per_game=True, actions=True, final_grid=True, reward=True, start=b)
    torch.cuda.synchronize()
    ref = oracle.connect_rollout_from(K, b.grid.cpu().numpy(), b.player.cpu().numpy(), b.winner.cpu().numpy(),
                                      gid0=1000, seed=21)
    for got, key in ((res.length, "length"), (res.winner, "winner"), (res.actions, "actions"),
                     (res.final_grid, "final_grid"), (res.reward, "reward"), (res.stats, "stats")):
        np.testing.assert_array_equal(got.cpu().numpy(), ref[key], err_msg=key)
    # from the empty board it is the ordinary rollout
    e = batch.ConnectBatch.initial(cfg, 2000)
    a = batch.connect_rollout(cfg, 2000, 3, 7, actions=True, start=e)
    c = batch.connect_rollout(cfg, 2000, 3, 7, actions=True)
    assert torch.equal(a.actions, c.actions) and torch.equal(a.stats, c.stats)


@pytest.mark.parametrize("cfg", [(6, 7, 4), (8, 9, 5), (5, 5, 3), (2, 3, 2), (1, 5, 2)])
def test_outputs_stay_inside_their_buffers(cfg):
    """compute-sanitizer is closed on this pool, so bounds are checked by hand: every output lives
    between two canary regions of one allocation, the C ABI gets raw interior pointers, and the
    canaries must come back untouched (this covers the in-place trajectory expansion, the packed
    board records and the tail group of the row export)."""
    from simulator import _native as N

    H, W, K = cfg
    n, HW, pad = 1003, H * W, 4096
    L = N.lib()
    pw = L.bgs_connect_packed_words(H, W)
    sizes = {"actions": n * HW, "length": n, "winner": n, "packed": n * pw * 8, "grid": n * HW, "reward": n * 8}
    off, total = {}, pad
    for k, v in sizes.items():
        off[k] = total
        total += (v + pad + 255) // 256 * 256
    buf = torch.full((total,), 0x5A, dtype=torch.uint8, device="cuda")
    base = buf.data_ptr()
    stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device="cuda")
    st = N.stream_ptr(torch)
    N.check(L.bgs_connect_rollout(H, W, K, n, 5, 9, base + off["actions"], base + off["length"], base + off["winner"],
                                  base + off["packed"], N.ptr(stats), st))
    N.check(L.bgs_connect_export(H, W, n, base + off["packed"], base + off["winner"], base + off["grid"],
                                 base + off["reward"], st))
    torch.cuda.synchronize()
    used = torch.zeros(total, dtype=torch.bool, device="cuda")
    for k, v in sizes.items():
        used[off[k]: off[k] + v] = True
    assert bool((buf[~used] == 0x5A).all()), "a kernel wrote outside its output buffer"
    assert int(stats[0]) == n
    acts = buf[off["actions"]: off["actions"] + n * HW].view(n, HW)
    length = buf[off["length"]: off["length"] + n].to(torch.int64)
    assert bool(((acts != 0xFF).sum(dim=1) == length).all())


def test_c_abi_argument_validation():
    """Bad arguments come back as status codes with a message -- never a crash, never an exception
    across the ABI (SURVEY.md 8b error contract)."""
    from simulator import _native as N

    L = N.lib()
    buf = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    st = N.stream_ptr(torch)
    # trajectories need the lengths to be expanded
    assert L.bgs_connect_rollout(6, 7, 4, 8, 0, 0, N.ptr(buf), None, None, None, None, st) == -1
    assert "length" in N.last_error()
    # unsupported boards
    assert L.bgs_connect_rollout(16, 16, 4, 8, 0, 0, None, None, None, None, None, st) == -2  # 256 cells
    assert L.bgs_connect_rollout(4, 33, 4, 8, 0, 0, None, None, None, None, None, st) == -2  # 33 columns
    assert L.bgs_connect_rollout_from(12, 12, 4, 4, 0, 0, N.ptr(buf), N.ptr(buf), None, N.ptr(buf), None, None, None, None, None, st) == -2
    assert L.bgs_connect_trajectory_grids(12, 12, 1, N.ptr(buf), N.ptr(buf), N.ptr(buf), st) == -2
    assert L.bgs_connect_step(0, 7, 4, 1, *([N.ptr(buf)] * 12), st) == -2
    # missing required pointers
    assert L.bgs_connect_step(6, 7, 4, 1, None, None, None, None, None, None, None, None, None, None, None, st) == -1
    assert L.bgs_connect_query(6, 7, 1, None, None, None, None, None, st) == -1
    assert L.bgs_connect_export(6, 7, 4, None, None, N.ptr(buf), None, st) == -1
    assert L.bgs_connect_rollout_from(6, 7, 4, 4, 0, 0, None, None, None, None, None, None, None, None, None, st) == -1
    assert L.bgs_bounce_rollout(None, 9, 6, 0, 64, 4, 0, 0, None, None, None, None, None, None, st) == -1
    assert L.bgs_bounce_moves(12, 11, 0, 1, N.ptr(buf), N.ptr(buf), None, None, N.ptr(buf), None, st) == -2
    # alignment the kernels rely on (include/bgs_b200.h, "Alignment"): rewards are written as float pairs, the
    # trajectories of even boards as 16-bit blocks
    assert buf.data_ptr() % 16 == 0
    assert L.bgs_connect_query(6, 7, 1, N.ptr(buf), N.ptr(buf), None, None, buf.data_ptr() + 4, st) == -1
    assert "8-byte aligned" in N.last_error()
    assert L.bgs_connect_export(6, 7, 1, N.ptr(buf), N.ptr(buf), None, buf.data_ptr() + 4, st) == -1
    assert L.bgs_connect_rollout(6, 7, 4, 8, 0, 0, buf.data_ptr() + 1, N.ptr(buf), None, None, None, st) == -1
    assert "2-byte aligned" in N.last_error()
    # zero games is a no-op
    assert L.bgs_connect_rollout(6, 7, 4, 0, 0, 0, None, None, None, None, None, st) == 0
    torch.cuda.synchronize()
    assert int(buf.sum()) == 0


@pytest.mark.parametrize("cfg", [(6, 7, 4), (8, 9, 5), (10, 12, 6), (4, 5, 3), (5, 5, 4)])
def test_per_ply_grids_equal_oracle_replay(oracle, cfg):
    """connect_trajectory_grids: the position after every ply of every game equals the oracle's
    transition applied move by move."""
    from simulator import batch

    H, W, K = cfg
    n = 300
    res = batch.connect_rollout(cfg, n, 8, 40, per_game=True, actions=True, final_grid=True)
    grids = batch.connect_trajectory_grids(cfg, res.actions, res.length)
    torch.cuda.synchronize()
    assert tuple(grids.shape) == (n, H * W + 1, H, W) and grids.dtype == torch.int8
    got = grids.cpu().numpy()
    acts, lens = res.actions.cpu().numpy(), res.length.cpu().numpy()
    for i in range(n):
        g, pl, w = np.full((H, W), -1, np.int8), 0, -1
        np.testing.assert_array_equal(got[i, 0], g)
        for t in range(int(lens[i])):
            g, pl, w = oracle.connect_next(g, K, pl, w, int(acts[i, t]))
            np.testing.assert_array_equal(got[i, t + 1], g, err_msg=f"game {i} ply {t + 1}")
        for t in range(int(lens[i]) + 1, H * W + 1):
            np.testing.assert_array_equal(got[i, t], g)
    np.testing.assert_array_equal(got[np.arange(n), lens.astype(np.int64)], res.final_grid.cpu().numpy())


@pytest.mark.parametrize("cfg", [(6, 7, 4), (8, 9, 5), (10, 12, 6)])
@pytest.mark.parametrize("n", [1, 7, 1031])
def test_per_ply_grids_ragged_batches(cfg, n):
    """The cell-stationary per-ply kernel packs 3 (8x9) / 2 (10x12) games into a warp, the word-stationary kernel
    (6x7) PAIRS of games (the last, odd game is a 2-byte aligned stream of its own): batch sizes that are not a
    multiple of that, checked against a host replay of the recorded trajectories."""
    from simulator import batch

    H, W, K = cfg
    res = batch.connect_rollout(cfg, n, 5, 77, per_game=True, actions=True, final_grid=True)
    canary = torch.full((n + 1, H * W + 1, H, W), 7, dtype=torch.int8, device="cuda")
    grids = batch.connect_trajectory_grids(cfg, res.actions, res.length, out=canary[:n])
    torch.cuda.synchronize()
    assert int((canary[n] != 7).sum()) == 0  # nothing written past the last game
    got = grids.cpu().numpy()
    acts, lens = res.actions.cpu().numpy(), res.length.cpu().numpy()
    for i in range(n):
        g = np.full((H, W), -1, np.int8)
        heights = [0] * W
        np.testing.assert_array_equal(got[i, 0], g)
        for t in range(int(lens[i])):
            c = int(acts[i, t])
            g[heights[c], c] = t & 1
            heights[c] += 1
            np.testing.assert_array_equal(got[i, t + 1], g, err_msg=f"game {i} ply {t + 1}")
        for t in range(int(lens[i]) + 1, H * W + 1):
            np.testing.assert_array_equal(got[i, t], g)
    np.testing.assert_array_equal(got[np.arange(n), lens.astype(np.int64)], res.final_grid.cpu().numpy())


@pytest.mark.parametrize("cfg", [(6, 7, 4), (8, 9, 5), (10, 12, 6)])
def test_per_ply_grids_with_unaligned_pointers_take_the_fallback_kernel(cfg):
    """The word- / cell-stationary per-ply kernels need a 16-byte aligned output and (6x7) 4-byte aligned
    trajectories; a C-ABI caller may pass anything, and the row kernel behind it must give the same bytes."""
    from simulator import batch

    H, W, K = cfg
    n = 333
    res = batch.connect_rollout(cfg, n, 9, 5, per_game=True, actions=True)
    want = batch.connect_trajectory_grids(cfg, res.actions, res.length)
    for a_off, o_off in ((1, 0), (0, 8), (2, 4), (3, 1)):
        abuf = torch.zeros(n * H * W + 16, dtype=torch.uint8, device="cuda")
        acts = abuf[a_off:a_off + n * H * W].view(n, H * W)
        acts.copy_(res.actions)
        obuf = torch.full((n * (H * W + 1) * H * W + 32,), 7, dtype=torch.int8, device="cuda")
        out = obuf[o_off:o_off + n * (H * W + 1) * H * W].view(n, H * W + 1, H, W)
        assert acts.data_ptr() % 16 == a_off % 16 and out.data_ptr() % 16 == o_off % 16
        got = batch.connect_trajectory_grids(cfg, acts, res.length, out=out)
        torch.cuda.synchronize()
        assert torch.equal(got, want), (a_off, o_off)
        assert int((obuf[:o_off] != 7).sum()) == 0 and int((obuf[o_off + out.numel():] != 7).sum()) == 0


@pytest.mark.parametrize("cfg", [(6, 7, 4), (10, 12, 6), (12, 12, 5)])
def test_batched_step_with_unaligned_grids_equals_the_aligned_call(cfg):
    """connect_step_kernel prefetches the next tile with cp.async when both grid pointers are 16-byte aligned (boards
    up to 128 cells); other pointers / boards take the synchronous copy -- same outputs, ragged last group included."""
    from simulator import batch

    H, W, K = cfg
    n = 32 * 37 + 5
    b = batch.ConnectBatch.initial(cfg, n)
    gen = torch.Generator(device="cuda").manual_seed(3)
    for _ in range(6):
        b, _ = b.step(torch.randint(0, W, (n,), device="cuda", generator=gen))
    acts = torch.randint(0, W, (n,), device="cuda", generator=gen).to(torch.int32)
    want, wstatus = b.step(acts)
    gbuf = torch.zeros(n * H * W + 16, dtype=torch.int8, device="cuda")
    g2 = gbuf[1:1 + n * H * W].view(n, H, W)
    g2.copy_(b.grid)
    assert g2.data_ptr() % 16 == 1
    b2 = batch.ConnectBatch(cfg, g2, b.player, b.winner)
    got, gstatus = b2.step(acts)
    torch.cuda.synchronize()
    assert torch.equal(got.grid, want.grid) and torch.equal(got.player, want.player)
    assert torch.equal(got.winner, want.winner) and torch.equal(gstatus, wstatus)


def test_packed_host_results_equal_oracle(oracle):
    """HostRollout(packed=True): one byte per game over PCIe, unpacked on the host = the oracle's results."""
    from simulator import batch

    n, k = 20011, 3
    host = batch.HostRollout((6, 7, 4), n, packed=True)
    for i, (st, result) in enumerate(host.stream(13, 500, k)):
        ref = oracle.connect_rollout(6, 7, 4, n, gid0=500 + i * n, seed=13, want_actions=False, want_grid=False)
        length, winner = batch.HostRollout.unpack(result)
        np.testing.assert_array_equal(length.numpy(), ref["length"])
        np.testing.assert_array_equal(winner.numpy(), ref["winner"])
        np.testing.assert_array_equal(st.numpy(), ref["stats"])
    with pytest.raises(ValueError):
        batch.HostRollout((12, 12, 5), 10, packed=True)  # 144 cells: a length does not fit 7 bits


@pytest.mark.parametrize("cfg", [(8, 9, 5), (10, 12, 6), (7, 9, 4)])
def test_packed_host_results_on_boards_of_more_than_63_cells(oracle, cfg):
    """64..127 cells: one byte = length | draw << 7 (a decided game's winner is the parity of its length)."""
    from simulator import batch

    n = 6001
    host = batch.HostRollout(cfg, n, packed=True)
    assert host.d2h_bytes == (n + 15) // 16 * 16 + 2048
    for i, (st, result) in enumerate(host.stream(5, 100, 3)):
        ref = oracle.connect_rollout(*cfg, n, gid0=100 + i * n, seed=5, want_actions=False, want_grid=False)
        length, winner = host.unpack_results(result)
        np.testing.assert_array_equal(length.numpy(), ref["length"])
        np.testing.assert_array_equal(winner.numpy(), ref["winner"])
        np.testing.assert_array_equal(st.numpy(), ref["stats"])
    if cfg[2] > 4:
        assert (ref["winner"] == -1).sum() > 0  # the sample holds draws (the bit-7 flag is exercised)


def test_leaf_rollouts_from_host_positions_pipelined(oracle):
    """HostLeafRollout: packed positions in pinned host memory -> rollouts -> packed results in pinned host
    memory, three batches in flight; every batch equals the oracle's rollouts from those positions."""
    from simulator import batch

    cfg, n, k = (6, 7, 4), 7000, 5
    g = torch.Generator(device="cuda").manual_seed(3)
    positions = []
    leaf = batch.HostLeafRollout(cfg, n, depth=3)
    assert leaf.h2d_bytes == n * 16 + (n + 15) // 16 * 16 and leaf.d2h_bytes == (n + 15) // 16 * 16 + 2048
    for j in range(k):
        b = batch.ConnectBatch.initial(cfg, n)
        for _ in range(4 + 3 * j):
            b, _ = b.step(torch.randint(0, 7, (n,), device="cuda", generator=g))
        positions.append(b)
    tickets = []
    for j in range(k):
        if j >= 3:  # the slot is about to be reused: consume its previous batch first
            stats, result = leaf.result(tickets[j - 3])
            _check_leaf(oracle, positions[j - 3], stats, result, 50 + (j - 3))
        slot = j % 3
        pk = positions[j].pack()
        hp, hm = leaf.host_inputs(slot)
        hp.copy_(pk.packed.cpu())
        hm.copy_(pk.meta.cpu())
        tickets.append(leaf.submit(9, (50 + j) * n, slot))
    for j in range(k - 3, k):
        stats, result = leaf.result(tickets[j])
        _check_leaf(oracle, positions[j], stats, result, 50 + j)


def _check_leaf(oracle, b, stats, result, batch_no):
    from simulator import batch

    n = b.n
    ref = oracle.connect_rollout_from(4, b.grid.cpu().numpy(), b.player.cpu().numpy(), b.winner.cpu().numpy(),
                                      gid0=batch_no * n, seed=9)
    length, winner = batch.HostRollout.unpack(result)
    np.testing.assert_array_equal(length.numpy(), ref["length"])
    np.testing.assert_array_equal(winner.numpy(), ref["winner"])
    np.testing.assert_array_equal(stats.numpy(), ref["stats"])


def test_dlpack_export():
    from simulator import batch

    res = batch.connect_rollout((6, 7, 4), 1000, actions=True)
    cap = torch.utils.dlpack.to_dlpack(res.actions)
    back = torch.utils.dlpack.from_dlpack(cap)
    assert back.data_ptr() == res.actions.data_ptr() and back.shape == (1000, 42)


@pytest.mark.parametrize("cfg", [(12, 12, 5), (15, 15, 5), (16, 7, 4), (5, 32, 4), (17, 15, 6), (3, 20, 3), (51, 5, 4)])
def test_boards_beyond_the_bitboard_limits_equal_oracle(oracle, cfg):
    """More than 128 cells / 16 columns / 15 rows (up to 255 cells, 32 columns): the byte-board fallback
    kernel gives the oracle's trajectories, lengths, winners, final grids, rewards and statistics."""
    from simulator import batch

    H, W, K = cfg
    n = 700
    res = batch.connect_rollout(cfg, n, seed=5, game_id0=31, per_game=True, actions=True, final_grid=True, reward=True)
    torch.cuda.synchronize()
    ref = oracle.connect_rollout(H, W, K, n, gid0=31, seed=5)
    for got, key in ((res.actions, "actions"), (res.length, "length"), (res.winner, "winner"),
                     (res.final_grid, "final_grid"), (res.reward, "reward"), (res.stats, "stats")):
        np.testing.assert_array_equal(got.cpu().numpy(), ref[key], err_msg=f"{key} {cfg}")
    # no trajectory requested: lengths / winners / statistics only
    res2 = batch.connect_rollout(cfg, n, seed=5, game_id0=31, per_game=True)
    np.testing.assert_array_equal(res2.length.cpu().numpy(), ref["length"])
    np.testing.assert_array_equal(res2.stats.cpu().numpy(), ref["stats"])


def test_large_board_step_and_object_api(oracle):
    """ConnectBatch.step / query and the reference's object loop on a 15x15x5 board (225 cells)."""
    from simulator import batch
    from simulator.game.connect import Config

    cfg = (15, 15, 5)
    H, W, K = cfg
    n = 256
    rng = np.random.default_rng(2)
    b = batch.ConnectBatch.initial(cfg, n)
    grids = [np.full((H, W), -1, np.int8) for _ in range(n)]
    players, winners = [0] * n, [-1] * n
    for ply in range(40):
        acts = rng.integers(-1, W + 1, size=n).astype(np.int32)
        nb, status = b.step(torch.from_numpy(acts).cuda())
        st = status.cpu().numpy()
        for i in range(n):
            nxt = oracle.connect_next(grids[i], K, players[i], winners[i], int(acts[i])) if 0 <= acts[i] < W else None
            if nxt is None:
                assert st[i] == 1, (ply, i)
                continue
            assert st[i] == 0, (ply, i)
            grids[i], players[i], winners[i] = nxt
        np.testing.assert_array_equal(nb.grid.cpu().numpy(), np.stack(grids))
        np.testing.assert_array_equal(nb.winner.cpu().numpy(), np.array(winners, dtype=np.int8))
        b = nb
    state = Config(*cfg).sample_initial_state()
    for col in (7, 7, 8, 7):
        state = state.action_at(col).sample_next_state()
    assert state.grid[0, 7] == 0 and state.grid[1, 7] == 1 and state.grid[0, 8] == 0 and state.grid[2, 7] == 1
    assert len(state.actions) == W and not state.has_ended


@pytest.mark.parametrize("cfg", [(8, 9, 5), (10, 12, 6), (6, 7, 4), (4, 4, 3)])
def test_single_pass_export_every_output_combination(oracle, cfg):
    """bgs_connect_rollout_export (BASELINE.json configs[3]): trajectory rows, final grids and rewards in
    the reference's layouts (tensor.hpp:69-87) from one call -- on 8x9x5 / 10x12x6 the rollout kernel
    writes them itself.  Every subset of the optional outputs, ragged batch sizes around the 64-id
    claim chunks, all bit-exact against the oracle."""
    from simulator import batch

    H, W, K = cfg
    for n in (1, 31, 63, 64, 65, 129, 1003, 4099):
        ref = oracle.connect_rollout(H, W, K, n, gid0=17, seed=3)
        for acts, grid, rew in ((1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1)):
            res = batch.connect_rollout(cfg, n, 3, 17, per_game=True, actions=bool(acts), final_grid=bool(grid),
                                        reward=bool(rew))
            torch.cuda.synchronize()
            np.testing.assert_array_equal(res.length.cpu().numpy(), ref["length"])
            np.testing.assert_array_equal(res.winner.cpu().numpy(), ref["winner"])
            np.testing.assert_array_equal(res.stats.cpu().numpy(), ref["stats"])
            if acts:
                np.testing.assert_array_equal(res.actions.cpu().numpy(), ref["actions"], err_msg=f"n={n}")
            if grid:
                np.testing.assert_array_equal(res.final_grid.cpu().numpy(), ref["final_grid"], err_msg=f"n={n}")
            if rew:
                np.testing.assert_array_equal(res.reward.cpu().numpy(), ref["reward"])


@pytest.mark.parametrize("cfg", [(8, 9, 5), (10, 12, 6), (6, 7, 4)])
@pytest.mark.parametrize("misalign", [0, 8])
def test_single_pass_export_stays_inside_its_buffers(oracle, cfg, misalign):
    """Raw C-ABI call with interior pointers between canaries; `misalign` = 8 puts `actions` off the
    16-byte boundary the fused kernel needs, so the two-step path must take over -- same bytes."""
    from simulator import _native as N

    H, W, K = cfg
    n, HW, pad = 1003, H * W, 4096
    L = N.lib()
    sizes = {"actions": n * HW, "length": n, "winner": n, "grid": n * HW, "reward": n * 8}
    off, total = {}, pad
    for k, v in sizes.items():
        off[k] = total + (misalign if k == "actions" else 0)
        total += (v + pad + 255) // 256 * 256
    buf = torch.full((total,), 0x5A, dtype=torch.uint8, device="cuda")
    base = buf.data_ptr()
    stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device="cuda")
    N.check(L.bgs_connect_rollout_export(H, W, K, n, 5, 9, base + off["actions"], base + off["length"],
                                         base + off["winner"], base + off["grid"], base + off["reward"],
                                         N.ptr(stats), N.stream_ptr(torch)))
    torch.cuda.synchronize()
    used = torch.zeros(total, dtype=torch.bool, device="cuda")
    for k, v in sizes.items():
        used[off[k]: off[k] + v] = True
    assert bool((buf[~used] == 0x5A).all()), "a kernel wrote outside its output buffer"
    ref = oracle.connect_rollout(H, W, K, n, gid0=5, seed=9)
    host = buf.cpu().numpy()
    np.testing.assert_array_equal(host[off["actions"]: off["actions"] + n * HW].reshape(n, HW), ref["actions"])
    np.testing.assert_array_equal(host[off["grid"]: off["grid"] + n * HW].view(np.int8).reshape(n, H, W), ref["final_grid"])
    np.testing.assert_array_equal(host[off["reward"]: off["reward"] + n * 8].view(np.float32).reshape(n, 2), ref["reward"])
    np.testing.assert_array_equal(host[off["length"]: off["length"] + n], ref["length"])
    np.testing.assert_array_equal(stats.cpu().numpy(), ref["stats"])


def test_full_size_launch_equals_oracle(oracle):
    """BASELINE.json configs[1] at full size: the statistics of the 16 Mi-game <6,7,4> launch (no trajectory)
    equal the oracle's over the same global ids (all host threads; ctypes releases the GIL), and a 1 Mi-game
    slice of length / winner is identical byte for byte."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    from simulator import batch

    n, seed, gid0 = 16 * 2**20, 20261018, 3 * 2**33
    res = batch.connect_rollout((6, 7, 4), n, seed, gid0, per_game=True)
    torch.cuda.synchronize()
    chunk = 2**18
    oracle.lib()

    def part(k):
        return oracle.connect_rollout(6, 7, 4, chunk, gid0=gid0 + k * chunk, seed=seed, want_actions=False, want_grid=False)

    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        parts = list(ex.map(part, range(n // chunk)))
    want = np.sum([p["stats"] for p in parts], axis=0)
    np.testing.assert_array_equal(res.stats.cpu().numpy(), want)
    lo = 5 * 2**20  # a 1 Mi slice in the middle of the batch
    length, winner = res.length.cpu().numpy(), res.winner.cpu().numpy()
    for k in range(lo // chunk, (lo + 2**20) // chunk):
        np.testing.assert_array_equal(length[k * chunk:(k + 1) * chunk], parts[k]["length"])
        np.testing.assert_array_equal(winner[k * chunk:(k + 1) * chunk], parts[k]["winner"])


@pytest.mark.parametrize("cfg", [(8, 9, 5), (10, 12, 6)])
def test_line_kernels_many_games_per_lane_equal_oracle(oracle, cfg):
    """BASELINE.json configs[3] boards with MANY games per lane: the line kernels clear only part of a lane's line
    words when a game starts (diagonals shorter than K keep the bits of earlier games -- LineGeo::per_game_reset),
    and a launch of a few thousand games gives every lane two games at most.  Here every lane of the resident grid
    plays ~8 games in a row, plain and with the fused export: statistics and every game's length / winner equal the
    oracle's over the same global ids, trajectories and final grids on a slice."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    from simulator import batch

    n, seed, gid0 = 5 * 2**18 + 77, 977, 7 * 2**34
    chunk = 2**16
    oracle.lib()
    bounds = [(a, min(chunk, n - a)) for a in range(0, n, chunk)]

    def part(b):
        a, cnt = b
        return oracle.connect_rollout(*cfg, cnt, gid0=gid0 + a, seed=seed, want_actions=a == 0, want_grid=a == 0)

    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        parts = list(ex.map(part, bounds))
    want_stats = np.sum([p["stats"] for p in parts], axis=0)
    want_len = np.concatenate([p["length"] for p in parts])
    want_win = np.concatenate([p["winner"] for p in parts])
    for kw in (dict(), dict(actions=True, final_grid=True, reward=True)):
        res = batch.connect_rollout(cfg, n, seed, gid0, per_game=True, **kw)
        np.testing.assert_array_equal(res.stats.cpu().numpy(), want_stats)
        np.testing.assert_array_equal(res.length.cpu().numpy(), want_len)
        np.testing.assert_array_equal(res.winner.cpu().numpy(), want_win)
        if kw:
            np.testing.assert_array_equal(res.actions[:chunk].cpu().numpy(), parts[0]["actions"])
            np.testing.assert_array_equal(res.final_grid[:chunk].cpu().numpy(), parts[0]["final_grid"])
            # a later slice, replayed: lanes are several games into the launch by then
            lo = n - chunk
            bad, first = oracle.connect_replay(
                *cfg, res.actions[lo:].cpu().numpy(), res.length[lo:].cpu().numpy(), res.winner[lo:].cpu().numpy(),
                res.final_grid[lo:].cpu().numpy(), res.reward[lo:].cpu().numpy())
            assert (bad, first) == (0, -1)


@pytest.mark.parametrize("cfg", [(6, 7, 4), (4, 5, 3), (8, 9, 5)])
def test_dense_host_results_equal_oracle(oracle, cfg):
    """HostRollout(packed="dense"): several games per 16-bit word in base S (6x7x4: 3 games, 5.33 bits each)."""
    from simulator import batch

    n = 10007
    host = batch.HostRollout(cfg, n, packed="dense")
    G, lmin, S = host.dense
    assert S ** G <= 65536 < S ** (G + 1) and lmin == 2 * cfg[2] - 1
    assert host.d2h_bytes == (2 * ((n + G - 1) // G) + 15) // 16 * 16 + 2048
    if cfg == (6, 7, 4):
        assert (G, S) == (3, 37)
    for i, (st, result) in enumerate(host.stream(5, 100, 3)):
        ref = oracle.connect_rollout(*cfg, n, gid0=100 + i * n, seed=5, want_actions=False, want_grid=False)
        length, winner = host.unpack_results(result)
        np.testing.assert_array_equal(length.numpy(), ref["length"])
        np.testing.assert_array_equal(winner.numpy(), ref["winner"])
        np.testing.assert_array_equal(st.numpy(), ref["stats"])
