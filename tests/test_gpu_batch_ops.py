"""Batched equality / keys / dedup, packed positions, weighted-policy stepping and the bulk JSON wire
format against the oracle.  Everything goes through the public Python API,
which calls the C ABI of libbgs_b200.so."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import DEFAULT_BOUNCE_GRID

pytestmark = pytest.mark.gpu


def _random_connect_batch(cfg, n, seed, max_plies=None):
    """n positions reached by random play for a per-state number of plies (illegal moves are no-ops)."""
    from simulator import batch

    H, W, _ = cfg
    g = torch.Generator(device="cuda").manual_seed(seed)
    b = batch.ConnectBatch.initial(cfg, n)
    plies = torch.randint(0, (max_plies or H * W) + 1, (n,), device="cuda", generator=g)
    for t in range(max_plies or H * W):
        acts = torch.randint(0, W, (n,), device="cuda", generator=g)
        b, _ = b.step(torch.where(plies > t, acts, torch.full_like(acts, -1)))
    return b


def _u64(t):
    return t.cpu().numpy().view(np.uint64)


# --------------------------------------------------------------------------------------------- keys
@pytest.mark.parametrize("cfg", [(6, 7, 4), (2, 3, 2), (8, 9, 5), (10, 12, 6), (7, 9, 4)])
def test_connect_keys_equal_oracle_and_host_equality(oracle, cfg):
    """keys are the oracle's, and equal keys <=> equal host _key() (grid, player, winner) on >= 1e5 states
    (many duplicates: short random play from the empty board)."""
    from simulator.game.connect import Config, State

    n = 120_000 if cfg == (6, 7, 4) else 20_000
    b = _random_connect_batch(cfg, n, 3, max_plies=6 if cfg == (6, 7, 4) else 5)
    keys = _u64(b.key())
    grid, player, winner = b.grid.cpu().numpy(), b.player.cpu().numpy(), b.winner.cpu().numpy()
    for i in range(0, n, max(1, n // 500)):
        assert tuple(int(x) for x in keys[i]) == oracle.state_key(1, grid[i], int(player[i]), int(winner[i]))
    config = Config(*cfg)
    host = {}
    for i in range(n):
        hk = State(config, grid[i], int(player[i]), int(winner[i]))._key()
        dk = (int(keys[i, 0]), int(keys[i, 1]))
        assert host.setdefault(hk, dk) == dk  # equal states -> equal keys
    assert len(set(host.values())) == len(host)  # different states -> different keys
    assert len(host) < n  # the sample did contain duplicates
    # same grid, other player / winner: different key
    b2 = type(b)(b.config, b.grid, 1 - b.player, b.winner)
    assert not bool(b.equal(b2).any())
    assert bool((b == b).all())


def test_connect_unique_is_a_dedup(oracle):
    cfg = (6, 7, 4)
    b = _random_connect_batch(cfg, 50_000, 9, max_plies=4)
    u, first, inverse = b.unique()
    assert u.n < b.n and int(inverse.max()) == u.n - 1
    assert bool(u.select(inverse).equal(b).all())
    assert torch.equal(u.grid, b.grid[first])
    seen = {(g.tobytes(), int(p), int(w)) for g, p, w in zip(b.grid.cpu().numpy(), b.player.cpu().numpy(), b.winner.cpu().numpy())}
    assert len(seen) == u.n
    # first[i] is the FIRST occurrence
    inv = inverse.cpu().numpy()
    fo = np.full(u.n, b.n)
    np.minimum.at(fo, inv, np.arange(b.n))
    np.testing.assert_array_equal(first.cpu().numpy(), fo)


def test_bounce_keys_and_unique(oracle):
    from simulator import batch

    grid0 = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
    n = 30_000
    b = batch.BounceBatch.initial(grid0, n)
    b.ply = torch.zeros(n, dtype=torch.int32, device="cuda")
    probs = torch.ones((n, 6, 54), device="cuda")
    for _ in range(3):
        b, _, _ = b.sample_step(probs, seed=5, game_id0=0)
    keys = _u64(b.key())
    grid, player, winner = b.grid.cpu().numpy(), b.player.cpu().numpy(), b.winner.cpu().numpy()
    for i in range(0, n, 60):
        assert tuple(int(x) for x in keys[i]) == oracle.state_key(2, grid[i], int(player[i]), int(winner[i]))
    host = {}
    for i in range(n):
        hk = (grid[i].tobytes(), int(player[i]), int(winner[i]))
        dk = (int(keys[i, 0]), int(keys[i, 1]))
        assert host.setdefault(hk, dk) == dk
    assert len(set(host.values())) == len(host) and len(host) < n
    u, first, inverse = b.unique()
    assert u.n == len(host) and bool(u.select(inverse).equal(b).all())


# ----------------------------------------------------------------------------------- packed positions
@pytest.mark.parametrize("cfg", [(6, 7, 4), (8, 9, 5), (10, 12, 6), (4, 5, 3)])
def test_rollouts_from_packed_positions_equal_oracle(oracle, cfg):
    """ConnectBatch.pack() -> connect_rollout(start=ConnectPacked): same games as from the int8 grids and as
    the oracle's bgso_connect_rollout_from; the packed form is 17 / 33 bytes per position."""
    from simulator import batch

    H, W, K = cfg
    n = 5000
    b = _random_connect_batch(cfg, n, 5)
    pk = b.pack()
    assert pk.packed.shape == (n, 2 if H * W <= 64 else 4) and pk.meta.dtype == torch.uint8
    res = batch.connect_rollout(cfg, n, 21, 1000, per_game=True, actions=True, final_grid=True, reward=True, start=pk)
    torch.cuda.synchronize()
    ref = oracle.connect_rollout_from(K, b.grid.cpu().numpy(), b.player.cpu().numpy(), b.winner.cpu().numpy(), gid0=1000, seed=21)
    for got, key in ((res.length, "length"), (res.winner, "winner"), (res.actions, "actions"),
                     (res.final_grid, "final_grid"), (res.reward, "reward"), (res.stats, "stats")):
        np.testing.assert_array_equal(got.cpu().numpy(), ref[key], err_msg=key)
    # the packed words are the public packed-board format: exporting them gives the grids back
    from simulator import _native as N

    grid = torch.empty_like(b.grid)
    N.check(N.lib().bgs_connect_export(H, W, n, N.ptr(pk.packed), None, N.ptr(grid), None, N.stream_ptr(torch)))
    assert torch.equal(grid, b.grid)
    meta = pk.meta.cpu().numpy()
    np.testing.assert_array_equal(meta & 1, b.player.cpu().numpy())
    np.testing.assert_array_equal(((meta >> 1) & 3).astype(np.int8) - 1, b.winner.cpu().numpy())


# --------------------------------------------------------------------------------- weighted stepping
@pytest.mark.parametrize("cfg", [(6, 7, 4), (8, 9, 5), (3, 4, 3)])
def test_connect_sample_step_with_equal_weights_replays_the_rollout_kernel(cfg):
    """Stepping n boards from empty with constant weights = the trajectories of connect_rollout, ply by ply."""
    from simulator import batch

    H, W, K = cfg
    n = 4000
    res = batch.connect_rollout(cfg, n, 11, 500, per_game=True, actions=True, final_grid=True)
    b = batch.ConnectBatch.initial(cfg, n)
    probs = torch.full((n, W), 0.25, device="cuda")
    acts = res.actions.to(torch.int32)
    for t in range(H * W):
        b, a, status = b.sample_step(probs, seed=11, game_id0=500)
        want = torch.where(acts[:, t] == 255, torch.full_like(acts[:, t], -1), acts[:, t])
        assert torch.equal(a, want), f"ply {t}"
        assert torch.equal(status == 1, want < 0)
    assert torch.equal(b.grid, res.final_grid) and torch.equal(b.winner, torch.where(res.winner < 0, -1, res.winner).to(torch.int8))
    assert bool(b.has_ended.all())


@pytest.mark.parametrize("cfg", [(6, 7, 4), (10, 12, 6)])
def test_connect_sample_step_equals_oracle_for_arbitrary_weights(oracle, cfg):
    from simulator import batch

    H, W, K = cfg
    n = 3000
    b = _random_connect_batch(cfg, n, 2)
    g = torch.Generator(device="cuda").manual_seed(1)
    probs = torch.rand((n, W), device="cuda", generator=g) ** 4
    probs[::7] = 0.0                      # all-zero rows: uniform fallback
    probs[1::7, 0] = float("nan")         # NaN counts as 0
    probs[2::7, W - 1] = float("inf")     # inf counts as FLT_MAX
    probs[3::7] = -1.0
    didx = torch.randint(0, 60, (n,), device="cuda", generator=g, dtype=torch.int32)
    gids = torch.randint(0, 2**40, (n,), device="cuda", generator=g)
    nb, a, status = b.sample_step(probs, seed=77, game_ids=gids, draw_index=didx)
    nb2, a2, _ = b.sample_step(probs, seed=77, game_id0=9)   # default draw index = stones on the board
    grid, winner, player = b.grid.cpu().numpy(), b.winner.cpu().numpy(), b.player.cpu().numpy()
    pr, di, gi = probs.cpu().numpy(), didx.cpu().numpy(), gids.cpu().numpy()
    a, a2, status = a.cpu().numpy(), a2.cpu().numpy(), status.cpu().numpy()
    ng, nw = nb.grid.cpu().numpy(), nb.winner.cpu().numpy()
    for i in range(n):
        col = oracle.connect_sample(grid[i], int(winner[i]), pr[i], 77, int(gi[i]), int(di[i]))
        assert a[i] == col and (status[i] == 1) == (col < 0), i
        stones = int((grid[i] >= 0).sum())
        assert a2[i] == oracle.connect_sample(grid[i], int(winner[i]), pr[i], 77, 9 + i, stones)
        if col >= 0:
            g2, p2, w2 = oracle.connect_next(grid[i], K, int(player[i]), int(winner[i]), col)
            np.testing.assert_array_equal(ng[i], g2)
            assert nw[i] == w2
    assert (a >= 0).sum() > n // 2


@pytest.mark.parametrize("rules", [0, 5])
def test_bounce_sample_step_with_equal_weights_replays_the_rollout_kernel(rules):
    from simulator import batch

    grid0 = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
    n, T = 2000, 40
    res = batch.bounce_rollout(grid0, n, 13, 200, max_plies=T, rules=rules, moves=True, final_grid=True)
    b = batch.BounceBatch.initial(grid0, n, rules=rules)
    b.ply = torch.zeros(n, dtype=torch.int32, device="cuda")
    probs = torch.full((n, 6, 54), 3.0, device="cuda")
    mv = res.actions.to(torch.int32)
    for t in range(T):
        b, move, status = b.sample_step(probs, seed=13, game_id0=200)
        played = mv[:, t, 0] != 255
        assert torch.equal(status == 0, played), f"ply {t}"
        src = move[:, 1] * 6 + move[:, 0]
        tgt = move[:, 3] * 6 + move[:, 2]
        assert torch.equal(src[played], mv[played, t, 0]) and torch.equal(tgt[played], mv[played, t, 1])
    done = res.winner != -2
    assert torch.equal(b.grid[done], res.final_grid[done]) and torch.equal(b.winner[done], torch.where(res.winner < 0, -1, res.winner).to(torch.int8)[done])


def test_bounce_sample_step_equals_oracle_for_arbitrary_weights(oracle):
    from simulator import batch

    grid0 = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
    n = 1500
    b = batch.BounceBatch.initial(grid0, n)
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(12):  # random mid-game positions (both sides to move)
        b, _, _ = b.sample_step(torch.rand((n, 6, 54), device="cuda", generator=g), seed=1, game_id0=0,
                                draw_index=torch.full((n,), t, dtype=torch.int32, device="cuda"))
    probs = torch.rand((n, 6, 54), device="cuda", generator=g) ** 3
    probs[::5] = 0.0
    didx = torch.randint(0, 300, (n,), device="cuda", generator=g, dtype=torch.int32)
    nb, move, status = b.sample_step(probs, seed=99, game_id0=1234, draw_index=didx)
    grid, player, ended = b.grid.cpu().numpy(), b.player.cpu().numpy(), b.has_ended.cpu().numpy()
    pr, di, move, status = probs.cpu().numpy(), didx.cpu().numpy(), move.cpu().numpy(), status.cpu().numpy()
    ng, nw, ne = nb.grid.cpu().numpy(), nb.winner.cpu().numpy(), nb.has_ended.cpu().numpy()
    for i in range(n):
        want = oracle.bounce_sample(grid[i], int(player[i]), bool(ended[i]), pr[i], 99, 1234 + i, int(di[i]))
        if want is None:
            assert status[i] == 1 and tuple(move[i]) == (-1, -1, -1, -1)
            continue
        assert tuple(int(x) for x in move[i]) == want and status[i] == 0, i
        g2, p2, w2, e2 = oracle.bounce_next(grid[i], int(player[i]), bool(ended[i]), *want)
        np.testing.assert_array_equal(ng[i], g2)
        assert nw[i] == w2 and bool(ne[i]) == e2
    assert (status == 0).sum() > n // 2


# ------------------------------------------------------------------------------------------ bulk JSON
def test_connect_bulk_json_round_trip(golden, oracle):
    """ConnectBatch.to_json() / from_json(list[dict], config) use the reference's State schema
    (tests/test_connect.py:130-139) and agree with the per-object API and the golden fixture."""
    from simulator import batch
    from simulator.game.connect import Config, State

    cfg = (6, 7, 4)
    b = _random_connect_batch(cfg, 300, 4)
    js = b.to_json()
    assert json.loads(json.dumps(js)) == js and set(js[0]) == {"grid", "player", "winner"}
    config = Config(*cfg)
    for i in (0, 17, 299):
        s = State.from_json(js[i], config)
        assert s.to_json() == js[i]
        np.testing.assert_array_equal(s.grid, b.grid[i].cpu().numpy())
    back = batch.ConnectBatch.from_json(js, config)
    assert bool(back.equal(b).all()) and torch.equal(back.has_ended, b.has_ended) and torch.equal(back.legal, b.legal)
    # the reference's own pictured / asserted JSON (tests/test_connect.py:118-145)
    gj = golden["connect"]["test_json"]["json"]
    one = batch.ConnectBatch.from_json([gj["state"]], Config.from_json(gj["config"]))
    assert one.to_json() == [gj["state"]]
    # a rollout as list-of-dicts: final State + the Action dicts of every ply
    res = batch.connect_rollout(cfg, 50, 3, 7, per_game=True, actions=True, final_grid=True)
    rj = res.to_json(config)
    ref = oracle.connect_rollout(*cfg, 50, gid0=7, seed=3)
    for i, d in enumerate(rj):
        assert d["actions"] == [{"column": int(c)} for c in ref["actions"][i, : ref["length"][i]]]
        assert d["state"]["grid"] == ref["final_grid"][i].tolist() and d["state"]["winner"] == int(ref["winner"][i])
        state = config.sample_initial_state()
        for a in d["actions"]:
            state = config.State.Action.from_json(a, state).sample_next_state() if i < 3 else state
        if i < 3:
            assert state.to_json() == d["state"]


def test_bounce_bulk_json_round_trip(golden):
    from simulator import batch
    from simulator.game.bounce import Config, State

    grid0 = np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)
    n = 200
    b = batch.BounceBatch.initial(grid0, n)
    g = torch.Generator(device="cuda").manual_seed(8)
    for t in range(25):
        b, _, _ = b.sample_step(torch.rand((n, 6, 54), device="cuda", generator=g), seed=2, game_id0=0,
                                draw_index=torch.full((n,), t, dtype=torch.int32, device="cuda"))
    js = b.to_json()
    assert json.loads(json.dumps(js)) == js and set(js[0]) == {"grid", "player", "winner"}
    config = Config(grid0)
    back = batch.BounceBatch.from_json(js, config)
    assert bool(back.equal(b).all()) and torch.equal(back.has_ended, b.has_ended)
    for i in (0, 5, 199):
        s = State.from_json(js[i], config)
        assert s.to_json() == js[i] and s.has_ended == bool(b.has_ended[i])
    gj = golden["bounce"]["test_json"]["json"]
    one = batch.BounceBatch.from_json([gj["state"]], Config.from_json(gj["config"]))
    assert one.to_json() == [gj["state"]]
    res = batch.bounce_rollout(grid0, 20, 3, 0, max_plies=64, moves=True, final_grid=True)
    rj = res.to_json(config)
    state = config.sample_initial_state()
    for a in rj[0]["actions"]:
        state = State.Action.from_json(a, state).sample_next_state()
    assert state.to_json()["grid"] == rj[0]["state"]["grid"]
