"""Host-side logic of the product's drop-in package (no GPU): surface, JSON schemas, equality,
ordering / hashing, argument errors, and loud failure of everything that needs the rules."""
import inspect

import numpy as np
import pytest
import torch

from simulator.game import ActionLike, ConfigLike, StateLike  # noqa: F401  (reference game/__init__.py:1)
from simulator.game import bounce, connect

from conftest import DEFAULT_BOUNCE_GRID

NO_GPU = not torch.cuda.is_available()


def test_surface_matches_reference_stubs():
    # reference connect.pyi / bounce.pyi / connect.cpp:24-61 / bounce.cpp:24-60
    for mod in (connect, bounce):
        assert mod.Config.num_players == 2
        assert mod.Config.State is mod.State and mod.State.Action is mod.Action
        for name in ("sample_initial_state", "to_json", "from_json"):
            assert hasattr(mod.Config, name)
        for name in ("has_ended", "player", "reward", "grid", "actions", "action_at", "to_json", "from_json"):
            assert hasattr(mod.State, name)
        for name in ("sample_next_state", "to_json", "from_json"):
            assert hasattr(mod.Action, name)
        for cls in (mod.Config, mod.State, mod.Action):
            for op in ("__eq__", "__ne__", "__lt__", "__le__", "__gt__", "__ge__", "__hash__"):
                assert op in vars(cls)
    assert hasattr(bounce.State, "actions_at")
    assert isinstance(inspect.getattr_static(bounce.Action, "source"), property)
    assert isinstance(inspect.getattr_static(bounce.Action, "target"), property)
    # Config(height, width, count) is positional-only in the reference (connect.cpp:26-27)
    with pytest.raises(TypeError):
        connect.Config(height=6, width=7, count=4)


def test_protocols_are_satisfied():
    # reference src/simulator/game/protocol.py:8-29: ConfigLike / StateLike / ActionLike
    c = connect.Config(6, 7, 4)
    s = c.sample_initial_state()
    assert isinstance(c, ConfigLike) and isinstance(connect.Action(s, 0), ActionLike)
    b = bounce.Config(np.array(DEFAULT_BOUNCE_GRID))
    assert isinstance(b, ConfigLike) and isinstance(bounce.Action(b.sample_initial_state(), (0, 1), (0, 2)), ActionLike)
    for cls in (connect.State, bounce.State):
        for name in ("config", "has_ended", "player", "reward", "actions", "Action"):
            assert hasattr(cls, name) or name == "config"


def test_connect_config_and_json(golden):
    rec = golden["connect"]["test_json"]["json"]
    c = connect.Config(2, 3, 2)
    assert (c.height, c.width, c.count) == (2, 3, 2)
    assert c.to_json() == rec["config"]
    assert connect.Config.from_json(c.to_json()) == c
    assert c != connect.Config(2, 3, 3) and c < connect.Config(2, 3, 3) and hash(c) == hash(connect.Config(2, 3, 2))
    s = connect.State.from_json(rec["state"], c)
    assert s.to_json() == rec["state"] and s.player == 1 and s.config is c
    assert s == connect.State.from_json(rec["state"], c) and hash(s) == hash(connect.State.from_json(rec["state"], c))
    g = s.grid
    g[0, 0] = 5
    assert s.grid[0, 0] == 0  # arrays cross the boundary by copy (reference tensor.hpp:69-87)
    a = connect.Action.from_json(rec["action"], s)
    assert a.to_json() == rec["action"] and a.column == 1 and a.state is s
    assert a == connect.Action(s, 1) and a != connect.Action(s, 2)
    s0 = c.sample_initial_state()
    assert s0.player == 0 and (s0.grid == -1).all() and s0.grid.dtype == np.int8 and s0 != s
    with pytest.raises(TypeError):
        connect.State(c, np.zeros((3, 3), np.int8), 0, -1)


def test_bounce_config_and_json(golden):
    rec = golden["bounce"]["test_json"]["json"]
    c = bounce.Config(np.array(rec["config"]["grid"]))
    assert c.to_json() == rec["config"] and bounce.Config.from_json(c.to_json()) == c
    assert c.grid.dtype == np.int8
    s = c.sample_initial_state()
    assert s.to_json() == rec["state"] and s.player == 0
    assert bounce.State.from_json(s.to_json(), c) == s
    a = bounce.Action.from_json(rec["action"], s)
    assert a.to_json() == rec["action"]
    assert a.source.tolist() == [1, 1] and a.target.tolist() == [0, 2]  # (x, y) order
    assert a == bounce.Action(s, (1, 1), (0, 2)) and a != bounce.Action(s, (1, 1), (1, 3))
    with pytest.raises(TypeError):
        bounce.Config(np.zeros(5))
    d = bounce.Config(np.array(DEFAULT_BOUNCE_GRID))
    assert d != c and {c: 1, d: 2}[bounce.Config(np.array(DEFAULT_BOUNCE_GRID))] == 2


@pytest.mark.skipif(not NO_GPU, reason="checks the behaviour WITHOUT a GPU")
def test_rules_need_the_gpu_and_fail_loudly():
    s = connect.Config(6, 7, 4).sample_initial_state()
    for fn in (lambda: s.actions, lambda: s.has_ended, lambda: s.reward, lambda: s.action_at(0),
               lambda: connect.Action(s, 0).sample_next_state(), lambda: s.config.rollout(10)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn()
    b = bounce.Config(np.array(DEFAULT_BOUNCE_GRID)).sample_initial_state()
    for fn in (lambda: b.actions, lambda: b.actions_at(np.array([0, 1])), lambda: b.config.rollout(10),
               lambda: bounce.Action(b, (0, 1), (0, 2)).sample_next_state()):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn()


def test_shard_range_partitions_exactly():
    from simulator.batch import shard_range

    for n in (0, 1, 7, 16 * 2**20, 10**6 + 3):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (s0, c0), (s1, _) in zip(parts, parts[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
