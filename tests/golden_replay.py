"""Replays the golden fixture (tests/golden/reference_positions.json) through any module pair that
implements the reference API -- used with the oracle stand-in on CPU and with the product on GPU."""
import numpy as np


def replay_connect(mod, golden):
    n = 0
    for name, rec in golden["connect"].items():
        if not rec["steps"]:
            continue
        config = mod.Config(*rec["config"])
        state = config.sample_initial_state()
        for step in rec["steps"]:
            np.testing.assert_array_equal(np.array(step["grid"]), state.grid)
            if state.has_ended:
                assert step["column"] is None
            else:
                assert step["player"] == state.player
            n += 1
            if step["column"] is not None:
                state = state.action_at(step["column"]).sample_next_state()
        assert state.has_ended
        np.testing.assert_array_equal(state.reward, rec["final"]["reward"])
    return n


def replay_connect_json(mod, golden):
    rec = golden["connect"]["test_json"]
    config = mod.Config(*rec["config"])
    assert config.to_json() == rec["json"]["config"]
    assert mod.Config.from_json(config.to_json()) == config
    state = config.sample_initial_state().action_at(0).sample_next_state()
    assert state.to_json() == rec["json"]["state"]
    assert mod.State.from_json(state.to_json(), config) == state
    action = state.action_at(1)
    assert action.to_json() == rec["json"]["action"]
    assert mod.Action.from_json(action.to_json(), state) == action


def replay_bounce(mod, golden):
    n = 0
    for name, rec in golden["bounce"].items():
        state = None
        for step in rec["steps"]:
            grid = np.array(step["grid"], dtype=np.int8)
            if state is None:
                state = mod.Config(grid).sample_initial_state()
            np.testing.assert_array_equal(grid, state.grid)
            assert step["player"] == state.player
            n += 1
            action = None
            if step["source"] is not None:
                actions = {tuple(int(v) for v in a.target): a for a in state.actions_at(np.array(step["source"]))}
                assert set(map(tuple, step["targets"])) == set(actions), (name, step["source"])
                # the same pairs must be listed by state.actions and accepted by action_at
                listed = {(tuple(int(v) for v in a.source), tuple(int(v) for v in a.target)) for a in state.actions}
                assert {(tuple(step["source"]), t) for t in actions} <= listed
                if step["target"] is not None:
                    action = actions[tuple(step["target"])]
                    assert state.action_at(np.array(step["source"]), np.array(step["target"])) == action
            if action is not None:
                state = action.sample_next_state()
        fin = rec["final"]
        if fin:
            assert state.has_ended == fin["has_ended"]
            assert len(state.actions) == fin["n_actions"]
            assert state.reward.tolist() == fin["reward"]
    return n


def replay_bounce_json(mod, golden):
    rec = golden["bounce"]["test_json"]
    step = rec["steps"][0]
    config = mod.Config(np.array(step["grid"], dtype=np.int8))
    state = config.sample_initial_state()
    assert config.to_json() == rec["json"]["config"]
    assert mod.Config.from_json(config.to_json()) == config
    assert state.to_json() == rec["json"]["state"]
    assert mod.State.from_json(state.to_json(), config) == state
    action = state.action_at(np.array(step["source"]), np.array(step["target"]))
    assert action.to_json() == rec["json"]["action"]
    assert mod.Action.from_json(action.to_json(), state) == action
