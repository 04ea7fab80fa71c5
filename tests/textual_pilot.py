"""Headless pilot of the reference's OWN Textual front ends against a `simulator.game` implementation.

    python tests/textual_pilot.py --backend dropin|standin

Runs the reference's widgets and example apps unmodified and in place
(/root/reference/src/simulator/textual/{connect,bounce}.py, examples/arena.py) with Textual's
App.run_test() pilot.  `simulator.game` is either the GPU-backed drop-in (needs a CUDA device) or the
oracle's object-API stand-in (CPU); `simulator.textual` is spliced in from the reference tree by extending
the package's __path__ -- nothing is copied.  The callers exercised are the ones SURVEY.md 3.3 lists:
  * ConnectBoard.select -> state.action_at(column) guarded by `except RuntimeError`  (textual/connect.py:111-119)
  * BounceBoard.render_line -> state.actions_at(...) per rendered line, Offset(*action.target) as (x, y)
    (textual/bounce.py:118-128, 222-223)
  * ArenaApp._handle_board -> agent.predict(state) -> random.choices(actions, weights) -> sample_next_state
    (textual/examples/arena.py:60-69)
Exit code 0 = every check passed.
"""
import argparse
import asyncio
import os
import random
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF_PKG = "/root/reference/src/simulator"


def setup(backend: str):
    if backend == "dropin":
        sys.path.insert(0, os.path.join(ROOT, "board-game-simulator-python_b200"))
    else:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "oracle", "pyapi"))
    import simulator

    simulator.__path__.append(REF_PKG)  # simulator.textual comes from the reference, simulator.game stays ours
    import simulator.game.connect as game_connect

    assert REF_PKG not in os.path.abspath(game_connect.__file__)
    return simulator


async def pilot_connect():
    import numpy as np
    from simulator.textual.connect import ConnectBoard, ExampleApp

    app = ExampleApp()
    async with app.run_test() as pilot:
        board = app.query_one(ConnectBoard)
        board.focus()
        await pilot.pause()
        assert board.state is not None and board.state.player == 0
        await pilot.press("right", "right", "enter")  # player 0 drops in column 2
        await pilot.pause()
        g = board.state.grid
        assert g[0, 2] == 0 and (g >= 0).sum() == 1 and board.state.player == 1, g
        for _ in range(5):  # fill column 2: 6 stones
            await pilot.press("enter")
            await pilot.pause()
        assert (board.state.grid[:, 2] >= 0).all()
        before = board.state
        await pilot.press("enter")  # full column: action_at raises RuntimeError, the widget swallows it
        await pilot.pause()
        assert board.state is before
        await pilot.press("left", "enter")
        await pilot.pause()
        assert board.state.grid[0, 1] == 0
        await pilot.press("r")  # reset
        await pilot.pause()
        assert (board.state.grid == -1).all() and board.state.player == 0
        # play a whole game from the keyboard: alternate columns 0 / 1 -> player 0 wins vertically in column... 
        board.cursor_column = 0
        for i in range(7):
            board.cursor_column = i % 2
            await pilot.press("enter")
            await pilot.pause()
        assert board.state.has_ended and list(board.state.reward) == [1.0, -1.0]
        final = board.state
        await pilot.press("enter")  # a finished game accepts no move
        await pilot.pause()
        assert board.state is final
        assert np.asarray(board.render_line(0)).size >= 0  # the widget renders the final position
    return "connect ok"


async def pilot_bounce():
    from textual.geometry import Offset

    from simulator.textual.bounce import BounceBoard, ExampleApp

    app = ExampleApp()
    async with app.run_test() as pilot:
        board = app.query_one(BounceBoard)
        board.focus()
        await pilot.pause()
        s0 = board.state
        assert s0.player == 0 and s0.grid.shape == (9, 6)
        # cursor starts at (0, 0): no piece there -> selecting does nothing
        await pilot.press("enter")
        await pilot.pause()
        assert board.source_offset is None and board.state is s0
        await pilot.press("up")  # (0, 1): the value-1 piece of player 0
        await pilot.press("enter")
        await pilot.pause()
        assert board.source_offset == Offset(0, 1)
        targets = {Offset(*a.target) for a in board.actions_at(Offset(0, 1))}
        assert Offset(0, 2) in targets  # one step forward
        for y in range(9):  # every line renders through actions_at (bounce.py:222-223)
            board.render_line(y)
        await pilot.press("up")  # (0, 2)
        await pilot.press("enter")
        await pilot.pause()
        s1 = board.state
        assert s1 is not s0 and s1.player == 1 and s1.grid[2, 0] == 1 and s1.grid[1, 0] == 0
        # player 1 must move a piece of row 7: a piece of player 0's row is not selectable
        board.cursor_offset = Offset(1, 1)
        await pilot.press("enter")
        await pilot.pause()
        assert board.source_offset is None
        board.cursor_offset = Offset(0, 7)
        await pilot.press("enter")
        await pilot.pause()
        assert board.source_offset == Offset(0, 7)
        await pilot.press("down", "enter")
        await pilot.pause()
        assert board.state.player == 0 and board.state.grid[6, 0] == 1
        await pilot.press("r")
        await pilot.pause()
        assert (board.state.grid == s0.grid).all()
    return "bounce ok"


async def pilot_arena():
    """ArenaApp._handle_board (arena.py:58-71) itself, on stub boards: `agent.predict` runs in pool threads
    (state.actions called off the main thread), the loop samples with random.choices and steps.  The App is
    not mounted: Textual 8 cannot render the reference's *disabled* boards (style-less segments meet the
    opacity filter -- a reference / Textual version matter, independent of the game API)."""
    import types
    from concurrent.futures import ThreadPoolExecutor

    from simulator.textual.examples import arena

    class StubBoard:
        def __init__(self):
            self.states = []
            self.styles = types.SimpleNamespace(border=None)

        state = property(lambda self: self.states[-1] if self.states else None,
                         lambda self, value: self.states.append(value))

    time_sleep = time.sleep
    arena.time.sleep = lambda s: None  # RandomAgent.predict sleeps up to 1 s per move
    try:
        with ThreadPoolExecutor(max_workers=4) as executor:
            app = arena.ArenaApp.__new__(arena.ArenaApp)
            app.agent, app.executor = arena.RandomAgent(), executor
            boards = [StubBoard() for _ in range(4)]
            tasks = [asyncio.create_task(arena.ArenaApp._handle_board(app, b)) for b in boards]
            for _ in range(600):
                await asyncio.sleep(0.05)
                if all(any(s.has_ended for s in b.states) for b in boards):
                    break
            for t in tasks:
                t.cancel()
            await asyncio.gather(*tasks, return_exceptions=True)
        for b in boards:
            ended = [s for s in b.states if s.has_ended]
            assert ended, "no game finished"
            assert sorted(ended[0].reward) in ([-1.0, 1.0], [0.0, 0.0]) and ended[0].actions == []
            assert b.styles.border == ("round", "red") or len(b.states) > 1
    finally:
        arena.time.sleep = time_sleep
    return "arena ok"


def agent_loop():
    """examples/agent.py + arena.py:60-69 without the UI: predict -> random.choices -> sample_next_state."""
    from simulator.textual.bounce import BounceBoard
    from simulator.textual.examples import arena

    arena.time.sleep = lambda s: None
    agent = arena.RandomAgent()
    random.seed(0)
    for _ in range(2):
        state = BounceBoard.DEFAULT_CONFIG.sample_initial_state()
        plies = 0
        while not state.has_ended and plies < 200:
            policy = agent.predict(state)
            assert abs(sum(policy.values()) - 1.0) < 1e-9
            actions, weights = zip(*policy.items())
            [action] = random.choices(actions, weights)
            state = action.sample_next_state()
            plies += 1
        assert plies > 0
    return "agent loop ok"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["dropin", "standin"], required=True)
    args = ap.parse_args()
    setup(args.backend)
    print(asyncio.run(pilot_connect()))
    print(asyncio.run(pilot_bounce()))
    print(agent_loop())
    print(asyncio.run(pilot_arena()))
    print("PILOT OK", args.backend)


if __name__ == "__main__":
    main()
