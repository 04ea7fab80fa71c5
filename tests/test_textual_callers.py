"""the reference's interactive callers on the drop-in.

The reference's Textual widgets and example apps live in /root/reference/src/simulator/textual -- they are
run unmodified and in place by tests/textual_pilot.py (never copied).  That tree exists in the build
container (no GPU) but not on the GPU boxes, so:
  * against the oracle's object-API stand-in the pilot runs HERE (CPU): it pins the harness and the API
    surface the widgets rely on;
  * against the GPU-backed drop-in it runs wherever both a CUDA device and the reference tree are present
    (a maintainer's machine; it skips on the driver's boxes);
  * on the GPU boxes the same call patterns are driven through a small Textual app of our own
    (test_own_textual_app_drives_the_dropin) and as plain calls (test_caller_patterns_on_the_dropin).
"""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from conftest import DEFAULT_BOUNCE_GRID, ROOT

REF_TEXTUAL = "/root/reference/src/simulator/textual"
PILOT = os.path.join(ROOT, "tests", "textual_pilot.py")

textual = pytest.importorskip("textual")
needs_reference = pytest.mark.skipif(not os.path.isdir(REF_TEXTUAL), reason="the reference tree is not on this machine")


def _run_pilot(backend):
    env = dict(os.environ, PYTHONPATH="")
    out = subprocess.run([sys.executable, PILOT, "--backend", backend], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and f"PILOT OK {backend}" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@needs_reference
def test_reference_widgets_on_the_oracle_standin():
    _run_pilot("standin")


@needs_reference
@pytest.mark.gpu
def test_reference_widgets_on_the_dropin():
    _run_pilot("dropin")


@pytest.mark.gpu
def test_caller_patterns_on_the_dropin():
    """The three caller patterns of SURVEY.md 3.3 as plain calls on simulator.game.* (GPU)."""
    from simulator.game import bounce, connect

    # textual/connect.py:111-119 -- action_at guarded by RuntimeError; a click on a full column is ignored
    state = connect.Config(6, 7, 4).sample_initial_state()
    for _ in range(6):
        state = state.action_at(3).sample_next_state()
    with pytest.raises(RuntimeError):
        state.action_at(3)
    with pytest.raises(RuntimeError):
        state.action_at(7)
    # textual/bounce.py:118-128, 222-223 -- actions_at per rendered line, action.target unpacked as (x, y)
    bstate = bounce.Config(np.array(DEFAULT_BOUNCE_GRID, dtype=np.int8)).sample_initial_state()
    for y in range(9):
        for x in (0, 5):
            try:
                acts = bstate.actions_at(np.array((x, y)))
            except RuntimeError:
                acts = []
            targets = {tuple(int(v) for v in a.target) for a in acts}
            assert (len(targets) > 0) == (y == 1)
            for tx, ty in targets:
                assert 0 <= tx < 6 and 0 <= ty < 9 and bstate.grid[ty, tx] == 0
    try:
        bad = bstate.action_at(np.array((0, 1)), np.array((5, 8)))
    except RuntimeError:
        bad = None
    assert bad is None
    # textual/examples/arena.py:60-69 + agent.py:13-27 -- policy dict -> random.choices -> sample_next_state
    random.seed(1)
    plies = 0
    while not bstate.has_ended and plies < 300:
        actions = bstate.actions
        policy = {a: 1 / len(actions) for a in actions}  # Action is hashable (helper.hpp:10-25)
        acts, weights = zip(*policy.items())
        [action] = random.choices(acts, weights)
        bstate = action.sample_next_state()
        plies += 1
    assert plies > 0 and (bstate.has_ended or plies == 300)
    if bstate.has_ended:
        assert bstate.actions == [] and sorted(bstate.reward.tolist()) in ([-1.0, 1.0], [0.0, 0.0])


@pytest.mark.gpu
def test_own_textual_app_drives_the_dropin():
    """A minimal Textual app of our own (not the reference's widget) under App.run_test(): key presses ->
    action_at / RuntimeError / sample_next_state on the GPU-backed objects, rendering from state.grid."""
    import asyncio

    from textual.app import App, ComposeResult
    from textual.widgets import Static

    from simulator.game.connect import Config

    class Board(Static, can_focus=True):
        BINDINGS = [("enter", "drop", "Drop"), ("right", "right", "Right"), ("r", "restart", "Restart")]

        def __init__(self):
            super().__init__("")
            self.config = Config(4, 5, 3)
            self.state = self.config.sample_initial_state()
            self.column, self.refused = 0, 0

        def picture(self):
            return "\n".join(" ".join(".OX"[int(v) + 1] for v in row) for row in self.state.grid[::-1])

        def on_mount(self):
            self.update(self.picture())

        def action_right(self):
            self.column = (self.column + 1) % self.config.width

        def action_restart(self):
            self.state = self.config.sample_initial_state()
            self.update(self.picture())

        def action_drop(self):
            try:
                action = self.state.action_at(self.column)
            except RuntimeError:
                self.refused += 1
                return
            self.state = action.sample_next_state()
            self.update(self.picture())

    class Mini(App):
        def compose(self) -> ComposeResult:
            yield Board()

    async def drive():
        app = Mini()
        async with app.run_test() as pilot:
            board = app.query_one(Board)
            board.focus()
            await pilot.press("enter", "enter", "enter", "enter")  # column 0 full after 4 stones
            await pilot.pause()
            assert (board.state.grid[:, 0] >= 0).all() and board.refused == 0
            await pilot.press("enter")
            await pilot.pause()
            assert board.refused == 1
            await pilot.press("r", "enter", "right", "enter", "right", "right")  # O . / X at 1
            for col in (0, 1, 0):  # O wins vertically? no: count 3 -> play O:0, X:1, O:0, X:1, O:0
                board.column = col
                await pilot.press("enter")
            await pilot.pause()
            assert board.state.has_ended and board.state.reward.tolist() == [1.0, -1.0]
            assert "O" in board.picture()
            await pilot.press("enter")
            await pilot.pause()
            assert board.refused == 2
        return True

    assert asyncio.run(drive())
