"""Oracle-backed ``simulator.game.bounce`` (API of reference bounce.cpp:24-60 / bounce.pyi).

TEST INFRASTRUCTURE ONLY: every rule decision is delegated to ``oracle/bgs_oracle.c``.
"""
from __future__ import annotations

import os

import numpy as np

from .._oracle import binding as _o

# rule-variant switches left unpinned by the reference tests; the env var exists so
# that tests can show the reference's tests pass under every variant.
RULES = int(os.environ.get("BGS_ORACLE_BOUNCE_RULES", "0"))


def _xy(a):
    a = np.asarray(a)
    if a.shape != (2,):
        raise TypeError("expected an (x, y) array of shape (2,)")
    return int(a[0]), int(a[1])


class Config:
    num_players = 2

    def __init__(self, grid, /):
        g = np.asarray(grid)
        if g.ndim != 2:
            raise TypeError("grid must be a 2-D array")
        self._grid = np.ascontiguousarray(g, dtype=np.int8)

    def _key(self):
        return (self._grid.shape, self._grid.tobytes())

    @property
    def grid(self):
        return self._grid.copy()

    def sample_initial_state(self):
        return State(self, self._grid, 0, -1, False)

    def to_json(self):
        return {"grid": self._grid.tolist()}

    @staticmethod
    def from_json(value):
        return Config(np.array(value["grid"], dtype=np.int8))

    def __eq__(self, other):
        return isinstance(other, Config) and self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __hash__(self):
        return hash(self._key())


class State:
    def __init__(self, config, grid, player, winner, ended):
        self.config = config
        self._grid = np.ascontiguousarray(grid, dtype=np.int8)
        self._player = int(player)
        self._winner = int(winner)
        self._ended = bool(ended) or self._winner >= 0

    def _key(self):
        return (self.config._key(), self._grid.tobytes(), self._player, self._winner, self._ended)

    @property
    def has_ended(self):
        return self._ended

    @property
    def player(self):
        return self._player

    @property
    def reward(self):
        return _o.reward(self._winner)

    @property
    def grid(self):
        return self._grid.copy()

    def _moves(self):
        return _o.bounce_actions(self._grid, self._player, self._ended, RULES)

    @property
    def actions(self):
        return [Action(self, (m[0], m[1]), (m[2], m[3])) for m in self._moves()]

    def actions_at(self, source):
        sx, sy = _xy(source)
        H, W = self._grid.shape
        if not (0 <= sx < W and 0 <= sy < H):
            raise RuntimeError(f"source {(sx, sy)} is outside the board")
        return [Action(self, (m[0], m[1]), (m[2], m[3])) for m in self._moves() if (m[0], m[1]) == (sx, sy)]

    def action_at(self, source, target):
        sx, sy = _xy(source)
        tx, ty = _xy(target)
        for m in self._moves():
            if tuple(m) == (sx, sy, tx, ty):
                return Action(self, (sx, sy), (tx, ty))
        raise RuntimeError(f"illegal action: {(sx, sy)} -> {(tx, ty)}")

    def to_json(self):
        return {"grid": self._grid.tolist(), "player": self._player, "winner": self._winner}

    @staticmethod
    def from_json(value, config):
        grid = np.array(value["grid"], dtype=np.int8).reshape(config._grid.shape)
        player, winner = int(value["player"]), int(value["winner"])
        # a draw is stored as winner == -1; "ended" is then recovered as "the mover has no action"
        ended = winner >= 0 or len(_o.bounce_actions(grid, player, False, RULES)) == 0
        return State(config, grid, player, winner, ended)

    def __eq__(self, other):
        return isinstance(other, State) and self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __hash__(self):
        return hash(self._key())


class Action:
    def __init__(self, state, source, target):
        self.state = state
        self._source = (int(source[0]), int(source[1]))
        self._target = (int(target[0]), int(target[1]))

    def _key(self):
        return (self.state._key(), self._source, self._target)

    @property
    def source(self):
        return np.array(self._source, dtype=np.int64)

    @property
    def target(self):
        return np.array(self._target, dtype=np.int64)

    def sample_next_state(self):
        s = self.state
        out = _o.bounce_next(s._grid, s._player, s._ended, *self._source, *self._target, RULES)
        if out is None:
            raise RuntimeError(f"illegal action: {self._source} -> {self._target}")
        return State(s.config, *out)

    def to_json(self):
        return {"source": list(self._source), "target": list(self._target)}

    @staticmethod
    def from_json(value, state):
        return state.action_at(np.array(value["source"]), np.array(value["target"]))

    def __eq__(self, other):
        return isinstance(other, Action) and self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __hash__(self):
        return hash(self._key())


Config.State = State
State.Action = Action
