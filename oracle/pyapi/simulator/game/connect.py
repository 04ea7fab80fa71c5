"""Oracle-backed ``simulator.game.connect`` (API of reference connect.cpp:24-61 / connect.pyi).

TEST INFRASTRUCTURE ONLY: every rule decision is delegated to ``oracle/bgs_oracle.c``.
"""
from __future__ import annotations

import numpy as np

from .._oracle import binding as _o


class Config:
    num_players = 2

    def __init__(self, height, width, count, /):
        self.height, self.width, self.count = int(height), int(width), int(count)

    def _key(self):
        return (self.height, self.width, self.count)

    def sample_initial_state(self):
        return State(self, np.full((self.height, self.width), -1, dtype=np.int8), 0, -1)

    def to_json(self):
        return {"height": self.height, "width": self.width, "count": self.count}

    @staticmethod
    def from_json(value):
        return Config(value["height"], value["width"], value["count"])

    def __eq__(self, other):
        return isinstance(other, Config) and self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __hash__(self):
        return hash(self._key())


class State:
    def __init__(self, config, grid, player, winner):
        self.config = config
        self._grid = np.ascontiguousarray(grid, dtype=np.int8)
        self._player = int(player)
        self._winner = int(winner)

    def _key(self):
        return (self.config._key(), self._grid.tobytes(), self._player, self._winner)

    @property
    def has_ended(self):
        return _o.connect_ended(self._grid, self._winner)

    @property
    def player(self):
        return self._player

    @property
    def reward(self):
        return _o.reward(self._winner)

    @property
    def grid(self):
        return self._grid.copy()

    @property
    def actions(self):
        return [Action(self, c) for c in _o.connect_actions(self._grid, self._winner)]

    def action_at(self, column):
        column = int(column)
        if column not in _o.connect_actions(self._grid, self._winner):
            raise RuntimeError(f"illegal action: column {column}")
        return Action(self, column)

    def to_json(self):
        return {"grid": self._grid.tolist(), "player": self._player, "winner": self._winner}

    @staticmethod
    def from_json(value, config):
        grid = np.array(value["grid"], dtype=np.int8).reshape(config.height, config.width)
        return State(config, grid, value["player"], value["winner"])

    def __eq__(self, other):
        return isinstance(other, State) and self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __hash__(self):
        return hash(self._key())


class Action:
    def __init__(self, state, column):
        self.state = state
        self.column = int(column)

    def _key(self):
        return (self.state._key(), self.column)

    def sample_next_state(self):
        s = self.state
        out = _o.connect_next(s._grid, s.config.count, s._player, s._winner, self.column)
        if out is None:
            raise RuntimeError(f"illegal action: column {self.column}")
        return State(s.config, *out)

    def to_json(self):
        return {"column": self.column}

    @staticmethod
    def from_json(value, state):
        return state.action_at(value["column"])

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
