"""``simulator.game.connect`` -- Connect-k with the reference's object API, computed on the GPU.

Mirrors the nanobind module of the reference (src/simulator/game/connect.cpp:24-61, typed by
connect.pyi): ``Config(height, width, count)``, ``State``, ``Action`` with the same attribute and
method names, immutability, copy-out arrays and ``RuntimeError`` on illegal actions.

Every rule decision (legal columns, drop, k-in-a-row, terminal, reward) is made by the CUDA kernels
of ``libbgs_b200.so`` through ``simulator.batch.ConnectBatch`` with a batch of one; this module only
holds host copies of the results.  There is no CPU implementation of the rules in this package: on a
machine without a GPU, constructors / JSON / equality work and anything that needs the rules raises
``RuntimeError``.  For throughput use ``simulator.batch`` (or ``Config.rollout``), not this API.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .. import _native as N
from .. import batch as _batch


class Config:
    """Game parameters (reference connect.cpp:24-34)."""

    num_players = 2

    def __init__(self, height: int, width: int, count: int, /) -> None:
        self.height, self.width, self.count = int(height), int(width), int(count)
        if self.height < 1 or self.width < 1 or self.count < 1:
            raise ValueError("height, width and count must be positive")

    def _key(self):
        return (self.height, self.width, self.count)

    def sample_initial_state(self) -> "State":
        """The empty board, player 0 to move (tests/test_connect.py:24,33-38)."""
        grid = np.full((self.height, self.width), -1, dtype=np.int8)
        return State(self, grid, 0, -1)

    def rollout(self, n_games: int, seed: int = 0, game_id0: int = 0, **kwargs) -> _batch.RolloutResult:
        """Batched random rollouts on the GPU; see :func:`simulator.batch.connect_rollout`."""
        return _batch.connect_rollout(self, n_games, seed, game_id0, **kwargs)

    def to_json(self) -> dict[str, Any]:
        return {"height": self.height, "width": self.width, "count": self.count}

    @staticmethod
    def from_json(value: dict[str, Any]) -> "Config":
        return Config(value["height"], value["width"], value["count"])

    def __eq__(self, other):
        return isinstance(other, Config) and self._key() == other._key()

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __gt__(self, other):
        return self._key() > other._key()

    def __ge__(self, other):
        return self._key() >= other._key()

    def __hash__(self):
        return hash(("connect.Config",) + self._key())

    def __repr__(self):
        return f"Config({self.height}, {self.width}, {self.count})"


class State:
    """An immutable position (reference connect.cpp:36-46)."""

    def __init__(self, config: Config, grid, player: int, winner: int, _info=None) -> None:
        g = np.asarray(grid)
        if g.shape != (config.height, config.width):
            raise TypeError(f"grid must have shape {(config.height, config.width)}")
        self.config = config
        self._grid = np.ascontiguousarray(g, dtype=np.int8)
        self._player = int(player)
        self._winner = int(winner)
        self._info = _info  # (has_ended, legal bit mask, reward) as computed on the GPU

    def _key(self):
        return (self.config._key(), self._grid.tobytes(), self._player, self._winner)

    # -- GPU-computed facts about this state -----------------------------------------------------
    def _facts(self):
        if self._info is None:
            torch = N.require_cuda()
            b = _batch.ConnectBatch(
                self.config,
                torch.from_numpy(self._grid[None]).cuda(),
                torch.tensor([self._player], dtype=torch.int8, device="cuda"),
                torch.tensor([self._winner], dtype=torch.int8, device="cuda"),
            )
            self._info = (bool(b.has_ended.item()), int(b.legal.item()) & 0xFFFFFFFF, b.reward[0].cpu().numpy())
        return self._info

    @property
    def has_ended(self) -> bool:
        return self._facts()[0]

    @property
    def player(self) -> int:
        return self._player

    @property
    def reward(self) -> np.ndarray:
        return self._facts()[2].copy()

    @property
    def grid(self) -> np.ndarray:
        """A fresh int8[H,W] copy, row 0 = bottom, -1 / 0 / 1 (reference tensor.hpp:69-87)."""
        return self._grid.copy()

    @property
    def actions(self) -> list["Action"]:
        legal = self._facts()[1]
        return [Action(self, c) for c in range(self.config.width) if (legal >> c) & 1]

    def action_at(self, column: int) -> "Action":
        column = int(column)
        legal = self._facts()[1]
        if not (0 <= column < self.config.width) or not (legal >> column) & 1:
            raise RuntimeError(f"illegal action: column {column}")
        return Action(self, column)

    def to_json(self) -> dict[str, Any]:
        return {"grid": self._grid.tolist(), "player": self._player, "winner": self._winner}

    @staticmethod
    def from_json(value: dict[str, Any], config: Config) -> "State":
        grid = np.array(value["grid"], dtype=np.int8)
        return State(config, grid, value["player"], value["winner"])

    def __eq__(self, other):
        return isinstance(other, State) and self._key() == other._key()

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __gt__(self, other):
        return self._key() > other._key()

    def __ge__(self, other):
        return self._key() >= other._key()

    def __hash__(self):
        return hash(("connect.State",) + self._key())


class Action:
    """Dropping a stone in ``column`` (reference connect.cpp:48-54)."""

    def __init__(self, state: State, column: int) -> None:
        self.state = state
        self.column = int(column)

    def _key(self):
        return (self.state._key(), self.column)

    def sample_next_state(self) -> State:
        torch = N.require_cuda()
        s = self.state
        b = _batch.ConnectBatch(
            s.config,
            torch.from_numpy(s._grid[None]).cuda(),
            torch.tensor([s._player], dtype=torch.int8, device="cuda"),
            torch.tensor([s._winner], dtype=torch.int8, device="cuda"),
            has_ended=False,  # facts of the OLD state are not needed for the transition
        )
        nxt, status = b.step(torch.tensor([self.column], dtype=torch.int32, device="cuda"))
        if int(status.item()) != 0:
            raise RuntimeError(f"illegal action: column {self.column}")
        info = (bool(nxt.has_ended.item()), int(nxt.legal.item()) & 0xFFFFFFFF, nxt.reward[0].cpu().numpy())
        return State(s.config, nxt.grid[0].cpu().numpy(), int(nxt.player.item()), int(nxt.winner.item()), info)

    def to_json(self) -> dict[str, Any]:
        return {"column": self.column}

    @staticmethod
    def from_json(value: dict[str, Any], state: State) -> "Action":
        return Action(state, value["column"])

    def __eq__(self, other):
        return isinstance(other, Action) and self._key() == other._key()

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __gt__(self, other):
        return self._key() > other._key()

    def __ge__(self, other):
        return self._key() >= other._key()

    def __hash__(self):
        return hash(("connect.Action",) + self._key())


Config.State = State
State.Action = Action
