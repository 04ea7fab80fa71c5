"""``simulator.game.bounce`` -- Bounce with the reference's object API, computed on the GPU.

Mirrors the nanobind module of the reference (src/simulator/game/bounce.cpp:24-60, typed by
bounce.pyi): ``Config(grid)``, ``State`` (``actions``, ``actions_at(source)``, ``action_at(source,
target)``), ``Action`` (``source``, ``target``, ``sample_next_state``); coordinates are ``(x, y)``
arrays (tests/test_bounce.py:33-35,406-409).

Move generation and transitions run in ``libbgs_b200.so`` through ``simulator.batch.BounceBatch``
with a batch of one.  There is no CPU implementation of the rules in this package.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .. import _native as N
from .. import batch as _batch

#: rule-variant switches the reference tests leave unpinned (include/bgs_b200.h BGS_BOUNCE_*)
RULES = 0


def _xy(a) -> tuple[int, int]:
    arr = np.asarray(a)
    if arr.shape != (2,):
        raise TypeError("expected an (x, y) array of shape (2,)")
    return int(arr[0]), int(arr[1])


class Config:
    """The start position (reference bounce.cpp:24-31)."""

    num_players = 2

    def __init__(self, grid, /) -> None:
        g = np.asarray(grid)
        if g.ndim != 2:
            raise TypeError("grid must be a 2-D array")
        self._grid = np.ascontiguousarray(g, dtype=np.int8)

    def _key(self):
        return (self._grid.shape, self._grid.tobytes())

    @property
    def grid(self) -> np.ndarray:
        return self._grid.copy()

    def sample_initial_state(self) -> "State":
        """The config grid, player 0 to move (tests/test_bounce.py:50-60)."""
        return State(self, self._grid, 0, -1)

    def rollout(self, n_games: int, seed: int = 0, game_id0: int = 0, **kwargs) -> _batch.RolloutResult:
        """Batched random rollouts on the GPU; see :func:`simulator.batch.bounce_rollout`."""
        kwargs.setdefault("rules", RULES)
        return _batch.bounce_rollout(self, n_games, seed, game_id0, **kwargs)

    def to_json(self) -> dict[str, Any]:
        return {"grid": self._grid.tolist()}

    @staticmethod
    def from_json(value: dict[str, Any]) -> "Config":
        return Config(np.array(value["grid"], dtype=np.int8))

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
        return hash(("bounce.Config",) + self._key())


class State:
    """An immutable position (reference bounce.cpp:33-44).

    ``ended`` may be None (unknown, e.g. after ``from_json`` of a drawn game, where winner == -1):
    it is then recovered on the GPU as "the side to move has no action"."""

    def __init__(self, config: Config, grid, player: int, winner: int, ended: bool | None = False) -> None:
        g = np.asarray(grid)
        if g.shape != config._grid.shape:
            raise TypeError(f"grid must have shape {config._grid.shape}")
        self.config = config
        self._grid = np.ascontiguousarray(g, dtype=np.int8)
        self._player = int(player)
        self._winner = int(winner)
        self._ended = True if self._winner >= 0 else ended
        self._moves = None  # (source_row, [target mask per source column]) from the GPU

    def _key(self):
        return (self.config._key(), self._grid.tobytes(), self._player, self._winner)

    def _movegen(self):
        if self._moves is None:
            torch = N.require_cuda()
            b = _batch.BounceBatch(
                torch.from_numpy(self._grid[None]).cuda(),
                torch.tensor([self._player], dtype=torch.int8, device="cuda"),
                torch.tensor([self._winner], dtype=torch.int8, device="cuda"),
                torch.tensor([1 if self._ended else 0], dtype=torch.uint8, device="cuda"),
                RULES,
            )
            _batch._bounce_check(N.lib(), self._grid)
            row, targets, count = b.moves()
            t0 = targets[0].cpu().tolist()
            if targets.dim() == 3:  # boards on 128-bit words: (low, high) per source column
                masks = [(int(lo) & 0xFFFFFFFFFFFFFFFF) | ((int(hi) & 0xFFFFFFFFFFFFFFFF) << 64) for lo, hi in t0]
            else:
                masks = [int(m) & 0xFFFFFFFFFFFFFFFF for m in t0]
            self._moves = (int(row.item()), masks)
            if self._ended is None:
                self._ended = int(count.item()) == 0
        return self._moves

    @property
    def has_ended(self) -> bool:
        if self._ended is None:
            self._movegen()
        return bool(self._ended)

    @property
    def player(self) -> int:
        return self._player

    @property
    def reward(self) -> np.ndarray:
        w = self._winner
        return np.array([1 if w == 0 else (-1 if w == 1 else 0), 1 if w == 1 else (-1 if w == 0 else 0)], dtype=np.float32)

    @property
    def grid(self) -> np.ndarray:
        return self._grid.copy()

    def _pairs(self, only_sx: int | None = None):
        if self.has_ended:
            return []
        row, masks = self._movegen()
        H, W = self._grid.shape
        out = []
        for sx, m in enumerate(masks):
            if only_sx is not None and sx != only_sx:
                continue
            cell = 0
            while m:
                if m & 1:
                    out.append(((sx, row), (cell % W, cell // W)))
                m >>= 1
                cell += 1
        return out

    @property
    def actions(self) -> list["Action"]:
        """Every (source, target) of the mover, ascending (sy, sx, ty, tx); empty when ended
        (tests/test_bounce.py:151)."""
        return [Action(self, s, t) for s, t in self._pairs()]

    def actions_at(self, source) -> list["Action"]:
        sx, sy = _xy(source)
        H, W = self._grid.shape
        if not (0 <= sx < W and 0 <= sy < H):
            raise RuntimeError(f"source {(sx, sy)} is outside the board")
        if self.has_ended or sy != self._movegen()[0]:
            return []
        return [Action(self, s, t) for s, t in self._pairs(sx)]

    def action_at(self, source, target) -> "Action":
        s, t = _xy(source), _xy(target)
        H, W = self._grid.shape
        ok = 0 <= s[0] < W and 0 <= s[1] < H and 0 <= t[0] < W and 0 <= t[1] < H and not self.has_ended
        if ok:
            row, masks = self._movegen()
            ok = s[1] == row and (masks[s[0]] >> (t[1] * W + t[0])) & 1
        if not ok:
            raise RuntimeError(f"illegal action: {s} -> {t}")
        return Action(self, s, t)

    def to_json(self) -> dict[str, Any]:
        return {"grid": self._grid.tolist(), "player": self._player, "winner": self._winner}

    @staticmethod
    def from_json(value: dict[str, Any], config: Config) -> "State":
        grid = np.array(value["grid"], dtype=np.int8)
        winner = int(value["winner"])
        return State(config, grid, value["player"], winner, True if winner >= 0 else None)

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
        return hash(("bounce.State",) + self._key())


class Action:
    """Moving the piece on ``source`` to ``target`` (reference bounce.cpp:46-53)."""

    def __init__(self, state: State, source, target) -> None:
        self.state = state
        self._source = (int(source[0]), int(source[1]))
        self._target = (int(target[0]), int(target[1]))

    def _key(self):
        return (self.state._key(), self._source, self._target)

    @property
    def source(self) -> np.ndarray:
        return np.array(self._source, dtype=np.int64)

    @property
    def target(self) -> np.ndarray:
        return np.array(self._target, dtype=np.int64)

    def sample_next_state(self) -> State:
        torch = N.require_cuda()
        s = self.state
        _batch._bounce_check(N.lib(), s._grid)
        b = _batch.BounceBatch(
            torch.from_numpy(s._grid[None]).cuda(),
            torch.tensor([s._player], dtype=torch.int8, device="cuda"),
            torch.tensor([s._winner], dtype=torch.int8, device="cuda"),
            torch.tensor([1 if s.has_ended else 0], dtype=torch.uint8, device="cuda"),
            RULES,
        )
        move = torch.tensor([[*self._source, *self._target]], dtype=torch.int32, device="cuda")
        nxt, status = b.step(move)
        if int(status.item()) != 0:
            raise RuntimeError(f"illegal action: {self._source} -> {self._target}")
        return State(
            s.config, nxt.grid[0].cpu().numpy(), int(nxt.player.item()), int(nxt.winner.item()),
            bool(nxt.has_ended.item()),
        )

    def to_json(self) -> dict[str, Any]:
        return {"source": list(self._source), "target": list(self._target)}

    @staticmethod
    def from_json(value: dict[str, Any], state: State) -> "Action":
        return Action(state, value["source"], value["target"])

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
        return hash(("bounce.Action",) + self._key())


Config.State = State
State.Action = Action
