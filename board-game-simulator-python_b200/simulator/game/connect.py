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

import threading
from typing import Any

import numpy as np

from .. import _native as N
from .. import batch as _batch


class _OneState:
    """Staging for the single-object API: ONE pinned host record in, ONE out, per call.

    Every property of the reference's objects is computed by the CUDA kernels with a batch of one.  A naive
    wrapper pays a small host<->device copy and a synchronisation per field (grid, player, winner, column in;
    grid, player, winner, ended, legal, reward, status out: ~10 of them, ~230 us per ply); here the inputs of a
    call travel as one pinned record, the outputs come back as one, and the C ABI gets raw pointers into the two
    device records.  Layout (bytes): grid[HW] | player | winner | pad -> 16-byte aligned int32 action;
    outputs: grid[HW] | player | winner | ended | pad -> aligned legal u32 | status i32 | reward f32[2]."""

    _cache: dict = {}

    def __init__(self, H: int, W: int):
        torch = N.require_cuda()
        self.torch = torch
        HW = H * W
        self.HW = HW
        self.o_act = (HW + 2 + 15) // 16 * 16
        self.in_bytes = self.o_act + 16
        self.o_words = (HW + 3 + 15) // 16 * 16
        self.out_bytes = self.o_words + 16
        self.in_host = torch.zeros(self.in_bytes, dtype=torch.uint8).pin_memory()
        self.out_host = torch.zeros(self.out_bytes, dtype=torch.uint8).pin_memory()
        self.in_dev = torch.zeros(self.in_bytes, dtype=torch.uint8, device="cuda")
        self.out_dev = torch.zeros(self.out_bytes, dtype=torch.uint8, device="cuda")
        self.in_np = self.in_host.numpy()
        self.out_np = self.out_host.numpy()
        #: the staging records are shared by every object of this board size: one call at a time (the
        #: reference's example apps call state.actions from pool threads, textual/examples/agent.py:61)
        self.lock = threading.Lock()

    @classmethod
    def get(cls, H: int, W: int) -> "_OneState":
        torch = N.require_cuda()
        key = (H, W, torch.cuda.current_device())
        if key not in cls._cache:
            cls._cache[key] = cls(H, W)
        return cls._cache[key]

    def load(self, grid, player, winner, column=0):
        HW = self.HW
        self.in_np[:HW] = grid.reshape(-1).view(np.uint8)
        self.in_np[HW] = player & 0xFF
        self.in_np[HW + 1] = winner & 0xFF
        self.in_np[self.o_act: self.o_act + 4].view(np.int32)[0] = column
        self.in_dev.copy_(self.in_host, non_blocking=True)

    def fetch(self):
        """(grid bytes, player, winner, ended, legal, status, reward) of the last call (synchronises once)."""
        self.out_host.copy_(self.out_dev, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        o, HW, w = self.out_np, self.HW, self.o_words
        words = o[w: w + 16]
        return (o[:HW].view(np.int8).copy(), int(o[HW:HW + 1].view(np.int8)[0]), int(o[HW + 1:HW + 2].view(np.int8)[0]),
                bool(o[HW + 2]), int(words[0:4].view(np.uint32)[0]), int(words[4:8].view(np.int32)[0]),
                words[8:16].view(np.float32).copy())


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
            cfg = self.config
            H, W = cfg.height, cfg.width
            if not N.lib().bgs_connect_supported(H, W, cfg.count):
                raise RuntimeError(f"Connect {H}x{W} k={cfg.count} is not supported by the CUDA kernels")
            st = _OneState.get(H, W)
            with st.lock:
                st.load(self._grid, self._player, self._winner)
                i, o, HW = st.in_dev.data_ptr(), st.out_dev.data_ptr(), st.HW
                N.check(N.lib().bgs_connect_query(H, W, 1, i, i + HW + 1, o + HW + 2, o + st.o_words, o + st.o_words + 8,
                                                  N.stream_ptr(st.torch)))
                _, _, _, ended, legal, _, reward = st.fetch()
            self._info = (ended, legal, reward)
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
        s = self.state
        cfg = s.config
        H, W, K = cfg.height, cfg.width, cfg.count
        if not N.lib().bgs_connect_supported(H, W, K):
            raise RuntimeError(f"Connect {H}x{W} k={K} is not supported by the CUDA kernels")
        st = _OneState.get(H, W)
        with st.lock:
            st.load(s._grid, s._player, s._winner, self.column)
            i, o, HW = st.in_dev.data_ptr(), st.out_dev.data_ptr(), st.HW
            ow = o + st.o_words
            N.check(N.lib().bgs_connect_step(H, W, K, 1, i, i + HW, i + HW + 1, i + st.o_act, o, o + HW, o + HW + 1,
                                             o + HW + 2, ow + 8, ow, ow + 4, N.stream_ptr(st.torch)))
            grid, player, winner, ended, legal, status, reward = st.fetch()
        if status != 0:
            raise RuntimeError(f"illegal action: column {self.column}")
        return State(cfg, grid.reshape(H, W), player, winner, (ended, legal, reward))

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
