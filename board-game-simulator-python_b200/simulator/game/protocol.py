"""Structural types every game module of this package satisfies.

These three names are the ones the reference re-exports from ``simulator.game``
(reference src/simulator/game/__init__.py:1, defined in protocol.py:8-29); agents and tools written
against the reference annotate with them, so they are provided here with the same attribute names.
They are ``runtime_checkable`` so that ``isinstance(state, StateLike)`` works on the GPU-backed
classes of ``simulator.game.connect`` and ``simulator.game.bounce``.
"""
from __future__ import annotations

from typing import ClassVar, Protocol, Sequence, runtime_checkable

import numpy as np


@runtime_checkable
class ActionLike(Protocol):
    """One legal move of one state.  Applying it yields a NEW state (states are immutable)."""

    @property
    def state(self) -> "StateLike":
        """The state this action belongs to."""

    def sample_next_state(self) -> "StateLike":
        """The successor state ("sample": a game may be stochastic; Connect and Bounce are not)."""


@runtime_checkable
class StateLike(Protocol):
    """A position: whose turn it is, whether the game is over, what it paid, what can be played."""

    Action: ClassVar[type]

    @property
    def config(self) -> "ConfigLike": ...

    @property
    def has_ended(self) -> bool: ...

    @property
    def player(self) -> int: ...

    @property
    def reward(self) -> np.ndarray:
        """One entry per player; meaningful once ``has_ended``."""

    @property
    def actions(self) -> Sequence[ActionLike]:
        """Every legal action; empty once ``has_ended``."""


@runtime_checkable
class ConfigLike(Protocol):
    """Game-wide parameters; the factory of initial states."""

    State: ClassVar[type]
    num_players: int

    def sample_initial_state(self) -> StateLike: ...
