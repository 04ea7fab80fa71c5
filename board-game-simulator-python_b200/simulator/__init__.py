"""B200-native drop-in for the rollout hot path of ``jojolebarjos/board-game-simulator-python``.

``simulator.game.connect`` / ``simulator.game.bounce`` keep the reference's object API
(reference src/simulator/game/connect.cpp:24-61, bounce.cpp:24-60); ``simulator.batch`` is the new
batched entry point that runs millions of games per call on the GPU and returns torch tensors.
"""

__version__ = "0.1.0"
