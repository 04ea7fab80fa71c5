"""Oracle-backed stand-in for ``simulator.game`` (tests only): the ``connect`` and ``bounce`` submodules."""
