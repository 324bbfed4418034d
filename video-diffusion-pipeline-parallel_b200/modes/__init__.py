"""Command-line entry points with the reference's flags and output lines:
``python -m src.modes.simulator`` (reference src/modes/simulator.py) and ``python -m src.modes.benchmark``
(reference src/modes/benchmark.py), so the reference's shell scripts can drive this build."""
