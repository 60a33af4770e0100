"""Bundled example models (mirror of ``pyroboplan.models`` for the collision path)."""
