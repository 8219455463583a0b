"""Import-path alias: `import saturn` resolves to the B200-native implementation in saturn_b200.

Keeps the reference's entry points (saturn/__init__.py:1, saturn/solver/__init__.py:1-2,
saturn/core/representations/__init__.py:1-2) importable without PuLP, Ray or Gurobi.
"""
from saturn_b200.orchestrator import orchestrate  # noqa: F401
