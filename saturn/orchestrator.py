from saturn_b200.orchestrator import orchestrate  # noqa: F401
