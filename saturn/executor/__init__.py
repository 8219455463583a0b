from saturn_b200.orchestrator import forecast  # noqa: F401
