from saturn_b200.solver import convert_into_comprehensible, solve  # noqa: F401
