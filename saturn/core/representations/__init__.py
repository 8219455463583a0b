from saturn_b200.representations import HParams, Strategy, Task, Techniques  # noqa: F401
