from saturn_b200.representations import HParams, Strategy, Task, Techniques  # noqa: F401


def __getattr__(name):
    # `search` / `execute` (the reference re-exports them from Strategy.py) belong to the trial runner /
    # Ray executor, which this drop-in does not replace
    if name in ("search", "execute"):
        from saturn_b200._alias import reference_attr
        return reference_attr(__name__, "Strategy.py", name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
