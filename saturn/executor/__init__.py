from saturn_b200.orchestrator import forecast  # noqa: F401


def __getattr__(name):
    if name == "execute":        # the Ray executor (saturn/executor/executor.py:88-129) is out of scope here
        from saturn_b200._alias import reference_attr
        return reference_attr(__name__, "executor.py", name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
