"""Minimal stand-in for the slice of Ray the reference's import chain touches.

ORACLE INFRASTRUCTURE ONLY (see oracle/shims/pulp.py).  ORACLE_RAY_NODES nodes (default one) of
8 GPUs each — 8 per node is what the reference solver assumes anyway (milp.py:57-62, DEBUG = True).
"""


def is_initialized():
    return True


def init(*a, **k):
    return None


def nodes():
    import os
    n = int(os.environ.get("ORACLE_RAY_NODES", "1"))    # multi-node fixtures: N nodes x 8 GPUs
    return [{"Resources": {"GPU": 8, "CPU": 8}} for _ in range(n)]


def get(x):
    return x


def kill(x):
    return None


def get_gpu_ids():
    return []


class _Remote:
    def __init__(self, fn):
        self._fn = fn

    def options(self, **k):
        return self

    def remote(self, *a, **k):
        return self._fn(*a, **k)

    def __call__(self, *a, **k):
        return self._fn(*a, **k)


def remote(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return _Remote(args[0])

    def deco(fn):
        return _Remote(fn)
    return deco
