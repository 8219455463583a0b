"""Minimal stand-in for the slice of Ray the reference's import chain touches.

ORACLE INFRASTRUCTURE ONLY (see oracle/shims/pulp.py).  One node, 8 GPUs —
which is what the reference solver assumes anyway (milp.py:57-62, DEBUG = True).
"""


def is_initialized():
    return True


def init(*a, **k):
    return None


def nodes():
    return [{"Resources": {"GPU": 8, "CPU": 8}}]


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
