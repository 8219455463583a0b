from saturn_b200._alias import fall_through as _fall_through

__path__ = _fall_through(__name__, __path__)   # saturn.core.executors lives in the reference only
