"""No-op patches."""
class _P:
    def __init__(self, *a, **k): pass
    def __getattr__(self, n): return lambda *a, **k: 0.0
Rectangle = Circle = _P
