"""No-op pyplot: every attribute is a function that does nothing (visualisation is out of scope)."""
def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
