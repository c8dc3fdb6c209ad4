class SummaryWriter:
    """No-op stand-in (train.py:26 only)."""
    def __init__(self, *a, **k): pass
    def add_scalar(self, *a, **k): pass
    def close(self): pass
