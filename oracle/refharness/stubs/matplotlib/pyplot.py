def _unavailable(*a, **k):
    raise NotImplementedError("matplotlib stub: rendering is outside the oracle's scope")

subplots = close = draw = _unavailable
