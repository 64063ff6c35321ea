"""Stub: the reference imports matplotlib for render() only (track_1v1.py:2-3,66-68)."""
from matplotlib import colors, pyplot
