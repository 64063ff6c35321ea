"""Stub: environment.py:8-9 imports gym_unrealcv at module top; the 2D path never uses it."""
