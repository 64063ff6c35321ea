"""Stub: generators.py:4 imports skimage.draw.circle and never calls it."""
