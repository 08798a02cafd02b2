"""Minimal `keras` namespace: only what the H-DenseUNet scripts import."""
from . import backend  # noqa: F401
