"""Shim for loss.py:5,27."""
import _root  # noqa: F401
from h_denseunet_b200 import weighted_crossentropy, weighted_crossentropy_2ddense  # noqa: F401
