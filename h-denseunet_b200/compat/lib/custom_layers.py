"""Shim for lib/custom_layers.py:10."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _root  # noqa: F401,E402
from h_denseunet_b200 import Scale  # noqa: F401,E402
