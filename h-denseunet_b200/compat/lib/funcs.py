"""Shim for lib/funcs.py:4 (predict_tumor_inwindow); the CPU post-processing helpers of that file are out of scope."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _root  # noqa: F401,E402
from h_denseunet_b200 import predict_tumor_inwindow  # noqa: F401,E402
