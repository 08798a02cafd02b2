"""Shim for preprocessing.py:7-85: the same function names on h_denseunet_b200.preprocessing (importing the reference module
runs its pipeline at import time, preprocessing.py:80-85; this shim only exposes the functions)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _root  # noqa: F401,E402
from h_denseunet_b200.preprocessing import (proprecessing, generate_livertxt, generate_tumortxt, generate_txt)  # noqa: F401,E402
