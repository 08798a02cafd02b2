"""Shim for denseunet3d.py: denseunet_3d(args) (denseunet3d.py:393)."""
import _root  # noqa: F401
from h_denseunet_b200 import denseunet_3d, DenseNet3D  # noqa: F401
