"""Shim for the reference module of the same name: DenseUNet(reduction=0.5, args=args) as train_2ddense.py:178 calls it
(densenet.py:10 topology; pass skip=True for the denseunet.py:190-210 skip-add decoder)."""
import _root  # noqa: F401
from h_denseunet_b200 import DenseUNet  # noqa: F401
from h_denseunet_b200 import weighted_crossentropy_2ddense as weighted_crossentropy  # noqa: F401  (denseunet.py:108-127)
